/* C restatement of the mask-scan definition — TEST INFRASTRUCTURE.
 *
 * The reference leaves the instance mask as a -1 placeholder (gcd.py:1908-1910); the scan
 * is defined by the numpy oracle (oracle/labels.py: per slot, the pixels whose id maps to it,
 * count and inclusive extents, absent -> {0, W, H, -1, -1}).  This file states the same thing
 * as plain loops so the CUDA kernel is checked against two independently written CPU
 * implementations, and it is fast enough to verify full-size batches.
 */
#include <stdint.h>

void oracle_mask_scan(const uint32_t *mask, int B, int H, int W, const int32_t *id2slot, int lut_len,
                      int64_t lut_stride, int N, int32_t *out /* [B][N][5] */) {
    for (int b = 0; b < B; ++b) {
        int32_t *o = out + (int64_t)b * N * 5;
        const int32_t *lut = id2slot + (int64_t)b * lut_stride;
        for (int n = 0; n < N; ++n) {
            o[n * 5 + 0] = 0;
            o[n * 5 + 1] = W;
            o[n * 5 + 2] = H;
            o[n * 5 + 3] = -1;
            o[n * 5 + 4] = -1;
        }
        const uint32_t *m = mask + (int64_t)b * H * W;
        for (int y = 0; y < H; ++y) {
            for (int x = 0; x < W; ++x) {
                uint32_t id = m[(int64_t)y * W + x];
                if (id >= (uint32_t)lut_len) continue;
                int32_t s = lut[id];
                if (s < 0 || s >= N) continue;
                int32_t *e = o + s * 5;
                e[0] += 1;
                if (x < e[1]) e[1] = x;
                if (y < e[2]) e[2] = y;
                if (x > e[3]) e[3] = x;
                if (y > e[4]) e[4] = y;
            }
        }
    }
}
