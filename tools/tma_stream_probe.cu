// Streaming-rate probe: how fast can one persistent CTA per SM pull a [rows x W] u32 image
// through a shared-memory ring with 2-D tensor-map TMA boxes of a given geometry?
// Consumers do nothing (wait full -> arrive empty), so this is the ceiling for the mask scan's
// load side.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_stream_probe.cu
// Run:   ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(s32(dst)), "l"(map), "r"(x), "r"(y), "r"(s32(bar)) : "memory");
}

__device__ __forceinline__ void tma3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(s32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(s32(bar)) : "memory");
}

// tile = bx boxes side by side (each box_w x box_h); tiles enumerate (row block, column block)
// three_d: ONE 3-D box {32 px, box_h rows, bx strips} per tile over the view [W/32 strips][rows][32 px]
// (strip stride 128 B): same shared-memory layout as bx 2-D boxes of 32 x box_h, a single TMA instruction
__global__ void __launch_bounds__(64, 2) probe(const __grid_constant__ CUtensorMap map, int W, int rows, int box_w,
                                               int box_h, int bx, int stages, int three_d, long long* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int box_bytes = box_w * box_h * 4;
  const int tile_bytes = box_bytes * bx;
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * tile_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tiles_x = (W + box_w * bx - 1) / (box_w * bx);
  const int tiles_y = (rows + box_h - 1) / box_h;
  const long long total = (long long)tiles_x * tiles_y;
  const long long t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 32) {  // producer
    int it = 0;
    for (long long t = t0; t < t1; ++t, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
      mbar_expect(&full[s], tile_bytes);
      if (three_d)
        tma3d(smem + (size_t)s * tile_bytes, &map, 0, ty * box_h, tx * bx, &full[s]);
      else
        for (int b = 0; b < bx; ++b)
          tma2d(smem + (size_t)s * tile_bytes + (size_t)b * box_bytes, &map, (tx * bx + b) * box_w, ty * box_h, &full[s]);
    }
  } else if (threadIdx.x == 0) {  // consumer
    int it = 0; long long acc = 0;
    for (long long t = t0; t < t1; ++t, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[s], ph);
      acc += *(volatile int*)(smem + (size_t)s * tile_bytes);
      mbar_arrive(&empty[s]);
    }
    if (acc == 0x7fffffffffffll) *sink = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int W = 1920, H = 1080, B = 64;
  const long long rows = (long long)H * B;
  uint32_t* d; long long* sink;
  CK(cudaMalloc(&d, rows * W * 4)); CK(cudaMemset(d, 0, rows * W * 4)); CK(cudaMalloc(&sink, 8));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  EncodeFn encode = (EncodeFn)fn;
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  struct Cfg { int box_w, box_h, bx, stages, swz, three_d, ctas; } cfgs[] = {
      {32, 32, 16, 3, 1, 0, 1}, {32, 64, 8, 3, 1, 0, 1}, {32, 128, 4, 3, 1, 0, 1}, {32, 256, 2, 3, 1, 0, 1},
      {32, 64, 4, 6, 1, 0, 1}, {256, 32, 1, 4, 0, 0, 1}, {256, 64, 1, 3, 0, 0, 1}, {256, 16, 1, 8, 0, 0, 1},
      {128, 64, 2, 3, 0, 0, 1}, {64, 64, 4, 3, 0, 0, 1}, {32, 64, 8, 3, 0, 0, 1}, {256, 40, 1, 4, 0, 0, 1},
      {192, 64, 1, 4, 0, 0, 1}, {240, 32, 2, 3, 0, 0, 1}, {160, 64, 1, 5, 0, 0, 1},
      // the shipped geometry (2 CTAs per SM, 4 boxes of 32 x 64, 3 stages) and its single-instruction 3-D forms
      {32, 64, 4, 3, 1, 0, 2}, {32, 64, 4, 3, 1, 1, 2}, {32, 64, 8, 3, 1, 1, 1}, {32, 64, 8, 3, 1, 0, 1},
      {32, 128, 2, 3, 1, 1, 2}, {32, 32, 8, 3, 1, 1, 2}, {32, 64, 4, 3, 1, 0, 2}, {32, 64, 4, 3, 1, 1, 2}};
  for (auto c : cfgs) {
    CUtensorMap map;
    CUresult r;
    if (c.three_d) {
      cuuint64_t gdim[3] = {32, (cuuint64_t)rows, (cuuint64_t)(W / 32)};
      cuuint64_t gstr[2] = {(cuuint64_t)W * 4, 128};
      cuuint32_t box[3] = {32, (cuuint32_t)c.box_h, (cuuint32_t)c.bx};
      cuuint32_t estr[3] = {1, 1, 1};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)rows};
      cuuint64_t gstr[1] = {(cuuint64_t)W * 4};
      cuuint32_t box[2] = {(cuuint32_t)c.box_w, (cuuint32_t)c.box_h};
      cuuint32_t estr[2] = {1, 1};
      r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d for box %dx%d (3-D %d)\n", (int)r, c.box_w, c.box_h, c.three_d); continue; }
    const size_t smem = (size_t)c.stages * c.box_w * c.box_h * 4 * c.bx + 256;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) probe<<<sms * c.ctas, 64, smem>>>(map, W, (int)rows, c.box_w, c.box_h, c.bx, c.stages, c.three_d, sink);
    CK(cudaDeviceSynchronize());
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) probe<<<sms * c.ctas, 64, smem>>>(map, W, (int)rows, c.box_w, c.box_h, c.bx, c.stages, c.three_d, sink);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    printf("box %3d px x %3d rows, %2d boxes/tile (%6.1f KB tile), %d stages, swz %d, %s, %d CTA/SM: %.4f ms  %.0f GB/s\n",
           c.box_w, c.box_h, c.bx, c.box_w * c.box_h * 4 * c.bx / 1024.0, c.stages, c.swz,
           c.three_d ? "one 3-D box" : "2-D boxes", c.ctas, ms, rows * W * 4.0 / ms / 1e6);
  }
  return 0;
}
