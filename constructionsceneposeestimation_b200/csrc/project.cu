// K2 — per-object transform, 3D-box corner projection and poses.
//
// Replaces the reference's per-object Python loop (gcd.py:1924-1950) and its callee
// bboxDict_to_transform (gcd.py:553-584): world centre, world size, Euler xyz of the pure
// rotation — and adds what the reference leaves out: the 8 projected box corners and the
// object-in-camera 6-DoF pose.
//
// One thread per (frame, slot).  This is 4x4-transform work on <= a few thousand objects: FP64
// FMA-free scalar math (the library is built with -fmad=false so every operation rounds exactly
// like the numpy oracle's), no tensor cores.  Latency bound; amortised by batching frames into
// one launch and hidden behind the mask scan on a side stream.
#include <math.h>

#include "cspe_common.cuh"

namespace cspe {
namespace {

struct Mat3 {
  double m[3][3];
};

__device__ __forceinline__ double det3(const Mat3& a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// cofactor matrix C with inverse-transpose = C / det
__device__ __forceinline__ void cofactor3(const Mat3& a, Mat3& c) {
  c.m[0][0] = a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1];
  c.m[0][1] = a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2];
  c.m[0][2] = a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0];
  c.m[1][0] = a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2];
  c.m[1][1] = a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0];
  c.m[1][2] = a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1];
  c.m[2][0] = a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1];
  c.m[2][1] = a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2];
  c.m[2][2] = a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0];
}

__device__ __forceinline__ double frob2(const Mat3& a) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) s += a.m[i][j] * a.m[i][j];
  return s;
}

// Orthogonal polar factor of a (== U @ Vt of its SVD, gcd.py:573-574) by scaled Newton
// iteration X <- (g X + X^-T / g) / 2.  Returns false when a is singular / non-finite.
__device__ bool polar3(const Mat3& a, Mat3& x) {
  x = a;
  for (int it = 0; it < 32; ++it) {
    const double d = det3(x);
    if (!(fabs(d) > 0.0) || !isfinite(d)) return false;
    Mat3 c;
    cofactor3(x, c);
    const double inv_d = 1.0 / d;
    const double nx = frob2(x);
    double ny = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        c.m[i][j] *= inv_d;
        ny += c.m[i][j] * c.m[i][j];
      }
    if (!(nx > 0.0) || !isfinite(ny)) return false;
    const double g = sqrt(sqrt(ny / nx));
    const double ig = 1.0 / g;
    double diff = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double nv = 0.5 * (g * x.m[i][j] + ig * c.m[i][j]);
        const double dl = nv - x.m[i][j];
        diff += dl * dl;
        x.m[i][j] = nv;
      }
    if (diff <= 1e-30 * 3.0) break;  // ||dX||_F <= 1e-15 * ||Q||_F (Q orthogonal: ||Q||_F^2 = 3)
  }
  return true;
}

// scipy Rotation.as_euler('xyz', degrees=True) (extrinsic), incl. its gimbal-lock rule
// (third angle := 0 when |cos b| <= 1e-7); gcd.py:576.
__device__ void euler_xyz_deg(const Mat3& r, double* e) {
  const double kRad2Deg = 57.295779513082320876798154814105;
  const double cb = hypot(r.m[0][0], r.m[1][0]);
  const double b = atan2(-r.m[2][0], cb);
  double a, c;
  if (cb > 1e-7) {
    a = atan2(r.m[2][1], r.m[2][2]);
    c = atan2(r.m[1][0], r.m[0][0]);
  } else {
    c = 0.0;
    a = r.m[2][0] < 0.0 ? atan2(r.m[0][1], r.m[1][1]) : atan2(-r.m[0][1], r.m[1][1]);
  }
  e[0] = a * kRad2Deg;
  e[1] = b * kRad2Deg;
  e[2] = c * kRad2Deg;
}

// rotation matrix -> unit quaternion xyzw with w >= 0 (Shepperd's branch on the largest term)
__device__ void quat_xyzw(const Mat3& r, double* q) {
  const double t = r.m[0][0] + r.m[1][1] + r.m[2][2];
  double x, y, z, w;
  if (t >= r.m[0][0] && t >= r.m[1][1] && t >= r.m[2][2]) {
    w = 1.0 + t;
    x = r.m[2][1] - r.m[1][2];
    y = r.m[0][2] - r.m[2][0];
    z = r.m[1][0] - r.m[0][1];
  } else if (r.m[0][0] >= r.m[1][1] && r.m[0][0] >= r.m[2][2]) {
    x = 1.0 - t + 2.0 * r.m[0][0];
    y = r.m[1][0] + r.m[0][1];
    z = r.m[2][0] + r.m[0][2];
    w = r.m[2][1] - r.m[1][2];
  } else if (r.m[1][1] >= r.m[2][2]) {
    x = r.m[1][0] + r.m[0][1];
    y = 1.0 - t + 2.0 * r.m[1][1];
    z = r.m[2][1] + r.m[1][2];
    w = r.m[0][2] - r.m[2][0];
  } else {
    x = r.m[2][0] + r.m[0][2];
    y = r.m[2][1] + r.m[1][2];
    z = 1.0 - t + 2.0 * r.m[2][2];
    w = r.m[1][0] - r.m[0][1];
  }
  const double n = sqrt(x * x + y * y + z * z + w * w);
  double s = 1.0 / n;
  if (w < 0.0) s = -s;
  q[0] = x * s;
  q[1] = y * s;
  q[2] = z * s;
  q[3] = w * s;
}

// One THREAD per (frame, slot), 64-thread blocks: the work per object is a serial FP64 chain
// (polar iteration, atan2, sqrt), so latency — not throughput — sets the kernel time; small
// blocks with a modest register footprint co-reside with the persistent mask-scan CTA on every
// SM when the caller forks K2 onto a side stream (pipeline.py), which hides K2 entirely.
constexpr int kProjThreads = 64;

__device__ __forceinline__ void project_one(const unsigned char* __restrict__ records, int rec_stride, int recs_per_frame,
                                            const int32_t* __restrict__ obj_record, const double* __restrict__ cam,
                                            long long obj, int N, double* __restrict__ uv, double* __restrict__ zc_out,
                                            double* __restrict__ pose, double* __restrict__ loose,
                                            uint8_t* __restrict__ flags);

__global__ void __launch_bounds__(kProjThreads)
    project_objects_kernel(const unsigned char* __restrict__ records, int rec_stride, int recs_per_frame,
                           const int32_t* __restrict__ obj_record, const double* __restrict__ cam, long long total, int N,
                           double* __restrict__ uv, double* __restrict__ zc_out, double* __restrict__ pose,
                           double* __restrict__ loose, uint8_t* __restrict__ flags, int overlap_previous) {
  pdl_launch_dependents();  // let the next kernel (K4) get its blocks placed; it waits for us before reading
  // default: behave like an ordinary serialised launch (inputs may come from the previous kernel)
  if (!overlap_previous) pdl_wait();
  const long long obj = static_cast<long long>(blockIdx.x) * kProjThreads + threadIdx.x;
  if (obj < total) project_one(records, rec_stride, recs_per_frame, obj_record, cam, obj, N, uv, zc_out, pose, loose, flags);
  // overlap mode: the grid may have started while the kernel before it in the stream (the mask
  // scan) was still running and nothing above depends on that kernel; waiting HERE makes this
  // grid's completion imply its completion, so plain stream order still holds for whatever is
  // queued next (K4 reads both).
  if (overlap_previous) pdl_wait();
}

__device__ __forceinline__ void project_one(const unsigned char* __restrict__ records, int rec_stride, int recs_per_frame,
                                            const int32_t* __restrict__ obj_record, const double* __restrict__ cam,
                                            long long obj, int N, double* __restrict__ uv, double* __restrict__ zc_out,
                                            double* __restrict__ pose, double* __restrict__ loose,
                                            uint8_t* __restrict__ flags) {
  const int frame = static_cast<int>(obj / N);
  const double kNaN = __longlong_as_double(0x7ff8000000000000ll);
  double* po = pose + obj * CSPE_POSE_STRIDE;

  // bit 30 of the record index marks a record that only approximates the object (a mesh record standing in
  // for an object whose root prim has none): passed through to the flags, stripped from the index
  // in overlapped mode every input is read before the PDL wait: L1 bypass (cspe_common.cuh, PDL rule)
  const int rec_raw = __ldcg(obj_record + obj);
  const bool approx = rec_raw >= 0 && (rec_raw & CSPE_OBJ_RECORD_APPROX_BIT);
  const int rec = rec_raw >= 0 ? (rec_raw & ~CSPE_OBJ_RECORD_APPROX_BIT) : rec_raw;
  if (rec < 0 || rec >= recs_per_frame) {
    // no record for this slot: defined outputs, flags 0
#pragma unroll
    for (int i = 0; i < 16; ++i) uv[obj * 16 + i] = kNaN;
#pragma unroll
    for (int i = 0; i < 8; ++i) zc_out[obj * 8 + i] = kNaN;
#pragma unroll
    for (int i = 0; i < CSPE_POSE_STRIDE; ++i) po[i] = kNaN;
#pragma unroll
    for (int i = 0; i < 4; ++i) loose[obj * 4 + i] = kNaN;
    flags[obj] = 0;
    return;
  }

  const float* rp = reinterpret_cast<const float*>(records + (static_cast<long long>(frame) * recs_per_frame + rec) *
                                                                 static_cast<long long>(rec_stride));
  // [0] semanticId, [1..6] extents, [7..22] transform (row-vector convention), [23] occlusionRatio
  const float ext_min[3] = {__ldcg(rp + 1), __ldcg(rp + 2), __ldcg(rp + 3)};
  const float ext_max[3] = {__ldcg(rp + 4), __ldcg(rp + 5), __ldcg(rp + 6)};
  float T32[4][3];
  double T[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T32[i][j] = __ldcg(rp + 7 + i * 4 + j);
      T[i][j] = static_cast<double>(T32[i][j]);
    }

  double cm[17];
#pragma unroll
  for (int j = 0; j < 17; ++j) cm[j] = __ldcg(cam + static_cast<long long>(frame) * CSPE_CAM_STRIDE + j);
  const double t[3] = {cm[0], cm[1], cm[2]};
  double Rcw[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rcw[i][j] = cm[3 + i * 3 + j];
  const double fx = cm[12], fy = cm[13], cx0 = cm[14], cy0 = cm[15], nearc = cm[16];

  // ---- the 8 corners: bit0 -> x, bit1 -> y, bit2 -> z picks max over min ----
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
  double umin = kInf, vmin = kInf, umax = -kInf, vmax = -kInf;
  unsigned fm = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double c[3] = {static_cast<double>((k & 1) ? ext_max[0] : ext_min[0]),
                         static_cast<double>((k & 2) ? ext_max[1] : ext_min[1]),
                         static_cast<double>((k & 4) ? ext_max[2] : ext_min[2])};
    double d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double pw = ((c[0] * T[0][j] + c[1] * T[1][j]) + c[2] * T[2][j]) + T[3][j];
      d[j] = pw - t[j];
    }
    double pc[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) pc[i] = (Rcw[0][i] * d[0] + Rcw[1][i] * d[1]) + Rcw[2][i] * d[2];
    const double z = -pc[2];
    const double u = cx0 + (fx * pc[0]) / z;
    const double v = cy0 - (fy * pc[1]) / z;
    uv[obj * 16 + k * 2 + 0] = u;
    uv[obj * 16 + k * 2 + 1] = v;
    zc_out[obj * 8 + k] = z;
    if (z > nearc) {
      fm |= 1u << k;
      umin = fmin(umin, u);
      vmin = fmin(vmin, v);
      umax = fmax(umax, u);
      vmax = fmax(vmax, v);
    }
  }
  if (fm) {
    loose[obj * 4 + 0] = umin;
    loose[obj * 4 + 1] = vmin;
    loose[obj * 4 + 2] = umax;
    loose[obj * 4 + 3] = vmax;
  } else {
    loose[obj * 4 + 0] = loose[obj * 4 + 1] = loose[obj * 4 + 2] = loose[obj * 4 + 3] = kNaN;
  }

  // ---- pose ----
  // gcd.py:566: centre of the local box, float32 mean like np.mean on the f32 corner array
  double cl[3], cw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) cl[i] = static_cast<double>(__fmul_rn(__fadd_rn(ext_min[i], ext_max[i]), 0.5f));
#pragma unroll
  for (int j = 0; j < 3; ++j) cw[j] = ((cl[0] * T[0][j] + cl[1] * T[1][j]) + cl[2] * T[2][j]) + T[3][j];
  // gcd.py:578-582: size_world = ||rot[:,k]|| * |max - min| (f32 norm, f32 difference)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a0 = T32[i][0], a1 = T32[i][1], a2 = T32[i][2];
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2));
    const float sc = __fsqrt_rn(n2);
    const float ext = fabsf(__fsub_rn(ext_max[i], ext_min[i]));
    po[10 + i] = static_cast<double>(sc) * static_cast<double>(ext);
  }
  // rot = M[:3,:3] with M = T^T (gcd.py:568,572): rot[a][b] = T[b][a]
  Mat3 rot, rwo;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) rot.m[a][b] = T[b][a];
  const double dr = det3(rot);
  bool pose_ok = isfinite(dr) && dr > 0.0 && polar3(rot, rwo);
  pose_ok = pose_ok && isfinite(cw[0]) && isfinite(cw[1]) && isfinite(cw[2]);

  const double d0 = cw[0] - t[0], d1 = cw[1] - t[1], d2 = cw[2] - t[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) po[i] = (Rcw[0][i] * d0 + Rcw[1][i] * d1) + Rcw[2][i] * d2;
  if (pose_ok) {
    Mat3 rco;  // Rcw^T * R_wo
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        rco.m[i][j] = (Rcw[0][i] * rwo.m[0][j] + Rcw[1][i] * rwo.m[1][j]) + Rcw[2][i] * rwo.m[2][j];
    quat_xyzw(rco, po + 3);
    euler_xyz_deg(rwo, po + 13);
  } else {
    po[3] = po[4] = po[5] = po[6] = kNaN;
    po[13] = po[14] = po[15] = kNaN;
  }
  po[7] = cw[0];
  po[8] = cw[1];
  po[9] = cw[2];

  uint8_t fl = CSPE_OBJ_HAS_RECORD;
  if (fm) fl |= CSPE_OBJ_ANY_FRONT;
  if (fm == 0xffu) fl |= CSPE_OBJ_ALL_FRONT;
  if (pose_ok) fl |= CSPE_OBJ_POSE_VALID;
  if (approx) fl |= CSPE_OBJ_APPROX_RECORD;
  flags[obj] = fl;
}


// ---- object-level records for multi-mesh objects (record_fallback = "union") ------------------------------------
// The reference reads the bound of an object whose root prim has no bbox3d record from the live USD stage
// (UsdGeom.BBoxCache.ComputeWorldBound(...).ComputeAlignedRange(), gcd.py:2000-2009): the world-axis-aligned range of
// everything under the prim.  Outside Isaac Sim the same range is the min / max of the eight world-space corners of every
// mesh record of the object.  One thread per (frame, union object): corners exactly as K2 computes them
// (p_w = ((c0*T0j + c1*T1j) + c2*T2j) + T3j in double, -fmad=false), fmin / fmax over them (NaN corners drop out), the
// range rounded to float32 and written as a record with an identity transform at records[frame][base + u] — K2 then
// treats it like any other record (centre = middle of the range, size = its extent, rotation 0).
__global__ void union_records_kernel(unsigned char* __restrict__ records, int rec_stride, int recs_per_frame, int base,
                                     const int32_t* __restrict__ offsets, long long off_stride,
                                     const int32_t* __restrict__ members, long long mem_stride, int B, int U) {
  // launched as an ordinary kernel (no programmatic attribute, no early release of dependents): whatever follows in
  // the stream — also a kernel that starts early behind the mask scan — sees these records complete
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * U) return;
  const int frame = static_cast<int>(i / U), u = static_cast<int>(i % U);
  const int32_t* off = offsets + frame * off_stride;
  const int32_t* mem = members + frame * mem_stride;
  const int lo = off[u], hi = off[u + 1];
  unsigned char* frame_recs = records + static_cast<long long>(frame) * recs_per_frame * static_cast<long long>(rec_stride);
  float* out = reinterpret_cast<float*>(frame_recs + static_cast<long long>(base + u) * rec_stride);
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);
  double mn[3] = {kInf, kInf, kInf}, mx[3] = {-kInf, -kInf, -kInf};
  float sem = 0.0f, occ = 0.0f;
  bool first = true;
  for (int m = lo; m < hi; ++m) {
    const int rec = mem[m];
    if (rec < 0 || rec >= base) continue;
    const float* rp = reinterpret_cast<const float*>(frame_recs + static_cast<long long>(rec) * rec_stride);
    if (first) {
      sem = rp[0];   // semanticId bits / occlusionRatio of the first mesh record travel with the object
      occ = rp[23];
      first = false;
    }
    const float e0[3] = {rp[1], rp[2], rp[3]}, e1[3] = {rp[4], rp[5], rp[6]};
    double T[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < 3; ++j) T[a][j] = static_cast<double>(rp[7 + a * 4 + j]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double c0 = static_cast<double>((k & 1) ? e1[0] : e0[0]), c1 = static_cast<double>((k & 2) ? e1[1] : e0[1]),
                   c2 = static_cast<double>((k & 4) ? e1[2] : e0[2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double pw = ((c0 * T[0][j] + c1 * T[1][j]) + c2 * T[2][j]) + T[3][j];
        mn[j] = fmin(mn[j], pw);
        mx[j] = fmax(mx[j], pw);
      }
    }
  }
  const float kNaNf = __int_as_float(0x7fc00000);
  const bool any = mn[0] <= mx[0] && mn[1] <= mx[1] && mn[2] <= mx[2];   // false when no finite corner was seen
  out[0] = sem;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    out[1 + j] = any ? __double2float_rn(mn[j]) : kNaNf;
    out[4 + j] = any ? __double2float_rn(mx[j]) : kNaNf;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[7 + a * 4 + j] = (a == j) ? 1.0f : 0.0f;
  out[23] = occ;
}

}  // namespace
}  // namespace cspe

using namespace cspe;

static int project_objects_impl(const void* records, int rec_stride, int recs_per_frame, const int32_t* obj_record,
                                const double* cam, int B, int N, double* uv, double* z, double* pose, double* loose,
                                uint8_t* flags, void* stream, int overlap_previous) {
  CSPE_REQUIRE(B >= 0 && N >= 0 && recs_per_frame >= 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_project_objects: negative size (B=%d N=%d recs_per_frame=%d)", B, N, recs_per_frame);
  if (B == 0 || N == 0) return CSPE_OK;
  CSPE_REQUIRE(rec_stride >= CSPE_BBOX3D_RECORD_BYTES && rec_stride % 4 == 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_project_objects: rec_stride %d (need >= %d and a multiple of 4)", rec_stride,
               CSPE_BBOX3D_RECORD_BYTES);
  CSPE_REQUIRE(obj_record && cam && uv && z && pose && loose && flags, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_project_objects: null pointer");
  CSPE_REQUIRE(recs_per_frame == 0 || records != nullptr, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_project_objects: records is null");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 3) == 0 && (reinterpret_cast<uintptr_t>(cam) & 7) == 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_project_objects: misaligned records/cam");
  const long long total = static_cast<long long>(B) * N;
  const long long blocks = (total + kProjThreads - 1) / kProjThreads;
  CSPE_REQUIRE(blocks < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_project_objects: too many objects");
  // Same shared-memory carve-out as the mask scan (max shared): an SM cannot host kernels with
  // different L1/shared splits at the same time, so without this K2 would wait for K1's persistent
  // CTAs to leave instead of running beside them on the caller's side stream.
  static const cudaError_t carve = cudaFuncSetAttribute(project_objects_kernel,
                                                        cudaFuncAttributePreferredSharedMemoryCarveout,
                                                        cudaSharedmemCarveoutMaxShared);
  (void)carve;
  // Programmatic stream serialisation: if the kernel before this one in `stream` releases its
  // dependents early (the mask scan does, right after its CTAs are resident), K2 starts beside it;
  // after any other kernel this is an ordinary serialised launch.
  CSPE_CUDA_OK(launch_pdl(project_objects_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kProjThreads), 0,
                          static_cast<cudaStream_t>(stream), static_cast<const unsigned char*>(records), rec_stride,
                          recs_per_frame, obj_record, cam, total, N, uv, z, pose, loose, flags, overlap_previous));
  return CSPE_OK;
}

extern "C" int cspe_project_objects(const void* records, int rec_stride, int recs_per_frame,
                                    const int32_t* obj_record, const double* cam, int B, int N, double* uv,
                                    double* z, double* pose, double* loose, uint8_t* flags, void* stream) {
  return project_objects_impl(records, rec_stride, recs_per_frame, obj_record, cam, B, N, uv, z, pose, loose, flags,
                              stream, 0);
}

extern "C" int cspe_project_objects_overlapped(const void* records, int rec_stride, int recs_per_frame,
                                               const int32_t* obj_record, const double* cam, int B, int N, double* uv,
                                               double* z, double* pose, double* loose, uint8_t* flags, void* stream) {
  return project_objects_impl(records, rec_stride, recs_per_frame, obj_record, cam, B, N, uv, z, pose, loose, flags,
                              stream, 1);
}

extern "C" int cspe_union_records(void* records, int rec_stride, int recs_per_frame, int base, const int32_t* offsets,
                                  int64_t offsets_stride, const int32_t* members, int64_t members_stride, int B, int U,
                                  void* stream) {
  CSPE_REQUIRE(B >= 0 && U >= 0 && base >= 0 && recs_per_frame >= 0 && offsets_stride >= 0 && members_stride >= 0,
               CSPE_ERR_INVALID_ARGUMENT, "cspe_union_records: negative size (B=%d U=%d base=%d recs_per_frame=%d)", B, U,
               base, recs_per_frame);
  if (B == 0 || U == 0) return CSPE_OK;
  CSPE_REQUIRE(rec_stride >= CSPE_BBOX3D_RECORD_BYTES && rec_stride % 4 == 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_union_records: rec_stride %d (need >= %d and a multiple of 4)", rec_stride, CSPE_BBOX3D_RECORD_BYTES);
  CSPE_REQUIRE(static_cast<long long>(base) + U <= recs_per_frame, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_union_records: base %d + U %d exceeds recs_per_frame %d", base, U, recs_per_frame);
  CSPE_REQUIRE(records && offsets && members, CSPE_ERR_INVALID_ARGUMENT, "cspe_union_records: null pointer");
  CSPE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 3) == 0, CSPE_ERR_INVALID_ARGUMENT,
               "cspe_union_records: misaligned records");
  const long long total = static_cast<long long>(B) * U;
  const int threads = 64;
  const long long blocks = (total + threads - 1) / threads;
  CSPE_REQUIRE(blocks < (1ll << 31), CSPE_ERR_UNSUPPORTED, "cspe_union_records: too many objects");
  union_records_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned char*>(records), rec_stride, recs_per_frame, base, offsets, offsets_stride, members,
      members_stride, B, U);
  CSPE_LAUNCH_OK("union_records_kernel");
  return CSPE_OK;
}
