// Does a PDL secondary start while the primary (persistent, trigger at entry) is still running?
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(288, 2) primary(unsigned long long* t, int spin_us, int trigger) {
  extern __shared__ unsigned char sm[];
  if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  unsigned long long t0 = gtime();
  if (blockIdx.x == 0 && threadIdx.x == 0) t[0] = t0;
  while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) { sm[threadIdx.x] = (unsigned char)t0; }
  if (blockIdx.x == 0 && threadIdx.x == 0) t[1] = gtime();
}
__global__ void secondary(unsigned long long* t, int wait_at_end) {
  unsigned long long t0 = gtime();
  if (blockIdx.x == 0 && threadIdx.x == 0) t[2] = t0;
  while (gtime() - t0 < 5000ull) {}
  if (blockIdx.x == 0 && threadIdx.x == 0) t[3] = gtime();
  if (wait_at_end) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) t[4] = gtime();
}
int main() {
  unsigned long long *d, h[5];
  cudaMalloc(&d, 40);
  int smem = 100 * 1024;
  cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int carve = 0; carve < 2; ++carve)
  for (int trig = 0; trig < 2; ++trig) {
    if (carve) cudaFuncSetAttribute(secondary, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d, 0, 40);
      primary<<<296, 288, smem>>>(d, 100, trig);
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(100); cfg.blockDim = dim3(64); cfg.stream = 0;
      cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = a; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, secondary, d, 1);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
      printf("carve=%d trigger=%d: primary [0, %.1f] us; secondary start %.1f, work done %.1f, exit %.1f (%s)\n", carve, trig,
             (h[1] - h[0]) / 1e3, ((long long)h[2] - (long long)h[0]) / 1e3, ((long long)h[3] - (long long)h[0]) / 1e3,
             ((long long)h[4] - (long long)h[0]) / 1e3, cudaGetErrorString(e));
    }
  }
  return 0;
}
