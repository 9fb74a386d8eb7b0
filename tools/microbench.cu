// microbench.cu -- B200 memory-system numbers that bound the H*v kernel design:
// HBM read / copy bandwidth, L2-resident read bandwidth, shared-memory gather throughput (fp64).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_read(const double2 *__restrict__ x, size_t n2, double *out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 v = x[i]; s += v.x + v.y;
  }
  if (s == 1.2345e-300) out[0] = s;
}
__global__ void k_copy(const double2 *__restrict__ x, double2 *__restrict__ y, size_t n2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) y[i] = x[i];
}
// each block re-reads its own slice of a small (L2-resident) buffer `reps` times
__global__ void k_l2(const double2 *__restrict__ x, size_t n2, int reps, double *out) {
  double s = 0;
  for (int r = 0; r < reps; r++)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
      double2 v = __ldcg(&x[(i + (size_t)r * 977) % n2]); s += v.x + v.y;
    }
  if (s == 1.2345e-300) out[0] = s;
}
// shared-memory gathers: pattern 0 = contiguous (conflict-free), 1 = pseudo-random indices
__global__ void k_smem(const int *__restrict__ idx, int n, int reps, int pattern, double *out) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = i * 0.5;
  __syncthreads();
  double s = 0;
  int base = threadIdx.x;
  for (int r = 0; r < reps; r++) {
#pragma unroll 8
    for (int k = 0; k < 8; k++) {
      int j = pattern ? idx[(base + k * 1024 + r * 31) % n] : (base + k * 1031 + r * 17) % n;
      s += sm[j];
    }
  }
  if (s == 1.2345e-300) out[0] = s;
}
__global__ void k_smem_reg(int n, int reps, int pattern, double *out) {
  // indices generated arithmetically (no index loads): isolates LDS.64 gather throughput
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = i * 0.5;
  __syncthreads();
  double s = 0;
  unsigned h = threadIdx.x * 2654435761u;
  for (int r = 0; r < reps; r++) {
#pragma unroll 8
    for (int k = 0; k < 8; k++) {
      int j;
      if (pattern == 0) j = (threadIdx.x + k * 1024 + r) % n;
      else { h = h * 1664525u + 1013904223u; j = (h >> 8) % n; }
      s += sm[j];
    }
  }
  if (s == 1.2345e-300) out[0] = s;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm %d.%d SMs %d L2 %d MB smem/block optin %zu\n", p.name, p.major, p.minor, p.multiProcessorCount, p.l2CacheSize >> 20, p.sharedMemPerBlockOptin);
  size_t n = (size_t)1 << 29;   // 4 GiB of doubles
  double *x, *y, *out; CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8)); CK(cudaMalloc(&out, 8));
  CK(cudaMemset(x, 0, n * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
    for (int it = 0; it < 3; it++) { cudaEventRecord(e0); k_read<<<blocks, 512>>>((double2 *)x, n / 2, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("HBM read  blocks %5d: %.1f GB/s\n", blocks, n * 8 / ms / 1e6);
    for (int it = 0; it < 3; it++) { cudaEventRecord(e0); k_copy<<<blocks, 512>>>((double2 *)x, (double2 *)y, n / 2); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("HBM copy  blocks %5d: %.1f GB/s (read+write)\n", blocks, 2 * n * 8 / ms / 1e6);
  }
  for (size_t mb : {8, 16, 32, 64, 96, 128, 256}) {
    size_t m = mb << 20 >> 3; int reps = 40;
    for (int it = 0; it < 2; it++) { cudaEventRecord(e0); k_l2<<<148 * 8, 512>>>((double2 *)x, m / 2, reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    printf("L2 read   %4zu MB buffer: %.1f GB/s\n", mb, (double)m * 8 * reps / ms / 1e6);
  }
  int nsm = 24 * 1024; int *idx; CK(cudaMalloc(&idx, nsm * 4));
  { int *h = (int *)malloc(nsm * 4); unsigned s = 12345; for (int i = 0; i < nsm; i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % nsm; } cudaMemcpy(idx, h, nsm * 4, cudaMemcpyHostToDevice); free(h); }
  CK(cudaFuncSetAttribute(k_smem_reg, cudaFuncAttributeMaxDynamicSharedMemorySize, nsm * 8));
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, nsm * 8));
  for (int pattern = 0; pattern < 2; pattern++) {
    int reps = 4000;
    for (int it = 0; it < 2; it++) { cudaEventRecord(e0); k_smem_reg<<<148, 1024, nsm * 8>>>(nsm, reps, pattern, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    double g = 148.0 * 1024 * reps * 8;
    printf("smem LDS.64 gather pattern %d (%s): %.2f Ggather/s chip, %.2f TB/s, %.2f gathers/clk/SM @1.9GHz\n", pattern, pattern ? "random" : "contiguous", g / ms / 1e6, g * 8 / ms / 1e9, g / ms / 1e6 / 148 / 1.9);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
