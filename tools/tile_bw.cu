// tile_bw.cu -- how fast can row-block tiles of a column-major double matrix be streamed through
// shared memory with 2-D TMA boxes (R rows x BC columns), in the item order k_srow uses?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tile_bw.bin tile_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}
// items: (row block rb, column chunk ci), rb-major like k_srow; each item = R rows x C columns, loaded as C/BC boxes.
// mode 0: load only; mode 1: also read y (R x C, coalesced R-row segments) and write y.
__global__ void __launch_bounds__(512, 1) k_stream(const __grid_constant__ CUtensorMap tm, int n, int nf, int R, int BC, int C, int mode,
                                                   const double *__restrict__ yin, double *__restrict__ yout, double *sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  double *t0 = (double *)sm; const int tsz = R * C;
  uint64_t *bar = (uint64_t *)(t0 + 2 * tsz);
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int nrb = (n + R - 1) / R, nch = (nf + C - 1) / C;
  const long nitems = (long)nrb * nch, G = gridDim.x;
  const int nops = C / BC;
  auto issue = [&](long t, int b) {
    const int rb = (int)(t / nch), ci = (int)(t % nch);
    if (lane == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect(&bar[b], (uint32_t)nops * R * BC * 8); }
    __syncwarp();
    for (int q = lane; q < nops; q += 32) tma2d(t0 + (size_t)b * tsz + (size_t)q * R * BC, &tm, rb * R, ci * C + q * BC, &bar[b]);
  };
  if (tid < 32) for (int b = 0; b < 2; b++) { long t = blockIdx.x + b * G; if (t < nitems) issue(t, b); }
  double s = 0;
  for (long it = 0;; it++) {
    const long t = blockIdx.x + it * G;
    if (t >= nitems) break;
    const int b = it & 1;
    mbar_wait(&bar[b], (it >> 1) & 1);
    const double *tl = t0 + (size_t)b * tsz;
    const int rb = (int)(t / nch), ci = (int)(t % nch);
    // consume: every thread touches R*C/512 tile values
    for (int i = tid; i < tsz; i += 512) s += tl[i];
    if (mode == 1) {
      const int rpw = R;                       // rows per segment
      for (int i = tid; i < tsz; i += 512) {
        const int c = i / rpw, r = i % rpw;
        const long col = (long)ci * C + c, row = (long)rb * R + r;
        if (col < nf && row < n) { const size_t o = (size_t)col * n + row; yout[o] = yin[o] + tl[i]; }
      }
    }
    __syncthreads();
    const long t2 = t + 2 * G;
    if (tid < 32 && t2 < nitems) issue(t2, b);
  }
  if (s == 1.2345e-300) sink[0] = s;
}
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int n = 12870, nf = 12870;
  const size_t tot = (size_t)n * nf;
  double *x, *y, *y2, *sink; CK(cudaMalloc(&x, tot * 8 + 4096)); CK(cudaMalloc(&y, tot * 8)); CK(cudaMalloc(&y2, tot * 8)); CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(x, 0, tot * 8)); CK(cudaMemset(y, 0, tot * 8));
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q)); PFN enc = (PFN)p;
  CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  struct Cfg { int R, BC, C; } cfgs[] = {{32, 32, 352}, {64, 32, 160}, {128, 16, 80}, {16, 64, 704}, {32, 32, 160}, {256, 8, 40}, {32, 32, 256}, {16, 32, 640}, {16, 32, 320}, {8, 32, 1280}, {8, 64, 1280}, {64, 32, 96}, {64, 16, 128}};
  for (auto cf : cfgs) for (int mode = 0; mode < 2; mode++) {
    CUtensorMap tm; cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)nf}, str[1] = {(cuuint64_t)n * 8}; cuuint32_t box[2] = {(cuuint32_t)cf.R, (cuuint32_t)cf.BC}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const size_t smem = (size_t)2 * cf.R * cf.C * 8 + 64;
    for (int it = 0; it < 3; it++) {
      cudaEventRecord(e0);
      k_stream<<<148, 512, smem>>>(tm, n, nf, cf.R, cf.BC, cf.C, mode, y, y2, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    CK(cudaGetLastError());
    const double bytes = tot * 8.0 * (mode ? 3 : 1);
    printf("R=%3d BC=%2d C=%3d smem=%zu mode=%d: %.3f ms  %.0f GB/s\n", cf.R, cf.BC, cf.C, smem, mode, ms, bytes / ms / 1e6);
  }
  return 0;
}
