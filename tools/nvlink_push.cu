// nvlink_push.cu -- what does the NVLink fabric deliver for the halo exchange pattern of the sharded H*v?
// One process, G GPUs with peer access.  Every GPU g owns NCOL columns of N doubles and sends column blocks to the
// peers g^1, g^2, g^4 (those that exist), all GPUs at once -- the traffic pattern of hxv_fast.cu's k_halo_push at
// P = 8 (C3: 265 MB out per GPU in 103 KB columns).  Variants:
//   st   : 16-byte register stores to the peer-mapped address (k_halo_push), CTAS CTAs of 512 threads
//   bulk : TMA engine both ways: cp.async.bulk global -> shared (local read), cp.async.bulk shared -> global (peer
//          write), one elected thread per CTA, ring of 16 KB stages
//   pull : 16-byte loads from the peer (the round-1 scheme), for reference
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlink_push.bin nvlink_push.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define MAXG 8
struct Args {
  const double *src;      // local columns
  double *dst[3];         // peers' receive buffers (this GPU's region inside each)
  const double *psrc[3];  // peers' local columns (pull)
  double *ldst;           // local receive buffer (pull)
  int npeer, n, ncol;     // columns of n doubles; column c goes to peer c % npeer
};
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int UNROLL>
__global__ void __launch_bounds__(512) k_st(Args a) {
  const int ROWS = 512 * UNROLL * 2;
  const int nrc = (a.n + ROWS - 1) / ROWS;
  const long nitems = (long)a.ncol * nrc;
  for (long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int c = (int)(it / nrc), r0 = (int)(it % nrc) * ROWS;
    const double2 *s = reinterpret_cast<const double2 *>(a.src + (size_t)c * a.n + r0);
    double2 *d = reinterpret_cast<double2 *>(a.dst[c % a.npeer] + (size_t)(c / a.npeer) * a.n + r0);
    double2 v[UNROLL];
#pragma unroll
    for (int q = 0; q < UNROLL; q++) { const int i = threadIdx.x + q * 512; v[q] = (r0 + 2 * i < a.n) ? __ldg(s + i) : make_double2(0, 0); }
#pragma unroll
    for (int q = 0; q < UNROLL; q++) { const int i = threadIdx.x + q * 512; if (r0 + 2 * i < a.n) d[i] = v[q]; }
  }
}
template <int UNROLL>
__global__ void __launch_bounds__(512) k_pull(Args a) {
  const int ROWS = 512 * UNROLL * 2;
  const int nrc = (a.n + ROWS - 1) / ROWS;
  const long nitems = (long)a.ncol * nrc;
  for (long it = blockIdx.x; it < nitems; it += gridDim.x) {
    const int c = (int)(it / nrc), r0 = (int)(it % nrc) * ROWS;
    const double2 *s = reinterpret_cast<const double2 *>(a.psrc[c % a.npeer] + (size_t)c * a.n + r0);
    double2 *d = reinterpret_cast<double2 *>(a.ldst + (size_t)c * a.n + r0);
    double2 v[UNROLL];
#pragma unroll
    for (int q = 0; q < UNROLL; q++) { const int i = threadIdx.x + q * 512; v[q] = (r0 + 2 * i < a.n) ? s[i] : make_double2(0, 0); }
#pragma unroll
    for (int q = 0; q < UNROLL; q++) { const int i = threadIdx.x + q * 512; if (r0 + 2 * i < a.n) d[i] = v[q]; }
  }
}
// TMA both ways: one thread per CTA drives a ring of STAGES buffers of SB bytes
#define SB 16384
template <int STAGES>
__global__ void __launch_bounds__(32) k_bulk(Args a) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + (size_t)STAGES * SB);
  if (threadIdx.x != 0) return;
  for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[s])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const long colbytes = (long)a.n * 8;
  const int npc = (int)((colbytes + SB - 1) / SB);
  const long nitems = (long)a.ncol * npc;
  long issued = 0, done = 0;
  auto issue = [&](long it, int s) {
    const int c = (int)(it / npc);
    const long off = (long)(it % npc) * SB;
    const uint32_t bytes = (uint32_t)((colbytes - off) < SB ? (colbytes - off) : SB);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * SB)),
                 "l"(reinterpret_cast<const char *>(a.src + (size_t)c * a.n) + off), "r"(bytes), "r"(s32(&bar[s]))
                 : "memory");
  };
  long my = 0;
  for (long it = blockIdx.x; it < nitems; it += gridDim.x) my++;
  for (; issued < my && issued < STAGES; issued++) issue(blockIdx.x + issued * gridDim.x, (int)(issued % STAGES));
  for (; done < my; done++) {
    const int s = (int)(done % STAGES);
    const uint32_t ph = (uint32_t)((done / STAGES) & 1);
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bar[s])), "r"(ph) : "memory");
    const long it = blockIdx.x + done * gridDim.x;
    const int c = (int)(it / npc);
    const long off = (long)(it % npc) * SB;
    const uint32_t bytes = (uint32_t)((colbytes - off) < SB ? (colbytes - off) : SB);
    char *d = reinterpret_cast<char *>(a.dst[c % a.npeer] + (size_t)(c / a.npeer) * a.n) + off;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(s32(sm + (size_t)s * SB)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (issued < my) {                                             // the next load reuses this stage: the store must have read it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue(blockIdx.x + issued * gridDim.x, s);
      issued++;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
  int G = 0;
  CK(cudaGetDeviceCount(&G));
  if (argc > 1) G = atoi(argv[1]) < G ? atoi(argv[1]) : G;
  if (G > MAXG) G = MAXG;
  const int n = argc > 2 ? atoi(argv[2]) : 12870;
  const int ncol = argc > 3 ? atoi(argv[3]) : 2570;
  int npeer = 0;
  for (int b = 1; b < G; b <<= 1) npeer++;
  if (npeer == 0) { printf("need >= 2 GPUs\n"); return 0; }
  printf("G=%d n=%d ncol=%d npeer=%d: %.1f MB out per GPU\n", G, n, ncol, npeer, (double)n * ncol * 8 / 1e6);
  double *src[MAXG], *rcv[MAXG];
  cudaStream_t st[MAXG];
  cudaEvent_t e0[MAXG], e1[MAXG];
  const size_t colsz = (size_t)n * 8, per = ((size_t)ncol / npeer + 2) * colsz;   // region of one sender in a receive buffer
  for (int g = 0; g < G; g++) {
    CK(cudaSetDevice(g));
    for (int p = 0; p < G; p++) if (p != g) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("no peer access %d->%d\n", g, p); return 1; } cudaGetLastError(); }
    CK(cudaMalloc(&src[g], (size_t)ncol * colsz + 64));
    CK(cudaMalloc(&rcv[g], per * MAXG + (size_t)ncol * colsz));
    CK(cudaMemset(src[g], 1, (size_t)ncol * colsz));
    CK(cudaStreamCreate(&st[g]));
    CK(cudaEventCreate(&e0[g])); CK(cudaEventCreate(&e1[g]));
  }
  Args a[MAXG];
  for (int g = 0; g < G; g++) {
    a[g].src = src[g]; a[g].npeer = npeer; a[g].n = n; a[g].ncol = ncol; a[g].ldst = rcv[g] + per * MAXG / 8;
    for (int k = 0; k < npeer; k++) { const int p = g ^ (1 << k); a[g].dst[k] = rcv[p] + (per / 8) * g; a[g].psrc[k] = src[p]; }
  }
  auto run = [&](const char *name, int ctas, int variant) {
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
      for (int g = 0; g < G; g++) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(st[g])); }
      for (int g = 0; g < G; g++) {
        CK(cudaSetDevice(g));
        CK(cudaEventRecord(e0[g], st[g]));
        if (variant == 0) k_st<8><<<ctas, 512, 0, st[g]>>>(a[g]);
        else if (variant == 1) k_st<4><<<ctas, 512, 0, st[g]>>>(a[g]);
        else if (variant == 2) k_pull<8><<<ctas, 512, 0, st[g]>>>(a[g]);
        else if (variant == 3) k_bulk<4><<<ctas, 32, 4 * SB + 64, st[g]>>>(a[g]);
        else if (variant == 4) k_bulk<8><<<ctas, 32, 8 * SB + 64, st[g]>>>(a[g]);
        CK(cudaEventRecord(e1[g], st[g]));
      }
      float worst = 0;
      for (int g = 0; g < G; g++) { CK(cudaSetDevice(g)); CK(cudaEventSynchronize(e1[g])); float ms; CK(cudaEventElapsedTime(&ms, e0[g], e1[g])); if (ms > worst) worst = ms; }
      if (rep > 0 && worst < best) best = worst;
    }
    printf("%-28s ctas %4d : %.3f ms  %.0f GB/s out per GPU\n", name, ctas, best, (double)n * ncol * 8 / best / 1e6);
  };
  for (int g = 0; g < G; g++) {
    CK(cudaSetDevice(g));
    CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SB + 64));
    CK(cudaFuncSetAttribute(k_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SB + 64));
  }
  const int cs[] = {16, 32, 64, 148, 296};
  for (int c : cs) run("st 16B x8 (k_halo_push)", c, 0);
  for (int c : cs) run("st 16B x4", c, 1);
  for (int c : cs) run("pull 16B x8", c, 2);
  for (int c : cs) run("bulk TMA 4 x 16KB", c, 3);
  for (int c : cs) run("bulk TMA 8 x 16KB", c, 4);
  const int cs2[] = {592, 1184};
  for (int c : cs2) run("bulk TMA 4 x 16KB", c, 3);
  return 0;
}
