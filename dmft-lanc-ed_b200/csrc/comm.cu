// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
//
// Replaces vector_transpose_MPI (ED_HAMILTONIAN_COMMON.f90:53-118: ceil(DimDw/P) MPI_AllToAllV
// calls of one column each + local_transpose) with ONE grouped all-to-all of dense sub-blocks per
// transpose, the per-dot MPI_Allreduce of SciFortran's MPI Lanczos with a device-resident
// ncclAllReduce, and allgather_vector_MPI (ED_SETUP.f90:687-733) with a grouped exchange.
//
// NCCL is resolved at run time with dlopen so that the library shares the copy already loaded by
// the host process (torch's bundled libnccl in the tests/bench, the system one under a Fortran/MPI
// host) instead of linking a second one.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <vector>

#include "engine.h"

namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int nccl_load() {
  if (g_nccl.h) return EDGPU_OK;
  const char *env = getenv("EDGPU_NCCL_LIB");
  void *h = nullptr;
  if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // copy already in the process
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return edgpu_set_err(EDGPU_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                          \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                     \
  if (!g_nccl.field) return edgpu_set_err(EDGPU_ERR_NCCL, "libnccl lacks %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.h = h;
  return EDGPU_OK;
}
}  // namespace

#define NK(call)                                                                              \
  do {                                                                                        \
    ncclResult_t r_ = (call);                                                                 \
    if (r_ != ncclSuccess)                                                                    \
      return edgpu_set_err(EDGPU_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call,         \
                           g_nccl.GetErrorString(r_));                                        \
  } while (0)

extern "C" int edgpu_comm_unique_id(char id[128]) {
  TRY(nccl_load());
  ncclUniqueId u;
  NK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return EDGPU_OK;
}

extern "C" int edgpu_comm_init(edgpu_ctx *c, int rank, int nranks, const char id[128]) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "comm_init while a sector is live");
  if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return edgpu_set_err(EDGPU_ERR_INVALID, "bad rank/nranks");
  if (c->comm) TRY(edgpu_comm_finalize(c));
  c->rank = rank; c->nranks = nranks;
  if (nranks == 1) return EDGPU_OK;
  TRY(nccl_load());
  CK(cudaSetDevice(c->device));
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  ncclComm_t comm;
  NK(g_nccl.CommInitRank(&comm, nranks, u, rank));
  c->comm = (void *)comm;
  return EDGPU_OK;
}

extern "C" int edgpu_comm_finalize(edgpu_ctx *c) {
  if (c && c->comm) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
  }
  if (c) { c->rank = 0; c->nranks = 1; }
  return EDGPU_OK;
}

// ---- work vectors ------------------------------------------------------------------------------------
int vec_alloc(edgpu_ctx *c, double **p, int64_t n) {
  (void)c;
  if (*p) return EDGPU_OK;
  const size_t bytes = (((size_t)(n > 0 ? n : 1) + 2) * sizeof(double) + 255) & ~(size_t)255;
  CK(cudaMalloc(p, bytes));
  return EDGPU_OK;
}
void vec_free(edgpu_ctx *c, double **p) {
  (void)c;
  if (!*p) return;
  cudaFree(*p);
  *p = nullptr;
}
// ---- symmetric slab ----------------------------------------------------------------------------------
int comm_barrier(edgpu_ctx *c) {
  if (c->nranks == 1 || !c->comm) return EDGPU_OK;
  NK(g_nccl.AllReduce(c->d_partials + 4000, c->d_partials + 4000, 1, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  return EDGPU_OK;
}
// Collective: every rank allocates `bytes`, the CUDA IPC handles travel with one ncclAllGather, every rank maps its
// peers' slabs.  On any failure (e.g. peers not visible to this process) every rank ends with sym_ok = false and
// H*v uses the all-to-all transposes instead.
int comm_symm_setup(edgpu_ctx *c, size_t bytes) {
  c->sym_ok = false;
  if (c->nranks == 1 || !c->comm || bytes == 0) return EDGPU_OK;
  const int P = c->nranks, me = c->rank;
  bytes = (bytes + 255) & ~(size_t)255;
  int good = 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&c->sym_slab, bytes) != cudaSuccess) { good = 0; c->sym_slab = nullptr; cudaGetLastError(); }
  if (good && cudaIpcGetMemHandle(&mine, c->sym_slab) != cudaSuccess) { good = 0; cudaGetLastError(); }
  // exchange the handles (and the success flags) with NCCL: no host-side plumbing needed
  const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
  char *d_all = nullptr;
  if (cudaMalloc(&d_all, rec * (size_t)P) != cudaSuccess) {       // cannot even take part in the exchange: fatal for the job
    if (c->sym_slab) { cudaFree(c->sym_slab); c->sym_slab = nullptr; }
    return edgpu_set_err(EDGPU_ERR_CUDA, "comm_symm_setup: cudaMalloc of the handle table failed");
  }
  std::vector<char> h_all(rec * (size_t)P, 0);
  memcpy(h_all.data() + rec * me, &mine, sizeof(mine));
  h_all[rec * me + sizeof(mine)] = (char)good;
  CK(cudaMemcpyAsync(d_all + rec * me, h_all.data() + rec * me, rec, cudaMemcpyHostToDevice, c->stream));
  NK(g_nccl.AllGather(d_all + rec * me, d_all, rec, ncclChar, (ncclComm_t)c->comm, c->stream));
  CK(cudaMemcpyAsync(h_all.data(), d_all, rec * (size_t)P, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_all);
  for (int p = 0; p < P; p++) if (!h_all[rec * p + sizeof(mine)]) good = 0;
  int opened = 0;
  if (good) {
    for (int p = 0; p < P; p++) {
      if (p == me) { c->sym_peer[p] = c->sym_slab; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, h_all.data() + rec * p, sizeof(h));
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = 0; cudaGetLastError(); break; }
      c->sym_peer[p] = reinterpret_cast<char *>(ptr);
      opened++;
    }
  }
  // agree on the outcome
  double flag = good ? 0.0 : 1.0;
  CK(cudaMemcpyAsync(c->d_partials + 4001, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NK(g_nccl.AllReduce(c->d_partials + 4001, c->d_partials + 4001, 1, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  CK(cudaMemcpyAsync(&flag, c->d_partials + 4001, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (flag != 0.0) {
    for (int p = 0; p < P; p++) { if (p != me && c->sym_peer[p]) cudaIpcCloseMemHandle(c->sym_peer[p]); c->sym_peer[p] = nullptr; }
    if (c->sym_slab) cudaFree(c->sym_slab);
    c->sym_slab = nullptr;
    return EDGPU_OK;
  }
  CK(cudaMemsetAsync(c->sym_slab, 0, bytes, c->stream));
  c->sym_bytes = bytes; c->sym_ok = true;
  TRY(comm_barrier(c));
  CK(cudaStreamSynchronize(c->stream));
  return EDGPU_OK;
}
int comm_symm_teardown(edgpu_ctx *c) {
  if (!c->sym_slab) { c->sym_ok = false; return EDGPU_OK; }
  if (c->comm) { comm_barrier(c); cudaStreamSynchronize(c->stream); }   // nobody reads my slab any more
  for (int p = 0; p < c->nranks; p++) { if (p != c->rank && c->sym_peer[p]) cudaIpcCloseMemHandle(c->sym_peer[p]); c->sym_peer[p] = nullptr; }
  if (c->comm) { comm_barrier(c); cudaStreamSynchronize(c->stream); }   // every peer has closed its mapping
  cudaFree(c->sym_slab);
  c->sym_slab = nullptr; c->sym_bytes = 0; c->sym_ok = false;
  return EDGPU_OK;
}

int comm_allreduce_scalar(edgpu_ctx *c, double *d_scalar) {
  if (c->nranks == 1) return EDGPU_OK;
  NK(g_nccl.AllReduce(d_scalar, d_scalar, 1, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  return EDGPU_OK;
}

int comm_allreduce_array(edgpu_ctx *c, double *d_a, int n) {
  if (c->nranks == 1) return EDGPU_OK;
  NK(g_nccl.AllReduce(d_a, d_a, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  return EDGPU_OK;
}

// out[b + nb*a] = A[(a0 + a) + lda*b],  a < na, b < nb   (32x32 tiles through shared memory)
__global__ void k_pack_transpose(const double *__restrict__ A, int64_t lda, int64_t a0, int64_t na, int64_t nb,
                                 double *__restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t tiles_a = (na + 31) / 32, tiles_b = (nb + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_a * tiles_b; t += gridDim.x) {
    const int64_t ta = t % tiles_a, tb = t / tiles_a;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
    for (int k = ty; k < 32; k += 8) {
      int64_t a = ta * 32 + tx, b = tb * 32 + k;
      if (a < na && b < nb) tile[k][tx] = A[(a0 + a) + lda * b];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      int64_t b = tb * 32 + tx, a = ta * 32 + k;
      if (a < na && b < nb) out[b + nb * a] = tile[tx][k];
    }
    __syncthreads();
  }
}
// B[(x0 + x) + ldb*y] (+)= in[x + nx*y]
template <bool ACC>
__global__ void k_unpack(const double *__restrict__ in, int64_t nx, int64_t ny, double *__restrict__ B, int64_t ldb,
                         int64_t x0) {
  for (int64_t y = blockIdx.y; y < ny; y += gridDim.y)
    for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < nx; x += (int64_t)gridDim.x * blockDim.x) {
      if (ACC) B[(x0 + x) + ldb * y] += in[x + nx * y];
      else B[(x0 + x) + ldb * y] = in[x + nx * y];
    }
}

static int ensure(double **p, int64_t n) {
  if (*p) return EDGPU_OK;
  CK(cudaMalloc(p, (size_t)(n > 0 ? n : 1) * sizeof(double)));
  return EDGPU_OK;
}

int comm_exchange(edgpu_ctx *c, const double *send, const int64_t *soff, const int64_t *scnt, double *recv,
                  const int64_t *roff, const int64_t *rcnt) {
  const int P = c->nranks, me = c->rank;
  NK(g_nccl.GroupStart());
  for (int p = 0; p < P; p++) {
    if (p == me) continue;
    if (scnt[p]) NK(g_nccl.Send(send + soff[p], (size_t)scnt[p], ncclDouble, p, (ncclComm_t)c->comm, c->stream));
    if (rcnt[p]) NK(g_nccl.Recv(recv + roff[p], (size_t)rcnt[p], ncclDouble, p, (ncclComm_t)c->comm, c->stream));
  }
  NK(g_nccl.GroupEnd());
  if (scnt[me])
    CK(cudaMemcpyAsync(recv + roff[me], send + soff[me], (size_t)scnt[me] * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return EDGPU_OK;
}

// Block geometry of the two all-to-all transposes (pure host arithmetic, exported so that the
// world_size-2 gloo tests can drive the exact offsets the CUDA path uses).
//   dir = 0: V(DimUp, qdw) -> Vt(DimDw, qup).  Block (me -> p) = V[rows(p), my columns] stored
//            transposed, [c + qdw_me * r];      received block (p -> me) = [c + qdw_p * r], r < qup_me.
//   dir = 1: Hvt(DimDw, qup) -> Hv(DimUp, qdw). Block (me -> p) = Hvt[columns(p), my rows] stored
//            transposed, [r + qup_me * c];      received block (p -> me) = [r + qup_p * c], c < qdw_me.
extern "C" void edgpu_transpose_plan(int64_t dimup, int64_t dimdw, int nranks, int rank, int dir,
                                     int64_t *soff, int64_t *scnt, int64_t *roff, int64_t *rcnt) {
  int64_t qdw, qup;
  edgpu_split(dimdw, nranks, rank, &qdw, nullptr);
  edgpu_split(dimup, nranks, rank, &qup, nullptr);
  for (int p = 0; p < nranks; p++) {
    int64_t qr, ro, qc, co;
    edgpu_split(dimup, nranks, p, &qr, &ro);
    edgpu_split(dimdw, nranks, p, &qc, &co);
    if (dir == 0) {
      soff[p] = ro * qdw; scnt[p] = qr * qdw;      // rows(p) x my columns
      roff[p] = co * qup; rcnt[p] = qc * qup;      // columns(p) x my rows
    } else {
      soff[p] = co * qup; scnt[p] = qc * qup;      // dw-rows(p) x my up-rows
      roff[p] = ro * qdw; rcnt[p] = qr * qdw;      // up-rows(p) x my columns
    }
  }
}

// V(DimUp, qdw) -> Vt(DimDw, qup): block (me -> d) = V[rows(d), :] transposed
int comm_transpose_fwd(edgpu_ctx *c, const double *d_x, double *d_vt) {
  const int P = c->nranks;
  TRY(ensure(&c->d_send, c->nel > c->dimdw * c->qup ? c->nel : c->dimdw * c->qup));
  TRY(ensure(&c->d_recv, c->nel > c->dimdw * c->qup ? c->nel : c->dimdw * c->qup));
  int64_t soff[64] = {0}, scnt[64] = {0}, roff[64] = {0}, rcnt[64] = {0};
  edgpu_transpose_plan(c->dimup, c->dimdw, P, c->rank, 0, soff, scnt, roff, rcnt);
  for (int p = 0; p < P; p++) {
    int64_t qr, ro;
    edgpu_split(c->dimup, P, p, &qr, &ro);
    if (scnt[p]) {
      int64_t tiles = ((qr + 31) / 32) * ((c->qdw + 31) / 32);
      int grid = (int)(tiles < (int64_t)c->sm_count * 8 ? tiles : (int64_t)c->sm_count * 8);
      k_pack_transpose<<<grid, 256, 0, c->stream>>>(d_x, c->dimup, ro, qr, c->qdw, c->d_send + soff[p]);
      CKL(c);
    }
  }
  TRY(comm_exchange(c, c->d_send, soff, scnt, c->d_recv, roff, rcnt));
  for (int p = 0; p < P; p++) {
    int64_t qc, co;
    edgpu_split(c->dimdw, P, p, &qc, &co);
    if (!rcnt[p]) continue;
    dim3 grid((unsigned)((qc + 255) / 256), (unsigned)(c->qup < 4096 ? c->qup : 4096));
    k_unpack<false><<<grid, 256, 0, c->stream>>>(c->d_recv + roff[p], qc, c->qup, d_vt, c->dimdw, co);
    CKL(c);
  }
  return EDGPU_OK;
}

// Hv(DimUp, qdw) += (Hvt(DimDw, qup))^T : block (me -> s) = Hvt[cols(s), :] transposed
int comm_transpose_bwd_add(edgpu_ctx *c, const double *d_hvt, double *d_y) {
  const int P = c->nranks;
  int64_t soff[64] = {0}, scnt[64] = {0}, roff[64] = {0}, rcnt[64] = {0};
  edgpu_transpose_plan(c->dimup, c->dimdw, P, c->rank, 1, soff, scnt, roff, rcnt);
  for (int p = 0; p < P; p++) {
    int64_t qc, co;
    edgpu_split(c->dimdw, P, p, &qc, &co);
    if (scnt[p]) {
      int64_t tiles = ((qc + 31) / 32) * ((c->qup + 31) / 32);
      int grid = (int)(tiles < (int64_t)c->sm_count * 8 ? tiles : (int64_t)c->sm_count * 8);
      k_pack_transpose<<<grid, 256, 0, c->stream>>>(d_hvt, c->dimdw, co, qc, c->qup, c->d_send + soff[p]);
      CKL(c);
    }
  }
  TRY(comm_exchange(c, c->d_send, soff, scnt, c->d_recv, roff, rcnt));
  for (int p = 0; p < P; p++) {
    int64_t qr, ro;
    edgpu_split(c->dimup, P, p, &qr, &ro);
    if (!rcnt[p]) continue;
    dim3 grid((unsigned)((qr + 255) / 256), (unsigned)(c->qdw < 4096 ? c->qdw : 4096));
    k_unpack<true><<<grid, 256, 0, c->stream>>>(c->d_recv + roff[p], qr, c->qdw, d_y, c->dimup, ro);
    CKL(c);
  }
  return EDGPU_OK;
}

// allgather_vector_MPI (ED_SETUP.f90:687-733): shards are contiguous slices of the full vector
int comm_allgather(edgpu_ctx *c, const double *d_x, double *d_full) {
  const int P = c->nranks, me = c->rank;
  NK(g_nccl.GroupStart());
  for (int p = 0; p < P; p++) {
    int64_t qc, co;
    edgpu_split(c->dimdw, P, p, &qc, &co);
    if (p == me) continue;
    NK(g_nccl.Send(d_x, (size_t)c->nel, ncclDouble, p, (ncclComm_t)c->comm, c->stream));
    NK(g_nccl.Recv(d_full + co * c->dimup, (size_t)(qc * c->dimup), ncclDouble, p, (ncclComm_t)c->comm, c->stream));
  }
  NK(g_nccl.GroupEnd());
  CK(cudaMemcpyAsync(d_full + c->coloff * c->dimup, d_x, (size_t)c->nel * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return EDGPU_OK;
}
