// Generic COO -> CSC: the device restatement of SparseArrays.sparse(I,J,V,m,n) (call site AssemblyModule.jl:319-325)
// for triplets that do not come with mesh structure (the startassembly!/assemble!/makematrix! protocol used by any
// other caller, dof maps that are not injective, elements that repeat a node).
//
//   key = (col-1) << 32 | (row-1)  -> stable LSD radix sort (8-bit digits, only the digits the matrix size needs) of
//   (key, triplet id) -> segment heads -> colptr / rowval -> nzval[k] = sum of V[id] over the segment in ascending
//   triplet id (stable sort => the reference's left-to-right sum).  Explicit zeros are kept.
#include <cstdlib>

#include "fegpu_internal.h"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // keys per block; warp w owns a contiguous 256-key span

__global__ void k_make_keys(int64_t n, const int64_t *__restrict__ I, const int64_t *__restrict__ J, int64_t nrows, int64_t ncols,
                            unsigned long long *__restrict__ keys, uint32_t *__restrict__ ids, int *err) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t i = I[k], j = J[k];
  // first violated check in the reference's order (column first): AssemblyModule.jl:268-273
  int code = 0;
  if (j < 1) code = 1;
  else if (j > ncols) code = 2;
  else if (i < 1) code = 3;
  else if (i > nrows) code = 4;
  if (code) {
    atomicCAS(err, 0, code);
    i = 1;
    j = 1;
  }
  keys[k] = ((unsigned long long)(j - 1) << 32) | (unsigned long long)(i - 1);
  ids[k] = (uint32_t)k;
}

// digit histogram of every tile: hist[digit][tile]
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long *__restrict__ keys, int64_t n, int shift,
                                                        int32_t *__restrict__ hist, int64_t ntiles) {
  __shared__ int32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  for (int i = 0; i < RS_ITEMS; i++) {
    int64_t k = base + (int64_t)i * RS_THREADS + threadIdx.x;
    if (k < n) atomicAdd(&h[(keys[k] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: rank of a key among the tile's keys with the same digit, in original order
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long *__restrict__ keys_in, const uint32_t *__restrict__ ids_in,
                                                           unsigned long long *__restrict__ keys_out, uint32_t *__restrict__ ids_out,
                                                           int64_t n, int shift, const int64_t *__restrict__ offs, int64_t ntiles) {
  __shared__ int32_t wcount[RS_THREADS / 32][256];  // per warp digit counts -> exclusive prefix over warps
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&wcount[0][0])[i] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
  unsigned long long key[RS_ITEMS];
  int32_t lrank[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int64_t k = wbase + r * 32 + lane;
    const bool valid = k < n;
    key[r] = valid ? keys_in[k] : ~0ull;
    const unsigned d = valid ? (unsigned)((key[r] >> shift) & 255u) : 256u + lane;  // invalid lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    int32_t basecnt = 0;
    if (valid && lane == leader) {
      basecnt = wcount[w][d];
      wcount[w][d] = basecnt + __popc(peers);
    }
    basecnt = __shfl_sync(0xffffffffu, basecnt, leader);
    lrank[r] = basecnt + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: exclusive prefix of digit d over the warps of this tile
    const int d = threadIdx.x;
    int32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < RS_THREADS / 32; ww++) {
      int32_t c = wcount[ww][d];
      wcount[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    const int64_t k = wbase + r * 32 + lane;
    if (k < n) {
      const unsigned d = (unsigned)((key[r] >> shift) & 255u);
      const int64_t pos = offs[(int64_t)d * ntiles + blockIdx.x] + wcount[w][d] + lrank[r];
      keys_out[pos] = key[r];
      ids_out[pos] = ids_in[k];
    }
  }
}

__global__ void k_heads(const unsigned long long *__restrict__ keys, int64_t n, int32_t *__restrict__ head) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  head[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

// per segment: rowval, column count; segstart[seg] = first sorted position
__global__ void k_segments(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ head, const int64_t *__restrict__ segid,
                           int64_t n, int64_t *__restrict__ rowval, int64_t *__restrict__ colcount, int64_t *__restrict__ segstart) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || !head[k]) return;
  const int64_t s = segid[k];
  rowval[s] = (int64_t)(keys[k] & 0xffffffffull) + 1;
  segstart[s] = k;
  atomicAdd((unsigned long long *)&colcount[keys[k] >> 32], 1ull);
}

__global__ void k_segsum(const uint32_t *__restrict__ ids, const double *__restrict__ V, const int64_t *__restrict__ segstart, int64_t nseg,
                         int64_t n, double *__restrict__ nzval) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int64_t b = segstart[s], e = (s + 1 < nseg) ? segstart[s + 1] : n;
  double v = V[ids[b]];
  for (int64_t k = b + 1; k < e; k++) v = v + V[ids[k]];
  nzval[s] = v;
}

// perm (optional): output position r takes the element of slot perm[r] (the caller's element order for the raw-COO export)
__global__ void k_emit_ij(const int32_t *__restrict__ conn, const int32_t *__restrict__ elem_list, int64_t nactive, int nne, int ndn,
                          int64_t nnodes, const int32_t *__restrict__ dof, int64_t *__restrict__ I, int64_t *__restrict__ J,
                          const int32_t *__restrict__ perm) {
  const int EM = nne * ndn;
  const int64_t EM2 = (int64_t)EM * EM;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nactive * EM2) return;
  const int64_t slot = perm ? (int64_t)perm[t / EM2] : t / EM2;
  const int loc = (int)(t - (t / EM2) * EM2);
  const int c = loc / EM, r = loc - c * EM;
  const int64_t e = elem_list ? elem_list[slot] : slot;
  const int32_t *cn = conn + e * nne;
  I[t] = (int64_t)dof[(int64_t)(r % ndn) * nnodes + cn[r / ndn]] + 1;
  J[t] = (int64_t)dof[(int64_t)(c % ndn) * nnodes + cn[c / ndn]] + 1;
}

int bits_for(int64_t n) {  // bits needed to represent values 0..n-1
  int b = 0;
  while (b < 32 && ((int64_t)1 << b) < n) b++;
  return b;
}

}  // namespace

int32_t fe_reserve_bytes(fegpu_ctx *ctx, void **buf, size_t *cap, size_t need) {
  if (*cap >= need && *buf) return FEGPU_OK;
  if (*buf) CUDA_TRY(ctx, cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  size_t want = need ? need : 8;
  cudaError_t e = cudaMalloc(buf, want);
  if (e != cudaSuccess) {  // the symbolic phase's block cache may be sitting on freed blocks: hand them back and retry once
    cudaGetLastError();
    fe_dev_cache_trim(ctx);
    e = cudaMalloc(buf, want);
  }
  if (e != cudaSuccess) {
    *buf = nullptr;
    return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string("cudaMalloc of ") + std::to_string(want) + " bytes: " + cudaGetErrorString(e));
  }
  *cap = want;
  return FEGPU_OK;
}

int32_t fe_asm_reserve(fegpu_asm *as, double **buf, size_t *cap, size_t need_doubles) {
  size_t capb = *cap * sizeof(double);
  int32_t s = fe_reserve_bytes(as->ctx, (void **)buf, &capb, need_doubles * sizeof(double));
  *cap = capb / sizeof(double);
  return s;
}

int32_t fe_emit_ij(fegpu_dofmap *dm, int64_t *d_I, int64_t *d_J, const int32_t *d_perm) {
  fegpu_mesh *mesh = dm->mesh;
  const int EM = mesh->nne * dm->ndn;
  const int64_t n = mesh->nactive * EM * EM;
  if (n == 0) return FEGPU_OK;
  k_emit_ij<<<grid_for(n, 256), 256, 0, dm->ctx->stream>>>(mesh->conn_act(), mesh->d_elem_list, mesh->nactive, mesh->nne, dm->ndn, mesh->nnodes,
                                                          dm->d_dof, d_I, d_J, d_perm);
  dm->ctx->launches++;
  CUDA_TRY(dm->ctx, cudaGetLastError());
  return FEGPU_OK;
}

static const char *dof_msg(int code) {
  switch (code) {
    case 1: return "Column degree of freedom < 1";
    case 2: return "Column degree of freedom > size";
    case 3: return "Row degree of freedom < 1";
    default: return "Row degree of freedom > size";
  }
}

// Stable LSD radix sort of (key, id) pairs over the given digit shifts; *kin/*iin end up pointing at the sorted buffers.
static int32_t radix_sort_pairs(fegpu_ctx *ctx, int64_t n, const std::vector<int> &shifts, unsigned long long **kin, unsigned long long **kout,
                                uint32_t **iin, uint32_t **iout, int32_t *hist, int64_t *offs) {
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  for (int sh : shifts) {
    k_rs_hist<<<(unsigned)ntiles, RS_THREADS, 0, ctx->stream>>>(*kin, n, sh, hist, ntiles);
    ctx->launches++;
    FE_TRY(fe_exclusive_scan_i32_to_i64(ctx, hist, offs, 256 * ntiles, 0, false, nullptr));
    k_rs_scatter<<<(unsigned)ntiles, RS_THREADS, 0, ctx->stream>>>(*kin, *iin, *kout, *iout, n, sh, offs, ntiles);
    ctx->launches++;
    std::swap(*kin, *kout);
    std::swap(*iin, *iout);
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

namespace {
// 30-bit Morton code of a node's position inside the mesh bounding box (10 bits per axis)
__device__ __forceinline__ unsigned spread10(unsigned v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__global__ void k_morton_keys(const double *__restrict__ xyz, int64_t nnodes, const int32_t *__restrict__ nodes, int64_t count, int sdim,
                              double lx, double ly, double lz, double sx, double sy, double sz, unsigned long long *__restrict__ keys,
                              uint32_t *__restrict__ ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t n = nodes ? (int64_t)nodes[i] : i;
  const unsigned qx = (unsigned)fmin(fmax((xyz[n] - lx) * sx, 0.0), 1023.0);
  const unsigned qy = sdim > 1 ? (unsigned)fmin(fmax((xyz[nnodes + n] - ly) * sy, 0.0), 1023.0) : 0u;
  const unsigned qz = sdim > 2 ? (unsigned)fmin(fmax((xyz[2 * nnodes + n] - lz) * sz, 0.0), 1023.0) : 0u;
  keys[i] = (unsigned long long)(spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2));
  ids[i] = (uint32_t)n;
}
__global__ void k_ids_to_i32(const uint32_t *__restrict__ ids, int32_t *__restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)ids[i];
}
}  // namespace

// Node visiting order with spatial locality (Morton order of the coordinates): nodes that share elements are processed close
// together in time, so element data read by several column nodes is still in L2 the second time.  d_order: [nnodes].
int32_t fe_morton_order(fegpu_mesh *mesh, const int32_t *d_nodes, int64_t n, int32_t *d_order) {
  fegpu_ctx *ctx = mesh->ctx;
  cudaStream_t st = ctx->stream;
  if (n == 0) return FEGPU_OK;
  unsigned long long *kA = nullptr, *kB = nullptr;
  uint32_t *iA = nullptr, *iB = nullptr;
  int32_t *hist = nullptr;
  int64_t *offs = nullptr;
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  auto cleanup = [&]() {
    void *ptrs[] = {kA, kB, iA, iB, hist, offs};
    for (void *q : ptrs)
      if (q) fe_dev_free(ctx, q, st);
  };
#define MC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
#define MT(expr) do { int32_t _s = (expr); if (_s != FEGPU_OK) { cleanup(); return _s; } } while (0)
  MT(fe_dev_alloc(ctx, (void **)&kA, sizeof(unsigned long long) * n, st));
  MT(fe_dev_alloc(ctx, (void **)&kB, sizeof(unsigned long long) * n, st));
  MT(fe_dev_alloc(ctx, (void **)&iA, sizeof(uint32_t) * n, st));
  MT(fe_dev_alloc(ctx, (void **)&iB, sizeof(uint32_t) * n, st));
  MT(fe_dev_alloc(ctx, (void **)&hist, sizeof(int32_t) * 256 * ntiles, st));
  MT(fe_dev_alloc(ctx, (void **)&offs, sizeof(int64_t) * (256 * ntiles + 1), st));
#undef MT
#undef MC
  // FEGPU_GATHER_ORDER_BITS (1..10, default 10): resolution of the Morton grid per axis.  The sort is a stable LSD radix sort,
  // so with a coarse grid the nodes of a cell keep their natural (ascending id) order -- bricks of nodes visited one after the
  // other, each walked in the mesh's own numbering -- and the sort needs ceil(3 bits / 8) passes instead of four.
  static const int bits = [] {
    const char *e = std::getenv("FEGPU_GATHER_ORDER_BITS");
    const int b = e ? std::atoi(e) : 10;
    return (b >= 1 && b <= 10) ? b : 10;
  }();
  double s[3];
  for (int d = 0; d < 3; d++) {
    const double ext = mesh->bbox_hi[d] - mesh->bbox_lo[d];
    s[d] = ext > 0 ? ((double)(1 << bits) - 0.001) / ext : 0.0;
  }
  k_morton_keys<<<grid_for(n, 256), 256, 0, st>>>(mesh->d_xyz, mesh->nnodes, d_nodes, n, mesh->sdim, mesh->bbox_lo[0], mesh->bbox_lo[1], mesh->bbox_lo[2], s[0], s[1], s[2], kA, iA);
  ctx->launches++;
  unsigned long long *kin = kA, *kout = kB;
  uint32_t *iin = iA, *iout = iB;
  std::vector<int> shifts;
  for (int sh = 0; sh < 3 * bits; sh += 8) shifts.push_back(sh);
  int32_t rc = radix_sort_pairs(ctx, n, shifts, &kin, &kout, &iin, &iout, hist, offs);
  if (rc == FEGPU_OK) {
    k_ids_to_i32<<<grid_for(n, 256), 256, 0, st>>>(iin, d_order, n);
    ctx->launches++;
  }
  cleanup();
  return rc;
}

namespace {
__global__ void k_elem_minnode_keys(const int32_t *__restrict__ conn, int64_t nelem, int nne, unsigned long long *__restrict__ keys,
                                    uint32_t *__restrict__ ids) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  int mn = conn[e * nne];
  for (int k = 1; k < nne; k++) mn = min(mn, conn[e * nne + k]);
  keys[e] = (unsigned long long)(unsigned)mn;
  ids[e] = (uint32_t)e;
}
__global__ void k_permute_conn(const int32_t *__restrict__ conn_in, const uint32_t *__restrict__ ids, int64_t nelem, int nne,
                               int32_t *__restrict__ conn_out, int32_t *__restrict__ orig) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nelem * nne) return;
  const int64_t i = t / nne;
  const int k = (int)(t - i * nne);
  const uint32_t e = ids[i];
  conn_out[t] = conn_in[(int64_t)e * nne + k];
  if (k == 0) orig[i] = (int32_t)e;
}
__global__ void k_orig_keys(const int32_t *__restrict__ orig, const int32_t *__restrict__ elem_list, int64_t elem_base, int64_t nactive,
                            unsigned long long *__restrict__ keys, uint32_t *__restrict__ ids) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nactive) return;
  const int64_t e = elem_list ? (int64_t)elem_list[s] : s + elem_base;
  keys[s] = (unsigned long long)(unsigned)(orig ? orig[e] : (int32_t)e);
  ids[s] = (uint32_t)s;
}
__global__ void k_copy_u32(const uint32_t *__restrict__ in, int32_t *__restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)in[i];
}

// scratch of one radix sort of n (key, id) pairs
struct SortScratch {
  fegpu_ctx *ctx;
  unsigned long long *kA = nullptr, *kB = nullptr;
  uint32_t *iA = nullptr, *iB = nullptr;
  int32_t *hist = nullptr;
  int64_t *offs = nullptr;
  explicit SortScratch(fegpu_ctx *c) : ctx(c) {}
  int32_t alloc(int64_t n) {
    const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    cudaStream_t st = ctx->stream;
    FE_TRY(fe_dev_alloc(ctx, (void **)&kA, sizeof(unsigned long long) * n, st));
    FE_TRY(fe_dev_alloc(ctx, (void **)&kB, sizeof(unsigned long long) * n, st));
    FE_TRY(fe_dev_alloc(ctx, (void **)&iA, sizeof(uint32_t) * n, st));
    FE_TRY(fe_dev_alloc(ctx, (void **)&iB, sizeof(uint32_t) * n, st));
    FE_TRY(fe_dev_alloc(ctx, (void **)&hist, sizeof(int32_t) * 256 * ntiles, st));
    FE_TRY(fe_dev_alloc(ctx, (void **)&offs, sizeof(int64_t) * (256 * ntiles + 1), st));
    return FEGPU_OK;
  }
  ~SortScratch() {
    void *ptrs[] = {kA, kB, iA, iB, hist, offs};
    for (void *q : ptrs)
      if (q) fe_dev_free(ctx, q, ctx->stream);
  }
};
}  // namespace

// Internal element order = ascending smallest node id (stable: ties keep the caller's order).  The outputs of the assembly are in
// node / dof order, so this is the order in which element records are produced and consumed with locality: the a-th adjacent
// element of node n and of node n+1 then sit next to each other in the element-value array, whatever order the caller's FESet
// lists its elements in (H8block numbers elements z-fastest and nodes x-fastest: neighbouring nodes' elements are 65 536
// records apart).  conn is permuted in place (through a temporary); orig[i] = the caller's id of internal element i.
int32_t fe_order_elements(fegpu_mesh *mesh) {
  fegpu_ctx *ctx = mesh->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t n = mesh->nelem;
  static const bool off = std::getenv("FEGPU_ELEM_ORDER") && std::atoi(std::getenv("FEGPU_ELEM_ORDER")) == 0;  // A/B knob
  if (n == 0 || off) return FEGPU_OK;
  SortScratch S(ctx);
  FE_TRY(S.alloc(n));
  k_elem_minnode_keys<<<grid_for(n, 256), 256, 0, st>>>(mesh->d_conn, n, mesh->nne, S.kA, S.iA);
  ctx->launches++;
  std::vector<int> shifts;
  for (int sh = 0; sh < bits_for(mesh->nnodes); sh += 8) shifts.push_back(sh);
  unsigned long long *kin = S.kA, *kout = S.kB;
  uint32_t *iin = S.iA, *iout = S.iB;
  FE_TRY(radix_sort_pairs(ctx, n, shifts, &kin, &kout, &iin, &iout, S.hist, S.offs));
  int32_t *conn_new = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void **)&conn_new, sizeof(int32_t) * (size_t)n * mesh->nne));
  if (!mesh->d_orig) {
    cudaError_t e = cudaMalloc((void **)&mesh->d_orig, sizeof(int32_t) * (size_t)n);
    if (e != cudaSuccess) { cudaFree(conn_new); return fegpu_fail(ctx, FEGPU_ERR_CUDA, cudaGetErrorString(e)); }
  }
  k_permute_conn<<<grid_for(n * mesh->nne, 256), 256, 0, st>>>(mesh->d_conn, iin, n, mesh->nne, conn_new, mesh->d_orig);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  cudaFree(mesh->d_conn);
  mesh->d_conn = conn_new;
  return FEGPU_OK;
}

// d_perm[r] = slot (position in the active element list) of the r-th active element in the CALLER's element order: the raw-COO
// export (AssemblyModule.jl:261-280: triplets in element call order) walks the element-value array through it.
int32_t fe_emission_order(fegpu_mesh *mesh, int32_t *d_perm) {
  fegpu_ctx *ctx = mesh->ctx;
  cudaStream_t st = ctx->stream;
  const int64_t n = mesh->nactive;
  if (n == 0) return FEGPU_OK;
  SortScratch S(ctx);
  FE_TRY(S.alloc(n));
  k_orig_keys<<<grid_for(n, 256), 256, 0, st>>>(mesh->d_orig, mesh->d_elem_list, mesh->elem_base, n, S.kA, S.iA);
  ctx->launches++;
  std::vector<int> shifts;
  for (int sh = 0; sh < bits_for(mesh->nelem); sh += 8) shifts.push_back(sh);
  unsigned long long *kin = S.kA, *kout = S.kB;
  uint32_t *iin = S.iA, *iout = S.iB;
  FE_TRY(radix_sort_pairs(ctx, n, shifts, &kin, &kout, &iin, &iout, S.hist, S.offs));
  k_copy_u32<<<grid_for(n, 256), 256, 0, st>>>(iin, d_perm, n);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(st));  // the scratch goes back to the block cache behind this point
  return FEGPU_OK;
}

int32_t fe_coo_to_csc(fegpu_asm *as, int64_t n, const int64_t *d_I, const int64_t *d_J, const double *d_V, int64_t nrows, int64_t ncols) {
  fegpu_ctx *ctx = as->ctx;
  cudaStream_t st = ctx->stream;
  if (nrows >= ((int64_t)1 << 32) || ncols >= ((int64_t)1 << 32) || n >= ((int64_t)1 << 32))
    return fegpu_fail(ctx, FEGPU_ERR_ARG, "generic sort path is limited to 2^32 rows / columns / triplets");
  {
    size_t capb = as->own_colptr_cap * sizeof(int64_t);
    FE_TRY(fe_reserve_bytes(ctx, (void **)&as->own_colptr, &capb, sizeof(int64_t) * (size_t)(ncols + 1)));
    as->own_colptr_cap = capb / sizeof(int64_t);
  }
  CUDA_TRY(ctx, cudaMemsetAsync(as->own_colptr, 0, sizeof(int64_t) * (size_t)(ncols + 1), st));
  as->nrows = nrows;
  as->ncols = ncols;
  if (n == 0) {
    int64_t tot = 0;
    FE_TRY(fe_exclusive_scan_i64(ctx, as->own_colptr, as->own_colptr, ncols, 1, true, &tot));
    as->nnz = 0;
    as->d_colptr = as->own_colptr;
    as->d_rowval = as->own_rowval;
    return FEGPU_OK;
  }
  unsigned long long *kA = nullptr, *kB = nullptr;
  uint32_t *iA = nullptr, *iB = nullptr;
  int32_t *hist = nullptr, *head = nullptr;
  int64_t *offs = nullptr, *segid = nullptr, *segstart = nullptr;
  int *d_err = nullptr;
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  auto cleanup = [&]() {
    cudaFree(kA); cudaFree(kB); cudaFree(iA); cudaFree(iB); cudaFree(hist); cudaFree(head); cudaFree(offs); cudaFree(segid);
    cudaFree(segstart); cudaFree(d_err);
  };
#define SC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
#define ST(expr) do { int32_t _s = (expr); if (_s != FEGPU_OK) { cleanup(); return _s; } } while (0)
  SC(cudaMalloc((void **)&kA, sizeof(unsigned long long) * n));
  SC(cudaMalloc((void **)&kB, sizeof(unsigned long long) * n));
  SC(cudaMalloc((void **)&iA, sizeof(uint32_t) * n));
  SC(cudaMalloc((void **)&iB, sizeof(uint32_t) * n));
  SC(cudaMalloc((void **)&hist, sizeof(int32_t) * 256 * ntiles));
  SC(cudaMalloc((void **)&offs, sizeof(int64_t) * (256 * ntiles + 1)));
  SC(cudaMalloc((void **)&d_err, sizeof(int)));
  SC(cudaMemsetAsync(d_err, 0, sizeof(int), st));
  k_make_keys<<<grid_for(n, 256), 256, 0, st>>>(n, d_I, d_J, nrows, ncols, kA, iA, d_err);
  ctx->launches++;
  int h_err = 0;
  SC(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  SC(cudaStreamSynchronize(st));
  if (h_err) {
    cleanup();
    return fegpu_fail(ctx, FEGPU_ERR_COL_LT1 - (h_err - 1), dof_msg(h_err));
  }
  // digits: row bits in [0, rb), column bits in [32, 32+cb)
  const int rb = bits_for(nrows), cb = bits_for(ncols);
  std::vector<int> shifts;
  for (int s = 0; s < rb; s += 8) shifts.push_back(s);
  for (int s = 0; s < cb; s += 8) shifts.push_back(32 + s);
  unsigned long long *kin = kA, *kout = kB;
  uint32_t *iin = iA, *iout = iB;
  ST(radix_sort_pairs(ctx, n, shifts, &kin, &kout, &iin, &iout, hist, offs));
  // segments
  SC(cudaMalloc((void **)&head, sizeof(int32_t) * n));
  SC(cudaMalloc((void **)&segid, sizeof(int64_t) * (n + 1)));
  k_heads<<<grid_for(n, 256), 256, 0, st>>>(kin, n, head);
  ctx->launches++;
  int64_t nseg = 0;
  ST(fe_exclusive_scan_i32_to_i64(ctx, head, segid, n, 0, true, &nseg));
  {
    size_t capb = as->own_rowval_cap * sizeof(int64_t);
    ST(fe_reserve_bytes(ctx, (void **)&as->own_rowval, &capb, sizeof(int64_t) * (size_t)nseg));
    as->own_rowval_cap = capb / sizeof(int64_t);
  }
  ST(fe_asm_reserve(as, &as->d_nzval, &as->nz_cap, (size_t)nseg));
  SC(cudaMalloc((void **)&segstart, sizeof(int64_t) * (nseg + 1)));
  k_segments<<<grid_for(n, 256), 256, 0, st>>>(kin, head, segid, n, as->own_rowval, as->own_colptr, segstart);
  ctx->launches++;
  int64_t tot = 0;
  ST(fe_exclusive_scan_i64(ctx, as->own_colptr, as->own_colptr, ncols, 1, true, &tot));
  k_segsum<<<grid_for(nseg, 256), 256, 0, st>>>(iin, d_V, segstart, nseg, n, as->d_nzval);
  ctx->launches++;
  SC(cudaGetLastError());
  SC(cudaStreamSynchronize(st));
  cleanup();
#undef SC
#undef ST
  as->nnz = nseg;
  as->d_colptr = as->own_colptr;
  as->d_rowval = as->own_rowval;
  return FEGPU_OK;
}
