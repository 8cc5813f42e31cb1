// Mesh-structured CSC construction: replaces SparseArrays.sparse(I,J,V,m,n) (AssemblyModule.jl:319-325) for
// assemblies whose triplets come from element matrices scattered through one dof map.
//
// Symbolic phase (once per mesh + dof map + partition; cached in the dofmap):
//   node -> element adjacency (CSR, ascending element order)            k_count_adj / k_fill_adj / k_sort_adj
//   node -> sorted unique neighbour nodes (warp per node, bitonic sort) k_nbr<false> (count) / k_nbr<true> (fill)
//   colptr from per-column counts (all ndn columns of a node share one row set), rowval = neighbour dofs sorted,
//   per (node, neighbour) the list of (adjacent element, local row node) sources in ascending triplet order.
// Numeric phase (every assembly): k_gather -- one warp per column node sums, for every stored entry, its source
//   values in ascending element order (the reference's left-to-right duplicate sum), no atomics => bit-reproducible.
//
// The pattern equals sparse()'s: one entry per (row dof, col dof) pair that shares an element, explicit zeros kept,
// rows strictly increasing in a column, 1-based int64 colptr/rowval.
#include "fegpu_internal.h"

struct Pattern {
  int64_t nnz = 0, ncols = 0, nrows = 0;
  int64_t *d_colptr = nullptr;  // [ncols+1] 1-based
  int64_t *d_rowval = nullptr;  // [nnz] 1-based
  int64_t *d_adjptr = nullptr;  // [nnodes+1]
  int32_t *d_adj_slot = nullptr;  // active-element slot
  uint8_t *d_adj_lc = nullptr;    // local node index of this node in that element
  int64_t *d_nbrptr = nullptr;    // [nnodes+1]
  uint16_t *d_srcoff = nullptr;   // per node nnbr+1 entries at nbrptr[n] + n
  uint16_t *d_src = nullptr;      // per node at adjptr[n]*nne: (adj index << 5) | local row node
  uint16_t *d_rank = nullptr;     // per node nnbr*ndn entries at nbrptr[n]*ndn, nullptr when identity
  int maxcand = 0;
};

namespace {

constexpr int WPB = 4;  // warps per block in the per-node kernels

struct SymParams {
  const int32_t *conn;
  const int32_t *elem_list;
  int64_t nactive;
  int nne;
  int64_t nnodes;
  const uint8_t *rowowned;
  const int32_t *dof;  // [ndn][nnodes]
  int ndn;
};

__global__ void k_count_adj(SymParams S, int32_t *deg, int *degenerate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nactive * S.nne) return;
  int64_t slot = i / S.nne;
  int lc = (int)(i % S.nne);
  int64_t e = S.elem_list ? S.elem_list[slot] : slot;
  const int32_t *c = S.conn + e * S.nne;
  int n = c[lc];
  atomicAdd(&deg[n], 1);
  for (int k = 0; k < lc; k++)
    if (c[k] == n) *degenerate = 1;
}

__global__ void k_fill_adj(SymParams S, const int64_t *adjptr, int32_t *cursor, int32_t *adj_slot, uint8_t *adj_lc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nactive * S.nne) return;
  int64_t slot = i / S.nne;
  int lc = (int)(i % S.nne);
  int64_t e = S.elem_list ? S.elem_list[slot] : slot;
  int n = S.conn[e * S.nne + lc];
  int64_t pos = adjptr[n] + atomicAdd(&cursor[n], 1);
  adj_slot[pos] = (int32_t)slot;
  adj_lc[pos] = (uint8_t)lc;
}

// ascending slot order inside every node's list (the atomics above deliver an arbitrary order)
__global__ void k_sort_adj(int64_t nnodes, const int64_t *adjptr, int32_t *adj_slot, uint8_t *adj_lc) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnodes) return;
  int64_t b = adjptr[n], e = adjptr[n + 1];
  for (int64_t i = b + 1; i < e; i++) {
    int32_t s = adj_slot[i];
    uint8_t l = adj_lc[i];
    int64_t j = i - 1;
    while (j >= b && adj_slot[j] > s) {
      adj_slot[j + 1] = adj_slot[j];
      adj_lc[j + 1] = adj_lc[j];
      j--;
    }
    adj_slot[j + 1] = s;
    adj_lc[j + 1] = l;
  }
}

// in-warp bitonic sort of n (power of two) keys in shared memory, ascending
template <typename K>
__device__ void warp_bitonic(K *a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        int ixj = i ^ j;
        if (ixj > i) {
          K x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// One warp per node.  Shared per warp: keys[cap] (uint64; reused as the three work lists), cap = pow2 >= maxcand*max(1,ndn)
// FILL == false: nnbr[n] only.  FILL == true: rowval, rank, srcoff, src.
template <bool FILL>
__global__ void __launch_bounds__(WPB * 32) k_nbr(SymParams S, const int64_t *adjptr, const int32_t *adj_slot, const uint8_t *adj_lc,
                                                  int cap, int32_t *nnbr_out, const int64_t *nbrptr, const int64_t *colptr,
                                                  int64_t *rowval, uint16_t *rank, uint16_t *srcoff, uint16_t *src, int *rank_nonident) {
  extern __shared__ unsigned long long sk[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long *keys = sk + (size_t)w * cap * 2;  // [cap] work keys
  unsigned long long *uniq = keys + cap;                // [cap] unique neighbour node ids (as u64) / second list
  const int nne = S.nne, ndn = S.ndn;
  for (int64_t n = (int64_t)blockIdx.x * WPB + w; n < S.nnodes; n += (int64_t)gridDim.x * WPB) {
    const int64_t ab = adjptr[n];
    const int deg = (int)(adjptr[n + 1] - ab);
    if (deg == 0) {
      if (!FILL && lane == 0) nnbr_out[n] = 0;
      continue;
    }
    const int ncand = deg * nne;
    const int p2 = next_pow2(ncand);
    // candidates: key = (node << 16) | k, k = a*nne + li ; dropped (not an owned row) -> all ones
    for (int k = lane; k < p2; k += 32) {
      unsigned long long key = ~0ull;
      if (k < ncand) {
        int a = k / nne, li = k - a * nne;
        int64_t slot = adj_slot[ab + a];
        int64_t e = S.elem_list ? S.elem_list[slot] : slot;
        int m = S.conn[e * nne + li];
        if (!S.rowowned || S.rowowned[m]) key = ((unsigned long long)(unsigned)m << 16) | (unsigned)k;
      }
      keys[k] = key;
    }
    __syncwarp();
    warp_bitonic(keys, p2, lane);
    // heads of runs of equal node id -> unique list; every candidate learns its neighbour slot
    // pass 1: count heads (ballot prefix)
    int nu = 0;
    for (int base = 0; base < p2; base += 32) {
      int k = base + lane;
      bool valid = (k < p2) && (keys[k] != ~0ull);
      bool head = valid && (k == 0 || (keys[k - 1] >> 16) != (keys[k] >> 16));
      unsigned bal = __ballot_sync(0xffffffffu, head);
      if (head) uniq[nu + __popc(bal & ((1u << lane) - 1))] = keys[k] >> 16;
      nu += __popc(bal);
    }
    __syncwarp();
    if (!FILL) {
      if (lane == 0) nnbr_out[n] = nu;
      continue;
    }
    // ---- sources: sorted keys are already grouped by neighbour (ascending node id = ascending slot s) and, inside a
    // group, ascending k = ascending (adjacent element, local row node) = the reference's triplet order.
    const int64_t nb = nbrptr[n];
    uint16_t *so = srcoff + nb + n;
    uint16_t *sr = src + ab * nne;
    int s_run = 0;  // number of heads seen before this chunk
    int nvalid = 0;
    for (int base = 0; base < p2; base += 32) {
      int k = base + lane;
      bool valid = (k < p2) && (keys[k] != ~0ull);
      bool head = valid && (k == 0 || (keys[k - 1] >> 16) != (keys[k] >> 16));
      unsigned bal = __ballot_sync(0xffffffffu, head);
      if (valid) {
        unsigned kk = (unsigned)(keys[k] & 0xffffu);
        unsigned a = kk / nne, li = kk - a * nne;
        sr[k] = (uint16_t)((a << 5) | li);
        if (head) so[s_run + __popc(bal & ((1u << lane) - 1))] = (uint16_t)k;
      }
      s_run += __popc(bal);
      nvalid += __popc(__ballot_sync(0xffffffffu, valid));
    }
    if (lane == 0) so[nu] = (uint16_t)nvalid;
    __syncwarp();
    // ---- rows: dofs of (neighbour s, component p), sorted ascending -> rowval of every column of this node, and rank
    const int nr = nu * ndn;
    const int q2 = next_pow2(nr);
    for (int i = lane; i < q2; i += 32) {
      unsigned long long key = ~0ull;
      if (i < nr) {
        int s = i / ndn, p = i - s * ndn;
        int m = (int)uniq[s];
        key = ((unsigned long long)(unsigned)S.dof[(int64_t)p * S.nnodes + m] << 16) | (unsigned)i;
      }
      keys[i] = key;
    }
    __syncwarp();
    warp_bitonic(keys, q2, lane);
    bool nonident = false;
    for (int pos = lane; pos < nr; pos += 32) {
      unsigned i = (unsigned)(keys[pos] & 0xffffu);
      int64_t rdof = (int64_t)(keys[pos] >> 16) + 1;
      rank[nb * ndn + i] = (uint16_t)pos;
      if ((int)i != pos) nonident = true;
      for (int q = 0; q < ndn; q++) {
        int64_t J = S.dof[(int64_t)q * S.nnodes + n];
        rowval[colptr[J] - 1 + pos] = rdof;
      }
    }
    if (__any_sync(0xffffffffu, nonident) && lane == 0) *rank_nonident = 1;
    __syncwarp();
  }
}

__global__ void k_col_counts(SymParams S, const int32_t *nnbr, int64_t *colcount) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nnodes * S.ndn) return;
  int64_t n = i % S.nnodes;
  int q = (int)(i / S.nnodes);
  if (nnbr[n] > 0) colcount[S.dof[(int64_t)q * S.nnodes + n]] = (int64_t)nnbr[n] * S.ndn;
}

// ---------------------------------------------------------------------------------------------- numeric gather
struct GatherParams {
  int64_t nnodes;
  int nne, ndn;
  const int64_t *adjptr;
  const int32_t *adj_slot;
  const uint8_t *adj_lc;
  const int64_t *nbrptr;
  const uint16_t *srcoff;
  const uint16_t *src;
  const uint16_t *rank;
  const int32_t *dof;
  const int64_t *colptr;
  const double *V;
  double *nzval;
};

template <int NDN>
__global__ void __launch_bounds__(256) k_gather(const GatherParams G) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int ndn = (NDN > 0) ? NDN : G.ndn;
  const int EM = G.nne * ndn;
  const int64_t EM2 = (int64_t)EM * EM;
  for (int64_t n = warp; n < G.nnodes; n += nwarps) {
    const int64_t nb = G.nbrptr[n];
    const int nn = (int)(G.nbrptr[n + 1] - nb);
    if (nn == 0) continue;
    const int64_t ab = G.adjptr[n];
    const uint16_t *so = G.srcoff + nb + n;
    const uint16_t *sr = G.src + ab * G.nne;
    const int per_col = nn * ndn;
    const int total = per_col * ndn;
    for (int idx = lane; idx < total; idx += 32) {
      const int q = idx / per_col;
      const int rem = idx - q * per_col;
      const int s = rem / ndn;
      const int p = rem - s * ndn;
      const int j0 = so[s], j1 = so[s + 1];
      double v = 0.0;
      for (int j = j0; j < j1; j++) {
        const unsigned code = sr[j];
        const unsigned a = code >> 5, li = code & 31u;
        const int64_t slot = G.adj_slot[ab + a];
        const int lc = G.adj_lc[ab + a];
        v += G.V[slot * EM2 + (int64_t)(lc * ndn + q) * EM + (li * ndn + p)];
      }
      const int64_t J = G.dof[(int64_t)q * G.nnodes + n];
      const int pos = G.rank ? G.rank[nb * ndn + rem] : rem;
      G.nzval[G.colptr[J] - 1 + pos] = v;
    }
  }
}

template <typename T>
int32_t dalloc(fegpu_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CUDA_TRY(ctx, cudaMalloc((void **)p, sizeof(T) * n));
  return FEGPU_OK;
}

}  // namespace

void fe_pattern_free(Pattern *p) {
  if (!p) return;
  cudaFree(p->d_colptr); cudaFree(p->d_rowval); cudaFree(p->d_adjptr); cudaFree(p->d_adj_slot); cudaFree(p->d_adj_lc);
  cudaFree(p->d_nbrptr); cudaFree(p->d_srcoff); cudaFree(p->d_src); cudaFree(p->d_rank);
  delete p;
}
int64_t fe_pattern_nnz(const Pattern *p) { return p->nnz; }
const int64_t *fe_pattern_colptr(const Pattern *p) { return p->d_colptr; }
const int64_t *fe_pattern_rowval(const Pattern *p) { return p->d_rowval; }

bool fe_pattern_usable(const fegpu_dofmap *dm) {
  return dm->injective && !dm->mesh->degenerate && dm->row_nall == dm->col_nall && dm->mesh->nne <= 32;
}

int32_t fe_pattern_build(fegpu_dofmap *dm) {
  fegpu_ctx *ctx = dm->ctx;
  fegpu_mesh *mesh = dm->mesh;
  cudaStream_t st = ctx->stream;
  if (dm->pat) { fe_pattern_free(dm->pat); dm->pat = nullptr; }
  Pattern *P = new Pattern();
  dm->pat = P;  // owned by the dofmap from here on (freed with it, also on error paths)
  P->ncols = dm->col_nall;
  P->nrows = dm->row_nall;
  const int64_t nn = mesh->nnodes;
  SymParams S{mesh->d_conn, mesh->d_elem_list, mesh->nactive, mesh->nne, nn, mesh->d_rowowned, dm->d_dof, dm->ndn};
  const int64_t nadj = mesh->nactive * mesh->nne;

  int32_t *d_deg = nullptr, *d_cursor = nullptr, *d_nnbr = nullptr;
  int *d_flags = nullptr;  // [0] degenerate, [1] rank non-identity
  FE_TRY(dalloc(ctx, &d_deg, nn));
  FE_TRY(dalloc(ctx, &d_cursor, nn));
  FE_TRY(dalloc(ctx, &d_nnbr, nn));
  FE_TRY(dalloc(ctx, &d_flags, 2));
  auto cleanup = [&]() { cudaFree(d_deg); cudaFree(d_cursor); cudaFree(d_nnbr); cudaFree(d_flags); };
#define PT(expr) do { int32_t _s = (expr); if (_s != FEGPU_OK) { cleanup(); return _s; } } while (0)
#define PC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  PC(cudaMemsetAsync(d_deg, 0, sizeof(int32_t) * nn, st));
  PC(cudaMemsetAsync(d_cursor, 0, sizeof(int32_t) * nn, st));
  PC(cudaMemsetAsync(d_flags, 0, sizeof(int) * 2, st));
  if (nadj > 0) {
    k_count_adj<<<grid_for(nadj, 256), 256, 0, st>>>(S, d_deg, d_flags);
    ctx->launches++;
  }
  PT(dalloc(ctx, &P->d_adjptr, nn + 1));
  PT(fe_exclusive_scan_i32_to_i64(ctx, d_deg, P->d_adjptr, nn, 0, true, nullptr));
  int32_t maxdeg = 0;
  PT(fe_max_i32(ctx, d_deg, nn, &maxdeg));
  int h_flags[2] = {0, 0};
  PC(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  if (h_flags[0]) {
    mesh->degenerate = true;
    cleanup();
    fe_pattern_free(P);
    dm->pat = nullptr;
    return FEGPU_OK;  // caller re-checks fe_pattern_usable and takes the sort path
  }
  if (maxdeg < 0) maxdeg = 0;
  P->maxcand = maxdeg * mesh->nne;
  int cap = 32;
  while (cap < P->maxcand * std::max(1, dm->ndn)) cap <<= 1;
  // limits of the packed encodings: adjacency index < 2048 (11 bits), candidate index < 65536, rows per column < 65536
  if (maxdeg >= 2048 || P->maxcand >= 65536 || (int64_t)P->maxcand * dm->ndn >= 65536 || (size_t)cap * 2 * 8 > 200 * 1024) {
    mesh->degenerate = true;  // valence beyond the fast path's encodings: generic sort path handles it
    cleanup();
    fe_pattern_free(P);
    dm->pat = nullptr;
    return FEGPU_OK;
  }
  PT(dalloc(ctx, &P->d_adj_slot, nadj));
  PT(dalloc(ctx, &P->d_adj_lc, nadj));
  if (nadj > 0) {
    k_fill_adj<<<grid_for(nadj, 256), 256, 0, st>>>(S, P->d_adjptr, d_cursor, P->d_adj_slot, P->d_adj_lc);
    k_sort_adj<<<grid_for(nn, 128), 128, 0, st>>>(nn, P->d_adjptr, P->d_adj_slot, P->d_adj_lc);
    ctx->launches += 2;
  }
  // neighbour counts
  int wpb = WPB;
  size_t smem = (size_t)wpb * cap * 2 * sizeof(unsigned long long);
  PC(cudaFuncSetAttribute(k_nbr<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  PC(cudaFuncSetAttribute(k_nbr<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  if (smem > 220 * 1024) {
    cleanup();
    return fegpu_fail(ctx, FEGPU_ERR_ARG, "internal: neighbour work list exceeds shared memory");
  }
  unsigned gridn = (unsigned)std::min<int64_t>((nn + WPB - 1) / WPB, (int64_t)ctx->sm_count * 32);
  if (gridn == 0) gridn = 1;
  k_nbr<false><<<gridn, WPB * 32, smem, st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, cap, d_nnbr, nullptr, nullptr, nullptr, nullptr,
                                              nullptr, nullptr, nullptr);
  ctx->launches++;
  PT(dalloc(ctx, &P->d_nbrptr, nn + 1));
  int64_t total_nbr = 0;
  PT(fe_exclusive_scan_i32_to_i64(ctx, d_nnbr, P->d_nbrptr, nn, 0, true, &total_nbr));
  // column pointers
  PT(dalloc(ctx, &P->d_colptr, P->ncols + 1));
  PC(cudaMemsetAsync(P->d_colptr, 0, sizeof(int64_t) * (P->ncols + 1), st));
  if (nn * dm->ndn > 0) {
    k_col_counts<<<grid_for(nn * dm->ndn, 256), 256, 0, st>>>(S, d_nnbr, P->d_colptr);
    ctx->launches++;
  }
  int64_t tot = 0;
  PT(fe_exclusive_scan_i64(ctx, P->d_colptr, P->d_colptr, P->ncols, 1, true, &tot));
  P->nnz = tot - 1;
  if (P->nnz != total_nbr * dm->ndn * dm->ndn) {
    cleanup();
    return fegpu_fail(ctx, FEGPU_ERR_STATE, "internal: pattern size mismatch");
  }
  PT(dalloc(ctx, &P->d_rowval, (size_t)P->nnz));
  PT(dalloc(ctx, &P->d_rank, (size_t)(total_nbr * dm->ndn)));
  PT(dalloc(ctx, &P->d_srcoff, (size_t)(total_nbr + nn)));
  PT(dalloc(ctx, &P->d_src, (size_t)nadj * mesh->nne));
  k_nbr<true><<<gridn, WPB * 32, smem, st>>>(S, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, cap, nullptr, P->d_nbrptr, P->d_colptr, P->d_rowval,
                                             P->d_rank, P->d_srcoff, P->d_src, d_flags + 1);
  ctx->launches++;
  PC(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  PC(cudaGetLastError());
  if (!h_flags[1]) {  // node-major row order already sorted: the gather can skip the rank table
    cudaFree(P->d_rank);
    P->d_rank = nullptr;
  }
  cleanup();
#undef PT
#undef PC
  dm->pat_topo_version = mesh->topo_version;
  return FEGPU_OK;
}

int32_t fe_gather(fegpu_dofmap *dm, const double *d_V, double *d_nzval) {
  fegpu_ctx *ctx = dm->ctx;
  Pattern *P = dm->pat;
  fegpu_mesh *mesh = dm->mesh;
  if (!P) return fegpu_fail(ctx, FEGPU_ERR_STATE, "no pattern");
  if (P->nnz == 0) return FEGPU_OK;
  GatherParams G{mesh->nnodes, mesh->nne, dm->ndn, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, P->d_nbrptr, P->d_srcoff,
                 P->d_src, P->d_rank, dm->d_dof, P->d_colptr, d_V, d_nzval};
  const int64_t warps_wanted = mesh->nnodes;
  unsigned grid = (unsigned)std::min<int64_t>((warps_wanted + 7) / 8, (int64_t)ctx->sm_count * 64);
  if (grid == 0) grid = 1;
  switch (dm->ndn) {
    case 1: k_gather<1><<<grid, 256, 0, ctx->stream>>>(G); break;
    case 2: k_gather<2><<<grid, 256, 0, ctx->stream>>>(G); break;
    case 3: k_gather<3><<<grid, 256, 0, ctx->stream>>>(G); break;
    default: k_gather<0><<<grid, 256, 0, ctx->stream>>>(G); break;
  }
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}
