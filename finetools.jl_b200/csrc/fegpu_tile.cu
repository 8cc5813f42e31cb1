// Thread-per-node kernels of the CSC construction for small stencils (H8, Q4, T3, T4 with few elements per node): the
// replacement of SparseArrays.sparse(I,J,V,m,n) (AssemblyModule.jl:319-325) on the path BASELINE.json benchmarks.
//
// The group / warp kernels of fegpu_pattern.cu spend most of their instructions on cross-lane traffic (shuffle network,
// ballots, per-key address arithmetic: 260 warp instructions per node on the 256^3 H8 block).  Here ONE THREAD owns a
// node: its <= 64 candidate neighbours are sorted by a compile-time sorting network on registers (fegpu_sortnet.h, 543
// comparators = 1086 VIMNMX), heads are found by a sequential scan of the sorted registers, and a CTA of 128 consecutive
// nodes writes its outputs -- which are contiguous in rowval / nzval when the dof map is node-major affine -- with flat,
// fully coalesced loops.  Three kernels per fresh assembly:
//
//   k_adj_table   one pass over the connectivity: count the elements at every node (atomicAdd, whose return value is the
//                 position) and drop (slot << 5 | local index) into a fixed-capacity row of the node.  Replaces the
//                 count + fill pair of the general path (no second read of conn, no rank array).
//   k_sym_tile    per node: sort the adjacency row, write it to the CSR arrays (ascending element order = the reference's
//                 left-to-right duplicate sum), load the element rows, sort the candidate keys (node << 6 | k), count the
//                 unique neighbours.  Per CTA: block scan + decoupled look-back over the tiles (single pass: the global
//                 prefix of the neighbour counts IS nbrptr and, for an affine dof map, colptr).  Then cslot, the neighbour
//                 lists and rowval go out from shared memory.  Replaces k_nbr_group + scans + k_col_counts + k_rows_sorted
//                 and the intermediate list U (written and read once each in the general path).
//   k_gather_tile numeric phase: thread per (node, column component) walks the node's adjacent elements in ascending order
//                 and adds every value into the CTA's shared-memory image of its slice of nzval (laid out exactly as in
//                 memory, so the write-out is a flat copy).  No atomics, fixed order => bit-reproducible
//                 (test/test_basics.jl:3039-3045).  ~15 instead of ~170 warp instructions per node for scalar H8.
//
// Preconditions, checked on the device and read back with the build's first host round trip: every node has at most MAXDEG
// elements, no element lists a node twice, node ids < 2^26 - 2, and the dof map is node-major affine on the node window
// (dof[p][n] = dof[0][lo] + (n - lo) ndn + p: the default numberdofs! without fixed dofs, FieldModule.jl:328-345).  When one
// fails the caller runs the general path of fegpu_pattern.cu instead (free-first numberings, T10 / H20 / H27, high valences).
// The pattern that comes out is the same set of arrays either way.
#include <cstdlib>

#include "fegpu_internal.h"
#include "fegpu_pattern.h"
#include "fegpu_sortnet.h"

namespace {

constexpr int TILE_T = 128;  // nodes (= threads) per tile of k_sym_tile
constexpr int TILE_KB = 6;   // low bits of a candidate key: k = a * nne + li < 64
constexpr uint32_t TILE_DROPPED = 0xffffffffu >> TILE_KB;  // node field of a candidate whose row this rank does not own / padding
constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_VMASK = (1ull << 62) - 1ull;

struct TileParams {
  const int32_t *conn;
  const int32_t *elem_list;
  int64_t nactive;
  int64_t nnodes;
  int64_t lo, nw;  // node window [lo, lo + nw)
  int32_t own_lo, own_hi;   // PART == 1: owned rows are the node range [own_lo, own_hi)
  const uint8_t *rowowned;  // PART == 2: byte map
  const int32_t *dof;       // [ndn][nnodes]
};

// dof[p][n] == dof[0][lo] + (n - lo) * ndn + p on the whole window?
__global__ void __launch_bounds__(256) k_dof_affine(const int32_t *__restrict__ dof, int64_t nnodes, int ndn, int64_t lo, int64_t nw, int *notaffine) {
  const int64_t d0 = dof[lo];
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x)
    for (int p = 0; p < ndn; p++) bad = bad || ((int64_t)dof[(int64_t)p * nnodes + lo + i] != d0 + i * ndn + p);
  if (bad) *notaffine = 1;
}

template <int NNE, int MAXDEG>
__global__ void __launch_bounds__(256) k_adj_table(const TileParams P, int32_t *__restrict__ deg, uint32_t *__restrict__ tab, int *degenerate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nactive * NNE) return;
  const int64_t slot = i / NNE;
  const int lc = (int)(i - slot * NNE);
  const int64_t e = P.elem_list ? (int64_t)P.elem_list[slot] : slot;
  const int32_t *c = P.conn + e * NNE;
  const int n = c[lc];
  const int pos = atomicAdd(&deg[n], 1);
  if (pos < MAXDEG) tab[((int64_t)n - P.lo) * MAXDEG + pos] = ((uint32_t)slot << 5) | (uint32_t)lc;
  for (int k = 0; k < lc; k++)
    if (c[k] == n) *degenerate = 1;
}

// prefix arrays outside the window: `before` ahead of it, the window's last value behind it (cf. k_fill_outside)
__global__ void k_tile_fill_outside(int64_t *__restrict__ a0, int64_t *__restrict__ a1, int64_t len, int64_t lo, int64_t hi, int64_t before) {
  const int64_t nout = len - (hi - lo + 1);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  const int64_t idx = (i < lo) ? i : i + (hi - lo + 1);
  if (a0) a0[idx] = (i < lo) ? before : a0[hi];
  if (a1) a1[idx] = (i < lo) ? before : a1[hi];
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long *>(p) = v; }

template <int NNE>
__device__ __forceinline__ void load_conn_row(const int32_t *__restrict__ row, int (&m)[NNE]) {
  if constexpr (NNE == 8) {
    const int4 a = __ldg(reinterpret_cast<const int4 *>(row)), b = __ldg(reinterpret_cast<const int4 *>(row) + 1);
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
  } else if constexpr (NNE == 4) {
    const int4 a = __ldg(reinterpret_cast<const int4 *>(row));
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w;
  } else {
#pragma unroll
    for (int k = 0; k < NNE; k++) m[k] = __ldg(row + k);
  }
}

// out[0] = total neighbour entries of the window, out[1] = largest neighbour count, out[2] = tile ticket
template <int NNE, int MAXDEG, int NDN, int PART>
__global__ void __launch_bounds__(TILE_T, 4)
    k_sym_tile(const TileParams P, const int64_t *__restrict__ adjptr, const uint32_t *__restrict__ tab, int32_t *__restrict__ adj_slot,
               uint8_t *__restrict__ adj_lc, int32_t *__restrict__ nnbr, int64_t *__restrict__ nbrptr, int64_t *__restrict__ colptr,
               uint16_t *__restrict__ cslot, int64_t *__restrict__ rowval, int32_t *__restrict__ nbr_out, unsigned long long *tile_state,
               unsigned long long *out) {
  constexpr int NKEY = NNE * MAXDEG;
  constexpr int CSW = NKEY / 2 + 1;  // words per thread of the staged cslot row; odd => conflict-free when the lanes write the same k
  static_assert(NKEY <= 64 && NKEY % 2 == 0 && CSW % 2 == 1, "candidate keys of a node must fit 6 bits");
  extern __shared__ uint32_t smem_u32[];
  uint32_t *U_sm = smem_u32;                   // [TILE_T * NKEY] unique neighbours of the tile's nodes, dense, node after node
  uint32_t *cs_sm = smem_u32 + TILE_T * NKEY;  // [TILE_T][CSW]
  __shared__ int s_tile;
  __shared__ long long s_base;
  __shared__ int s_wtot[TILE_T / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tiles are handed out in the order the CTAs start, so a tile only ever waits for tiles that are already running
  if (tid == 0) s_tile = atomicAdd(reinterpret_cast<int *>(out + 2), 1);
  __syncthreads();
  const int tile = s_tile;
  const int64_t i = (int64_t)tile * TILE_T + tid;  // window-relative node index
  const bool live = i < P.nw;
  const int64_t n = P.lo + i;
  int64_t ab = 0;
  int deg = 0;
  if (live) {
    ab = adjptr[n];
    deg = min((int)(adjptr[n + 1] - ab), MAXDEG);
  }
  // ---- adjacency row: sort by element slot, write the CSR arrays
  uint32_t adj[MAXDEG];
  {
    const uint4 *row = reinterpret_cast<const uint4 *>(tab + i * MAXDEG);
#pragma unroll
    for (int v = 0; v < MAXDEG / 4; v++) {
      uint4 x = make_uint4(~0u, ~0u, ~0u, ~0u);
      if (4 * v < deg) x = row[v];
      adj[4 * v + 0] = (4 * v + 0 < deg) ? x.x : ~0u;
      adj[4 * v + 1] = (4 * v + 1 < deg) ? x.y : ~0u;
      adj[4 * v + 2] = (4 * v + 2 < deg) ? x.z : ~0u;
      adj[4 * v + 3] = (4 * v + 3 < deg) ? x.w : ~0u;
    }
  }
  fesort::sort<MAXDEG>(adj);
  // ---- candidate keys: (neighbour node << 6) | k, k = a * NNE + li; rows of other ranks and padding carry the all-ones node.
  // All loads first (element ids, then the connectivity rows: 2 x MAXDEG independent requests in flight per thread); the
  // stores of the sorted adjacency come after the key sort so that nothing orders the loads behind them.
  const int32_t *__restrict__ conn = P.conn;
  const int32_t *__restrict__ elem_list = P.elem_list;
  uint32_t el[MAXDEG];
#pragma unroll
  for (int j = 0; j < MAXDEG; j++) {
    el[j] = adj[j] >> 5;
    if (PART != 0 && j < deg) el[j] = (uint32_t)__ldg(elem_list + el[j]);  // partitioned meshes keep a list of active elements
  }
  uint32_t keys[NKEY];
#pragma unroll
  for (int j = 0; j < MAXDEG; j++) {
    int m[NNE];
#pragma unroll
    for (int li = 0; li < NNE; li++) m[li] = -1;
    if (j < deg) load_conn_row<NNE>(conn + (int64_t)el[j] * NNE, m);
#pragma unroll
    for (int li = 0; li < NNE; li++) {
      uint32_t node = (uint32_t)m[li];
      if (PART == 1 && (m[li] < P.own_lo || m[li] >= P.own_hi)) node = TILE_DROPPED;
      if (PART == 2 && j < deg && !P.rowowned[m[li]]) node = TILE_DROPPED;
      if (j >= deg) node = TILE_DROPPED;
      keys[j * NNE + li] = (node << TILE_KB) | (uint32_t)(j * NNE + li);
    }
  }
  fesort::sort<NKEY>(keys);
#pragma unroll
  for (int j = 0; j < MAXDEG; j++)
    if (j < deg) {
      adj_slot[ab + j] = (int32_t)(adj[j] >> 5);
      adj_lc[ab + j] = (uint8_t)(adj[j] & 31u);
    }
  // ---- unique neighbours of this node
  int nu = 0;
  {
    uint32_t prev = TILE_DROPPED;
#pragma unroll
    for (int x = 0; x < NKEY; x++) {
      const uint32_t node = keys[x] >> TILE_KB;
      nu += (node != TILE_DROPPED && node != prev) ? 1 : 0;
      prev = node;
    }
  }
  // ---- prefix of the counts: inside the CTA, then over the tiles (decoupled look-back)
  int incl = nu;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < TILE_T / 32; w++) {
    const int t = s_wtot[w];
    if (w < warp) woff += t;
    total += t;
  }
  const int excl = woff + incl - nu;
  if (warp == 0) {
    long long run = 0;
    if (tile == 0) {
      if (lane == 0) st_state(tile_state, ST_PREFIX | (unsigned long long)total);
    } else {
      if (lane == 0) st_state(tile_state + tile, ST_AGG | (unsigned long long)total);
      int look = tile - 1;
      while (true) {
        const int idx = look - lane;
        unsigned long long st;
        do {
          st = (idx >= 0) ? ld_state(tile_state + idx) : ST_PREFIX;  // tiles before the first one: prefix 0
        } while (__any_sync(0xffffffffu, (st >> 62) == 0ull));
        const unsigned pm = __ballot_sync(0xffffffffu, (st >> 62) == 2ull);
        const int first = pm ? (__ffs(pm) - 1) : 32;
        long long v = (lane <= first) ? (long long)(st & ST_VMASK) : 0ll;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        run += v;
        if (pm) break;
        look -= 32;
      }
      if (lane == 0) st_state(tile_state + tile, ST_PREFIX | (unsigned long long)(run + total));
    }
    if (lane == 0) s_base = run;
  }
  // ---- neighbour slot of every candidate (staged row) and the unique list (dense, at the node's offset inside the tile)
  {
    uint16_t *cs16 = reinterpret_cast<uint16_t *>(cs_sm + tid * CSW);
    int slot = -1;
    uint32_t prev = TILE_DROPPED;
#pragma unroll
    for (int x = 0; x < NKEY; x++) {
      const uint32_t node = keys[x] >> TILE_KB, k = keys[x] & ((1u << TILE_KB) - 1u);
      const bool valid = node != TILE_DROPPED;
      const bool head = valid && node != prev;
      slot += head ? 1 : 0;
      cs16[k] = valid ? (uint16_t)slot : (uint16_t)0xffffu;
      if (head) U_sm[excl + slot] = node;
      prev = node;
    }
  }
  __syncthreads();
  const long long base = s_base;
  const int64_t dof0 = P.dof[P.lo];
  if (live) {
    const long long nb = base + excl;
    nnbr[n] = nu;
    nbrptr[n] = nb;
    const int64_t c0 = dof0 + i * NDN;
#pragma unroll
    for (int q = 0; q < NDN; q++) colptr[c0 + q] = 1 + (nb * NDN + (long long)q * nu) * NDN;
    if (i == P.nw - 1) {
      nbrptr[n + 1] = nb + nu;
      colptr[c0 + NDN] = 1 + (nb + nu) * (long long)(NDN * NDN);
      out[0] = (unsigned long long)(nb + nu);
    }
  }
  {
    int mx = nu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if (lane == 0 && mx > 0) atomicMax(out + 1, (unsigned long long)mx);
  }
  // ---- outputs of the tile: contiguous in memory
  if (nbr_out)
    for (int idx = tid; idx < total; idx += TILE_T) nbr_out[base + idx] = (int32_t)U_sm[idx];
  if (NDN == 1) {
    for (int idx = tid; idx < total; idx += TILE_T) rowval[base + idx] = dof0 + ((int64_t)U_sm[idx] - P.lo) + 1;
  } else {
    for (int t = 0; t < 32; t++) {
      const int nu_t = __shfl_sync(0xffffffffu, nu, t);
      const int ex_t = __shfl_sync(0xffffffffu, excl, t);
      if (nu_t == 0) continue;
      const long long rb = (base + ex_t) * (long long)(NDN * NDN);
      const int per_col = nu_t * NDN;
      for (int r = lane; r < per_col; r += 32) {
        const int s = r / NDN, p = r - s * NDN;
        const int64_t rd = dof0 + ((int64_t)U_sm[ex_t + s] - P.lo) * NDN + p + 1;
#pragma unroll
        for (int q = 0; q < NDN; q++) rowval[rb + (long long)q * per_col + r] = rd;
      }
    }
  }
  for (int t = 0; t < 32; t++) {
    const int deg_t = __shfl_sync(0xffffffffu, deg, t);
    const long long ab_t = __shfl_sync(0xffffffffu, (long long)ab, t);
    if (deg_t == 0) continue;
    const uint32_t *src = cs_sm + (warp * 32 + t) * CSW;
    if (NNE % 2 == 0) {
      uint32_t *dst = reinterpret_cast<uint32_t *>(cslot) + (ab_t * NNE) / 2;
      for (int w = lane; w < deg_t * NNE / 2; w += 32) dst[w] = src[w];
    } else {
      uint16_t *dst = cslot + ab_t * NNE;
      const uint16_t *s16 = reinterpret_cast<const uint16_t *>(src);
      for (int k = lane; k < deg_t * NNE; k += 32) dst[k] = s16[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ numeric phase
struct TileGatherParams {
  int64_t lo, nw, nnodes;
  const int64_t *adjptr;
  const int32_t *adj_slot;
  const uint8_t *adj_lc;
  const int32_t *nnbr;
  const uint16_t *cslot;
  const int64_t *colptr;
  const int32_t *dof;
  const double *V;
  double *nzval;
};

// the NNE neighbour slots (uint16) of one (node, adjacent element), packed two per register
template <int NNE>
struct CsRow {
  uint32_t w[(NNE + 1) / 2];
  __device__ __forceinline__ void fill() {
#pragma unroll
    for (int k = 0; k < (NNE + 1) / 2; k++) w[k] = 0xffffffffu;
  }
  __device__ __forceinline__ void load(const uint16_t *__restrict__ cp) {
    if constexpr (NNE == 8) {
      const uint4 x = *reinterpret_cast<const uint4 *>(cp);
      w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
    } else if constexpr (NNE == 4) {
      const uint2 x = *reinterpret_cast<const uint2 *>(cp);
      w[0] = x.x; w[1] = x.y;
    } else {
#pragma unroll
      for (int k = 0; k < NNE; k++) {
        const uint32_t u = cp[k];
        if (k & 1) w[k >> 1] = (w[k >> 1] & 0xffffu) | (u << 16);
        else w[k >> 1] = (w[k >> 1] & 0xffff0000u) | u;
      }
    }
  }
  __device__ __forceinline__ unsigned get(int li) const { return (li & 1) ? (w[li >> 1] >> 16) : (w[li >> 1] & 0xffffu); }
};

template <int NDN>
struct GatherShape {
  static constexpr int T = (NDN == 3) ? 96 : 128;  // threads per CTA: a multiple of NDN and of 32
  static constexpr int NPB = T / NDN;              // nodes per CTA
};

template <int NNE, int MAXDEG, int NDN, bool COMPACT>
__global__ void __launch_bounds__(GatherShape<NDN>::T) k_gather_tile(const TileGatherParams G) {
  extern __shared__ double acc[];  // the CTA's slice of nzval: columns of its nodes, exactly as in memory
  constexpr int T = GatherShape<NDN>::T, NPB = GatherShape<NDN>::NPB;
  constexpr int EM = NNE * NDN, ND2 = NDN * NDN;
  constexpr int64_t VPE = COMPACT ? (int64_t)(NNE * (NNE + 1) / 2) * ND2 : (int64_t)EM * EM;
  const int tid = threadIdx.x;
  const int ln = tid / NDN, q = tid - ln * NDN;
  const int64_t i0 = (int64_t)blockIdx.x * NPB;
  const int64_t i = i0 + ln;
  const bool live = i < G.nw;
  const int64_t n = G.lo + i;
  const int64_t dof0 = G.dof[G.lo];
  const int64_t iend = min(i0 + (int64_t)NPB, G.nw);
  const int64_t cb0 = G.colptr[dof0 + i0 * NDN] - 1, cbE = G.colptr[dof0 + iend * NDN] - 1;
  const int total = (int)(cbE - cb0);
  for (int idx = tid; idx < total; idx += T) acc[idx] = 0.0;
  int deg = 0, nu = 0, off = 0;
  int64_t ab = 0;
  if (live) {
    nu = G.nnbr[n];
    ab = G.adjptr[n];
    deg = min((int)(G.adjptr[n + 1] - ab), MAXDEG);
    off = (int)(G.colptr[dof0 + i * NDN + q] - 1 - cb0);
  }
  __syncthreads();
  if (nu > 0) {
    double *col = acc + off;
    const int32_t *__restrict__ adj_slot = G.adj_slot;
    const uint8_t *__restrict__ adj_lc = G.adj_lc;
    const uint16_t *__restrict__ cslot = G.cslot;
    const double *__restrict__ V = G.V;
    // metadata of every adjacent element first (3 x MAXDEG independent loads in flight), then per element: all its value
    // loads, then the adds.  Rows of other ranks (slot 0xffff) are loaded as well and dropped at the add: no predicate
    // between the loads.
    int64_t vb[MAXDEG];
    int lcs[MAXDEG];
    CsRow<NNE> cs[MAXDEG];
#pragma unroll
    for (int j = 0; j < MAXDEG; j++) {
      vb[j] = 0;
      lcs[j] = 0;
      cs[j].fill();
      if (j < deg) {
        vb[j] = (int64_t)adj_slot[ab + j] * VPE;
        lcs[j] = adj_lc[ab + j];
        cs[j].load(cslot + (ab + j) * NNE);
      }
    }
#pragma unroll
    for (int j = 0; j < MAXDEG; j++) {
      if (j < deg) {
        const double *Vb = V + vb[j];
        const int lc = lcs[j];
        double v[NNE][NDN];
#pragma unroll
        for (int li = 0; li < NNE; li++) {
          if (COMPACT) {
            // block (min, max) of the upper block triangle, entry (comp of min, comp of max) at comp_max * NDN + comp_min
            const bool tr = li > lc;
            const int blk = tr ? li * (li + 1) / 2 + lc : lc * (lc + 1) / 2 + li;
            const double *B = Vb + ND2 * blk;
#pragma unroll
            for (int p = 0; p < NDN; p++) v[li][p] = tr ? B[p * NDN + q] : B[q * NDN + p];
          } else {
            const double *B = Vb + (lc * NDN + q) * EM + li * NDN;  // emission order: column (lc, q), rows (li, p)
#pragma unroll
            for (int p = 0; p < NDN; p++) v[li][p] = B[p];
          }
        }
#pragma unroll
        for (int li = 0; li < NNE; li++) {
          const unsigned s = cs[j].get(li);
          if (s != 0xffffu) {
            double *dst = col + s * NDN;
#pragma unroll
            for (int p = 0; p < NDN; p++) dst[p] += v[li][p];
          }
        }
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < total; idx += T) G.nzval[cb0 + idx] = acc[idx];
}

template <typename T>
int32_t talloc(fegpu_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  return fe_dev_alloc(ctx, (void **)p, sizeof(T) * std::max<size_t>(n, 1), ctx->stream);
}

int tile_maxdeg_for(int nne) { return nne == 8 ? 8 : ((nne == 4 || nne == 3) ? 16 : 0); }

}  // namespace

// Thread-per-node symbolic phase.  *taken = false: the preconditions do not hold, nothing was built, the caller runs the
// general path.  *taken = true with dm->pat == nullptr: degenerate elements (the caller takes the sort path).
int32_t fe_tile_build(fegpu_dofmap *dm, const std::function<int32_t()> *fork, bool *taken) {
  *taken = false;
  fegpu_ctx *ctx = dm->ctx;
  fegpu_mesh *mesh = dm->mesh;
  cudaStream_t st = ctx->stream;
  static const bool tile_off = std::getenv("FEGPU_TILE") && std::atoi(std::getenv("FEGPU_TILE")) == 0;  // A/B knob
  const int nne = mesh->nne, ndn = dm->ndn;
  const int MD = tile_maxdeg_for(nne);
  const int64_t nn = mesh->nnodes;
  const int64_t lo = mesh->win_lo, hi = mesh->win_hi, nw = hi - lo;
  if (tile_off || MD == 0 || ndn > 3 || nw <= 0 || mesh->nactive <= 0) return FEGPU_OK;
  if (dm->tile_failed_version == mesh->topo_version) return FEGPU_OK;  // this mesh / numbering already failed the preconditions
  if (nn > (int64_t)TILE_DROPPED - 1 || mesh->nactive >= ((int64_t)1 << 27)) return FEGPU_OK;
  const int64_t nadj = mesh->nactive * nne;

  int32_t *d_deg = nullptr;
  uint32_t *d_tab = nullptr;
  int *d_flags = nullptr;  // [0] degenerate, [1] dof map not affine, [2] largest degree
  unsigned long long *d_state = nullptr, *d_out = nullptr;
  Pattern *P = nullptr;
  auto cleanup = [&]() {
    void *ptrs[] = {d_deg, d_tab, d_flags, d_state, d_out};
    for (void *q : ptrs)
      if (q) fe_dev_free(ctx, q, st);
    d_deg = nullptr; d_tab = nullptr; d_flags = nullptr; d_state = nullptr; d_out = nullptr;
  };
  auto drop_pattern = [&]() {
    if (P) fe_pattern_free(P);
    P = nullptr;
    dm->pat = nullptr;
  };
#define PT(expr) do { int32_t _s = (expr); if (_s != FEGPU_OK) { cleanup(); drop_pattern(); return _s; } } while (0)
#define PC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); drop_pattern(); return fegpu_fail(ctx, FEGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  FE_TRACE("tile build: enter");
  fe_mark(ctx, "sym:start");
  PT(talloc(ctx, &d_deg, (size_t)nn));
  PT(talloc(ctx, &d_tab, (size_t)nw * MD));
  PT(talloc(ctx, &d_flags, 4));
  PC(cudaMemsetAsync(d_deg + lo, 0, sizeof(int32_t) * nw, st));
  PC(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, st));
  TileParams TP{mesh->d_conn, mesh->d_elem_list, mesh->nactive, nn, lo, nw, (int32_t)mesh->own_lo, (int32_t)mesh->own_hi, mesh->d_rowowned, dm->d_dof};
  k_dof_affine<<<(unsigned)std::min<int64_t>(grid_for(nw, 256), (int64_t)ctx->sm_count * 8), 256, 0, st>>>(dm->d_dof, nn, ndn, lo, nw, d_flags + 1);
  switch (nne) {
    case 8: k_adj_table<8, 8><<<grid_for(nadj, 256), 256, 0, st>>>(TP, d_deg, d_tab, d_flags); break;
    case 4: k_adj_table<4, 16><<<grid_for(nadj, 256), 256, 0, st>>>(TP, d_deg, d_tab, d_flags); break;
    default: k_adj_table<3, 16><<<grid_for(nadj, 256), 256, 0, st>>>(TP, d_deg, d_tab, d_flags); break;
  }
  ctx->launches += 2;
  fe_mark(ctx, "sym:k_adj_table");
  if (dm->pat) { fe_pattern_free(dm->pat); dm->pat = nullptr; }
  P = new Pattern();
  dm->pat = P;
  P->ctx = ctx;
  P->stream = st;
  P->alloc_stream = st;
  P->ncols = dm->col_nall;
  P->nrows = dm->row_nall;
  PT(talloc(ctx, &P->d_adjptr, (size_t)nn + 1));
  PT(fe_exclusive_scan_i32_to_i64(ctx, d_deg + lo, P->d_adjptr + lo, nw, 0, true, nullptr));
  PT(fe_max_i32_dev(ctx, d_deg + lo, nw, d_flags + 2));
  int h_flags[4] = {0, 0, 0, 0};
  PC(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  FE_TRACE("tile build: sync A done");
  fe_mark(ctx, "sym:scan_adj");
  if (h_flags[0]) {  // an element lists a node twice: no structured path at all (same as the general build's bail)
    mesh->degenerate = true;
    cleanup();
    drop_pattern();
    *taken = true;
    return FEGPU_OK;
  }
  if (h_flags[1] || h_flags[2] > MD) {
    dm->tile_failed_version = mesh->topo_version;
    cleanup();
    drop_pattern();
    return FEGPU_OK;  // *taken stays false: general path
  }
  *taken = true;
  if (fork) PT((*fork)());
  const int maxdeg = std::max(h_flags[2], 1);
  P->maxdeg = maxdeg;
  P->maxcand = maxdeg * nne;
  const int64_t ntiles = (nw + TILE_T - 1) / TILE_T;
  const size_t nb_cap = (size_t)nadj * nne;  // upper bound of the neighbour entries: every candidate unique
  PT(talloc(ctx, &P->d_adj_slot, (size_t)nadj));
  PT(talloc(ctx, &P->d_adj_lc, (size_t)nadj));
  PT(talloc(ctx, &P->d_nnbr, (size_t)nn));
  PT(talloc(ctx, &P->d_nbrptr, (size_t)nn + 1));
  PT(talloc(ctx, &P->d_colptr, (size_t)P->ncols + 1));
  PT(talloc(ctx, &P->d_cslot, nb_cap));
  // rowval is sized before its length is known (single pass): the bound is what the reference's COO would hold per column
  // node, nnz is typically 0.42 of it (H8)
  PT(talloc(ctx, &P->d_rowval, nb_cap * ndn * ndn));
  if (ndn >= 2) PT(talloc(ctx, &P->d_nbr, nb_cap));
  PT(talloc(ctx, &d_state, (size_t)ntiles));
  PT(talloc(ctx, &d_out, 4));
  PC(cudaMemsetAsync(d_state, 0, sizeof(unsigned long long) * ntiles, st));
  PC(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long) * 4, st));
  if (nw < nn) {  // consumers that walk every node (the general gather as an A/B partner) must see empty nodes outside the window
    if (lo > 0) PC(cudaMemsetAsync(P->d_nnbr, 0, sizeof(int32_t) * lo, st));
    if (hi < nn) PC(cudaMemsetAsync(P->d_nnbr + hi, 0, sizeof(int32_t) * (nn - hi), st));
  }
  const int part = !mesh->d_rowowned ? 0 : (mesh->own_contig ? 1 : 2);
  const size_t smem = sizeof(uint32_t) * (size_t)TILE_T * (nne * MD + nne * MD / 2 + 1);
#define SYM_LAUNCH(NNE_, MD_, NDN_, PART_)                                                                                          \
  do {                                                                                                                              \
    PC(cudaFuncSetAttribute(k_sym_tile<NNE_, MD_, NDN_, PART_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    k_sym_tile<NNE_, MD_, NDN_, PART_><<<(unsigned)ntiles, TILE_T, smem, st>>>(TP, P->d_adjptr, d_tab, P->d_adj_slot, P->d_adj_lc, \
        P->d_nnbr, P->d_nbrptr, P->d_colptr, P->d_cslot, P->d_rowval, P->d_nbr, d_state, d_out);                                    \
  } while (0)
#define SYM_PART(NNE_, MD_, NDN_)                        \
  do {                                                   \
    if (part == 0) SYM_LAUNCH(NNE_, MD_, NDN_, 0);       \
    else if (part == 1) SYM_LAUNCH(NNE_, MD_, NDN_, 1);  \
    else SYM_LAUNCH(NNE_, MD_, NDN_, 2);                 \
  } while (0)
#define SYM_NDN(NNE_, MD_)                    \
  do {                                        \
    if (ndn == 1) SYM_PART(NNE_, MD_, 1);     \
    else if (ndn == 2) SYM_PART(NNE_, MD_, 2); \
    else SYM_PART(NNE_, MD_, 3);              \
  } while (0)
  switch (nne) {
    case 8: SYM_NDN(8, 8); break;
    case 4: SYM_NDN(4, 16); break;
    default: SYM_NDN(3, 16); break;
  }
#undef SYM_NDN
#undef SYM_PART
#undef SYM_LAUNCH
  ctx->launches++;
  PC(cudaGetLastError());
  fe_mark(ctx, "sym:k_sym_tile");
  if (nw < nn) {
    k_tile_fill_outside<<<grid_for(nn - nw, 256), 256, 0, st>>>(P->d_adjptr, P->d_nbrptr, nn + 1, lo, hi, 0);
    ctx->launches++;
  }
  // colptr: the window's columns are dof0 .. dof0 + nw*ndn; constants on both sides.  dof0 is read on the host below, so the
  // fill of the outside runs after the second round trip (it is tiny)
  unsigned long long h_out[4] = {0, 0, 0, 0};
  int32_t h_dof0 = 0;
  PC(cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, st));
  PC(cudaMemcpyAsync(&h_dof0, dm->d_dof + lo, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  FE_TRACE("tile build: sync B done");
  const int64_t dlo = h_dof0, dhi = dlo + nw * ndn;  // colptr[dlo .. dhi] written by the kernel
  if (dhi - dlo < P->ncols) {
    k_tile_fill_outside<<<grid_for(P->ncols - (dhi - dlo), 256), 256, 0, st>>>(P->d_colptr, nullptr, P->ncols + 1, dlo, dhi, 1);
    ctx->launches++;
  }
  P->total_nbr = (int64_t)h_out[0];
  P->nnz = P->total_nbr * ndn * ndn;
  P->maxnbr = std::max((int)h_out[1], 1);
  P->d_dof = dm->d_dof;
  P->ndn = ndn;
  P->nnodes = nn;
  P->tile = true;
  P->tile_lo = lo;
  P->tile_nw = nw;
  P->tile_md = MD;
  if (P->d_nbr && P->total_nbr == 0) {  // fe_pattern_compressed keys on d_nbr: nothing to compress
    fe_dev_free(ctx, P->d_nbr, st);
    P->d_nbr = nullptr;
  }
  PC(cudaEventCreateWithFlags(&P->ready, cudaEventDisableTiming));
  PC(cudaEventRecord(P->ready, st));
  fe_mark(ctx, "sym:finish");
  cleanup();
  FE_TRACE("tile build: done");
#undef PT
#undef PC
  dm->pat_topo_version = mesh->topo_version;
  return FEGPU_OK;
}

// Numeric phase on a pattern built by fe_tile_build.  *taken = false: not applicable (the general gather runs).
int32_t fe_tile_gather(fegpu_dofmap *dm, const double *d_V, bool compact, double *d_nzval, bool *taken) {
  *taken = false;
  fegpu_ctx *ctx = dm->ctx;
  Pattern *P = dm->pat;
  fegpu_mesh *mesh = dm->mesh;
  static const bool gather_off = std::getenv("FEGPU_TILE_GATHER") && std::atoi(std::getenv("FEGPU_TILE_GATHER")) == 0;  // A/B knob
  if (!P || !P->tile || gather_off) return FEGPU_OK;
  const int nne = mesh->nne, ndn = dm->ndn;
  if (tile_maxdeg_for(nne) != P->tile_md || ndn > 3) return FEGPU_OK;
  const int T = (ndn == 3) ? 96 : 128, npb = T / ndn;
  const size_t smem = sizeof(double) * (size_t)npb * P->maxnbr * ndn * ndn;
  if (smem > 200 * 1024) return FEGPU_OK;
  *taken = true;
  if (P->nnz == 0) return FEGPU_OK;
  TileGatherParams G{P->tile_lo, P->tile_nw, mesh->nnodes, P->d_adjptr, P->d_adj_slot, P->d_adj_lc, P->d_nnbr, P->d_cslot, P->d_colptr, dm->d_dof, d_V, d_nzval};
  const unsigned grid = (unsigned)((P->tile_nw + npb - 1) / npb);
#define G_LAUNCH(NNE_, MD_, NDN_, C_)                                                                                               \
  do {                                                                                                                              \
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_gather_tile<NNE_, MD_, NDN_, C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_gather_tile<NNE_, MD_, NDN_, C_><<<grid, GatherShape<NDN_>::T, smem, ctx->stream>>>(G);                                       \
  } while (0)
#define G_C(NNE_, MD_, NDN_)                        \
  do {                                              \
    if (compact) G_LAUNCH(NNE_, MD_, NDN_, true);   \
    else G_LAUNCH(NNE_, MD_, NDN_, false);          \
  } while (0)
#define G_NDN(NNE_, MD_)                  \
  do {                                    \
    if (ndn == 1) G_C(NNE_, MD_, 1);      \
    else if (ndn == 2) G_C(NNE_, MD_, 2); \
    else G_C(NNE_, MD_, 3);               \
  } while (0)
  switch (nne) {
    case 8: G_NDN(8, 8); break;
    case 4: G_NDN(4, 16); break;
    default: G_NDN(3, 16); break;
  }
#undef G_NDN
#undef G_C
#undef G_LAUNCH
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}
