// Thread-per-node kernels of the CSC construction for small stencils (H8, Q4, T3, T4 with few elements per node): the
// replacement of SparseArrays.sparse(I,J,V,m,n) (AssemblyModule.jl:319-325) on the path BASELINE.json benchmarks.
//
// The group / warp kernels of fegpu_pattern.cu spend most of their instructions on cross-lane traffic (shuffle network,
// ballots, per-key address arithmetic: 260 warp instructions per node on the 256^3 H8 block).  Here ONE THREAD owns a
// node: its <= 64 candidate neighbours are sorted by a compile-time sorting network on registers (fegpu_sortnet.h, 543
// comparators = 1086 VIMNMX), heads are found by a sequential scan of the sorted registers, and a CTA of 128 consecutive
// nodes writes its outputs -- which are contiguous in rowval / nzval when the dof map is node-major affine -- with flat,
// fully coalesced loops.
//
// Data layout: PLANES.  A thread-per-node kernel whose lanes each read their own record touches 32 cache lines per load
// instruction, and the L1 data pipe retires about one line per cycle: the first version of the numeric kernel (element
// records of 288 B, one per lane) ran at 88 % of the L1 wavefront peak and 32 % of DRAM (profiles/r02_ncu_tile_v1.txt).
// So everything a node reads lives in struct-of-arrays form, indexed [entry][node] or [entry][element slot]:
//   adj  uint32 [MAXDEG][nwp]   adjacent element j of window node i: (slot << 5) | local index, ascending slot
//   cs   word   [MAXDEG][nwp]   neighbour slot of each of the NNE candidates of (node, adjacent element): one byte each
//                               (0xff = row not owned by this rank), packed in one 64-bit (H8) / 32-bit (Q4, T3, T4) word
//   V    double [values per element][vstride]   element matrices as written by the integration kernels (FormArgs::planes): one plane
//                               per position of the (compact) element record
// Consecutive lanes = consecutive nodes read consecutive words; with elements stored in ascending-smallest-node order
// (fe_order_elements) the j-th adjacent elements of consecutive nodes are consecutive slots, so the V loads of a warp
// fall into one or two lines as well.
//
// Three kernels per fresh assembly:
//   k_adj_place   one thread per element: drop (slot << 5 | local index) into plane `local index` of each of its nodes' columns,
//                 no atomics (optimistic: verified by the entry count, see the kernel).  k_adj_table is the version with an
//                 atomicAdd per (element, node) whose return value is the position; triangles and colliding meshes use it.
//   k_sym_tile    per node: sort the adjacency column (ascending slot = the order of the duplicate sum), load the element
//                 rows, sort the candidate keys (node << 6 | k), count the unique neighbours.  Per CTA: block scan +
//                 decoupled look-back over the tiles (single pass: the global prefix of the neighbour counts IS nbrptr and,
//                 for an affine dof map, colptr).  Then the cs planes, the neighbour lists and rowval go out.
//   k_gather_tile numeric phase: thread per (node, column component) walks the node's adjacent elements in ascending order
//                 and adds every value into the CTA's shared-memory image of its slice of nzval (laid out exactly as in
//                 memory, so the write-out is a flat copy).  No atomics, fixed order => bit-reproducible
//                 (test/test_basics.jl:3039-3045).  The loads of the next element are in flight during the adds of this one;
//                 what it took to make that true (scoreboards, L1 carve-out) is in the kernel's comment.
//
// Preconditions, checked on the device and read back with the build's ONE host round trip (the kernels are launched
// optimistically and are safe when a precondition fails): every node has at most MAXDEG elements, no element lists a node
// twice, node ids < 2^26 - 2, and the dof map is node-major affine on the node window (dof[p][n] = dof[0][lo] + (n - lo) ndn
// + p: the default numberdofs! without fixed dofs, FieldModule.jl:328-345).  When one fails the caller runs the general path
// of fegpu_pattern.cu instead (free-first numberings, T10 / H20 / H27, high valences) and the failure is remembered.
#include <cstdlib>

#include "fegpu_internal.h"
#include "fegpu_pattern.h"
#include "fegpu_sortnet.h"

namespace {

// nodes (= threads) per tile of k_sym_tile: a template parameter.  Measured on config 4: 64 nodes (8 CTAs per SM) 3.34 ms, 128 nodes
// 3.11 ms (later 2.85), 256 nodes (2 CTAs per SM) 2.78 ms -- the default for H8 (FEGPU_TILE_T = 64 / 128: A/B knob)
constexpr int TILE_KB = 6;   // low bits of a candidate key: k = a * nne + li < 64
constexpr uint32_t TILE_DROPPED = 0xffffffffu >> TILE_KB;  // largest node field of a key; the field holds node id + 1 (0 = dropped candidate / padding)
constexpr unsigned long long ST_AGG = 1ull << 62, ST_PREFIX = 2ull << 62, ST_VMASK = (1ull << 62) - 1ull;

template <int NNE>
struct CsWord {
  using type = uint32_t;
};
template <>
struct CsWord<8> {
  using type = unsigned long long;
};

struct TileParams {
  const int32_t *conn;
  const int32_t *elem_list;
  int64_t nactive;
  int64_t nnodes;
  int64_t lo, nw, nwp;      // node window [lo, lo + nw); plane stride nwp >= nw
  int32_t own_lo, own_hi;   // PART == 1: owned rows are the node range [own_lo, own_hi)
  const uint8_t *rowowned;  // PART == 2: byte map
  const int32_t *dof;       // [ndn][nnodes]
  int64_t ncols;
};

// dof[p][n] == dof[0][lo] + (n - lo) * ndn + p on the whole window?
__global__ void __launch_bounds__(256) k_dof_affine(const int32_t *__restrict__ dof, int64_t nnodes, int ndn, int64_t lo, int64_t nw, int *notaffine) {
  const int64_t d0 = dof[lo];
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (int64_t)gridDim.x * blockDim.x)
    for (int p = 0; p < ndn; p++) bad = bad || ((int64_t)dof[(int64_t)p * nnodes + lo + i] != d0 + i * ndn + p);
  if (bad) *notaffine = 1;
}

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

// flags: [0] an element lists a node twice, [2] a node has more than MAXDEG elements
template <int NNE, int MAXDEG>
__global__ void __launch_bounds__(256) k_adj_table(const TileParams P, int32_t *__restrict__ deg, uint32_t *__restrict__ tab, int *flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nactive * NNE) return;
  const int64_t slot = i / NNE;
  const int lc = (int)(i - slot * NNE);
  const int64_t e = P.elem_list ? (int64_t)P.elem_list[slot] : slot;
  const int32_t *c = P.conn + e * NNE;
  const int n = c[lc];
  const int pos = atomicAdd(&deg[n - P.lo], 1);
  if (pos < MAXDEG) tab[(int64_t)pos * P.nwp + (n - P.lo)] = ((uint32_t)slot << 5) | (uint32_t)lc;
  else flags[2] = 1;
  for (int k = 0; k < lc; k++)
    if (c[k] == n) flags[0] = 1;
}

// The same table WITHOUT atomics, optimistically: position = the node's local index in the element.  Around a node of a mesh
// with regular topology (block / swept / mapped meshes) every element sees the node at a different local index, so the plain
// stores never collide; on any other mesh two elements may claim one position and the later store wins.  k_sym_tile counts the
// entries it finds, the host compares the total with nactive * nne after the build's round trip: a lost entry shows there, the
// mesh is remembered as colliding and the build is redone with k_adj_table.  The planes are pre-filled with the EMPTY marker.
// One thread per ELEMENT (the first version had one per (element, node): 134 instructions per entry, issue-bound at 0.68 ms on the
// 256^3 block): the row is loaded once, the duplicate test is NNE (NNE - 1) / 2 register compares, and the lanes of a warp --
// consecutive slots, i.e. mostly consecutive nodes -- write runs of one plane.
template <int NNE, int MAXDEG>
__global__ void __launch_bounds__(256) k_adj_place(const TileParams P, uint32_t *__restrict__ tab, int *flags) {
  static_assert(NNE <= MAXDEG, "one plane per local index");
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P.nactive) return;
  const int64_t e = P.elem_list ? (int64_t)P.elem_list[slot] : slot;
  int m[NNE];
  load_conn_row<NNE>(P.conn + e * NNE, m);
  bool dup = false;
#pragma unroll
  for (int a = 1; a < NNE; a++)
#pragma unroll
    for (int b = 0; b < a; b++) dup = dup || (m[a] == m[b]);
  if (dup) flags[0] = 1;
  const int nwp = (int)P.nwp, lo = (int)P.lo;  // plane indices fit 32 bits (MAXDEG * nwp <= 2^30)
#pragma unroll
  for (int lc = 0; lc < NNE; lc++) tab[lc * nwp + (m[lc] - lo)] = ((uint32_t)slot << 5) | (uint32_t)lc;
}

// prefix arrays outside the window: `before` ahead of it, the window's last value behind it
__global__ void k_tile_fill_outside(int64_t *__restrict__ a, int64_t len, int64_t lo, int64_t hi, int64_t before) {
  const int64_t nout = len - (hi - lo + 1);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nout) return;
  const int64_t idx = (i < lo) ? i : i + (hi - lo + 1);
  a[idx] = (i < lo) ? before : a[hi];
}
// colptr outside the window's dof range [dlo, dlo + nw * ndn], dlo = dof[0][lo] read on the device (no host round trip)
__global__ void k_tile_fill_colptr(int64_t *__restrict__ colptr, int64_t ncols, const int32_t *__restrict__ dof, int64_t lo, int64_t nw, int ndn) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > ncols) return;
  const int64_t dlo = dof[lo], dhi = min(dlo + nw * ndn, ncols);
  if (j < dlo) colptr[j] = 1;
  else if (j > dhi) colptr[j] = colptr[dhi];
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) { return *reinterpret_cast<const volatile unsigned long long *>(p); }
__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long *>(p) = v; }


// out[0] = total neighbour entries of the window, out[1] = largest neighbour count, out[2] = tile ticket, out[3] = adjacency
// entries found (= nactive * nne unless the optimistic placement lost one)
template <int NNE, int MAXDEG, int NDN, int PART, int TILE_T>
__global__ void __launch_bounds__(TILE_T, 512 / TILE_T)
    k_sym_tile(const TileParams P, int32_t *__restrict__ deg_out, uint32_t *__restrict__ adj_planes,
               typename CsWord<NNE>::type *__restrict__ cs_planes, int32_t *__restrict__ nnbr, int64_t *__restrict__ nbrptr,
               int64_t *__restrict__ colptr, int64_t *__restrict__ rowval, int32_t *__restrict__ nbr_out, unsigned long long *tile_state,
               unsigned long long *out) {
  constexpr int NKEY = NNE * MAXDEG;
  constexpr int CSB = NKEY + 4;  // bytes per thread of the staged slot row: NKEY/4 + 1 words, odd => conflict-free when the lanes write the same k
  static_assert(NKEY <= 64 && NKEY % 4 == 0 && (CSB / 4) % 2 == 1, "candidate keys of a node must fit 6 bits");
  using CsT = typename CsWord<NNE>::type;
  extern __shared__ uint32_t smem_u32[];
  uint32_t *U_sm = smem_u32;                                                  // [TILE_T * NKEY] unique neighbours of the tile's nodes, dense, node after node
  uint8_t *cs_sm = reinterpret_cast<uint8_t *>(smem_u32 + TILE_T * NKEY);   // [TILE_T][CSB]
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
  // plane indices fit 32 bits (MAXDEG * nwp <= 16 * 2^26): one integer multiply-add per access, no 64-bit address registers
  const int ii = (int)i, nwp = (int)P.nwp;
  const int64_t dof0 = P.dof[P.lo];  // requested here, needed after the look-back
  // ---- adjacency column (EMPTY = all ones where no element sits): sort by element slot (the order of the duplicate sum; EMPTY
  // entries go last), count, write it back in place
  uint32_t adj[MAXDEG];
#pragma unroll
  for (int j = 0; j < MAXDEG; j++) adj[j] = live ? adj_planes[j * nwp + ii] : ~0u;
  fesort::sort<MAXDEG>(adj);
  int deg = 0;
#pragma unroll
  for (int j = 0; j < MAXDEG; j++) deg += (adj[j] != ~0u) ? 1 : 0;
  if (live) deg_out[i] = deg;
#pragma unroll
  for (int j = 0; j < MAXDEG; j++)
    if (j < deg) adj_planes[j * nwp + ii] = adj[j];
  // ---- candidate keys: ((neighbour node + 1) << 6) | k, k = a * NNE + li; rows of other ranks and padding carry node field 0, so
  // they sort FIRST and are never a head: no validity test per key in the two scans below.
  // All loads first (element ids, then the connectivity rows: 2 x MAXDEG independent requests in flight per thread).
  const int32_t *__restrict__ conn = P.conn;
  const int32_t *__restrict__ elem_list = P.elem_list;
  uint32_t el[MAXDEG];
#pragma unroll
  for (int j = 0; j < MAXDEG; j++) {
    el[j] = adj[j] >> 5;
    if (elem_list && j < deg) el[j] = (uint32_t)__ldg(elem_list + el[j]);  // a partition whose active elements are not one contiguous range
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
      int mm = m[li];  // -1 where j >= deg
      if (PART == 1 && (mm < P.own_lo || mm >= P.own_hi)) mm = -1;
      if (PART == 2 && j < deg && !P.rowowned[mm]) mm = -1;
      keys[j * NNE + li] = ((uint32_t)(mm + 1) << TILE_KB) | (uint32_t)(j * NNE + li);
    }
  }
  fesort::sort<NKEY>(keys);
  // ---- unique neighbours of this node
  constexpr uint32_t KMASK = (1u << TILE_KB) - 1u;
  int nu = 0;
  {
    uint32_t prev = 0;  // node field 0: a dropped candidate
#pragma unroll
    for (int x = 0; x < NKEY; x++) {
      nu += ((keys[x] ^ prev) > KMASK) ? 1 : 0;  // the node field differs from the previous key's
      prev = keys[x];
    }
  }
  // ---- prefix of the counts inside the CTA; the tile's aggregate is published right away, the look-back over the earlier
  // tiles happens after the base-independent work below, when their prefixes have had time to arrive
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
  if (tid == 0) st_state(tile_state + tile, (tile == 0 ? ST_PREFIX : ST_AGG) | (unsigned long long)total);
  // ---- neighbour slot of every candidate (staged row, one byte each) and the unique list (dense, at the node's offset)
  {
    uint8_t *cs8 = cs_sm + tid * CSB;
    int slot = -1;  // dropped candidates come first and keep -1 = 0xff
    uint32_t prev = 0;
#pragma unroll
    for (int x = 0; x < NKEY; x++) {
      const bool head = (keys[x] ^ prev) > KMASK;
      slot += head ? 1 : 0;
      cs8[keys[x] & KMASK] = (uint8_t)slot;
      if (head) U_sm[excl + slot] = keys[x] >> TILE_KB;  // node + 1
      prev = keys[x];
    }
    // the thread's own row back from shared memory, one word per adjacent element, straight into the planes (coalesced)
    const uint32_t *row = reinterpret_cast<const uint32_t *>(cs8);
#pragma unroll
    for (int j = 0; j < MAXDEG; j++) {
      if (j < deg) {
        CsT w;
        if constexpr (NNE == 8) w = (unsigned long long)row[2 * j] | ((unsigned long long)row[2 * j + 1] << 32);
        else if constexpr (NNE == 4) w = row[j];
        else w = (uint32_t)cs8[3 * j] | ((uint32_t)cs8[3 * j + 1] << 8) | ((uint32_t)cs8[3 * j + 2] << 16) | 0xff000000u;
        cs_planes[j * nwp + ii] = w;
      }
    }
  }
  if (live) nnbr[n] = nu;
  {
    int mx = nu, dsum = deg;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      dsum += __shfl_xor_sync(0xffffffffu, dsum, d);
    }
    if (lane == 0 && mx > 0) atomicMax(out + 1, (unsigned long long)mx);
    if (lane == 0 && dsum > 0) atomicAdd(out + 3, (unsigned long long)dsum);
  }
  // ---- decoupled look-back (warp 0): sum the aggregates of the preceding tiles down to the first published prefix
  if (warp == 0) {
    long long run = 0;
    if (tile > 0) {
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
  __syncthreads();
  const long long base = s_base;
  if (live) {
    const long long nb = base + excl;
    nbrptr[n] = nb;
    const int64_t c0 = dof0 + i * NDN;
#pragma unroll
    for (int q = 0; q < NDN; q++)
      if (c0 + q <= P.ncols) colptr[c0 + q] = 1 + (nb * NDN + (long long)q * nu) * NDN;  // the guard only matters when the map is not affine
    if (i == P.nw - 1) {
      nbrptr[n + 1] = nb + nu;
      if (c0 + NDN <= P.ncols) colptr[c0 + NDN] = 1 + (nb + nu) * (long long)(NDN * NDN);
      out[0] = (unsigned long long)(nb + nu);
    }
  }
  // ---- outputs of the tile: contiguous in memory
  if (nbr_out)
    for (int idx = tid; idx < total; idx += TILE_T) nbr_out[base + idx] = (int32_t)U_sm[idx] - 1;
  if (NDN == 1) {
    const int64_t shift = dof0 - P.lo;  // U holds node + 1 and rowval is 1-based
    for (int idx = tid; idx < total; idx += TILE_T) rowval[base + idx] = shift + (int64_t)U_sm[idx];
  } else {
    for (int t = 0; t < 32; t++) {
      const int nu_t = __shfl_sync(0xffffffffu, nu, t);
      const int ex_t = __shfl_sync(0xffffffffu, excl, t);
      if (nu_t == 0) continue;
      const long long rb = (base + ex_t) * (long long)(NDN * NDN);
      const int per_col = nu_t * NDN;
      for (int r = lane; r < per_col; r += 32) {
        const int s = r / NDN, p = r - s * NDN;
        const int64_t rd = dof0 + ((int64_t)U_sm[ex_t + s] - 1 - P.lo) * NDN + p + 1;
#pragma unroll
        for (int q = 0; q < NDN; q++) rowval[rb + (long long)q * per_col + r] = rd;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ numeric phase
template <int NNE>
struct TileGatherParams {
  int64_t lo, nw, nwp, nnodes;
  const int32_t *deg;
  const uint32_t *adj;
  const typename CsWord<NNE>::type *cs;
  const int32_t *nnbr;
  const int64_t *colptr;
  const int32_t *dof;
  const double *V;
  int64_t vstride;  // PLANES: doubles between the planes of two consecutive value indices
  double *nzval;
  int acc_doubles;  // size of the accumulator image in shared memory; T * NDN dummy cells follow it
};

// One thread per (node, column component q).  Measured alternatives for vector fields (config 2, profiles/r02_gather_c2_variants.txt):
// a thread per (node, q, p) triples the warps in flight but also the per-candidate bookkeeping -- 2.4 G warp instructions, 61 %
// issue-active, 3.3 ms -- against 3.1 ms for this mapping, whose three adds per candidate share one slot lookup and one address.
template <int NDN>
struct GatherShape {
  static constexpr int T = (NDN == 3) ? 96 : 128;  // threads per CTA: a multiple of NDN and of 32
  static constexpr int NPB = T / NDN;              // nodes per CTA
};

// Value load of the numeric kernel: volatile with a memory clobber, so the compiler keeps the batch of loads of one element together
// and ahead of the shared-memory adds of the previous element (left to itself it sinks every load next to its use: one or two
// loads in flight per thread).
__device__ __forceinline__ double ld_value(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Position of value (block blk, entry e) of the element in slot `slot`:
//   element-major records     slot * VPE + ND2 * blk + e          (vector fields: the NDN lanes of a node read one 72-byte block)
//   planes (FormArgs::planes) (blk * ND2 + e) * vstride + slot     (value planes: consecutive nodes read consecutive words)
// The accumulator image in shared memory is exactly the CTA's slice of nzval (column after column, rows in order), so the
// write-out is a flat copy; lanes = (node, q) pairs hit distinct banks for a given (slot, p) because the column stride nu * NDN
// is odd for the 27-neighbour interior stencil.
// MODE (how the value loads of element j + 1 overlap the shared-memory adds of element j; measured in profiles/r02_gather_modes.txt):
//   0  loads of j + 1 issued, then the adds of j.  ptxas gives every load of this branchy body ONE scoreboard (profiles/
//      r02_scoreboards.txt), so the first add of j also waits for the loads of j + 1 that were just issued: nothing overlaps.
//      The 16-plane kernels (Q4, T3, T4: few of the 16 planes are occupied) keep this form.
//   1  straight-line body (no branch: missing elements re-read element 0, rows of other ranks go to a dummy cell) with warp-level
//      fences between the load batches and the add batches -- the default for H8
//   4  straight-line body without fences (ptxas interleaves single loads with the adds, rotating over four scoreboards)
// (Modes 2, 3, 5, 6 of the measurements -- first add before the next loads, L2 prefetches two elements ahead -- were removed after
// they lost; the code is in the history of this file.)
template <int NNE, int MAXDEG, int NDN, bool COMPACT, bool PLANES, int MODE>
__global__ void __launch_bounds__(GatherShape<NDN>::T) k_gather_tile(const TileGatherParams<NNE> G) {
  extern __shared__ double acc[];
  constexpr int T = GatherShape<NDN>::T, NPB = GatherShape<NDN>::NPB, ND2 = NDN * NDN;
  constexpr int EM = NNE * NDN;
  constexpr int64_t VPE = COMPACT ? (int64_t)(NNE * (NNE + 1) / 2) * ND2 : (int64_t)EM * EM;
  constexpr bool STRAIGHT = (MODE == 1 || MODE == 4);
  using CsT = typename CsWord<NNE>::type;
  const int tid = threadIdx.x;
  // element-major records: the NDN threads of a node are neighbours (they read one contiguous block); planes: component-major, the
  // threads of a warp are consecutive nodes with the same q, so every load is one contiguous run of a value plane
  const int q = PLANES ? tid / NPB : tid % NDN;
  const int ln = PLANES ? tid - q * NPB : tid / NDN;
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
  if (live) {
    nu = G.nnbr[n];
    deg = min(G.deg[i], MAXDEG);
    off = (int)(G.colptr[dof0 + i * NDN + q] - 1 - cb0);
  }
  __syncthreads();
  const unsigned wmask = __ballot_sync(0xffffffffu, nu > 0);
  if (nu > 0) {
    double *col = acc + off;
    const int ii = (int)i, nwp = (int)G.nwp;  // plane indices fit 32 bits
    const uint32_t *__restrict__ adjp = G.adj;
    const CsT *__restrict__ csp = G.cs;
    const double *__restrict__ V = G.V;
    // metadata of every adjacent element first (2 x MAXDEG independent, coalesced loads in flight).  Rows of other ranks (slot 0xff)
    // are loaded as well and dropped at the add.
    double *dummy = acc + G.acc_doubles + tid * NDN;
    uint32_t ad[MAXDEG];
    CsT cs[MAXDEG];
#pragma unroll
    for (int j = 0; j < MAXDEG; j++) {
      if (STRAIGHT) {
        const int jj = j < deg ? j : 0;
        ad[j] = __ldcs(adjp + (jj * nwp + ii));
        cs[j] = __ldcs(csp + (jj * nwp + ii));
        if (j >= deg) cs[j] = ~(CsT)0;
      } else {
        ad[j] = 0;
        cs[j] = ~(CsT)0;
        if (j < deg) {
          ad[j] = __ldcs(adjp + (j * nwp + ii));
          cs[j] = __ldcs(csp + (j * nwp + ii));
        }
      }
    }
    // address of value (row node li, row component p) of the element in adjacency position j
    auto value_ptr = [&](int j, int li, int p) -> const double * {
      const int64_t slot = ad[j] >> 5;
      const int lc = (int)(ad[j] & 31u);
      if (COMPACT) {
        // block (min, max) of the upper block triangle, entry (comp of min, comp of max) at comp_max * NDN + comp_min
        const bool tr = li > lc;
        const int blk = tr ? li * (li + 1) / 2 + lc : lc * (lc + 1) / 2 + li;
        const int e = tr ? q + p * NDN : q * NDN + p;  // transposed block: the row component strides by NDN
        return PLANES ? V + ((int64_t)blk * ND2 + e) * G.vstride + slot : V + slot * VPE + ND2 * blk + e;
      }
      // full matrix in emission order: column (lc, q), rows (li, p); planes: value index (lc * NNE + li) * ND2 + q * NDN + p
      return PLANES ? V + ((int64_t)(lc * NNE + li) * ND2 + q * NDN + p) * G.vstride + slot : V + slot * VPE + (lc * NDN + q) * EM + li * NDN + p;
    };
    auto load_vals = [&](int j, double (&v)[NNE][NDN]) {
#pragma unroll
      for (int li = 0; li < NNE; li++)
#pragma unroll
        for (int p = 0; p < NDN; p++) v[li][p] = (MODE == 1) ? ld_value(value_ptr(j, li, p)) : *value_ptr(j, li, p);  // plain ld.global: the .nc path measured 0.7 ms slower
    };
    auto add_vals = [&](int j, const double (&v)[NNE][NDN], int li0, int li1) {
#pragma unroll
      for (int li = 0; li < NNE; li++) {
        if (li < li0 || li >= li1) continue;
        const unsigned s = (unsigned)((cs[j] >> (8 * li)) & 0xffu);
        if (STRAIGHT) {
          double *dst = (s != 0xffu) ? col + s * NDN : dummy;
#pragma unroll
          for (int p = 0; p < NDN; p++) dst[p] += v[li][p];
        } else if (s != 0xffu) {
          double *dst = col + s * NDN;
#pragma unroll
          for (int p = 0; p < NDN; p++) dst[p] += v[li][p];
        }
      }
    };
    double va[NNE][NDN], vb[NNE][NDN];
    load_vals(0, va);  // deg >= 1 here (nu > 0)
#pragma unroll
    for (int j = 0; j < MAXDEG; j += 2) {
      if (MODE == 1) {
        load_vals(j + 1, vb);
        __syncwarp(wmask);
        add_vals(j, va, 0, NNE);
        __syncwarp(wmask);
        if (j + 2 < MAXDEG) load_vals(j + 2, va);
        __syncwarp(wmask);
        add_vals(j + 1, vb, 0, NNE);
        __syncwarp(wmask);
      } else if (STRAIGHT) {
        load_vals(j + 1, vb);
        add_vals(j, va, 0, NNE);
        if (j + 2 < MAXDEG) load_vals(j + 2, va);
        add_vals(j + 1, vb, 0, NNE);
      } else {
        if (j + 1 < deg) load_vals(j + 1, vb);
        if (j < deg) add_vals(j, va, 0, NNE);
        if (j + 2 < deg && j + 2 < MAXDEG) load_vals(j + 2, va);
        if (j + 1 < deg) add_vals(j + 1, vb, 0, NNE);
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < total; idx += T) __stcs(G.nzval + cb0 + idx, acc[idx]);
}

// Vector assembly on a thread-per-node pattern (cf. k_vec_gather): thread per (window node, component)
__global__ void k_vec_gather_tile(int64_t lo, int64_t nw, int64_t nwp, int64_t nnodes, int nne, int ndn, int maxdeg, const int32_t *__restrict__ deg,
                                  const uint32_t *__restrict__ adj, const int32_t *__restrict__ dof, const uint8_t *__restrict__ rowowned,
                                  const double *__restrict__ elvec, double *__restrict__ F) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nw * ndn) return;
  const int64_t i = t / ndn;
  const int p = (int)(t - i * ndn);
  const int64_t n = lo + i;
  if (rowowned && !rowowned[n]) return;
  const int d = min(deg[i], maxdeg);
  if (d == 0) return;
  const int EM = nne * ndn;
  double acc = 0.0;
  for (int j = 0; j < d; j++) {
    const uint32_t a = adj[(int64_t)j * nwp + i];
    acc += elvec[(int64_t)(a >> 5) * EM + (a & 31u) * ndn + p];
  }
  F[dof[(int64_t)p * nnodes + n]] = acc;
}

template <typename T>
int32_t talloc(fegpu_ctx *ctx, T **p, size_t n) {
  *p = nullptr;
  return fe_dev_alloc(ctx, (void **)p, sizeof(T) * std::max<size_t>(n, 1), ctx->stream);
}

int tile_maxdeg_for(int nne) { return nne == 8 ? 8 : ((nne == 4 || nne == 3) ? 16 : 0); }

}  // namespace

bool fe_tile_candidate(const fegpu_dofmap *dm) {
  static const bool tile_off = std::getenv("FEGPU_TILE") && std::atoi(std::getenv("FEGPU_TILE")) == 0;  // A/B knob
  const fegpu_mesh *mesh = dm->mesh;
  if (tile_off || tile_maxdeg_for(mesh->nne) == 0 || dm->ndn > 3) return false;
  if (mesh->win_hi - mesh->win_lo <= 0 || mesh->nactive <= 0) return false;
  if (dm->tile_failed_version == mesh->topo_version) return false;  // this mesh / numbering already failed the preconditions
  if (mesh->nnodes > (int64_t)TILE_DROPPED - 1 || mesh->nactive >= ((int64_t)1 << 27)) return false;
  return true;
}

// Thread-per-node symbolic phase.  *taken = false: the preconditions do not hold, nothing was built, the caller runs the
// general path.  *taken = true with dm->pat == nullptr: degenerate elements (the caller takes the sort path).
int32_t fe_tile_build(fegpu_dofmap *dm, const std::function<int32_t(bool)> *fork, bool *taken) {
  *taken = false;
  if (!fe_tile_candidate(dm)) return FEGPU_OK;
  fegpu_ctx *ctx = dm->ctx;
  fegpu_mesh *mesh = dm->mesh;
  cudaStream_t st = ctx->stream;
  const int nne = mesh->nne, ndn = dm->ndn;
  const int MD = tile_maxdeg_for(nne);
  const int64_t nn = mesh->nnodes;
  const int64_t lo = mesh->win_lo, hi = mesh->win_hi, nw = hi - lo;
  const int64_t nwp = (nw + 31) & ~(int64_t)31;
  const int64_t nadj = mesh->nactive * nne;

  int *d_flags = nullptr;  // [0] degenerate, [1] dof map not affine, [2] a node with more than MD elements
  unsigned long long *d_state = nullptr, *d_out = nullptr;
  Pattern *P = nullptr;
  auto cleanup = [&]() {
    void *ptrs[] = {d_flags, d_state, d_out};
    for (void *q : ptrs)
      if (q) fe_dev_free(ctx, q, st);
    d_flags = nullptr; d_state = nullptr; d_out = nullptr;
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
  if (dm->pat) { fe_pattern_free(dm->pat); dm->pat = nullptr; }
  P = new Pattern();
  dm->pat = P;
  P->ctx = ctx;
  P->stream = st;
  P->alloc_stream = st;
  P->ncols = dm->col_nall;
  P->nrows = dm->row_nall;
  // H8: 256 nodes per tile (two CTAs of eight warps per SM: the warps of a CTA run in step, so each scheduler's instruction cache
  // sees two code positions instead of four; 2.78 against 2.85 ms on config 4); the 16-plane kernels keep 128.  FEGPU_TILE_T = A/B knob
  static const int tile_env = std::getenv("FEGPU_TILE_T") ? std::atoi(std::getenv("FEGPU_TILE_T")) : 0;
  // (Two lanes per node -- half the registers, twice the warps -- was built and measured: a tie at best, profiles/r02_sym_pair.txt.)
  const int tile_t = (tile_env == 64 || tile_env == 128) ? tile_env : (nne == 8 ? 256 : 128);
  const int64_t ntiles = (nw + tile_t - 1) / tile_t;
  const size_t nb_cap = (size_t)nadj * nne;  // upper bound of the neighbour entries: every candidate unique
  const size_t cs_bytes = (nne == 8 ? sizeof(unsigned long long) : sizeof(uint32_t)) * (size_t)MD * nwp;
  PT(talloc(ctx, &P->t_deg, (size_t)nw));
  PT(talloc(ctx, &P->t_adj, (size_t)MD * nwp));
  PT(talloc(ctx, &d_flags, 4));
  PT(talloc(ctx, &P->d_nnbr, (size_t)nn));
  PT(talloc(ctx, &P->d_nbrptr, (size_t)nn + 1));
  PT(talloc(ctx, &P->d_colptr, (size_t)P->ncols + 1));
  PT(fe_dev_alloc(ctx, (void **)&P->t_cs, cs_bytes, st));
  // rowval is sized before its length is known (single pass): the bound is what the reference's COO would hold per column
  // node, nnz is typically 0.42 of it (H8)
  PT(talloc(ctx, &P->d_rowval, nb_cap * ndn * ndn));
  if (ndn >= 2) PT(talloc(ctx, &P->d_nbr, nb_cap));
  PT(talloc(ctx, &d_state, (size_t)ntiles));
  PT(talloc(ctx, &d_out, 4));
  // adjacency table: without atomics when the mesh has not shown local-index collisions (see k_adj_place); the planes start EMPTY
  static const bool place_off = std::getenv("FEGPU_ADJ_PLACE") && std::atoi(std::getenv("FEGPU_ADJ_PLACE")) == 0;  // A/B knob
  // (triangles: six elements around an interior node and three local indices -- they always collide, so they start with the atomics)
  const bool optimistic = !place_off && nne >= 4 && nne <= MD && mesh->adj_collide_version != mesh->topo_version;
  PC(cudaMemsetAsync(P->t_adj, 0xff, sizeof(uint32_t) * (size_t)MD * nwp, st));
  if (!optimistic) PC(cudaMemsetAsync(P->t_deg, 0, sizeof(int32_t) * nw, st));
  PC(cudaMemsetAsync(d_flags, 0, sizeof(int) * 4, st));
  PC(cudaMemsetAsync(d_state, 0, sizeof(unsigned long long) * ntiles, st));
  PC(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long) * 4, st));
  TileParams TP{mesh->conn_act(), mesh->d_elem_list, mesh->nactive, nn, lo, nw, nwp, (int32_t)mesh->own_lo, (int32_t)mesh->own_hi, mesh->d_rowowned, dm->d_dof, P->ncols};
  k_dof_affine<<<(unsigned)std::min<int64_t>(grid_for(nw, 256), (int64_t)ctx->sm_count * 8), 256, 0, st>>>(dm->d_dof, nn, ndn, lo, nw, d_flags + 1);
  if (optimistic) {
    switch (nne) {
      case 8: k_adj_place<8, 8><<<grid_for(mesh->nactive, 256), 256, 0, st>>>(TP, P->t_adj, d_flags); break;
      default: k_adj_place<4, 16><<<grid_for(mesh->nactive, 256), 256, 0, st>>>(TP, P->t_adj, d_flags); break;
    }
  } else {
    switch (nne) {
      case 8: k_adj_table<8, 8><<<grid_for(nadj, 256), 256, 0, st>>>(TP, P->t_deg, P->t_adj, d_flags); break;
      case 4: k_adj_table<4, 16><<<grid_for(nadj, 256), 256, 0, st>>>(TP, P->t_deg, P->t_adj, d_flags); break;
      default: k_adj_table<3, 16><<<grid_for(nadj, 256), 256, 0, st>>>(TP, P->t_deg, P->t_adj, d_flags); break;
    }
  }
  ctx->launches += 2;
  fe_mark(ctx, optimistic ? "sym:k_adj_place" : "sym:k_adj_table");
  // optimistic: the element integration may start now (plane layout); should a precondition turn out violated, the caller
  // integrates again in the layout the general path needs
  if (fork) PT((*fork)(true));
  const int part = !mesh->d_rowowned ? 0 : (mesh->own_contig ? 1 : 2);
  // (Co-residency with the integration kernel was measured twice, profiles/r02_coresidency.txt: capping this kernel at 2 or 3 CTAs per
  // SM so that a CTA of k_h8_diffusion fits beside them makes the fresh step slower, 10.1 / 9.3 against 8.9 ms.)
  const size_t smem = sizeof(uint32_t) * (size_t)tile_t * (nne * MD) + (size_t)tile_t * (nne * MD + 4);
#define SYM_LAUNCH_T(NNE_, MD_, NDN_, PART_, TT_)                                                                                   \
  do {                                                                                                                              \
    PC(cudaFuncSetAttribute(k_sym_tile<NNE_, MD_, NDN_, PART_, TT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    k_sym_tile<NNE_, MD_, NDN_, PART_, TT_><<<(unsigned)ntiles, TT_, smem, st>>>(TP, P->t_deg, P->t_adj,                            \
        reinterpret_cast<CsWord<NNE_>::type *>(P->t_cs), P->d_nnbr, P->d_nbrptr, P->d_colptr, P->d_rowval, P->d_nbr, d_state, d_out); \
  } while (0)
#define SYM_LAUNCH(NNE_, MD_, NDN_, PART_)                        \
  do {                                                            \
    if (tile_t == 128) SYM_LAUNCH_T(NNE_, MD_, NDN_, PART_, 128); \
    else if (NNE_ == 8 && tile_t == 256) SYM_LAUNCH_T(NNE_, MD_, NDN_, PART_, (NNE_ == 8 ? 256 : 128)); \
    else SYM_LAUNCH_T(NNE_, MD_, NDN_, PART_, 64);                \
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
#undef SYM_LAUNCH_T
  ctx->launches++;
  PC(cudaGetLastError());
  fe_mark(ctx, "sym:k_sym_tile");
  if (nw < nn && P->d_nbr) {
    // nbrptr stays a complete prefix array for the one consumer that walks every node: the result transport of vector fields
    // (neighbour lists instead of rowval on the link).  Scalar fields never read it (or nnbr) outside the window.
    k_tile_fill_outside<<<grid_for(nn - nw, 256), 256, 0, st>>>(P->d_nbrptr, nn + 1, lo, hi, 0);
    ctx->launches++;
  }
  if (nw * ndn < P->ncols) {
    k_tile_fill_colptr<<<grid_for(P->ncols + 1, 256), 256, 0, st>>>(P->d_colptr, P->ncols, dm->d_dof, lo, nw, ndn);
    ctx->launches++;
  }
  // the ONE host round trip of the build: preconditions + sizes
  int h_flags[4] = {0, 0, 0, 0};
  unsigned long long h_out[4] = {0, 0, 0, 0};
  PC(cudaMemcpyAsync(h_flags, d_flags, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
  PC(cudaMemcpyAsync(h_out, d_out, sizeof(h_out), cudaMemcpyDeviceToHost, st));
  PC(cudaStreamSynchronize(st));
  FE_TRACE("tile build: sync done");
  if (h_flags[0]) {  // an element lists a node twice: no structured path at all (same as the general build's bail)
    mesh->degenerate = true;
    cleanup();
    drop_pattern();
    *taken = true;
    return FEGPU_OK;
  }
  if (optimistic && !h_flags[1] && (int64_t)h_out[3] != nadj) {
    // two elements saw a node at the same local index: an adjacency entry was overwritten.  Remember it for this mesh and
    // build again with the atomic table (the integration that is already queued is not repeated: same layout).
    mesh->adj_collide_version = mesh->topo_version;
    cleanup();
    drop_pattern();
    return fe_tile_build(dm, fork, taken);
  }
  if (h_flags[1] || h_flags[2]) {
    dm->tile_failed_version = mesh->topo_version;
    cleanup();
    drop_pattern();
    return FEGPU_OK;  // *taken stays false: general path
  }
  *taken = true;
  P->maxdeg = MD;
  P->maxcand = MD * nne;
  P->total_nbr = (int64_t)h_out[0];
  P->nnz = P->total_nbr * ndn * ndn;
  P->maxnbr = std::max((int)h_out[1], 1);
  P->d_dof = dm->d_dof;
  P->ndn = ndn;
  P->nnodes = nn;
  P->tile = true;
  P->tile_lo = lo;
  P->tile_nw = nw;
  P->tile_nwp = nwp;
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

// Numeric phase on a pattern built by fe_tile_build.
int32_t fe_tile_gather(fegpu_dofmap *dm, const double *d_V, bool compact, bool planes, int64_t vstride, double *d_nzval) {
  fegpu_ctx *ctx = dm->ctx;
  Pattern *P = dm->pat;
  fegpu_mesh *mesh = dm->mesh;
  if (!P || !P->tile) return fegpu_fail(ctx, FEGPU_ERR_STATE, "internal: not a thread-per-node pattern");
  const int nne = mesh->nne, ndn = dm->ndn;
  const int npb = ((ndn == 3) ? 96 : 128) / ndn;
  const int acc_doubles = (int)((size_t)npb * P->maxnbr * ndn * ndn);
  const size_t smem = sizeof(double) * ((size_t)acc_doubles + (size_t)npb * ndn * ndn);  // + one dummy cell of ndn doubles per thread
  if (smem > 200 * 1024) return fegpu_fail(ctx, FEGPU_ERR_STATE, "internal: gather accumulators exceed shared memory");
  if (P->nnz == 0) return FEGPU_OK;
  const unsigned grid = (unsigned)((P->tile_nw + npb - 1) / npb);
  // shared-memory carve-out in percent of the SM's 228 KB (-1: the driver's choice).  What is not shared memory is L1, and the L1
  // holds the lines of the loads in flight: fewer resident CTAs with a larger L1 can move more bytes (profiles/r02_gather_modes.txt)
  // Measured (call R): 85 % (194 KB shared + 62 KB L1) is the best point for both field kinds; at 100 % (which the driver picks when
  // it maximises resident CTAs) the scalar gather takes 2.8 instead of 1.9 ms, the elasticity gather 3.5 instead of 2.5 ms.
  static const int carveout_env = std::getenv("FEGPU_GATHER_CARVEOUT") ? std::atoi(std::getenv("FEGPU_GATHER_CARVEOUT")) : 85;
  const int carveout = (smem + 1024 > (size_t)carveout_env * 228 * 1024 / 100) ? -1 : carveout_env;  // one CTA must still fit
  static const int gmode_env = std::getenv("FEGPU_GATHER_MODE") ? std::atoi(std::getenv("FEGPU_GATHER_MODE")) : 1;  // A/B knob
  const int gmode = (nne == 8 && (gmode_env == 0 || gmode_env == 1 || gmode_env == 4)) ? gmode_env : 0;  // the variants exist for H8 only
#define G_LAUNCH_M(NNE_, MD_, NDN_, C_, PL_, M_)                                                                                    \
  do {                                                                                                                              \
    TileGatherParams<NNE_> G{P->tile_lo, P->tile_nw, P->tile_nwp, mesh->nnodes, P->t_deg, P->t_adj,                                 \
                             reinterpret_cast<const CsWord<NNE_>::type *>(P->t_cs), P->d_nnbr, P->d_colptr, dm->d_dof, d_V, vstride, d_nzval, acc_doubles}; \
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_gather_tile<NNE_, MD_, NDN_, C_, PL_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    if (carveout >= 0) CUDA_TRY(ctx, cudaFuncSetAttribute(k_gather_tile<NNE_, MD_, NDN_, C_, PL_, M_>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout)); \
    k_gather_tile<NNE_, MD_, NDN_, C_, PL_, M_><<<grid, GatherShape<NDN_>::T, smem, ctx->stream>>>(G);                              \
  } while (0)
#define G_LAUNCH(NNE_, MD_, NDN_, C_, PL_)                         \
  do {                                                             \
    if constexpr (NNE_ == 8) {                                     \
      switch (gmode) {                                             \
        case 1: G_LAUNCH_M(NNE_, MD_, NDN_, C_, PL_, 1); break;    \
        case 4: G_LAUNCH_M(NNE_, MD_, NDN_, C_, PL_, 4); break;    \
        default: G_LAUNCH_M(NNE_, MD_, NDN_, C_, PL_, 0); break;   \
      }                                                            \
    } else {                                                       \
      G_LAUNCH_M(NNE_, MD_, NDN_, C_, PL_, 0);                     \
    }                                                              \
  } while (0)
#define G_PL(NNE_, MD_, NDN_, C_)                      \
  do {                                                 \
    if (planes) G_LAUNCH(NNE_, MD_, NDN_, C_, true);   \
    else G_LAUNCH(NNE_, MD_, NDN_, C_, false);         \
  } while (0)
#define G_C(NNE_, MD_, NDN_)                   \
  do {                                         \
    if (compact) G_PL(NNE_, MD_, NDN_, true);  \
    else G_PL(NNE_, MD_, NDN_, false);         \
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
#undef G_PL
#undef G_LAUNCH
#undef G_LAUNCH_M
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

int32_t fe_tile_vec_gather(fegpu_dofmap *dm, const double *d_elvec, double *d_F) {
  fegpu_ctx *ctx = dm->ctx;
  Pattern *P = dm->pat;
  fegpu_mesh *mesh = dm->mesh;
  const int64_t n = P->tile_nw * dm->ndn;
  if (n == 0) return FEGPU_OK;
  k_vec_gather_tile<<<grid_for(n, 256), 256, 0, ctx->stream>>>(P->tile_lo, P->tile_nw, P->tile_nwp, mesh->nnodes, mesh->nne, dm->ndn, P->tile_md, P->t_deg,
                                                              P->t_adj, dm->d_dof, mesh->d_rowowned, d_elvec, d_F);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}
