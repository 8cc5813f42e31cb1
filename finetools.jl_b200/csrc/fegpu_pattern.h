// The cached sparsity pattern of one (mesh, dof map, partition): what the symbolic phase builds and the numeric phase,
// the vector gather and the result transport read.  Shared by fegpu_pattern.cu (group / warp kernels, every element type)
// and fegpu_tile.cu (thread-per-node kernels for small stencils).
#pragma once
#include "fegpu_internal.h"

struct Pattern {
  fegpu_ctx *ctx = nullptr;
  int64_t nnz = 0, ncols = 0, nrows = 0;
  int64_t *d_colptr = nullptr;    // [ncols+1] 1-based
  int64_t *d_rowval = nullptr;    // [nnz] 1-based
  int64_t *d_adjptr = nullptr;    // [nnodes+1]
  int32_t *d_adj_slot = nullptr;  // active-element slot
  uint8_t *d_adj_lc = nullptr;    // local node index of this node in that element
  int32_t *d_nnbr = nullptr;      // [nnodes]
  int64_t *d_nbrptr = nullptr;    // [nnodes+1]
  uint16_t *d_cslot = nullptr;    // per node at adjptr[n]*nne + a*nne + li: neighbour slot of that candidate (0xffff = dropped)
  uint16_t *d_rank = nullptr;     // per node nnbr*ndn entries at nbrptr[n]*ndn, nullptr when identity everywhere
  int32_t *d_order = nullptr;     // node visiting order of the gather: the active nodes in Morton order of their coordinates
  int64_t norder = 0;             // (nullptr = all nodes, natural order)
  // compressed form of rowval for the result transport (vector fields whose node-major dof order is ascending everywhere):
  // the rows of every column of node n are { dof[p][nbr[nbrptr[n] + s]] + 1 : s ascending, p ascending }
  int32_t *d_nbr = nullptr;       // [total_nbr] neighbour nodes, ascending per node
  int64_t total_nbr = 0;
  const int32_t *d_dof = nullptr; // borrowed from the dof map that owns this pattern
  int ndn = 0;
  int64_t nnodes = 0;
  int maxdeg = 0, maxcand = 0, maxnbr = 0;
  cudaStream_t stream = 0;        // consumer stream (the numeric phase and the transport read the arrays here)
  cudaStream_t alloc_stream = 0;  // stream the arrays were allocated on (the build's); they are freed on it, see fe_pattern_free
  int refs = 1;                   // owners: the dof map + every assembler result that borrows colptr / rowval
  // built by the thread-per-node kernels (fegpu_tile.cu): node window, dof map affine on it, adjacency capacity per node
  // -- in plane (struct-of-arrays) form instead of the CSR adjacency / cslot arrays above, which stay nullptr:
  bool tile = false;
  int64_t tile_lo = 0, tile_nw = 0, tile_nwp = 0;  // node window [lo, lo + nw), plane stride nwp
  int tile_md = 0;                 // adjacency capacity per node (planes)
  int32_t *t_deg = nullptr;        // [nw] elements at window node i
  uint32_t *t_adj = nullptr;       // [md][nwp] (slot << 5) | local index, ascending slot per node
  void *t_cs = nullptr;            // [md][nwp] one word per (node, adjacent element): the NNE neighbour slots, a byte each (0xff = dropped)
  cudaEvent_t ready = nullptr;    // recorded when the build's last kernel is queued: the result transport may ship the pattern's
                                  // arrays while the integration and the numeric phase of the same call are still running
};
