// bilform_dot (FEMMBaseModule.jl:1335-1366) for scalar fields (ndn = 1) on the elements with up to 10 nodes (T3, Q4, T4, H8, T10):
// one thread per element.  Coordinates, the Jacobian and the upper triangle of the element matrix live in registers; the
// basis-function tables sit in shared memory.  factor = ((N_k N_m) Jac) w is symmetric in (k, m) bit for bit (IEEE
// multiplication commutes), so the mirrored entry equals what the reference's full double loop computes.
// The entry-per-thread generic kernel recomputed the Jacobian in all 32 threads of an element for 4 useful entries each
// (T10 mass, config 3: 7.8 ms); here it is computed once.
// Output: compact upper triangle (mesh-structured path) or the full matrix in emission order.
#include "fegpu_internal.h"

namespace {

struct DotParams {
  const int32_t *conn;
  const double *xyz;
  int64_t nnodes;
  const int32_t *elem_list;
  int64_t nactive;
  const double *tab;  // N [npts][NNE], then dN [npts][MDIM][NNE]
  const double *w;
  int npts;
  double *V;
  int compact;
  int64_t vstride;  // > 0: plane layout (FormArgs::planes)
  double c;         // the 1 x 1 coefficient
  int m;            // manifold dimension kwarg
  double otherdim;
};

template <int NNE, int MDIM, int SDIM>
__global__ void __launch_bounds__(128) k_dot_scalar(const DotParams P) {
  constexpr int NT = NNE * (NNE + 1) / 2;
  extern __shared__ double stab[];  // N | dN | w
  double *sN = stab, *sdN = sN + P.npts * NNE, *sw = sdN + P.npts * NNE * MDIM;
  for (int i = threadIdx.x; i < P.npts * NNE * (1 + MDIM); i += blockDim.x) stab[i] = P.tab[i];
  for (int i = threadIdx.x; i < P.npts; i += blockDim.x) sw[i] = P.w[i];
  __syncthreads();
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= P.nactive) return;
  const int64_t e = P.elem_list ? P.elem_list[slot] : slot;
  double X[NNE][SDIM];
#pragma unroll
  for (int a = 0; a < NNE; a++) {
    const int nd = P.conn[e * NNE + a];
#pragma unroll
    for (int s = 0; s < SDIM; s++) X[a][s] = P.xyz[(int64_t)s * P.nnodes + nd];
  }
  double acc[NT];
#pragma unroll
  for (int i = 0; i < NT; i++) acc[i] = 0.0;
  for (int j = 0; j < P.npts; j++) {
    const double *N = sN + j * NNE, *dN = sdN + j * NNE * MDIM;
    double J[SDIM * MDIM];
#pragma unroll
    for (int i = 0; i < SDIM * MDIM; i++) J[i] = 0.0;
#pragma unroll
    for (int a = 0; a < NNE; a++)
#pragma unroll
      for (int d = 0; d < MDIM; d++)
#pragma unroll
        for (int s = 0; s < SDIM; s++) J[s + SDIM * d] += X[a][s] * dN[d * NNE + a];
    double Jac;
    if (SDIM == 3 && MDIM == 3) {
      Jac = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[3] * (J[1] * J[8] - J[7] * J[2]) + J[6] * (J[1] * J[5] - J[4] * J[2]);
    } else if (SDIM == 2 && MDIM == 2) {
      Jac = J[0] * J[3] - J[1] * J[2];
    } else {  // surface in 3-D: |J1 x J2|   (Jacobian, FESetModule.jl:426-435)
      const double c0 = J[1] * J[5] - J[2] * J[4], c1 = J[2] * J[3] - J[0] * J[5], c2 = J[0] * J[4] - J[1] * J[3];
      Jac = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
    }
    if (MDIM == 2 && P.m == 3) Jac = Jac * P.otherdim;  // Jacobianmdim, IntegDomainModule.jl:504-517
    const double wj = sw[j];
#pragma unroll
    for (int mx = 0; mx < NNE; mx++)
#pragma unroll
      for (int k = 0; k <= mx; k++) {
        const double factor = N[k] * N[mx] * Jac * wj;  // FEMMBaseModule.jl:1356
        acc[mx * (mx + 1) / 2 + k] += factor * P.c;
      }
  }
  if (P.vstride > 0) {
    double *out = P.V + slot;
    if (P.compact) {
#pragma unroll
      for (int i = 0; i < NT; i++) out[(int64_t)i * P.vstride] = acc[i];
    } else {
#pragma unroll
      for (int c = 0; c < NNE; c++)
#pragma unroll
        for (int r = 0; r < NNE; r++) out[(int64_t)(c * NNE + r) * P.vstride] = (r <= c) ? acc[c * (c + 1) / 2 + r] : acc[r * (r + 1) / 2 + c];
    }
  } else if (P.compact) {
    double *Ve = P.V + slot * NT;
#pragma unroll
    for (int i = 0; i < NT; i++) Ve[i] = acc[i];
  } else {
    double *Ve = P.V + slot * (NNE * NNE);
#pragma unroll
    for (int c = 0; c < NNE; c++)
#pragma unroll
      for (int r = 0; r < NNE; r++) Ve[c * NNE + r] = (r <= c) ? acc[c * (c + 1) / 2 + r] : acc[r * (r + 1) / 2 + c];
  }
}

template <int NNE, int MDIM, int SDIM>
int32_t launch_dot(fegpu_mesh *mesh, const FormArgs &fa, double *d_V) {
  fegpu_ctx *ctx = mesh->ctx;
  if (mesh->nactive == 0) return FEGPU_OK;
  DotParams P{mesh->conn_act(), mesh->d_xyz, mesh->nnodes, mesh->d_elem_list, mesh->nactive, mesh->d_tab, mesh->d_w, mesh->npts, d_V,
              fa.compact ? 1 : 0, fa.planes ? fa.vstride : 0, fa.coef[0], fa.m, fa.otherdim};
  const size_t smem = sizeof(double) * ((size_t)mesh->npts * NNE * (1 + MDIM) + mesh->npts);
  auto kern = k_dot_scalar<NNE, MDIM, SDIM>;
  if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_for(mesh->nactive, 128), 128, smem, ctx->stream>>>(P);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return FEGPU_OK;
}

}  // namespace

bool fe_dot_scalar_applies(const fegpu_mesh *mesh, const FormArgs &fa) {
  if (fa.form != FORM_DOT || fa.ndn != 1) return false;
  switch (mesh->etype) {
    case FEGPU_T3: case FEGPU_Q4: case FEGPU_T4: case FEGPU_H8: case FEGPU_T10: return true;
  }
  return false;
}

int32_t fe_integrate_dot_scalar(fegpu_mesh *mesh, const FormArgs &fa, double *d_V) {
  switch (mesh->etype) {
    case FEGPU_T3: return mesh->sdim == 2 ? launch_dot<3, 2, 2>(mesh, fa, d_V) : launch_dot<3, 2, 3>(mesh, fa, d_V);
    case FEGPU_Q4: return mesh->sdim == 2 ? launch_dot<4, 2, 2>(mesh, fa, d_V) : launch_dot<4, 2, 3>(mesh, fa, d_V);
    case FEGPU_T4: return launch_dot<4, 3, 3>(mesh, fa, d_V);
    case FEGPU_H8: return launch_dot<8, 3, 3>(mesh, fa, d_V);
    case FEGPU_T10: return launch_dot<10, 3, 3>(mesh, fa, d_V);
  }
  return fegpu_fail(mesh->ctx, FEGPU_ERR_ARG, "internal: scalar dot kernel does not take this element type");
}
