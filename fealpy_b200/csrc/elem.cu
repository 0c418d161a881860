// K1: per-cell element matrices for Lagrange simplex elements, FP64, one thread per cell.
//
// Follows the arithmetic of the reference integrators (paths relative to the reference root):
//   ScalarDiffusionIntegrator.assembly          fem/scalar_diffusion_integrator.py:54-79
//   ScalarMassIntegrator.assembly               fem/scalar_mass_integrator.py:47-54
//   LinearElasticityIntegrator.assembly         fem/linear_elasticity_integrator.py:60-181
//   bilinear_integral (coefficient rules)       functional.py:68-106
//   cell measure / grad lambda                  backend/numpy_backend.py:413-421,586-598,619-629
//
// Design: geometry (vol, Dlambda) lives in registers; the reference-element tables (pre-
// contracted M tensors for constant coefficients, R / phi / weights for quadrature loops)
// are staged once per CTA in shared memory and read with warp-uniform (broadcast) LDS;
// results are staged per warp in shared memory and written back as contiguous segments so
// that the (NC, l, l) C-order output is stored with full sectors.
#include "common.cuh"
#include "elem.cuh"

namespace fb2 {

template <int TD>
struct Geo {
  static constexpr int NV = TD + 1;
  double cm;            // signed measure, det/TD! (never abs'ed, like the reference)
  double D[NV][TD];     // grad lambda
};

__device__ __forceinline__ void load_geo(const double* __restrict__ node, const int* __restrict__ cell, int64_t c, Geo<2>& g) {
  const int v0 = cell[3 * c], v1 = cell[3 * c + 1], v2 = cell[3 * c + 2];
  const double2 p0 = *reinterpret_cast<const double2*>(node + 2 * (int64_t)v0);
  const double2 p1 = *reinterpret_cast<const double2*>(node + 2 * (int64_t)v1);
  const double2 p2 = *reinterpret_cast<const double2*>(node + 2 * (int64_t)v2);
  const double e0x = p2.x - p1.x, e0y = p2.y - p1.y;
  const double e1x = p0.x - p2.x, e1y = p0.y - p2.y;
  const double e2x = p1.x - p0.x, e2y = p1.y - p0.y;
  const double nv = e0x * e1y - e0y * e1x;
  const double inv = 1.0 / nv;
  g.D[0][0] = -e0y * inv; g.D[0][1] = e0x * inv;
  g.D[1][0] = -e1y * inv; g.D[1][1] = e1x * inv;
  g.D[2][0] = -e2y * inv; g.D[2][1] = e2x * inv;
  // simplex_measure: det([p1-p0; p2-p1]) / 2
  g.cm = 0.5 * (e2x * e0y - e2y * e0x);
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ void load_geo(const double* __restrict__ node, const int* __restrict__ cell, int64_t c, Geo<3>& g) {
  const int4 v = *reinterpret_cast<const int4*>(cell + 4 * c);
  const int vid[4] = {v.x, v.y, v.z, v.w};
  double P[4][3];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double* q = node + 3 * (int64_t)vid[k];
    P[k][0] = q[0]; P[k][1] = q[1]; P[k][2] = q[2];
  }
  // volume = det([p1-p0; p2-p1; p3-p2]) / 6
  double a[3], b[3], cc[3], bc[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) { a[m] = P[1][m] - P[0][m]; b[m] = P[2][m] - P[1][m]; cc[m] = P[3][m] - P[2][m]; }
  cross3(b, cc, bc);
  const double det = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2];
  g.cm = det / 6.0;
  const double inv = 1.0 / det;      // 1 / (6 vol)
  // Dlambda_i = cross(v_jm, v_jk) / (6 vol), (j,k,m) = localFace[i]
  constexpr int LF[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = LF[i][0], k = LF[i][1], mm = LF[i][2];
    double vjk[3], vjm[3], cr[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { vjk[m] = P[k][m] - P[j][m]; vjm[m] = P[mm][m] - P[j][m]; }
    cross3(vjm, vjk, cr);
#pragma unroll
    for (int m = 0; m < 3; ++m) g.D[i][m] = cr[m] * inv;
  }
}

// warp-staged store of 32 cells' dense (LL doubles) blocks to out[(c0+lane)*LL + e]
template <int LL>
struct WarpStage {
  static constexpr int S = LL | 1;     // odd stride -> conflict-free 64-bit smem stores
  static constexpr size_t bytes_per_warp = (size_t)32 * S * sizeof(double);
};

template <int LL>
__device__ __forceinline__ void warp_flush(const double* stage, double* __restrict__ out, int64_t c0, int64_t NC, int accumulate = 0) {
  constexpr int S = WarpStage<LL>::S;
  const int lane = threadIdx.x & 31;
  const int64_t ncell = (NC - c0) < 32 ? (NC - c0) : 32;
  const int64_t tot = ncell * LL;
  double* dst = out + c0 * LL;
  if (accumulate) {
    for (int64_t idx = lane; idx < tot; idx += 32) {
      const int cl = (int)(idx / LL), e = (int)(idx - (int64_t)cl * LL);
      dst[idx] += stage[cl * S + e];
    }
    return;
  }
  for (int64_t idx = lane; idx < tot; idx += 32) {
    const int cl = (int)(idx / LL), e = (int)(idx - (int64_t)cl * LL);
    dst[idx] = stage[cl * S + e];
  }
}

// -------------------------------------------------------------------------------------
// constant-coefficient scalar kernel:  K = kd * (Ms : G)  +  km * cm * Mm
//   Ms[i][j][kl] = sum_q w R[q,i,k] R[q,j,l] (+ k<->l), kl over the upper triangle
//   G[kl] = cm * Dlam_k . Dlam_l ;  kd / km = scalar * optional per-cell array
// -------------------------------------------------------------------------------------
template <int TD, int L, bool STAGED>
__global__ void __launch_bounds__(128) elem_const_kernel(ElemConstArgs a) {
  constexpr int NV = TD + 1, NG = NV * (NV + 1) / 2, LL = L * L;
  extern __shared__ __align__(16) double sm[];
  double* sMs = sm;                               // [LL][NG] (only if diffusion)
  double* sMm = sMs + (a.has_diff ? LL * NG : 0); // [LL]
  double* sStage = sMm + (a.has_mass ? LL : 0);
  if (a.has_diff) for (int i = threadIdx.x; i < LL * NG; i += blockDim.x) sMs[i] = a.Ms[i];
  if (a.has_mass) for (int i = threadIdx.x; i < LL; i += blockDim.x) sMm[i] = a.Mm[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wid) * 32;
  if (c0 >= a.NC) return;
  const int64_t c = c0 + lane;
  const bool ok = c < a.NC;
  double* stage = STAGED ? sStage + (size_t)wid * 32 * WarpStage<LL>::S : nullptr;

  double G[NG], kd = 0.0, km = 0.0;
  if (ok) {
    Geo<TD> g;
    load_geo(a.node, a.cell, c, g);
    int t = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int l = k; l < NV; ++l) {
        double d = 0.0;
#pragma unroll
        for (int m = 0; m < TD; ++m) d += g.D[k][m] * g.D[l][m];
        G[t++] = d * g.cm;
      }
    kd = a.scal_d * (a.coef_d ? a.coef_d[c] : 1.0);
    km = a.scal_m * (a.coef_m ? a.coef_m[c] : 1.0) * g.cm;
  }
  if (ok) {
#pragma unroll 1
    for (int i = 0; i < L; ++i) {
#pragma unroll
      for (int j = 0; j < L; ++j) {
        double v = 0.0;
        if (a.has_diff) {
          const double* m = sMs + (i * L + j) * NG;
          double s = 0.0;
#pragma unroll
          for (int t = 0; t < NG; ++t) s += m[t] * G[t];
          v = kd * s;
        }
        if (a.has_mass) v += km * sMm[i * L + j];
        if (STAGED) stage[lane * WarpStage<LL>::S + i * L + j] = v;
        else if (a.accumulate) a.out[c * LL + i * L + j] += v;
        else a.out[c * LL + i * L + j] = v;
      }
    }
  }
  if (STAGED) {
    __syncwarp();
    warp_flush<LL>(stage, a.out, c0, a.NC, a.accumulate);
  }
}

// -------------------------------------------------------------------------------------
// quadrature-loop scalar kernel (variable coefficients):
//   diffusion: K[i][j] = cm * sum_q w_q * gphi_i^T C_q gphi_j,  gphi = R[q] Dlam
//              C_q = kappa[c,q] * I  (coef_kind 2)  or full matrix (NC,NQ,GD,GD) (coef_kind 3)
//   mass:      K[i][j] = cm * sum_q w_q kappa[c,q] phi_i phi_j
// rows are processed in passes of RB rows so that the accumulators stay in registers.
// -------------------------------------------------------------------------------------
// constant-coefficient terms folded into the quadrature kernel's final write (see ElemQuadArgs)
template <int TD>
struct QuadExtra {
  static constexpr int NG = (TD + 1) * (TD + 2) / 2;
  const double* sMs;     // shared-memory copies (null = absent)
  const double* sMm;
  double kd, km;         // km already carries |K|
};

// SYM (scalar coefficient fields and mass terms): K_e is symmetric -- products commute and (i, j), (j, i) are summed over
// (q, d) in the same order -- so only j >= i is accumulated (55 instead of 100 FMA chains per quadrature point on 10 local
// dofs, gradients formed only for the columns a pass still needs) and the block is mirrored on the way out.  ncu of the
// full version on config 3 (profiles/r02_ncu_asm_v6.txt): FP64 pipe 63.5 % busy, 252 registers -- the kernel is bound by
// its DFMA count, which is what this halves.  (DMMA cannot help on B200: its FP64 tensor peak equals the DFMA peak.)
template <int TD, int L, int RB, int R0, bool STAGED, bool SYM>
__device__ __forceinline__ void quad_pass(const ElemQuadArgs& a, const Geo<TD>& g, const double* sW, const double* sT,
                                          const double* coefc, double* stage, int64_t c, const QuadExtra<TD>& xt) {
  constexpr int NV = TD + 1, LL = L * L;
  constexpr int NR = (R0 + RB <= L) ? RB : (L - R0);     // rows in this pass
  constexpr int J0 = SYM ? R0 : 0;                       // first column this pass touches
  const int NQ = a.NQ, lane = threadIdx.x & 31;
  double acc[NR][L];
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < L; ++j) acc[i][j] = 0.0;
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    const double w = sW[q];
    if (a.is_mass) {
      const double wk = w * coefc[q];
      const double* ph = sT + q * L;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const double pi = wk * ph[R0 + i];
#pragma unroll
        for (int j = 0; j < L; ++j)
          if (!SYM || j >= R0 + i) acc[i][j] = fma(pi, ph[j], acc[i][j]);
      }
    } else {
      double gp[L][TD];                       // physical gradients of the basis functions at q (columns >= J0)
      const double* R = sT + q * L * NV;
#pragma unroll
      for (int j = J0; j < L; ++j)
#pragma unroll
        for (int m = 0; m < TD; ++m) {
          double s = 0.0;
#pragma unroll
          for (int b = 0; b < NV; ++b) s += R[j * NV + b] * g.D[b][m];
          gp[j][m] = s;
        }
      if constexpr (!SYM) {                   // full matrix coefficient: gphi_i^T C gphi_j (C need not be symmetric)
        double C[TD][TD];
#pragma unroll
        for (int d = 0; d < TD; ++d)
#pragma unroll
          for (int n = 0; n < TD; ++n) C[d][n] = w * coefc[(q * TD + d) * TD + n];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          double tv[TD];
#pragma unroll
          for (int n = 0; n < TD; ++n) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < TD; ++d) s += gp[R0 + i][d] * C[d][n];
            tv[n] = s;
          }
#pragma unroll
          for (int j = 0; j < L; ++j)
#pragma unroll
            for (int n = 0; n < TD; ++n) acc[i][j] = fma(tv[n], gp[j][n], acc[i][j]);
        }
      } else {
        const double wk = w * coefc[q];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          double tv[TD];
#pragma unroll
          for (int n = 0; n < TD; ++n) tv[n] = wk * gp[R0 + i][n];
#pragma unroll
          for (int j = R0 + i; j < L; ++j)
#pragma unroll
            for (int n = 0; n < TD; ++n) acc[i][j] = fma(tv[n], gp[j][n], acc[i][j]);
        }
      }
    }
  }
  double G[QuadExtra<TD>::NG];          // kd |K| grad(lambda_k).grad(lambda_l), upper triangle: formed here, after the
  if (xt.sMs) {                         // quadrature loop, so that it holds no registers while the accumulators are hot
    int t = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int l = k; l < NV; ++l) {
        double d = 0.0;
#pragma unroll
        for (int m = 0; m < TD; ++m) d += g.D[k][m] * g.D[l][m];
        G[t++] = d * g.cm * xt.kd;
      }
  }
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int j = 0; j < L; ++j) {
      if (SYM && j < R0 + i) continue;
      double v = acc[i][j] * g.cm;
      if (xt.sMs) {
        const double* m = xt.sMs + ((R0 + i) * L + j) * QuadExtra<TD>::NG;
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < QuadExtra<TD>::NG; ++t) s += m[t] * G[t];
        v += s;
      }
      if (xt.sMm) v += xt.km * xt.sMm[(R0 + i) * L + j];
      const int e0 = (R0 + i) * L + j, e1 = j * L + (R0 + i);       // (i, j) and its mirror image
      if (STAGED) {
        stage[lane * WarpStage<LL>::S + e0] = v;
        if (SYM && e1 != e0) stage[lane * WarpStage<LL>::S + e1] = v;
      } else if (a.accumulate) {
        a.out[c * LL + e0] += v;
        if (SYM && e1 != e0) a.out[c * LL + e1] += v;
      } else {
        a.out[c * LL + e0] = v;
        if (SYM && e1 != e0) a.out[c * LL + e1] = v;
      }
    }
  if constexpr (R0 + RB < L) quad_pass<TD, L, RB, R0 + RB, STAGED, SYM>(a, g, sW, sT, coefc, stage, c, xt);
}

template <int TD, int L, int RB, bool STAGED, bool SYM>
__global__ void __launch_bounds__(128) elem_quad_kernel(ElemQuadArgs a) {
  constexpr int NV = TD + 1, LL = L * L;
  extern __shared__ __align__(16) double sm[];
  const int NQ = a.NQ;
  double* sW = sm;                                     // [NQ]
  double* sT = sW + NQ;                                // diffusion: R[NQ][L][NV]; mass: phi[NQ][L]
  const int tab = a.is_mass ? NQ * L : NQ * L * NV;
  constexpr int NG = QuadExtra<TD>::NG;
  double* sXs = sT + tab;                              // folded constant terms: Ms [LL][NG], Mm [LL]
  double* sXm = sXs + (a.xMs ? LL * NG : 0);
  double* sStage = sXm + (a.xMm ? LL : 0);
  for (int i = threadIdx.x; i < NQ; i += blockDim.x) sW[i] = a.ws[i];
  for (int i = threadIdx.x; i < tab; i += blockDim.x) sT[i] = a.tab[i];
  if (a.xMs) for (int i = threadIdx.x; i < LL * NG; i += blockDim.x) sXs[i] = a.xMs[i];
  if (a.xMm) for (int i = threadIdx.x; i < LL; i += blockDim.x) sXm[i] = a.xMm[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wid) * 32;
  if (c0 >= a.NC) return;
  const int64_t c = c0 + lane;
  double* stage = STAGED ? sStage + (size_t)wid * 32 * WarpStage<LL>::S : nullptr;
  if (c < a.NC) {
    Geo<TD> g;
    load_geo(a.node, a.cell, c, g);
    const double* coefc = a.coef + c * (int64_t)NQ * (a.coef_kind == 3 ? TD * TD : 1);
    QuadExtra<TD> xt;
    xt.sMs = a.xMs ? sXs : nullptr;
    xt.sMm = a.xMm ? sXm : nullptr;
    xt.kd = a.x_scal_d * (a.x_coef_d ? a.x_coef_d[c] : 1.0);
    xt.km = a.x_scal_m * (a.x_coef_m ? a.x_coef_m[c] : 1.0) * g.cm;
    quad_pass<TD, L, RB, 0, STAGED, SYM>(a, g, sW, sT, coefc, stage, c, xt);
  }
  if (STAGED) {
    __syncwarp();
    warp_flush<LL>(stage, a.out, c0, a.NC, a.accumulate);
  }
}

// -------------------------------------------------------------------------------------
// linear elasticity (constant isotropic material), simplex cells:
//   A_ab[i][j] = cm * sum_{kl} M4[i][j][k][l] Dlam[k][a] Dlam[l][b]
//   KK(i a, j a) = d_diag * A_aa + d_shear * sum_{b != a} A_bb
//   KK(i a, j b) = d_lam * A_ab + d_shear * A_ba                       (a != b)
//   layout: interleaved row = GD*i + a (dof_priority False) or blocked row = a*L + i (True)
// -------------------------------------------------------------------------------------
// One thread per cell computes; the TD rows of local index i (TD*LD values per cell) are parked in a
// per-warp shared-memory stage (odd stride: conflict-free) and written out cooperatively, so the
// stores are contiguous runs of TD*LD (interleaved) or LD (dof_priority) doubles instead of one
// scattered 8-byte store per thread (the first version wrote 16.1 GB at 1.1 TB/s for config 4).
template <int TD, int L>
__global__ void __launch_bounds__(128) elem_elasticity_kernel(ElemElasticityArgs a) {
  constexpr int NV = TD + 1, LD = L * TD, ROWS = TD * LD, STRIDE = ROWS + 1;
  extern __shared__ __align__(16) double sm[];
  double* stage = sm + L * L * NV * NV + (threadIdx.x >> 5) * 32 * STRIDE;
  for (int i = threadIdx.x; i < L * L * NV * NV; i += blockDim.x) sm[i] = a.M4[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t cw = c - lane;                              // first cell of this warp
  const bool live = c < a.NC;
  Geo<TD> g;
  if (live) load_geo(a.node, a.cell, c, g);
  double* mine = stage + lane * STRIDE;
#pragma unroll 1
  for (int i = 0; i < L; ++i) {
    if (live) {
#pragma unroll 1
      for (int j = 0; j < L; ++j) {
        const double* m = sm + (i * L + j) * NV * NV;
        double A[TD][TD];
#pragma unroll
        for (int x = 0; x < TD; ++x)
#pragma unroll
          for (int y = 0; y < TD; ++y) A[x][y] = 0.0;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
          for (int l = 0; l < NV; ++l) {
            const double mv = m[k * NV + l];
#pragma unroll
            for (int x = 0; x < TD; ++x)
#pragma unroll
              for (int y = 0; y < TD; ++y) A[x][y] += mv * (g.D[k][x] * g.D[l][y]);
          }
#pragma unroll
        for (int x = 0; x < TD; ++x)
#pragma unroll
          for (int y = 0; y < TD; ++y) {
            double v;
            if (x == y) {
              double oth = 0.0;
#pragma unroll
              for (int z = 0; z < TD; ++z) if (z != x) oth += A[z][z];
              v = a.d_diag * A[x][x] + a.d_shear * oth;
            } else {
              v = a.d_lam * A[x][y] + a.d_shear * A[y][x];
            }
            v *= g.cm;
            const int col = a.dof_priority ? y * L + j : j * TD + y;
            mine[x * LD + col] = v;
          }
      }
    }
    __syncwarp();
    for (int idx = lane; idx < 32 * ROWS; idx += 32) {
      const int cl = idx / ROWS, k = idx - cl * ROWS;
      if (cw + cl < a.NC) {
        const int x = k / LD, col = k - x * LD;
        const int row = a.dof_priority ? x * L + i : i * TD + x;
        a.out[(cw + cl) * (int64_t)LD * LD + row * LD + col] = stage[cl * STRIDE + k];
      }
    }
    __syncwarp();
  }
}

// per-cell record (grad lambda (NV x TD), signed measure) for the fused P1 elasticity assembly
template <int TD>
__global__ void __launch_bounds__(256) cell_gradients_kernel(const double* __restrict__ node, const int* __restrict__ cell, int64_t NC,
                                                             double* __restrict__ out) {
  constexpr int NV = TD + 1, RS = NV * TD + 1;
  __shared__ double stage[256 * RS];
  const int64_t c0 = (int64_t)blockIdx.x * 256, c = c0 + threadIdx.x;
  if (c < NC) {
    Geo<TD> g;
    load_geo(node, cell, c, g);
    double* h = stage + threadIdx.x * RS;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int x = 0; x < TD; ++x) h[k * TD + x] = g.D[k][x];
    h[NV * TD] = g.cm;
  }
  __syncthreads();
  const int64_t ncell = (NC - c0) < 256 ? (NC - c0) : 256;
  for (int64_t t = threadIdx.x; t < ncell * RS; t += 256) out[c0 * RS + t] = stage[t];
}

int cell_gradients(int TD, int64_t NC, const double* node, const int* cell, double* out, cudaStream_t s) {
  if (NC <= 0) return OK;
  const unsigned grid = (unsigned)ceil_div(NC, 256);
  if (TD == 2) cell_gradients_kernel<2><<<grid, 256, 0, s>>>(node, cell, NC, out);
  else if (TD == 3) cell_gradients_kernel<3><<<grid, 256, 0, s>>>(node, cell, NC, out);
  else return fail(ERR_UNSUPPORTED, "cell_gradients: TD must be 2 or 3");
  FB2_LAUNCH_CHECK();
  return OK;
}

// physical gradients of the basis functions, the (NC, NQ, ldof, GD) array of LagrangeFESpace.grad_basis /
// mesh.grad_shape_function(variables='x') (mesh/mesh_base.py:713-749, mesh/triangle_mesh.py:142-153):
//   gphi[c][q][i][m] = sum_b R[q][i][b] * Dlambda[c][b][m],  Dlambda from the records of cell_gradients
// one thread per (c, q, i): consecutive threads write consecutive TD-vectors (coalesced)
template <int TD>
__global__ void __launch_bounds__(256) grad_basis_kernel(int64_t total, int NQL, const double* __restrict__ rec, const double* __restrict__ R,
                                                         double* __restrict__ out) {
  constexpr int NV = TD + 1, RS = NV * TD + 1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / NQL;
    const int qi = (int)(t - c * NQL);
    const double* D = rec + c * RS;
    const double* r = R + (int64_t)qi * NV;
    double g[TD];
#pragma unroll
    for (int m = 0; m < TD; ++m) g[m] = 0.0;
#pragma unroll
    for (int b = 0; b < NV; ++b)
#pragma unroll
      for (int m = 0; m < TD; ++m) g[m] += r[b] * D[b * TD + m];
#pragma unroll
    for (int m = 0; m < TD; ++m) out[t * TD + m] = g[m];
  }
}

int grad_basis(int TD, int64_t NC, int NQ, int L, const double* rec, const double* R, double* out, cudaStream_t s) {
  if (NC <= 0 || NQ <= 0 || L <= 0) return OK;
  const int64_t total = NC * NQ * L;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSM * 32);
  if (TD == 2) grad_basis_kernel<2><<<grid, 256, 0, s>>>(total, NQ * L, rec, R, out);
  else if (TD == 3) grad_basis_kernel<3><<<grid, 256, 0, s>>>(total, NQ * L, rec, R, out);
  else return fail(ERR_UNSUPPORTED, "grad_basis: TD must be 2 or 3");
  FB2_LAUNCH_CHECK();
  return OK;
}

// -------------------------------------------------------------------------------------
// host-side dispatch
// -------------------------------------------------------------------------------------
static int ldof_of(int TD, int p) { return TD == 2 ? (p + 1) * (p + 2) / 2 : (p + 1) * (p + 2) * (p + 3) / 6; }

template <int TD, int L>
static int launch_const(const ElemConstArgs& a, cudaStream_t s) {
  constexpr int NV = TD + 1, NG = NV * (NV + 1) / 2, LL = L * L;
  size_t tab = ((a.has_diff ? LL * NG : 0) + (a.has_mass ? LL : 0)) * sizeof(double);
  constexpr bool STAGED = WarpStage<LL>::bytes_per_warp <= 26 * 1024;
  const int warps = 4;
  size_t smem = tab + (STAGED ? warps * WarpStage<LL>::bytes_per_warp : 0);
  auto kern = elem_const_kernel<TD, L, STAGED>;
  FB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nb = ceil_div(a.NC, 32 * warps);
  kern<<<(unsigned)nb, 32 * warps, smem, s>>>(a);
  FB2_LAUNCH_CHECK();
  return OK;
}

template <int TD, int L>
static int launch_quad(const ElemQuadArgs& a, cudaStream_t s) {
  constexpr int NV = TD + 1, LL = L * L;
  constexpr int RB = L <= 6 ? L : (L <= 10 ? 5 : 2);
  constexpr int NG = NV * (NV + 1) / 2;
  size_t tab = ((size_t)a.NQ + (size_t)a.NQ * L * (a.is_mass ? 1 : NV) + (a.xMs ? LL * NG : 0) + (a.xMm ? LL : 0)) * sizeof(double);
  constexpr bool STAGED = WarpStage<LL>::bytes_per_warp <= 26 * 1024;
  const int warps = 4;
  size_t smem = tab + (STAGED ? warps * WarpStage<LL>::bytes_per_warp : 0);
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "elem_quad: quadrature table too large for shared memory (NQ=%d)", a.NQ);
  auto kern = a.coef_kind == 3 ? elem_quad_kernel<TD, L, RB, STAGED, false> : elem_quad_kernel<TD, L, RB, STAGED, true>;
  FB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nb = ceil_div(a.NC, 32 * warps);
  kern<<<(unsigned)nb, 32 * warps, smem, s>>>(a);
  FB2_LAUNCH_CHECK();
  return OK;
}

template <int TD, int L>
static int launch_elast(const ElemElasticityArgs& a, cudaStream_t s) {
  constexpr int NV = TD + 1;
  size_t smem = ((size_t)L * L * NV * NV + (size_t)4 * 32 * (TD * L * TD + 1)) * sizeof(double);      // M4 table + 4 warp stages
  auto kern = elem_elasticity_kernel<TD, L>;
  FB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)ceil_div(a.NC, 128), 128, smem, s>>>(a);
  FB2_LAUNCH_CHECK();
  return OK;
}

#define FB2_DISPATCH_TD_P(FN, TDv, Pv, ...)                                          \
  do {                                                                               \
    const int _key = (TDv) * 10 + (Pv);                                              \
    switch (_key) {                                                                  \
      case 21: return FN<2, 3>(__VA_ARGS__);                                         \
      case 22: return FN<2, 6>(__VA_ARGS__);                                         \
      case 23: return FN<2, 10>(__VA_ARGS__);                                        \
      case 31: return FN<3, 4>(__VA_ARGS__);                                         \
      case 32: return FN<3, 10>(__VA_ARGS__);                                        \
      case 33: return FN<3, 20>(__VA_ARGS__);                                        \
      default:                                                                       \
        return fail(ERR_UNSUPPORTED, "unsupported element TD=%d p=%d (simplex p=1..3)", (TDv), (Pv)); \
    }                                                                                \
  } while (0)

int elem_const(int TD, int p, const ElemConstArgs& a, cudaStream_t s) {
  if (a.NC <= 0) return OK;
  (void)ldof_of;
  FB2_DISPATCH_TD_P(launch_const, TD, p, a, s);
}
int elem_quad(int TD, int p, const ElemQuadArgs& a, cudaStream_t s) {
  if (a.NC <= 0) return OK;
  FB2_DISPATCH_TD_P(launch_quad, TD, p, a, s);
}
int elem_elasticity(int TD, int p, const ElemElasticityArgs& a, cudaStream_t s) {
  if (a.NC <= 0) return OK;
  const int key = TD * 10 + p;
  switch (key) {
    case 21: return launch_elast<2, 3>(a, s);
    case 22: return launch_elast<2, 6>(a, s);
    case 23: return launch_elast<2, 10>(a, s);
    case 31: return launch_elast<3, 4>(a, s);
    case 32: return launch_elast<3, 10>(a, s);
    default: return fail(ERR_UNSUPPORTED, "elasticity: unsupported element TD=%d p=%d", TD, p);
  }
}

}  // namespace fb2
