// "Next" rows of the path: the right-hand side and the Dirichlet step
// that sit between assembly() and cg() in every real caller.
//   LinearForm.assembly + ScalarSourceIntegrator   fem/linear_form.py:36-86, fem/scalar_source_integrator.py:13-57,
//                                                  functional.py linear_integral
//   DirichletBC.apply / apply_matrix / apply_vector fem/dirichlet_bc.py:101-235
#include "common.cuh"
#include "sort_scan.cuh"
#include "bc_source.cuh"

namespace fb2 {

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSM * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <int TD>
__device__ __forceinline__ double cell_measure(const double* __restrict__ node, const int* __restrict__ cell, int64_t c);

template <>
__device__ __forceinline__ double cell_measure<2>(const double* __restrict__ node, const int* __restrict__ cell, int64_t c) {
  const double2 p0 = *reinterpret_cast<const double2*>(node + 2 * (int64_t)cell[3 * c]);
  const double2 p1 = *reinterpret_cast<const double2*>(node + 2 * (int64_t)cell[3 * c + 1]);
  const double2 p2 = *reinterpret_cast<const double2*>(node + 2 * (int64_t)cell[3 * c + 2]);
  return 0.5 * ((p1.x - p0.x) * (p2.y - p1.y) - (p1.y - p0.y) * (p2.x - p1.x));
}

template <>
__device__ __forceinline__ double cell_measure<3>(const double* __restrict__ node, const int* __restrict__ cell, int64_t c) {
  const int4 v = *reinterpret_cast<const int4*>(cell + 4 * c);
  const double* q0 = node + 3 * (int64_t)v.x;
  const double* q1 = node + 3 * (int64_t)v.y;
  const double* q2 = node + 3 * (int64_t)v.z;
  const double* q3 = node + 3 * (int64_t)v.w;
  double a[3], b[3], cc[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) { a[m] = q1[m] - q0[m]; b[m] = q2[m] - q1[m]; cc[m] = q3[m] - q2[m]; }
  const double det = a[0] * (b[1] * cc[2] - b[2] * cc[1]) + a[1] * (b[2] * cc[0] - b[0] * cc[2]) + a[2] * (b[0] * cc[1] - b[1] * cc[0]);
  return det / 6.0;
}

// F_e[c][i] = vol_c * sum_q w_q phi_i(q) f_cq ;  f: scalar (kind 0), per cell (1), per quadrature point (2)
template <int TD>
__global__ void __launch_bounds__(256) elem_source_kernel(const double* __restrict__ node, const int* __restrict__ cell, int64_t NC, int L,
                                                          int NQ, const double* __restrict__ phiw /*[NQ][L] = w_q phi_i(q)*/,
                                                          int kind, double scal, const double* __restrict__ f, double* __restrict__ out) {
  extern __shared__ double sp[];
  for (int t = threadIdx.x; t < NQ * L; t += blockDim.x) sp[t] = phiw[t];
  __syncthreads();
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < NC; c += (int64_t)gridDim.x * blockDim.x) {
    const double cm = cell_measure<TD>(node, cell, c);
    for (int i = 0; i < L; ++i) {
      double s = 0.0;
      if (kind == 2) {
        for (int q = 0; q < NQ; ++q) s += sp[q * L + i] * f[c * NQ + q];
      } else {
        for (int q = 0; q < NQ; ++q) s += sp[q * L + i];
        s *= (kind == 1) ? f[c] : 1.0;
      }
      out[c * L + i] = cm * scal * s;
    }
  }
}

// F[d] = sum over the (cell, i) pairs of dof d, ascending (the order of the reference's index_add)
__global__ void __launch_bounds__(256) gather_vector_kernel(int64_t gdof, const int64_t* __restrict__ adj_ptr, const int* __restrict__ adj_pair,
                                                            const double* __restrict__ fe, double* __restrict__ F) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < gdof; d += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t q = adj_ptr[d]; q < adj_ptr[d + 1]; ++q) s += fe[adj_pair[q]];
    F[d] = s;
  }
}

int elem_source(int TD, int64_t NC, int L, int NQ, const double* node, const int* cell, const double* phiw, int kind, double scal,
                const double* f, double* out, cudaStream_t s) {
  if (NC <= 0) return OK;
  const size_t smem = (size_t)NQ * L * sizeof(double);
  if (smem > 48 * 1024) return fail(ERR_UNSUPPORTED, "elem_source: quadrature table too large");
  if (TD == 2) elem_source_kernel<2><<<grid_for(NC), 256, smem, s>>>(node, cell, NC, L, NQ, phiw, kind, scal, f, out);
  else if (TD == 3) elem_source_kernel<3><<<grid_for(NC), 256, smem, s>>>(node, cell, NC, L, NQ, phiw, kind, scal, f, out);
  else return fail(ERR_UNSUPPORTED, "elem_source: TD=%d", TD);
  FB2_LAUNCH_CHECK();
  return OK;
}

int gather_vector(int64_t gdof, const int64_t* adj_ptr, const int* adj_pair, const double* fe, double* F, cudaStream_t s) {
  if (gdof <= 0) return OK;
  gather_vector_kernel<<<grid_for(gdof), 256, 0, s>>>(gdof, adj_ptr, adj_pair, fe, F);
  FB2_LAUNCH_CHECK();
  return OK;
}

// ---- Dirichlet ---------------------------------------------------------------------------
// matrix: rows/columns of boundary dofs removed, unit diagonal on boundary rows (canonical
// sorted CSR; the reference reaches the same matrix through _mul + spdiags, dirichlet_bc.py:133-229)
// Eight lanes own a row and stride over its entries (coalesced 32-/64-byte runs of col / val instead of one thread
// walking a whole row: DirichletBC.apply 60.9 -> 4.5 ms on the 286 M-nonzero elasticity matrix, profiles/r02_cold_breakdown.txt); kept entries
// keep their order -- the output position of an entry is the number of kept entries before it, from a ballot.
constexpr int BC_G = 8;

__global__ void __launch_bounds__(256) bc_count_kernel(int64_t n, const int64_t* __restrict__ crow, const int* __restrict__ col,
                                                       const uint8_t* __restrict__ isbd, int* __restrict__ cnt) {
  const int g = threadIdx.x % BC_G;
  const int64_t ngroup = (int64_t)gridDim.x * (blockDim.x / BC_G);
  const int64_t nrow_pad = (n + 31) / 32 * 32;           // whole warps stay in the loop together (shuffles below)
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / BC_G; r < nrow_pad; r += ngroup) {
    int c = 0;
    const bool live = r < n;
    const bool bd = live && isbd[r];
    if (live && !bd)
      for (int64_t k = crow[r] + g; k < crow[r + 1]; k += BC_G) c += isbd[col[k]] ? 0 : 1;
#pragma unroll
    for (int o = BC_G / 2; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (live && g == 0) cnt[r] = bd ? 1 : c;
  }
}

__global__ void __launch_bounds__(256) bc_fill_kernel(int64_t n, const int64_t* __restrict__ crow, const int* __restrict__ col,
                                                      const double* __restrict__ val, const uint8_t* __restrict__ isbd,
                                                      const int64_t* __restrict__ crow_new, int* __restrict__ col_new,
                                                      double* __restrict__ val_new) {
  const int lane = threadIdx.x & 31, g = lane % BC_G, gsh = lane / BC_G * BC_G;
  const unsigned below = (1u << g) - 1u;
  const int64_t ngroup = (int64_t)gridDim.x * (blockDim.x / BC_G);
  const int64_t nrow_pad = (n + 31) / 32 * 32;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / BC_G; r < nrow_pad; r += ngroup) {
    const bool live = r < n;
    const bool bd = live && isbd[r];
    int64_t o = live ? crow_new[r] : 0;
    const int64_t k0 = live ? crow[r] : 0, k1 = (live && !bd) ? crow[r + 1] : k0;
    if (bd && g == 0) { col_new[o] = (int)r; val_new[o] = 1.0; }
    // rows of one warp have different lengths: every lane runs as many rounds as the warp's longest row needs
    int rounds = (int)((k1 - k0 + BC_G - 1) / BC_G);
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, of));
    for (int t = 0; t < rounds; ++t) {
      const int64_t k = k0 + (int64_t)t * BC_G + g;
      int c = 0;
      bool keep = false;
      if (k < k1) { c = col[k]; keep = !isbd[c]; }
      const unsigned kept = (__ballot_sync(0xffffffffu, keep) >> gsh) & ((1u << BC_G) - 1u);
      if (keep) {
        const int64_t dst = o + __popc(kept & below);
        col_new[dst] = c;
        val_new[dst] = val[k];
      }
      o += __popc(kept);
    }
  }
}

// f <- isbd ? uh : f   (f already holds f - A uh)
__global__ void __launch_bounds__(256) bc_vector_kernel(int64_t n, const uint8_t* __restrict__ isbd, const double* __restrict__ uh,
                                                        double* __restrict__ f) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    if (isbd[r]) f[r] = uh[r];
}

size_t bc_workspace_bytes(int64_t n) { return align_up((size_t)n * 4) + scan_workspace_bytes(n) + 1024; }

int bc_matrix_count(int64_t n, const int64_t* crow, const int* col, const uint8_t* isbd, int64_t* crow_new, int64_t* nnz_host, void* ws,
                    cudaStream_t s) {
  Carver c(ws);
  int* cnt = c.take<int>(n);
  void* scan_ws = c.take<char>(scan_workspace_bytes(n));
  if (n > 0) bc_count_kernel<<<grid_for(n * BC_G), 256, 0, s>>>(n, crow, col, isbd, cnt);
  FB2_LAUNCH_CHECK();
  FB2_TRY(exclusive_scan_i32(cnt, crow_new, n, true, scan_ws, s));
  FB2_CUDA(cudaMemcpyAsync(nnz_host, crow_new + n, 8, cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  return OK;
}

int bc_matrix_fill(int64_t n, const int64_t* crow, const int* col, const double* val, const uint8_t* isbd, const int64_t* crow_new,
                   int* col_new, double* val_new, cudaStream_t s) {
  if (n <= 0) return OK;
  bc_fill_kernel<<<grid_for(n * BC_G), 256, 0, s>>>(n, crow, col, val, isbd, crow_new, col_new, val_new);
  FB2_LAUNCH_CHECK();
  return OK;
}

int bc_vector(int64_t n, const uint8_t* isbd, const double* uh, double* f, cudaStream_t s) {
  if (n <= 0) return OK;
  bc_vector_kernel<<<grid_for(n), 256, 0, s>>>(n, isbd, uh, f);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2

// ---- matrix-free operator (row f4): v = sum_c P_c^T K_e[c] P_c u without the global matrix
// (BilinearForm.__matmul__ before assembly, fem/bilinear_form.py:126-158).  Row-owner form: dof d
// sums, over its (cell, i) pairs in canonical order, the dot of K_e's row i with u gathered at the
// cell's dofs -- deterministic, no atomics.
namespace fb2 {
__global__ void __launch_bounds__(256) matfree_apply_kernel(int64_t gdof, int L, const int64_t* __restrict__ adj_ptr,
                                                            const int* __restrict__ adj_pair, const int* __restrict__ c2d,
                                                            const double* __restrict__ ke, const double* __restrict__ u,
                                                            double* __restrict__ v) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < gdof; d += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t q = adj_ptr[d]; q < adj_ptr[d + 1]; ++q) {
      const int pair = adj_pair[q];
      const int64_t c = pair / L;
      const double* row = ke + (int64_t)pair * L;          // K_e[c][i][:] (pair = c*L + i)
      const int* dofs = c2d + c * L;
      double t = 0.0;
      for (int j = 0; j < L; ++j) t += row[j] * u[dofs[j]];
      s += t;
    }
    v[d] = s;
  }
}

int matfree_apply(int64_t gdof, int L, const int64_t* adj_ptr, const int* adj_pair, const int* c2d, const double* ke, const double* u,
                  double* v, cudaStream_t s) {
  if (gdof <= 0) return OK;
  matfree_apply_kernel<<<grid_for(gdof), 256, 0, s>>>(gdof, L, adj_ptr, adj_pair, c2d, ke, u, v);
  FB2_LAUNCH_CHECK();
  return OK;
}

// ---- bc_to_point -------------------------------------------------------------------------
// one thread per (cell, point): consecutive threads write consecutive points (coalesced), the
// vertices of a cell are shared through L1 by the NQ threads that need them
template <int TD>
__global__ void __launch_bounds__(256) bc_to_points_kernel(int64_t NC, int NQ, const double* __restrict__ node, const int* __restrict__ cell,
                                                           const double* __restrict__ bcs, double* __restrict__ out) {
  const int64_t total = NC * NQ;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / NQ;
    const int q = (int)(t - c * NQ);
    double x[TD];
#pragma unroll
    for (int m = 0; m < TD; ++m) x[m] = 0.0;
#pragma unroll
    for (int j = 0; j <= TD; ++j) {
      const int v = cell[c * (TD + 1) + j];
      const double b = bcs[q * (TD + 1) + j];
#pragma unroll
      for (int m = 0; m < TD; ++m) x[m] += b * node[(int64_t)v * TD + m];      // same j-ascending order as the reference's einsum
    }
#pragma unroll
    for (int m = 0; m < TD; ++m) out[t * TD + m] = x[m];
  }
}

int bc_to_points(int TD, int64_t NC, int NQ, const double* node, const int* cell, const double* bcs, double* out, cudaStream_t s) {
  if (NC <= 0 || NQ <= 0) return OK;
  const int64_t total = NC * NQ;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSM * 32);
  if (TD == 2) bc_to_points_kernel<2><<<grid, 256, 0, s>>>(NC, NQ, node, cell, bcs, out);
  else if (TD == 3) bc_to_points_kernel<3><<<grid, 256, 0, s>>>(NC, NQ, node, cell, bcs, out);
  else return fail(ERR_UNSUPPORTED, "bc_to_points: TD must be 2 or 3");
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2
