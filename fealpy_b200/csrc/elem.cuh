// argument blocks of the element-matrix kernels (elem.cu)
#pragma once
#include "common.cuh"

namespace fb2 {

struct ElemConstArgs {
  const double* node;    // (NN, GD)
  const int* cell;       // (NC, TD+1)
  int64_t NC;
  int has_diff, has_mass;
  const double* Ms;      // [L][L][NG]  pre-contracted diffusion tensor (device)
  const double* Mm;      // [L][L]      reference mass matrix (device)
  double scal_d, scal_m; // scalar factors
  const double* coef_d;  // optional per-cell factor (NC,) or null
  const double* coef_m;
  double* out;           // (NC, L, L)
  int accumulate;        // 1: out += K (a later term of the same group / form), 0: out = K
};

struct ElemQuadArgs {
  const double* node;
  const int* cell;
  int64_t NC;
  int is_mass;           // 0: diffusion, 1: mass
  int NQ;
  const double* ws;      // [NQ]
  const double* tab;     // diffusion: R[NQ][L][TD+1]; mass: phi[NQ][L]
  int coef_kind;         // 2: (NC,NQ) scalar field, 3: (NC,NQ,GD,GD) matrix field (diffusion only)
  const double* coef;
  double* out;
  int accumulate;        // 1: out += K, 0: out = K
  // optional constant / per-cell coefficient terms folded into the same pass (the K of elem_const_kernel):
  // K += x_kd * (xMs : G) + x_km * |K| * xMm -- a variable-coefficient diffusion plus a constant mass term writes K_e once
  const double* xMs;     // [L][L][NG] or null
  const double* xMm;     // [L][L] or null
  double x_scal_d, x_scal_m;
  const double* x_coef_d;
  const double* x_coef_m;
};

struct ElemElasticityArgs {
  const double* node;
  const int* cell;
  int64_t NC;
  const double* M4;      // [L][L][NV][NV] = sum_q w R[q,i,k] R[q,j,l]
  double d_diag, d_lam, d_shear;
  int dof_priority;
  double* out;           // (NC, GD*L, GD*L)
};

int elem_const(int TD, int p, const ElemConstArgs& a, cudaStream_t s);
int elem_quad(int TD, int p, const ElemQuadArgs& a, cudaStream_t s);
// out[c] = (grad lambda_k[x] for k, x; signed measure): (TD+1)*TD + 1 doubles per cell
int cell_gradients(int TD, int64_t NC, const double* node, const int* cell, double* out, cudaStream_t s);
int elem_elasticity(int TD, int p, const ElemElasticityArgs& a, cudaStream_t s);
// gphi (NC, NQ, L, TD) = R (NQ, L, TD+1) . Dlambda, Dlambda from the records of cell_gradients
int grad_basis(int TD, int64_t NC, int NQ, int L, const double* rec, const double* R, double* out, cudaStream_t s);

}  // namespace fb2
