#pragma once
#include "common.cuh"

namespace fb2 {
// physical points of NQ barycentric points in every cell: out (NC, NQ, TD) (mesh/mesh_base.py:454-478, backend/numpy_backend.py:401-407)
int bc_to_points(int TD, int64_t NC, int NQ, const double* node, const int* cell, const double* bcs, double* out, cudaStream_t s);
int elem_source(int TD, int64_t NC, int L, int NQ, const double* node, const int* cell, const double* phiw, int kind, double scal,
                const double* f, double* out, cudaStream_t s);
int gather_vector(int64_t gdof, const int64_t* adj_ptr, const int* adj_pair, const double* fe, double* F, cudaStream_t s);
size_t bc_workspace_bytes(int64_t n);
int bc_matrix_count(int64_t n, const int64_t* crow, const int* col, const uint8_t* isbd, int64_t* crow_new, int64_t* nnz_host, void* ws,
                    cudaStream_t s);
int bc_matrix_fill(int64_t n, const int64_t* crow, const int* col, const double* val, const uint8_t* isbd, const int64_t* crow_new,
                   int* col_new, double* val_new, cudaStream_t s);
int bc_vector(int64_t n, const uint8_t* isbd, const double* uh, double* f, cudaStream_t s);
int matfree_apply(int64_t gdof, int L, const int64_t* adj_ptr, const int* adj_pair, const int* c2d, const double* ke, const double* u,
                  double* v, cudaStream_t s);
}  // namespace fb2
