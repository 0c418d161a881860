#pragma once
#include "common.cuh"

namespace fb2 {
int coo_keys_from_c2d(const int* rdof, const int* cdof, int64_t NC, int Lr, int Lc, int cbits, uint64_t* keys, cudaStream_t s);
int coo_keys_from_coo(const void* row, const void* col, int index_bytes, int64_t n, int cbits, uint64_t* keys, cudaStream_t s);
size_t coo_symbolic_workspace_bytes(int64_t n);
int coo_symbolic(uint64_t* keys, uint32_t* perm, int64_t n, int nbits, void* ws, int64_t* nnz_host, cudaStream_t s);
int coo_fill(const uint64_t* keys, int64_t n, int cbits, int64_t nrow, void* ws, int64_t* crow, void* col, int col_bytes,
             int64_t* seg_start, cudaStream_t s);
int coo_reduce(const uint32_t* perm, const int64_t* seg_start, int64_t nnz, const double* vin, double* vout, cudaStream_t s);
}  // namespace fb2
