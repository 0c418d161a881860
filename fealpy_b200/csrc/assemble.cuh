#pragma once
#include "common.cuh"

namespace fb2 {

struct AsmConstArgs {
  const double* node;
  const int* cell;
  int64_t gdof, nnz;
  const int64_t* adj_ptr;     // (gdof+1)
  const int* adj_pair;        // (NC*L) pair ids c*L+i, ascending per dof
  const void* slots;          // (NC*L, L) position of the cell's j-th dof inside the row
  const int64_t* crow;        // (gdof+1)
  int has_diff, has_mass;
  const double* Ms;           // [L][L][NG]
  const double* Mm;           // [L][L]
  double scal_d, scal_m;
  const double* coef_d;       // per-cell or null
  const double* coef_m;
  double* values;             // (nnz)
  const int32_t* blk_row;     // (nblk+1) first row of every CTA tile (fb2_spmv_plan_build on crow)
  int nblk, tile, threads;    // tile = values per CTA the partition was built with
  int64_t NC;
};

struct AsmKeArgs {
  int64_t gdof;               // scalar dofs
  int64_t ncell;              // cells (the average number of (cell, i) pairs per row picks the gather's chunk size)
  int64_t nnz_out;
  int L, ncomp, dof_priority;
  const double* Ke;           // (NC, L*ncomp, L*ncomp)
  const int64_t* adj_ptr;
  const int* adj_pair;
  const void* slots;
  const int64_t* crow_s;      // scalar pattern
  const int64_t* crow_out;    // tensor pattern (== crow_s when ncomp == 1)
  double* values;
  const int32_t* blk_row;     // (nblk+1) tiling of the OUTPUT rows (crow_out)
  int nblk, tile, slot_stride;
  // fused P1 linear elasticity (Ke == nullptr): per-cell (grad lambda, measure) records instead of element matrices
  const double* geo;
  double d_diag, d_lam, d_shear, wsum;
};

int slot_stride(int L, int slot_bytes);
int assemble_elasticity_p1(int TD, AsmKeArgs a, int slot_bytes, int max_row, cudaStream_t s);
struct Asm4Args {
  const double* node;
  const int* cell;
  int64_t NC;
  const int64_t* crow;
  const int32_t* blk_row;       // (ntile+1) warp tiles (fb2_spmv_plan_build with the v4 tile size)
  int ntile, tile, max_row, acc_stride;
  const int64_t* batch_ptr;     // (ntile+1)
  const unsigned char* batch_i; // (nbatch) local index shared by the batch
  const uint32_t* blocks;       // (nbatch, 64 + 32*words + 4) packed batch blocks: 32 cell ids (-1 = padding) | 32 x (row offset in the
                                // tile (12 bits) | first-touch column mask) | 32 slot records | header word: local index i
  const double* Ms;             // device tables (null = term absent)
  const double* Mm;
  const double* Ms_host;        // host copies (go into the kernel parameter block)
  const double* Mm_host;
  double scal_d, scal_m;
  const double* coef_d;
  const double* coef_m;
  double* Hbuf;                 // (NC, HS) geometry workspace
  const double* H;
  double* values;
};
struct MatfreeArgs {
  const double* node;
  const int* cell;
  const int* c2d;
  int64_t NC;
  const double* Ms_host;        // host tables (null = term absent); folded into the kernel parameter block
  const double* Mm_host;
  double scal_d, scal_m;
  const double* coef_d;         // per-cell or null
  const double* coef_m;
  const double* u;              // (gdof)
  double* w;                    // (NC, L) per-cell products K_e u_e: pair order, or adjacency order when pair_pos is given
  const int* pair_pos;          // (NC*L) position of every (cell, i) pair in the adjacency lists, or null
};
int matfree_scalar_const(int TD, int p, const MatfreeArgs& a, cudaStream_t s);
int pair_positions(int64_t npos, const int* adj_pair, int* pair_pos, cudaStream_t s);
int segment_sum(int64_t gdof, const int64_t* adj_ptr, const double* w, double* F, cudaStream_t s);
size_t adjacency_workspace_bytes(int64_t gdof);
int build_adjacency(const int* c2d, int64_t NC, int L, int64_t gdof, int64_t* adj_ptr, int* adj_pair, void* ws, cudaStream_t s);
size_t asm4_workspace_bytes(int ntile);
int asm4_plan_count(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int* adj_pair, int L,
                    int64_t* batch_ptr, int64_t* nbatch_host, void* ws, cudaStream_t s);
int asm4_plan_fill(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int* adj_pair, int L,
                   const int64_t* batch_ptr, unsigned char* batch_i, uint32_t* blocks, const void* slots, int slot_bytes,
                   cudaStream_t s);
int assemble_v4(int TD, int p, const Asm4Args& a, int slot_bytes, cudaStream_t s);
size_t sym_workspace_bytes(int64_t NC, int L, int64_t gdof);
int sym_count(const int* c2d, int64_t NC, int L, int64_t gdof, int64_t* adj_ptr, int* adj_pair, int64_t* crow, int64_t* nnz_host,
              int* max_row_host, uint16_t* stash, void* ws, cudaStream_t s);
int sym_fill(const int* c2d, int64_t NC, int L, int64_t gdof, const int64_t* adj_ptr, const int* adj_pair, const int64_t* crow,
             int* col, void* slots, int slot_bytes, const uint16_t* stash, cudaStream_t s);
int assemble_const(int TD, int p, const AsmConstArgs& a, int slot_bytes, int max_row, cudaStream_t s);
int assemble_from_ke(AsmKeArgs a, int slot_bytes, int max_row, cudaStream_t s);
int expand_pattern(int64_t gdof, int nc, int prio, const int64_t* crow_s, const int* col_s, int64_t* crow_out, int* col_out,
                   cudaStream_t s);

}  // namespace fb2
