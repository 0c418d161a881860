#pragma once
#include "common.cuh"

namespace fb2 {
int tri_from_box(const double* box, int nx, int ny, double* node, int* cell, cudaStream_t s);
int tet_from_box(const double* box, int nx, int ny, int nz, double* node, int* cell, cudaStream_t s);
size_t entity_workspace_bytes(int64_t NC, int per_cell);
int build_entities(const int* cell, int64_t NC, int TD, int kind, int64_t NN, int* cell2ent, int64_t* count_host, void* ws,
                   cudaStream_t s);
int entities_emit(const int* cell, int64_t NC, int TD, int kind, int* cell2ent, int* ent, void* ws, cudaStream_t s);
int cell_to_dof(const int* cell, const int* cell2edge, const int* edge, const int* cell2face, int64_t NC, int TD, int p, int64_t NN,
                int64_t NE, int64_t NF, const unsigned char* mi_host, int L, int* c2d, cudaStream_t s);
int tensor_cell_to_dof(const int* c2d, int64_t NC, int L, int GD, int64_t gdof, int prio, int* out, cudaStream_t s);
int64_t box_edges_before_host(int nx, int ny, int nz, int i, int j, int k);
int tet_box_slab(const double* box, int nx, int ny, int nz, int cl0, int cl1, int p, double* node, int* cell, int* c2d,
                 cudaStream_t s);
}  // namespace fb2
