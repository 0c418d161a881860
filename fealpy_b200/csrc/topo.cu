// Mesh generation and topology -> DOF numbering on the device (integer-only work).
//
// The global DOF numbering must reproduce the reference's exactly (it defines the CSR
// pattern):   from_box            mesh/triangle_mesh.py:1386-1435, tetrahedron_mesh.py:1016-1086
//             unique edges/faces  mesh/mesh_data_structure.py:428-464 + mesh/utils.py:81-110 (flocc):
//                                 id = rank of the sorted vertex tuple in lexicographic order,
//                                 stored orientation = first occurrence in (cell, local entity) order
//             cell_to_ipoint      mesh/triangle_mesh.py:218-270, tetrahedron_mesh.py:388-441,
//                                 edge_to_ipoint mesh/mesh_base.py:188-210
// Here: sorted-tuple keys -> stable radix sort (payload = position) -> head flags -> scan.
#include "common.cuh"
#include "sort_scan.cuh"
#include "topo.cuh"

namespace fb2 {

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSM * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// numpy.linspace: start + i*step (two roundings, no fma), last point = stop exactly
__device__ __forceinline__ double linspace_at(double a, double b, int n, int i) {
  if (i == n) return b;
  const double step = (b - a) / (double)n;
  return __dadd_rn(__dmul_rn((double)i, step), a);
}

__global__ void tri_box_kernel(double x0, double x1, double y0, double y1, int nx, int ny, double* __restrict__ node,
                               int* __restrict__ cell) {
  const int64_t NN = (int64_t)(nx + 1) * (ny + 1), NQ = (int64_t)nx * ny;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < NN; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / (ny + 1)), j = (int)(t % (ny + 1));
    node[2 * t] = linspace_at(x0, x1, nx, i);
    node[2 * t + 1] = linspace_at(y0, y1, ny, j);
  }
  // squares enumerated with j slowest (the reference transposes before flattening)
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < NQ; t += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(t / nx), i = (int)(t % nx);
    const int n00 = i * (ny + 1) + j, n10 = n00 + (ny + 1), n01 = n00 + 1, n11 = n10 + 1;
    int* lo = cell + 3 * t;
    lo[0] = n10; lo[1] = n11; lo[2] = n00;
    int* up = cell + 3 * (NQ + t);
    up[0] = n01; up[1] = n00; up[2] = n11;
  }
}

__global__ void tet_box_kernel(double x0, double x1, double y0, double y1, double z0, double z1, int nx, int ny, int nz,
                               double* __restrict__ node, int* __restrict__ cell) {
  const int64_t nyz = (int64_t)(ny + 1) * (nz + 1);
  const int64_t NN = (int64_t)(nx + 1) * nyz, NB = (int64_t)nx * ny * nz;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < NN; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / nyz);
    const int rem = (int)(t - (int64_t)i * nyz);
    const int j = rem / (nz + 1), k = rem % (nz + 1);
    node[3 * t] = linspace_at(x0, x1, nx, i);
    node[3 * t + 1] = linspace_at(y0, y1, ny, j);
    node[3 * t + 2] = linspace_at(z0, z1, nz, k);
  }
  // 6 Kuhn tets per cube: corners 0..7 = (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1)
  constexpr int K[6][4] = {{0, 1, 2, 6}, {0, 5, 1, 6}, {0, 4, 5, 6}, {0, 7, 4, 6}, {0, 3, 7, 6}, {0, 2, 3, 6}};
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < NB; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / ((int64_t)ny * nz));
    const int rem = (int)(t - (int64_t)i * ny * nz);
    const int j = rem / nz, k = rem % nz;
    const int64_t c0 = (int64_t)i * nyz + (int64_t)j * (nz + 1) + k;
    const int64_t off[8] = {0, nyz, nyz + nz + 1, nz + 1, 1, nyz + 1, nyz + nz + 2, nz + 2};
    int4* out = reinterpret_cast<int4*>(cell + 24 * t);
#pragma unroll
    for (int q = 0; q < 6; ++q)
      out[q] = make_int4((int)(c0 + off[K[q][0]]), (int)(c0 + off[K[q][1]]), (int)(c0 + off[K[q][2]]), (int)(c0 + off[K[q][3]]));
  }
}

int tri_from_box(const double* box, int nx, int ny, double* node, int* cell, cudaStream_t s) {
  const int64_t NN = (int64_t)(nx + 1) * (ny + 1);
  tri_box_kernel<<<grid_for(NN), 256, 0, s>>>(box[0], box[1], box[2], box[3], nx, ny, node, cell);
  FB2_LAUNCH_CHECK();
  return OK;
}
int tet_from_box(const double* box, int nx, int ny, int nz, double* node, int* cell, cudaStream_t s) {
  const int64_t NN = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
  tet_box_kernel<<<grid_for(NN), 256, 0, s>>>(box[0], box[1], box[2], box[3], box[4], box[5], nx, ny, nz, node, cell);
  FB2_LAUNCH_CHECK();
  return OK;
}

// ---- unique sub-entities (edges: NVE=2, faces of tets: NVE=3) ----------------------------
struct LocalEnt {
  int n;            // local entities per cell
  int nve;          // vertices per entity
  int v[6][3];
};

__global__ void __launch_bounds__(256) entity_keys_kernel(const int* __restrict__ cell, int64_t NC, int NV, LocalEnt le, int vbits,
                                                          uint64_t* __restrict__ keys) {
  const int64_t n = NC * le.n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / le.n;
    const int e = (int)(t - c * le.n);
    uint64_t a = (uint32_t)cell[c * NV + le.v[e][0]], b = (uint32_t)cell[c * NV + le.v[e][1]];
    if (a > b) { uint64_t q = a; a = b; b = q; }
    if (le.nve == 2) {
      keys[t] = (a << vbits) | b;
    } else {
      uint64_t d = (uint32_t)cell[c * NV + le.v[e][2]];
      if (b > d) { uint64_t q = b; b = d; d = q; }
      if (a > b) { uint64_t q = a; a = b; b = q; }
      keys[t] = (a << (2 * vbits)) | (b << vbits) | d;
    }
  }
}

__global__ void __launch_bounds__(256) heads_kernel(const uint64_t* __restrict__ keys, int64_t n, uint8_t* __restrict__ head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// cell2ent[perm[i]] = id(i);  ent[id] = vertices of the first occurrence (original orientation)
__global__ void __launch_bounds__(256) entity_fill_kernel(const int* __restrict__ cell, int NV, LocalEnt le,
                                                          const uint32_t* __restrict__ perm, const uint8_t* __restrict__ head,
                                                          const int64_t* __restrict__ S, int64_t n, int* __restrict__ cell2ent,
                                                          int* __restrict__ ent /* may be null on the count pass */) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = S[i] + head[i] - 1;
    const uint32_t src = perm[i];
    cell2ent[src] = (int)id;
    if (head[i] && ent) {
      const int64_t c = src / le.n;
      const int e = (int)(src - c * le.n);
      for (int k = 0; k < le.nve; ++k) ent[id * le.nve + k] = cell[c * NV + le.v[e][k]];
    }
  }
}

static LocalEnt local_entities(int TD, int kind /*1 edge, 2 face*/) {
  LocalEnt le{};
  if (kind == 1 && TD == 2) {
    le.n = 3; le.nve = 2;
    const int v[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    for (int i = 0; i < 3; ++i) { le.v[i][0] = v[i][0]; le.v[i][1] = v[i][1]; }
  } else if (kind == 1 && TD == 3) {
    le.n = 6; le.nve = 2;
    const int v[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    for (int i = 0; i < 6; ++i) { le.v[i][0] = v[i][0]; le.v[i][1] = v[i][1]; }
  } else {
    le.n = 4; le.nve = 3;
    const int v[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) le.v[i][k] = v[i][k];
  }
  return le;
}

static int bits_for(int64_t n) { int b = 1; while (((int64_t)1 << b) < n) ++b; return b; }

size_t entity_workspace_bytes(int64_t NC, int per_cell) {
  const int64_t n = NC * per_cell;
  return align_up((size_t)n * 8) + align_up((size_t)n * 4) + sort_workspace_bytes(n) + align_up((size_t)n) +
         align_up((size_t)(n + 1) * 8) + scan_workspace_bytes(n) + 1024;
}

// step 1: sort/unique; fills cell2ent, returns the entity count, leaves its state in `ws`.
// step 2 (entities_emit, same ws): writes the entity vertex table once the caller has sized it.
int build_entities(const int* cell, int64_t NC, int TD, int kind, int64_t NN, int* cell2ent, int64_t* count_host, void* ws,
                   cudaStream_t s) {
  const LocalEnt le = local_entities(TD, kind);
  const int NV = TD + 1;
  const int64_t n = NC * le.n;
  if (n <= 0) { *count_host = 0; return OK; }
  const int vbits = bits_for(NN);
  if (vbits * le.nve > 63) return fail(ERR_UNSUPPORTED, "build_entities: %d-vertex keys need %d bits (NN=%lld too large)", le.nve, vbits * le.nve, (long long)NN);
  Carver c(ws);
  uint64_t* keys = c.take<uint64_t>(n);
  uint32_t* perm = c.take<uint32_t>(n);
  void* sort_ws = c.take<char>(sort_workspace_bytes(n));
  uint8_t* head = c.take<uint8_t>(n);
  int64_t* S = c.take<int64_t>(n + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(n));
  entity_keys_kernel<<<grid_for(n), 256, 0, s>>>(cell, NC, NV, le, vbits, keys);
  FB2_LAUNCH_CHECK();
  uint64_t* ks = nullptr;
  uint32_t* ps = nullptr;
  FB2_TRY(radix_sort_pairs(keys, perm, n, vbits * le.nve, sort_ws, s, &ks, &ps));
  if (ps != perm) FB2_CUDA(cudaMemcpyAsync(perm, ps, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  heads_kernel<<<grid_for(n), 256, 0, s>>>(ks, n, head);
  FB2_LAUNCH_CHECK();
  FB2_TRY(exclusive_scan_u8(head, S, n, true, scan_ws, s));
  entity_fill_kernel<<<grid_for(n), 256, 0, s>>>(cell, NV, le, perm, head, S, n, cell2ent, nullptr);
  FB2_LAUNCH_CHECK();
  FB2_CUDA(cudaMemcpyAsync(count_host, S + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  return OK;
}

int entities_emit(const int* cell, int64_t NC, int TD, int kind, int* cell2ent, int* ent, void* ws, cudaStream_t s) {
  const LocalEnt le = local_entities(TD, kind);
  const int64_t n = NC * le.n;
  if (n <= 0) return OK;
  Carver c(ws);
  c.take<uint64_t>(n);
  uint32_t* perm = c.take<uint32_t>(n);
  c.take<char>(sort_workspace_bytes(n));
  uint8_t* head = c.take<uint8_t>(n);
  int64_t* S = c.take<int64_t>(n + 1);
  entity_fill_kernel<<<grid_for(n), 256, 0, s>>>(cell, TD + 1, le, perm, head, S, n, cell2ent, ent);
  FB2_LAUNCH_CHECK();
  return OK;
}

// ---- cell -> dof (p <= 3) ----------------------------------------------------------------
struct MiTable { unsigned char a[20][4]; int L; };

__global__ void __launch_bounds__(256) cell_to_dof_kernel(const int* __restrict__ cell, const int* __restrict__ cell2edge,
                                                          const int* __restrict__ edge, const int* __restrict__ cell2face, int64_t NC,
                                                          int TD, int p, int64_t NN, int64_t NE, int64_t NF, MiTable mi,
                                                          int* __restrict__ c2d) {
  const int NV = TD + 1, L = mi.L, NEC = TD == 2 ? 3 : 6;
  const int64_t n = NC * L;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / L;
    const int i = (int)(t - c * L);
    int nz[4], cnt = 0;
    for (int b = 0; b < NV; ++b) if (mi.a[i][b]) nz[cnt++] = b;
    int64_t dof;
    if (cnt == 1) {
      dof = cell[c * NV + nz[0]];
    } else if (cnt == 2) {
      const int va = nz[0], vb = nz[1];
      int le;
      if (TD == 2) le = 3 - va - vb;                          // (0,1)->2 (0,2)->1 (1,2)->0
      else le = va == 0 ? vb - 1 : (va == 1 ? vb + 1 : 5);    // (0,1)0 (0,2)1 (0,3)2 (1,2)3 (1,3)4 (2,3)5
      const int64_t ge = cell2edge[c * NEC + le];
      const int aa = mi.a[i][va];
      const int tpos = (edge[2 * ge] == cell[c * NV + va]) ? p - aa : aa;
      dof = NN + (int64_t)(p - 1) * ge + (tpos - 1);
    } else if (cnt == 3 && TD == 3) {
      const int lf = 6 - nz[0] - nz[1] - nz[2];               // the missing vertex = local face id
      dof = NN + (int64_t)(p - 1) * NE + cell2face[c * 4 + lf];   // p == 3: one interior point per face
    } else {
      dof = NN + (int64_t)(p - 1) * NE + c;                   // triangle p == 3 cell interior
    }
    c2d[t] = (int)dof;
  }
}

int cell_to_dof(const int* cell, const int* cell2edge, const int* edge, const int* cell2face, int64_t NC, int TD, int p, int64_t NN,
                int64_t NE, int64_t NF, const unsigned char* mi_host, int L, int* c2d, cudaStream_t s) {
  if (p < 1 || p > 3) return fail(ERR_UNSUPPORTED, "cell_to_dof: p=%d not supported (1..3)", p);
  if (L > 20) return fail(ERR_INVALID, "cell_to_dof: L too large");
  MiTable mi{};
  mi.L = L;
  for (int i = 0; i < L; ++i) for (int b = 0; b <= TD; ++b) mi.a[i][b] = mi_host[i * (TD + 1) + b];
  if (NC <= 0) return OK;
  cell_to_dof_kernel<<<grid_for(NC * L), 256, 0, s>>>(cell, cell2edge, edge, cell2face, NC, TD, p, NN, NE, NF, mi, c2d);
  FB2_LAUNCH_CHECK();
  return OK;
}

// tensor-space dof map (functionspace/utils.py:83-95)
__global__ void __launch_bounds__(256) tensor_dof_kernel(const int* __restrict__ c2d, int64_t NC, int L, int GD, int64_t gdof, int prio,
                                                         int* __restrict__ out) {
  const int64_t n = NC * L * GD;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = t / (L * GD);
    const int rem = (int)(t - c * L * GD);
    int i, a;
    if (prio) { a = rem / L; i = rem - a * L; } else { i = rem / GD; a = rem - i * GD; }
    const int64_t d = c2d[c * L + i];
    out[t] = (int)(prio ? (int64_t)a * gdof + d : d * GD + a);
  }
}

int tensor_cell_to_dof(const int* c2d, int64_t NC, int L, int GD, int64_t gdof, int prio, int* out, cudaStream_t s) {
  if (NC <= 0) return OK;
  tensor_dof_kernel<<<grid_for(NC * L * GD), 256, 0, s>>>(c2d, NC, L, GD, gdof, prio, out);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2
