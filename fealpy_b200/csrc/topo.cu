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
#include <cstdlib>
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

// ---- faces of meshes with more than 2^21 nodes: the sorted vertex triple does not fit one 64-bit key --------------------
// (tet from_box 128^3 has 2 146 689 nodes: 3 x 22 bits.)  The stable LSD sort runs in two legs instead: by (b, d) with the
// original position as payload, then by a through that payload -- the order of the single wide key, first occurrences first.
__device__ __forceinline__ void sorted_triple(const int* __restrict__ cell, int NV, const LocalEnt& le, int64_t t, uint64_t& a,
                                              uint64_t& b, uint64_t& d) {
  const int64_t c = t / le.n;
  const int e = (int)(t - c * le.n);
  a = (uint32_t)cell[c * NV + le.v[e][0]];
  b = (uint32_t)cell[c * NV + le.v[e][1]];
  d = (uint32_t)cell[c * NV + le.v[e][2]];
  if (a > b) { uint64_t q = a; a = b; b = q; }
  if (b > d) { uint64_t q = b; b = d; d = q; }
  if (a > b) { uint64_t q = a; a = b; b = q; }
}

__global__ void __launch_bounds__(256) entity_keys_lo_kernel(const int* __restrict__ cell, int64_t NC, int NV, LocalEnt le, int vbits,
                                                             uint64_t* __restrict__ keys) {
  const int64_t n = NC * le.n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    uint64_t a, b, d;
    sorted_triple(cell, NV, le, t, a, b, d);
    keys[t] = (b << vbits) | d;
  }
}

__global__ void __launch_bounds__(256) entity_keys_hi_kernel(const int* __restrict__ cell, int NV, LocalEnt le,
                                                             const uint32_t* __restrict__ perm, int64_t n, uint64_t* __restrict__ keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t a, b, d;
    sorted_triple(cell, NV, le, perm[i], a, b, d);
    keys[i] = a;
  }
}

__global__ void __launch_bounds__(256) heads_wide_kernel(const int* __restrict__ cell, int NV, LocalEnt le,
                                                         const uint32_t* __restrict__ perm, int64_t n, uint8_t* __restrict__ head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint8_t h = 1;
    if (i > 0) {
      uint64_t a, b, d, a0, b0, d0;
      sorted_triple(cell, NV, le, perm[i], a, b, d);
      sorted_triple(cell, NV, le, perm[i - 1], a0, b0, d0);
      h = (a != a0 || b != b0 || d != d0) ? 1 : 0;
    }
    head[i] = h;
  }
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
  // faces of meshes with more than 2^21 nodes: two-leg sort (FB2_TOPO_FORCE_WIDE=1 takes that path on any mesh: tests)
  const char* fw = getenv("FB2_TOPO_FORCE_WIDE");
  const bool wide = vbits * le.nve > 63 || (le.nve == 3 && fw && fw[0] == '1');
  if (wide && (le.nve != 3 || 2 * vbits > 63))
    return fail(ERR_UNSUPPORTED, "build_entities: %d-vertex keys need %d bits (NN=%lld too large)", le.nve, vbits * le.nve, (long long)NN);
  Carver c(ws);
  uint64_t* keys = c.take<uint64_t>(n);
  uint32_t* perm = c.take<uint32_t>(n);
  void* sort_ws = c.take<char>(sort_workspace_bytes(n));
  uint8_t* head = c.take<uint8_t>(n);
  int64_t* S = c.take<int64_t>(n + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(n));
  uint64_t* ks = nullptr;
  uint32_t* ps = nullptr;
  if (!wide) {
    entity_keys_kernel<<<grid_for(n), 256, 0, s>>>(cell, NC, NV, le, vbits, keys);
    FB2_LAUNCH_CHECK();
    FB2_TRY(radix_sort_pairs(keys, perm, n, vbits * le.nve, sort_ws, s, &ks, &ps));
    if (ps != perm) FB2_CUDA(cudaMemcpyAsync(perm, ps, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    heads_kernel<<<grid_for(n), 256, 0, s>>>(ks, n, head);
  } else {
    entity_keys_lo_kernel<<<grid_for(n), 256, 0, s>>>(cell, NC, NV, le, vbits, keys);
    FB2_LAUNCH_CHECK();
    FB2_TRY(radix_sort_pairs(keys, perm, n, 2 * vbits, sort_ws, s, &ks, &ps));      // leg 1: by (b, d), payload = position
    if (ps != perm) FB2_CUDA(cudaMemcpyAsync(perm, ps, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    entity_keys_hi_kernel<<<grid_for(n), 256, 0, s>>>(cell, NV, le, perm, n, keys);
    FB2_LAUNCH_CHECK();
    ps = perm;                                                                      // leg 2: by a, payload carried (stable)
    FB2_TRY(radix_sort_pairs(keys, perm, n, vbits, sort_ws, s, &ks, &ps));
    if (ps != perm) FB2_CUDA(cudaMemcpyAsync(perm, ps, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    heads_wide_kernel<<<grid_for(n), 256, 0, s>>>(cell, NV, le, perm, n, head);
  }
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

// =====================================================================================
// x-slabs of TetrahedronMesh.from_box with the GLOBAL numbering in closed form (SURVEY.md
// appendix D): node id = i*(ny+1)(nz+1) + j*(nz+1) + k; every edge leaves its smaller node in
// one of 7 non-negative directions, ordered by node-id offset; edge id = number of edges
// leaving smaller nodes + rank of the direction among the ones that exist at that node.
// Used by the multi-GPU row partition: a rank generates only its slab of cube layers
// [cl0, cl1) and numbers DOFs in a "window" of the global numbering:
//   nodes of planes [cl0, cl1]             -> [0, nwin_nodes)
//   edges whose smaller node lies there    -> nwin_nodes + (global edge id - first such id)
// =====================================================================================
struct BoxDims { int nx, ny, nz; };

// number of edges leaving all nodes that precede node (i,j,k) in id order
__host__ __device__ inline int64_t box_edges_before(BoxDims d, int i, int j, int k) {
  const int64_t Sb = 2 * (int64_t)d.ny + 1, Sc = 2 * (int64_t)d.nz + 1;
  const int64_t plane_full = 2 * Sb * Sc - (int64_t)(d.ny + 1) * (d.nz + 1);   // a plane with i < nx
  int64_t cnt = (int64_t)i * plane_full;           // all planes before i have i' < nx
  const int64_t a1 = (i < d.nx) ? 2 : 1;
  cnt += (int64_t)j * (a1 * 2 * Sc - (d.nz + 1));   // rows j' < j (<= ny - 1 < ny): (1+b) = 2
  const int64_t b1 = (j < d.ny) ? 2 : 1;
  cnt += (int64_t)k * (a1 * b1 * 2 - 1);            // k' < k (< nz): (1+c) = 2
  return cnt;
}

// rank of direction (di,dj,dk) among the directions that exist at node (i,j,k)
__host__ __device__ inline int box_dir_rank(BoxDims d, int i, int j, int k, int di, int dj, int dk) {
  const bool a = i < d.nx, b = j < d.ny, c = k < d.nz;
  // directions in id-offset order: 001 010 011 100 101 110 111
  const bool ex[7] = {c, b, b && c, a, a && c, a && b, a && b && c};
  const int idx = di * 4 + dj * 2 + dk - 1;
  int r = 0;
  for (int q = 0; q < idx; ++q) r += ex[q] ? 1 : 0;
  return r;
}

__global__ void __launch_bounds__(256) tet_slab_kernel(double x0, double x1, double y0, double y1, double z0, double z1, BoxDims d,
                                                       int cl0, int cl1, int p, int64_t win_edge0, double* __restrict__ node,
                                                       int* __restrict__ cell, int* __restrict__ c2d) {
  const int64_t nyz = (int64_t)(d.ny + 1) * (d.nz + 1);
  const int nplanes = cl1 - cl0 + 1;
  const int64_t NNw = (int64_t)nplanes * nyz, NB = (int64_t)(cl1 - cl0) * d.ny * d.nz;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; node != nullptr && t < NNw; t += (int64_t)gridDim.x * blockDim.x) {
    const int il = (int)(t / nyz);
    const int rem = (int)(t - (int64_t)il * nyz);
    const int j = rem / (d.nz + 1), k = rem % (d.nz + 1);
    node[3 * t] = linspace_at(x0, x1, d.nx, cl0 + il);
    node[3 * t + 1] = linspace_at(y0, y1, d.ny, j);
    node[3 * t + 2] = linspace_at(z0, z1, d.nz, k);
  }
  constexpr int K[6][4] = {{0, 1, 2, 6}, {0, 5, 1, 6}, {0, 4, 5, 6}, {0, 7, 4, 6}, {0, 3, 7, 6}, {0, 2, 3, 6}};
  constexpr int CO[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  constexpr int LE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  constexpr int P2POS[6] = {1, 2, 3, 5, 6, 8};           // local dof of edge (a,b) in the P2 ordering
  constexpr int P2V[4] = {0, 4, 7, 9};
  const int L = p == 1 ? 4 : 10;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < NB * 6; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t cube = t / 6;
    const int q = (int)(t - cube * 6);
    const int il = (int)(cube / ((int64_t)d.ny * d.nz));
    const int rem = (int)(cube - (int64_t)il * d.ny * d.nz);
    const int j = rem / d.nz, k = rem % d.nz;
    int vi[4], vj[4], vk[4];
    int64_t vloc[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int cidx = K[q][v];
      vi[v] = il + CO[cidx][0]; vj[v] = j + CO[cidx][1]; vk[v] = k + CO[cidx][2];
      vloc[v] = (int64_t)vi[v] * nyz + (int64_t)vj[v] * (d.nz + 1) + vk[v];      // window-local node id
      if (cell) cell[4 * t + v] = (int)vloc[v];
    }
    if (p == 1) {
#pragma unroll
      for (int v = 0; v < 4; ++v) c2d[L * t + v] = (int)vloc[v];
    } else {
#pragma unroll
      for (int v = 0; v < 4; ++v) c2d[L * t + P2V[v]] = (int)vloc[v];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        int a = LE[e][0], b = LE[e][1];
        if (vloc[a] > vloc[b]) { const int s = a; a = b; b = s; }
        const int gi = cl0 + vi[a];                                              // global plane of the smaller node
        const int64_t eid = box_edges_before(d, gi, vj[a], vk[a]) +
                            box_dir_rank(d, gi, vj[a], vk[a], vi[b] - vi[a], vj[b] - vj[a], vk[b] - vk[a]);
        c2d[L * t + P2POS[e]] = (int)(NNw + (eid - win_edge0));
      }
    }
  }
}

int64_t box_edges_before_host(int nx, int ny, int nz, int i, int j, int k) { return box_edges_before(BoxDims{nx, ny, nz}, i, j, k); }

int tet_box_slab(const double* box, int nx, int ny, int nz, int cl0, int cl1, int p, double* node, int* cell, int* c2d,
                 cudaStream_t s) {
  if (p != 1 && p != 2) return fail(ERR_UNSUPPORTED, "tet_box_slab: closed-form numbering is available for p = 1, 2");
  if (cl0 < 0 || cl1 > nx || cl0 >= cl1) return fail(ERR_INVALID, "tet_box_slab: bad cube-layer range [%d, %d)", cl0, cl1);
  const BoxDims d{nx, ny, nz};
  const int64_t e0 = box_edges_before(d, cl0, 0, 0);
  const int64_t nwork = (int64_t)(cl1 - cl0) * ny * nz * 6;
  tet_slab_kernel<<<grid_for(nwork), 256, 0, s>>>(box[0], box[1], box[2], box[3], box[4], box[5], d, cl0, cl1, p, e0, node, cell, c2d);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2
