// Symbolic + fused numeric assembly: CSR straight from cell2dof, no COO.
//
// The CSR pattern of a BilinearForm is a pure function of cell2dof (row = cell2dof[c,i],
// col = cell2dof[c,j], fem/bilinear_form.py:69-72; kept with explicit zeros by
// COOTensor.coalesce, sparse/coo_tensor.py:184-213).  So:
//   symbolic (once per space, cached):  dof -> incident (cell, local index) list in ascending
//       cell order; per row the sorted set of columns; per (row, cell) the position ("slot")
//       of each of the cell's dofs inside that row.
//   numeric (every assembly): one thread owns one CSR row, walks its incident cells in
//       ascending cell order -- the order the reference's stable sort + np.add.at produces --
//       recomputes the element-matrix row in registers and accumulates into its private
//       segment of a shared-memory tile, which the CTA then streams out contiguously.
// Deterministic by construction: no floating-point atomics, fixed summation order.
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "sort_scan.cuh"
#include "assemble.cuh"

namespace fb2 {

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSM * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// exact division by a small run-time constant (ldof) without the ~25-instruction integer-divide sequence:
// floor(n / d) = mulhi(n, floor(2^w / d) + 1) for n < 2^16 (w = 32) resp. n < 2^32 (w = 64), d >= 2
struct FastDiv {
  uint32_t m32;
  uint64_t m64;
  int d;
  __device__ __forceinline__ explicit FastDiv(int dd) : m32(0xffffffffu / (uint32_t)dd + 1u), m64(~0ull / (uint64_t)dd + 1ull), d(dd) {}
  __device__ __forceinline__ int small(int n) const { return d == 1 ? n : (int)__umulhi((uint32_t)n, m32); }        // 0 <= n < 65536
  __device__ __forceinline__ int wide(int n) const { return d == 1 ? n : (int)__umul64hi((uint64_t)(uint32_t)n, m64); }   // 0 <= n < 2^31
};

// =====================================================================================
// symbolic
// =====================================================================================
__global__ void __launch_bounds__(256) deg_kernel(const int* __restrict__ c2d, int64_t npair, int* __restrict__ deg) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npair; t += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(deg + c2d[t], 1);
}

__global__ void __launch_bounds__(256) adj_fill_kernel(const int* __restrict__ c2d, int64_t npair, const int64_t* __restrict__ adj_ptr,
                                                       int* __restrict__ cursor, int* __restrict__ adj_pair) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npair; t += (int64_t)gridDim.x * blockDim.x) {
    const int d = c2d[t];
    const int k = atomicAdd(cursor + d, 1);
    adj_pair[adj_ptr[d] + k] = (int)t;        // pair id = c*L + i
  }
}

// per-dof lists are short: in-place insertion sort into the canonical order
// (local index i, then cell) -- grouping by i makes the element-table row warp-uniform in the
// numeric kernel; within a group cells ascend.  The order is a pure function of cell2dof.
__global__ void __launch_bounds__(256) adj_sort_kernel(int64_t gdof, int L, const int64_t* __restrict__ adj_ptr, int* __restrict__ adj_pair) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < gdof; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = adj_ptr[r], e = adj_ptr[r + 1];
    for (int64_t i = b + 1; i < e; ++i) {
      const int v = adj_pair[i];
      const int64_t kv = (int64_t)(v % L) * ((int64_t)1 << 32) + v / L;
      int64_t j = i - 1;
      while (j >= b) {
        const int w = adj_pair[j];
        const int64_t kw = (int64_t)(w % L) * ((int64_t)1 << 32) + w / L;
        if (kw <= kv) break;
        adj_pair[j + 1] = w;
        --j;
      }
      adj_pair[j + 1] = v;
    }
  }
}

constexpr int SYM_CAP = 1024;       // max candidate columns (valence*ldof) per row handled in shared memory
constexpr int SYM_KBITS = 10;
constexpr int SYM_WARPS = 4;

// bitonic sort of 32*K keys held K per lane (element e = k*32 + lane): strides below 32 are lane exchanges
// (shuffles), strides of 32 and more are register swaps -- no shared memory, no barriers
template <int K>
__device__ __forceinline__ void bitonic_sort_regs(uint64_t (&v)[K], int lane) {
  constexpr int N = 32 * K;
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int ks = stride >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if ((k & ks) == 0) {
            const bool up = (((k << 5) + lane) & size) == 0;
            const uint64_t x = v[k], y = v[k | ks];
            if ((x > y) == up) { v[k] = y; v[k | ks] = x; }
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const bool up = (((k << 5) + lane) & size) == 0;
          const uint64_t x = v[k], y = __shfl_xor_sync(0xffffffffu, x, stride);
          const bool lower = (lane & stride) == 0;
          v[k] = (lower == up) ? (x < y ? x : y) : (x < y ? y : x);
        }
      }
    }
  }
}

// ---- rows of up to SYM_HASH_CAP candidates: deduplicate first, sort only the unique columns --------
// A row's candidates are heavily duplicated (tet P2: 50-240 candidates, 28-80 distinct columns), and the
// sort is what the symbolic phase spends its instructions on.  So: (1) every candidate is inserted
// into a small open-addressing hash set in shared memory (atomicCAS; the inserting candidate is the
// column's representative), (2) the representatives' (column, table slot) keys are compacted and
// sorted in registers, (3) the sorted position = rank is written back to the table slot, where every
// candidate of that column picks it up.
constexpr int SYM_HASH_CAP = 256;           // candidates per row on this path
constexpr int SYM_HASH_SIZE = 512;          // table entries (load factor <= 0.5)
constexpr uint32_t SYM_EMPTY = 0xffffffffu;

template <int K>
__device__ __forceinline__ void sym_sort_unique(const uint64_t* __restrict__ ulist, int nu, int lane, uint32_t* __restrict__ trank,
                                                int* __restrict__ col_out) {
  uint64_t v[K];
#pragma unroll
  for (int it = 0; it < K; ++it) { const int p = it * 32 + lane; v[it] = p < nu ? ulist[p] : ~0ull; }
  bitonic_sort_regs<K>(v, lane);
#pragma unroll
  for (int it = 0; it < K; ++it) {
    const int p = it * 32 + lane;
    if (p < nu) {
      trank[(int)(v[it] & (SYM_HASH_SIZE - 1))] = (uint32_t)p;
      if (col_out) col_out[p] = (int)(v[it] >> SYM_KBITS);
    }
  }
}

// Rank by counting, for rows of up to 128 distinct columns (all of tet P2: 29 on edge dofs, 65 on vertex dofs).  The
// distinct columns of a row are few and unsorted; the rank of one is the number of smaller ones.  Every column is
// broadcast from shared memory once and compared with the K columns a lane owns: nu * (1 + 2K) instructions on 32-bit
// keys, against a bitonic network on 64-bit keys that had to be padded to a power of two (65 columns sorted as 128:
// ncu of the sorting version: 720 warp instructions per row, 57 % of them ISETP / SEL / IMAD of the network).
template <int K>
__device__ __forceinline__ void sym_rank_count(const uint64_t* __restrict__ ulist, int nu, int lane, uint32_t* __restrict__ trank,
                                               int* __restrict__ col_out) {
  uint32_t mine[K];
  int rk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int p = k * 32 + lane;
    mine[k] = p < nu ? (uint32_t)(ulist[p] >> SYM_KBITS) : 0xffffffffu;
    rk[k] = 0;
  }
  for (int q = 0; q < nu; ++q) {
    const uint32_t v = (uint32_t)(ulist[q] >> SYM_KBITS);
#pragma unroll
    for (int k = 0; k < K; ++k) rk[k] += (v < mine[k]) ? 1 : 0;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int p = k * 32 + lane;
    if (p < nu) {
      trank[(int)(ulist[p] & (SYM_HASH_SIZE - 1))] = (uint32_t)rk[k];
      if (col_out) col_out[rk[k]] = (int)mine[k];
    }
  }
}

template <bool FILL, typename SlotT>
__global__ void __launch_bounds__(SYM_WARPS * 32) sym_rows_kernel(const int* __restrict__ c2d, int L, int64_t gdof,
                                                                  const int64_t* __restrict__ adj_ptr, const int* __restrict__ adj_pair,
                                                                  const int64_t* __restrict__ crow, int* __restrict__ rowlen,
                                                                  int* __restrict__ col, SlotT* __restrict__ slots, int slot_stride,
                                                                  int* __restrict__ err, uint16_t* __restrict__ stash) {
  __shared__ uint64_t buf_all[SYM_WARPS][SYM_CAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t* buf = buf_all[wid];
  const int64_t nwarp = (int64_t)gridDim.x * SYM_WARPS;
  const uint32_t lt = (1u << lane) - 1u;
  const FastDiv fd(L);
  for (int t = lane; t < SYM_HASH_SIZE; t += 32) reinterpret_cast<uint32_t*>(buf)[t] = SYM_EMPTY;
  __syncwarp();
  for (int64_t r = (int64_t)blockIdx.x * SYM_WARPS + wid; r < gdof; r += nwarp) {
    const int64_t a0 = adj_ptr[r];
    const int deg = (int)(adj_ptr[r + 1] - a0);
    const int ncand = deg * L;
    if (ncand > SYM_CAP) {
      if (lane == 0) atomicExch(err, 1);
      if (!FILL && lane == 0) rowlen[r] = 0;
      continue;
    }
    if (ncand <= SYM_HASH_CAP) {
      // shared-memory carve-up of this warp's 8 KB: hash table | rank per table slot | unique keys
      uint32_t* tab = reinterpret_cast<uint32_t*>(buf);
      uint32_t* trank = tab + SYM_HASH_SIZE;
      uint64_t* ulist = buf + SYM_HASH_SIZE;                     // (2 * SYM_HASH_SIZE u32 = SYM_HASH_SIZE u64 in)
      constexpr int KMAX = SYM_HASH_CAP / 32;
      int hs[KMAX];                                              // per candidate: table slot << 1 | representative
      int nu = 0;
#pragma unroll
      for (int it = 0; it < KMAX; ++it) {
        hs[it] = 0;
        if (it * 32 < ncand) {                                   // warp-uniform
          const int k = it * 32 + lane;
          bool rep = false;
          uint32_t cv = 0;
          int h = 0;
          if (k < ncand) {
            const int pl = fd.small(k), j = k - pl * L;
            cv = (uint32_t)c2d[(int64_t)fd.wide(adj_pair[a0 + pl]) * L + j];
            h = (int)((cv * 2654435761u) >> 23);                 // 9 bits
            while (true) {
              const uint32_t prev = atomicCAS(tab + h, SYM_EMPTY, cv);
              if (prev == SYM_EMPTY) { rep = true; break; }
              if (prev == cv) break;
              h = (h + 1) & (SYM_HASH_SIZE - 1);
            }
            hs[it] = (h << 1) | (rep ? 1 : 0);
          }
          const uint32_t rb = __ballot_sync(0xffffffffu, rep);
          if (rep) ulist[nu + __popc(rb & lt)] = ((uint64_t)cv << SYM_KBITS) | (uint64_t)h;
          nu += __popc(rb);
        }
      }
      __syncwarp();
      const int64_t cbase = FILL ? crow[r] : 0;
      int* cout = FILL ? col + cbase : nullptr;
      if (nu <= 32) sym_rank_count<1>(ulist, nu, lane, trank, cout);
      else if (nu <= 64) sym_rank_count<2>(ulist, nu, lane, trank, cout);
      else if (nu <= 96) sym_rank_count<3>(ulist, nu, lane, trank, cout);
      else if (nu <= 128) sym_rank_count<4>(ulist, nu, lane, trank, cout);
      else sym_sort_unique<8>(ulist, nu, lane, trank, cout);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < KMAX; ++it) {
        const int k = it * 32 + lane;
        if (k < ncand) {
          const int h = hs[it] >> 1, rank = (int)trank[h];
          if (FILL) {
            const int pl = fd.small(k);
            slots[(a0 + pl) * slot_stride + (k - pl * L)] = (SlotT)rank;
          } else if (stash) {
            stash[a0 * L + k] = (uint16_t)((rank << 1) | (hs[it] & 1));
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < KMAX; ++it)
        if ((hs[it] & 1) && it * 32 + lane < ncand) tab[hs[it] >> 1] = SYM_EMPTY;      // leave the table empty for the next row
      if (!FILL && lane == 0) rowlen[r] = nu;
      __syncwarp();
      continue;
    }
    // ---- long rows: bitonic sort of all (column, candidate) keys in shared memory -----------
    int n2 = 32;
    while (n2 < ncand) n2 <<= 1;
    {
      for (int k = lane; k < n2; k += 32) {
        uint64_t key = ~0ull;
        if (k < ncand) {
          const int pl = fd.small(k), j = k - pl * L;
          const int pair = adj_pair[a0 + pl];
          const int64_t c = fd.wide(pair);
          key = ((uint64_t)(uint32_t)c2d[c * L + j] << SYM_KBITS) | (uint64_t)k;
        }
        buf[k] = key;
      }
      __syncwarp();
      for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = lane; t < (n2 >> 1); t += 32) {
            const int i = 2 * t - (t & (stride - 1));
            const int j = i + stride;
            const bool up = (i & size) == 0;
            const uint64_t x = buf[i], y = buf[j];
            if ((x > y) == up) { buf[i] = y; buf[j] = x; }
          }
          __syncwarp();
        }
      }
    }
    int running = 0;
    const int64_t cbase = FILL ? crow[r] : 0;
    for (int k0 = 0; k0 < ncand; k0 += 32) {
      const int k = k0 + lane;
      bool head = false;
      uint64_t key = 0;
      if (k < ncand) {
        key = buf[k];
        head = (k == 0) || ((key >> SYM_KBITS) != (buf[k - 1] >> SYM_KBITS));
      }
      const uint32_t hb = __ballot_sync(0xffffffffu, head);
      if (k < ncand) {
        const int rank = running + __popc(hb & lt) + (head ? 1 : 0) - 1;
        const int kk = (int)(key & ((1u << SYM_KBITS) - 1u));      // candidate index = pair_local * L + j
        if (FILL) {
          if (head) col[cbase + rank] = (int)(key >> SYM_KBITS);
          const int pl = fd.small(kk);
          slots[(a0 + pl) * slot_stride + (kk - pl * L)] = (SlotT)rank;
        } else if (stash) {                        // rank and representative flag per candidate for sym_replay_flat_kernel
          stash[a0 * L + kk] = (uint16_t)((rank << 1) | (head ? 1 : 0));
        }
      }
      running += __popc(hb);
    }
    if (!FILL && lane == 0) rowlen[r] = running;
    __syncwarp();
    for (int t = lane; t < SYM_HASH_SIZE; t += 32) reinterpret_cast<uint32_t*>(buf)[t] = SYM_EMPTY;      // the sort used the table's memory
    __syncwarp();
  }
}

// fill pass from the stash of the count pass: per candidate (rank << 1 | representative) -- no ranking again; the column
// of a representative is re-gathered from cell2dof.  Flat over the adjacency positions: position q = (row r, pair) knows
// its row without a search (r = cell2dof[pair]), so one THREAD replays the ldof candidates of one position -- contiguous
// 2*ldof-byte stash reads and ldof-byte slot writes per thread (a warp-per-row version measured 0.8 ms slower)
template <typename SlotT>
__global__ void __launch_bounds__(256) sym_replay_flat_kernel(const int* __restrict__ c2d, int L, int64_t npos,
                                                              const int* __restrict__ adj_pair, const int64_t* __restrict__ crow,
                                                              const uint16_t* __restrict__ stash, int* __restrict__ col,
                                                              SlotT* __restrict__ slots, int slot_stride) {
  const FastDiv fd(L);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npos; q += (int64_t)gridDim.x * blockDim.x) {
    const int pair = adj_pair[q];
    const int64_t cell = fd.wide(pair);
    const int64_t cbase = crow[c2d[pair]];
    const uint16_t* __restrict__ st = stash + q * L;
    SlotT* __restrict__ sl = slots + q * slot_stride;
    const int* __restrict__ dofs = c2d + cell * L;
    for (int j = 0; j < L; ++j) {
      const int e = st[j], rank = e >> 1;
      sl[j] = (SlotT)rank;
      if (e & 1) col[cbase + rank] = dofs[j];
    }
  }
}

__global__ void max_kernel(const int* __restrict__ v, int64_t n, int* out) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = max(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// elements per slot record: records are padded to 4-byte multiples (see SlotRec)
int slot_stride(int L, int slot_bytes) {
  const int per_word = 4 / slot_bytes;
  return (L + per_word - 1) / per_word * per_word;
}

size_t sym_workspace_bytes(int64_t NC, int L, int64_t gdof) {
  return align_up((size_t)gdof * 4) * 2 + scan_workspace_bytes(gdof + 1) + 1024;
}

int sym_count(const int* c2d, int64_t NC, int L, int64_t gdof, int64_t* adj_ptr, int* adj_pair, int64_t* crow, int64_t* nnz_host,
              int* max_row_host, uint16_t* stash, void* ws, cudaStream_t s) {
  const int64_t npair = NC * L;
  if (npair >= ((int64_t)1 << 31)) return fail(ERR_UNSUPPORTED, "sym_count: NC*ldof=%lld exceeds int32 pair ids", (long long)npair);
  FB2_TRY(build_adjacency(c2d, NC, L, gdof, adj_ptr, adj_pair, ws, s));
  Carver c(ws);                         // same carving as build_adjacency
  int* deg = c.take<int>(gdof);         // reused as rowlen
  int* cursor = c.take<int>(gdof);      // cursor[0..1] reused as flags at the end
  void* scan_ws = c.take<char>(scan_workspace_bytes(gdof + 1));
  FB2_CUDA(cudaMemsetAsync(cursor, 0, 8, s));
  int* rowlen = deg;
  const unsigned nb = (unsigned)std::min<int64_t>(ceil_div(gdof, SYM_WARPS), (int64_t)kNumSM * 16);
  if (gdof > 0) {
    sym_rows_kernel<false, uint8_t><<<nb, SYM_WARPS * 32, 0, s>>>(c2d, L, gdof, adj_ptr, adj_pair, nullptr, rowlen, nullptr, nullptr, 0, cursor, stash);
    max_kernel<<<grid_for(gdof), 256, 0, s>>>(rowlen, gdof, cursor + 1);
  }
  FB2_LAUNCH_CHECK();
  FB2_TRY(exclusive_scan_i32(rowlen, crow, gdof, true, scan_ws, s));
  int flags[2] = {0, 0};
  FB2_CUDA(cudaMemcpyAsync(flags, cursor, 8, cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaMemcpyAsync(nnz_host, crow + gdof, 8, cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  if (flags[0]) return fail(ERR_UNSUPPORTED, "sym_count: a dof touches more than %d candidate columns (valence*ldof)", SYM_CAP);
  *max_row_host = flags[1];
  return OK;
}

int sym_fill(const int* c2d, int64_t NC, int L, int64_t gdof, const int64_t* adj_ptr, const int* adj_pair, const int64_t* crow,
             int* col, void* slots, int slot_bytes, const uint16_t* stash, cudaStream_t s) {
  if (gdof <= 0) return OK;
  if (stash) {
    if (slot_bytes != 1 && slot_bytes != 2) return fail(ERR_INVALID, "sym_fill: slot_bytes must be 1 or 2");
    const int st = slot_stride(L, slot_bytes);
    const int64_t npos = NC * L;
    const unsigned g = grid_for(npos);
    if (slot_bytes == 1) sym_replay_flat_kernel<uint8_t><<<g, 256, 0, s>>>(c2d, L, npos, adj_pair, crow, stash, col, (uint8_t*)slots, st);
    else sym_replay_flat_kernel<uint16_t><<<g, 256, 0, s>>>(c2d, L, npos, adj_pair, crow, stash, col, (uint16_t*)slots, st);
    FB2_LAUNCH_CHECK();
    return OK;
  }
  const unsigned nb = (unsigned)std::min<int64_t>(ceil_div(gdof, SYM_WARPS), (int64_t)kNumSM * 16);
  const int stride = slot_stride(L, slot_bytes);
  if (slot_bytes == 1)
    sym_rows_kernel<true, uint8_t><<<nb, SYM_WARPS * 32, 0, s>>>(c2d, L, gdof, adj_ptr, adj_pair, crow, nullptr, col, (uint8_t*)slots, stride, nullptr, nullptr);
  else if (slot_bytes == 2)
    sym_rows_kernel<true, uint16_t><<<nb, SYM_WARPS * 32, 0, s>>>(c2d, L, gdof, adj_ptr, adj_pair, crow, nullptr, col, (uint16_t*)slots, stride, nullptr, nullptr);
  else
    return fail(ERR_INVALID, "sym_fill: slot_bytes must be 1 or 2");
  FB2_LAUNCH_CHECK();
  return OK;
}

// =====================================================================================
// numeric: scalar forms with constant / per-cell coefficients, element rows recomputed
// =====================================================================================
// ---- geometry from preloaded vertex coordinates ------------------------------------------
template <int TD>
struct CellX {                       // vertex ids + coordinates of one cell (software-pipelined loads)
  int v[TD + 1];
  double x[TD + 1][TD];
};

template <int TD>
__device__ __forceinline__ void load_verts(const int* __restrict__ cell, int64_t c, int (&v)[TD + 1]) {
  if constexpr (TD == 3) {
    const int4 t = *reinterpret_cast<const int4*>(cell + 4 * c);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = cell[3 * c]; v[1] = cell[3 * c + 1]; v[2] = cell[3 * c + 2];
  }
}

template <int TD>
__device__ __forceinline__ void load_coords(const double* __restrict__ node, const int (&v)[TD + 1], double (&x)[TD + 1][TD]) {
#pragma unroll
  for (int k = 0; k <= TD; ++k) {
    if constexpr (TD == 2) {
      const double2 p = *reinterpret_cast<const double2*>(node + 2 * (int64_t)v[k]);
      x[k][0] = p.x; x[k][1] = p.y;
    } else {
      const double* q = node + 3 * (int64_t)v[k];
      x[k][0] = q[0]; x[k][1] = q[1]; x[k][2] = q[2];
    }
  }
}

// G[kl] = vol * grad(lambda_k).grad(lambda_l) (upper triangle), cm = signed measure
__device__ __forceinline__ void geo_from_coords(const double (&P)[3][2], double& cm, double (&G)[6]) {
  const double e0x = P[2][0] - P[1][0], e0y = P[2][1] - P[1][1];
  const double e1x = P[0][0] - P[2][0], e1y = P[0][1] - P[2][1];
  const double e2x = P[1][0] - P[0][0], e2y = P[1][1] - P[0][1];
  const double nv = e0x * e1y - e0y * e1x, inv = 1.0 / nv;
  const double D[3][2] = {{-e0y * inv, e0x * inv}, {-e1y * inv, e1x * inv}, {-e2y * inv, e2x * inv}};
  cm = 0.5 * (e2x * e0y - e2y * e0x);
  int t = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int l = k; l < 3; ++l) G[t++] = (D[k][0] * D[l][0] + D[k][1] * D[l][1]) * cm;
}

__device__ __forceinline__ void geo_from_coords(const double (&P)[4][3], double& cm, double (&G)[10]) {
  double a[3], b[3], cc[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) { a[m] = P[1][m] - P[0][m]; b[m] = P[2][m] - P[1][m]; cc[m] = P[3][m] - P[2][m]; }
  const double bc0 = b[1] * cc[2] - b[2] * cc[1], bc1 = b[2] * cc[0] - b[0] * cc[2], bc2 = b[0] * cc[1] - b[1] * cc[0];
  const double det = a[0] * bc0 + a[1] * bc1 + a[2] * bc2;
  cm = det / 6.0;
  const double inv = 1.0 / det;
  constexpr int LF[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
  double D[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = LF[i][0], k = LF[i][1], mm = LF[i][2];
    double u[3], w[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { u[m] = P[mm][m] - P[j][m]; w[m] = P[k][m] - P[j][m]; }   // v_jm, v_jk
    D[i][0] = (u[1] * w[2] - u[2] * w[1]) * inv;
    D[i][1] = (u[2] * w[0] - u[0] * w[2]) * inv;
    D[i][2] = (u[0] * w[1] - u[1] * w[0]) * inv;
  }
  int t = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int l = k; l < 4; ++l) G[t++] = (D[k][0] * D[l][0] + D[k][1] * D[l][1] + D[k][2] * D[l][2]) * cm;
}

// slot records are padded to 4-byte multiples so that one pair's slots are a few 32-bit loads
template <typename SlotT, int L>
struct SlotRec {
  static constexpr int PER_WORD = 4 / (int)sizeof(SlotT);
  static constexpr int WORDS = (L + PER_WORD - 1) / PER_WORD;
  static constexpr int STRIDE = WORDS * PER_WORD;            // elements per record
  __device__ static __forceinline__ int get(const uint32_t (&w)[WORDS], int j) {
    if constexpr (sizeof(SlotT) == 1) return (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
    else return (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu;
  }
};

// CTA b owns the rows [blk_row[b], blk_row[b+1]) (equal value volume per CTA whatever the row
// lengths).  Inside, every warp owns 32 consecutive rows at a time and is fully autonomous:
// private accumulator segment, no CTA barrier after start-up, its own contiguous write-back.
// The dependent chain adjacency -> cell vertices -> node coordinates is software-pipelined
// three deep so that the global-load latency of pair q+1..q+3 hides behind the FP64 work of pair q.
#ifndef FB2_ASM_MINBLOCKS
#define FB2_ASM_MINBLOCKS 3
#endif
template <int TD, int L, typename SlotT>
__global__ void __launch_bounds__(128, FB2_ASM_MINBLOCKS) assemble_const_kernel(AsmConstArgs a) {
  constexpr int NV = TD + 1, NG = NV * (NV + 1) / 2;
  constexpr int MS_STRIDE = L * NG + 2;           // even (16-byte aligned rows for LDS.128) and != 0 mod 32 words
  constexpr int MM_STRIDE = (L + 3) & ~1;
  static_assert(NG % 2 == 0, "NG is even for triangles (6) and tetrahedra (10)");
  using SR = SlotRec<SlotT, L>;
  extern __shared__ __align__(16) double sm[];
  double* sMs = sm;                                            // [L][MS_STRIDE]
  double* sMm = sMs + (a.has_diff ? L * MS_STRIDE : 0);        // [L][MM_STRIDE]
  double* acc = sMm + (a.has_mass ? L * MM_STRIDE : 0);        // CTA tile of CSR values
  if (a.has_diff)
    for (int t = threadIdx.x; t < L * L * NG; t += blockDim.x) sMs[(t / (L * NG)) * MS_STRIDE + t % (L * NG)] = a.Ms[t];
  if (a.has_mass)
    for (int t = threadIdx.x; t < L * L; t += blockDim.x) sMm[(t / L) * MM_STRIDE + t % L] = a.Mm[t];

  __shared__ int next_rows;                                    // dynamic queue of 32-row blocks
  const int64_t r0 = a.blk_row[blockIdx.x], r1 = a.blk_row[blockIdx.x + 1];
  const int64_t v0 = a.crow[r0];
  const int nval = (int)(a.crow[r1] - v0);
  for (int t = threadIdx.x; t < nval; t += blockDim.x) acc[t] = 0.0;
  if (threadIdx.x == 0) next_rows = 0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const uint32_t* __restrict__ slot_words = static_cast<const uint32_t*>(a.slots);
  while (true) {
    int off = 0;
    if (lane == 0) off = atomicAdd(&next_rows, 32);
    off = __shfl_sync(0xffffffffu, off, 0);
    const int64_t rb = r0 + off;
    if (rb >= r1) break;
    const int64_t r = rb + lane;
    const bool act = r < r1;
    const int64_t q0 = act ? a.adj_ptr[r] : 0, q1 = act ? a.adj_ptr[r + 1] : 0;
    double* my = acc + (act ? (a.crow[r] - v0) : 0);
    if (q0 < q1) {
      const int64_t qlast = q1 - 1;
      auto clampq = [&](int64_t q) { return q < qlast ? q : qlast; };
      // prologue of the 3-deep pipeline
      int pair0 = a.adj_pair[q0], pair1 = a.adj_pair[clampq(q0 + 1)], pair2 = a.adj_pair[clampq(q0 + 2)];
      int v_cur[NV], v_nxt[NV];
      load_verts<TD>(a.cell, pair0 / L, v_cur);
      load_verts<TD>(a.cell, pair1 / L, v_nxt);
      double x_cur[NV][TD];
      load_coords<TD>(a.node, v_cur, x_cur);
      for (int64_t q = q0; q < q1; ++q) {
        // ---- issue the loads of the following pairs
        double x_nxt[NV][TD];
        load_coords<TD>(a.node, v_nxt, x_nxt);
        int v_nn[NV];
        load_verts<TD>(a.cell, pair2 / L, v_nn);
        const int pair3 = a.adj_pair[clampq(q + 3)];
        uint32_t sw[SR::WORDS];
#pragma unroll
        for (int w = 0; w < SR::WORDS; ++w) sw[w] = slot_words[q * SR::WORDS + w];
        // ---- FP64 work of the current pair
        const int64_t c = pair0 / L;
        const int i = pair0 - (int)c * L;
        double cm, G[NG];
        geo_from_coords(x_cur, cm, G);
        const double kd = a.scal_d * (a.coef_d ? a.coef_d[c] : 1.0);
        const double km = a.scal_m * (a.coef_m ? a.coef_m[c] : 1.0) * cm;
        const double2* mrow = reinterpret_cast<const double2*>(sMs + i * MS_STRIDE);
        const double* mm = sMm + i * MM_STRIDE;
        double val[L];
#pragma unroll
        for (int j = 0; j < L; ++j) {
          double v = 0.0;
          if (a.has_diff) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int t = 0; t < NG / 2; ++t) {
              const double2 m = mrow[j * (NG / 2) + t];
              s0 += m.x * G[2 * t];
              s1 += m.y * G[2 * t + 1];
            }
            v = kd * (s0 + s1);
          }
          if (a.has_mass) v += km * mm[j];
          val[j] = v;
        }
        // the dofs of one cell are distinct, so are their slots: independent read-modify-writes
        double old[L];
#pragma unroll
        for (int j = 0; j < L; ++j) old[j] = my[SR::get(sw, j)];
#pragma unroll
        for (int j = 0; j < L; ++j) my[SR::get(sw, j)] = old[j] + val[j];
        // ---- rotate the pipeline
        pair0 = pair1; pair1 = pair2; pair2 = pair3;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          v_nxt[k] = v_nn[k];
#pragma unroll
          for (int m = 0; m < TD; ++m) x_cur[k][m] = x_nxt[k][m];
        }
      }
    }
    __syncwarp();
    // the warp's 32 rows are a contiguous value range: stream it out
    const int64_t rend = (rb + 32 < r1) ? rb + 32 : r1;
    const int s0 = (int)(a.crow[rb] - v0), s1 = (int)(a.crow[rend] - v0);
    double* out = a.values + v0;
    for (int t = s0 + lane; t < s1; t += 32) out[t] = acc[t];
  }
}

// =====================================================================================
// numeric: generic gather of precomputed element-matrix rows (any integrator, tensor spaces)
//   lanes = columns: a group of lt = ldof*ncomp lanes owns one output (tensor) row at a time and walks
//   its (cell, i) pairs in order; every pair is ONE coalesced read of an element-matrix row (lt
//   doubles) and lt adds at distinct positions of the row's segment in the CTA tile; eight pairs are
//   in flight per group.  (The first version had one THREAD per row chasing 8-byte loads.)
//   scalar row r = row / ncomp (interleaved) or row % gdof (dof_priority)
// =====================================================================================
//   ETD > 0: P1 linear elasticity on the fly (ldof = ETD+1, ncomp = ETD): the element-matrix entry is
//   |K| w (d_lam D_i[a] D_j[b] + d_shear D_i[b] D_j[a]) (a != b) resp. |K| w (d_diag D_i[a] D_j[a] + d_shear sum_{z != a}
//   D_i[z] D_j[z]) from a 13-double (7 in 2-D) per-cell record of grad lambda -- the 1152-byte K_e per cell is never written
//   or read (fem/linear_elasticity_integrator.py:159-179 with grad phi_i = grad lambda_i)
template <int N>
__device__ __forceinline__ double pick(const double (&d)[N], int k) {
  double v = d[0];
#pragma unroll
  for (int z = 1; z < N; ++z) v = (k == z) ? d[z] : v;
  return v;
}

template <typename SlotT, int ETD, int NC, int U>
__global__ void __launch_bounds__(128) assemble_from_ke_kernel(AsmKeArgs a) {
  extern __shared__ __align__(16) double acc[];
  const int L = ETD > 0 ? ETD + 1 : a.L;                   // lanes per group: one per local (scalar) column dof j
  const int lt = L * NC;
  const int64_t R0 = a.blk_row[blockIdx.x], R1 = a.blk_row[blockIdx.x + 1];
  const int64_t v0 = a.crow_out[R0];
  const int nval = (int)(a.crow_out[R1] - v0);
  for (int t = threadIdx.x; t < nval; t += blockDim.x) acc[t] = 0.0;
  __syncthreads();
  const SlotT* __restrict__ slots = static_cast<const SlotT*>(a.slots);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int G = 32 / L;                                    // output rows in flight per warp (L <= 20)
  const int g = lane / L, j = lane - g * L;                // this lane's row group and column dof
  const uint32_t gmask = ((1u << L) - 1u) << (g * L);
  if (g < G) {
    for (int64_t R = R0 + (int64_t)wid * G + g; R < R1; R += (int64_t)nw * G) {
      int64_t r;
      int comp;
      if (NC == 1) { r = R; comp = 0; }
      else if (a.dof_priority) { comp = (int)(R / a.gdof); r = R - (int64_t)comp * a.gdof; }
      else { r = R / NC; comp = (int)(R - r * NC); }
      const int len = (int)(a.crow_s[r + 1] - a.crow_s[r]);
      double* my = acc + (a.crow_out[R] - v0);
      const int64_t q0 = a.adj_ptr[r], q1 = a.adj_ptr[r + 1];
      // the NC entries (row (i, comp), columns (j, b), b < NC) of one (cell, i) pair, and where the first goes
      auto operand = [&](int64_t q, int pair, double (&kv)[NC], int& sl) {
        const int64_t c = pair / L;
        const int i = pair - (int)c * L;
        if constexpr (ETD > 0) {
          constexpr int RS = (ETD + 1) * ETD + 1;
          const double* rec = a.geo + c * RS;
          double Di[ETD], Dj[ETD];
#pragma unroll
          for (int z = 0; z < ETD; ++z) { Di[z] = rec[i * ETD + z]; Dj[z] = rec[j * ETD + z]; }
          const double cm = rec[RS - 1], dic = pick(Di, comp), djc = pick(Dj, comp);
          double oth = 0.0;
#pragma unroll
          for (int z = 0; z < ETD; ++z) oth += (z == comp) ? 0.0 : a.wsum * (Di[z] * Dj[z]);
#pragma unroll
          for (int bb = 0; bb < NC; ++bb) {
            const double v = (bb == comp) ? a.d_diag * (a.wsum * (dic * djc)) + a.d_shear * oth
                                          : a.d_lam * (a.wsum * (dic * Dj[bb])) + a.d_shear * (a.wsum * (Di[bb] * djc));
            kv[bb] = v * cm;
          }
        } else {
          const int lrow = a.dof_priority ? comp * L + i : i * NC + comp;
          const double* krow = a.Ke + (c * lt + lrow) * (int64_t)lt;
#pragma unroll
          for (int bb = 0; bb < NC; ++bb) kv[bb] = krow[a.dof_priority ? bb * L + j : j * NC + bb];
        }
        sl = slots[q * a.slot_stride + j];
      };
      // pairs in chunks of U: the element-matrix rows of a chunk are independent loads (memory-level
      // parallelism), the adds then run in pair order; the next chunk's pair ids are fetched meanwhile
      // (U = 8 for scalar forms on meshes of high valence, 4 for tensor spaces; 2 where a dof meets few cells -- tri P3 has
      // 2.2 pairs per row, most of an 8-wide chunk would be predicated-off instructions: 2.71 -> see profiles/r02_tune_gather.txt)
      int pr[U], prn[U];
#pragma unroll
      for (int u = 0; u < U; ++u) pr[u] = q0 + u < q1 ? a.adj_pair[q0 + u] : 0;
      for (int64_t q = q0; q < q1; q += U) {
        double kv[U][NC];
        int sl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          sl[u] = 0;
#pragma unroll
          for (int bb = 0; bb < NC; ++bb) kv[u][bb] = 0.0;
          if (q + u < q1) operand(q + u, pr[u], kv[u], sl[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) prn[u] = q + U + u < q1 ? a.adj_pair[q + U + u] : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (q + u < q1) {
#pragma unroll
            for (int bb = 0; bb < NC; ++bb) my[a.dof_priority ? bb * len + sl[u] : sl[u] * NC + bb] += kv[u][bb];
          }
          __syncwarp(gmask);                               // the next pair may hit the same entry from another lane
        }
#pragma unroll
        for (int u = 0; u < U; ++u) pr[u] = prn[u];
      }
    }
  }
  __syncthreads();
  double* out = a.values + v0;
  for (int t = threadIdx.x; t < nval; t += blockDim.x) out[t] = acc[t];      // (st.global.cs measured no different here)
}

__global__ void __launch_bounds__(256) expand_crow_kernel(int64_t gdof, int nc, int prio, const int64_t* __restrict__ crow_s,
                                                          int64_t* __restrict__ crow_out) {
  // row lengths of the tensor pattern are nc * len(scalar row); rows are laid out so that
  // crow_out is available in closed form from crow_s
  const int64_t nrow = gdof * nc, nnz_s = crow_s[gdof];
  for (int64_t R = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; R <= nrow; R += (int64_t)gridDim.x * blockDim.x) {
    if (R == nrow) { crow_out[R] = nnz_s * nc * nc; continue; }
    if (prio) {
      const int comp = (int)(R / gdof);
      const int64_t r = R - (int64_t)comp * gdof;
      crow_out[R] = (int64_t)comp * nnz_s * nc + crow_s[r] * nc;
    } else {
      const int64_t r = R / nc;
      const int comp = (int)(R - r * nc);
      const int64_t len = crow_s[r + 1] - crow_s[r];
      crow_out[R] = crow_s[r] * nc * nc + (int64_t)comp * len * nc;
    }
  }
}

__global__ void __launch_bounds__(256) expand_col_kernel(int64_t gdof, int nc, int prio, const int64_t* __restrict__ crow_s,
                                                         const int* __restrict__ col_s, const int64_t* __restrict__ crow_out,
                                                         int* __restrict__ col_out) {
  const int64_t nrow = gdof * nc;
  for (int64_t R = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; R < nrow; R += (int64_t)gridDim.x * blockDim.x) {
    int64_t r;
    if (prio) r = R % gdof; else r = R / nc;
    const int64_t b = crow_s[r];
    const int len = (int)(crow_s[r + 1] - b);
    int* out = col_out + crow_out[R];
    for (int s = 0; s < len; ++s) {
      const int64_t cs = col_s[b + s];
      for (int q = 0; q < nc; ++q) {
        if (prio) out[q * len + s] = (int)((int64_t)q * gdof + cs);
        else out[s * nc + q] = (int)(cs * nc + q);
      }
    }
  }
}

// ---- host dispatch --------------------------------------------------------------------
template <int TD, int L>
static int launch_asm_const(AsmConstArgs a, int slot_bytes, int max_row, cudaStream_t s) {
  constexpr int NV = TD + 1, NG = NV * (NV + 1) / 2;
  // must mirror MS_STRIDE / MM_STRIDE of assemble_const_kernel
  const size_t tab = ((a.has_diff ? (size_t)L * (L * NG + 2) : 0) + (a.has_mass ? (size_t)L * ((L + 3) & ~1) : 0)) * sizeof(double);
  if (a.threads <= 0) a.threads = 128;
  const size_t smem = tab + (size_t)(a.tile + max_row) * 8;
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "assemble_const: row tile does not fit shared memory (max_row=%d)", max_row);
  const int64_t nb = a.nblk;
  if (nb <= 0) return OK;
  if (slot_bytes == 1) {
    auto k = assemble_const_kernel<TD, L, uint8_t>;
    FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)nb, a.threads, smem, s>>>(a);
  } else {
    auto k = assemble_const_kernel<TD, L, uint16_t>;
    FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<(unsigned)nb, a.threads, smem, s>>>(a);
  }
  FB2_LAUNCH_CHECK();
  return OK;
}

int assemble_const(int TD, int p, const AsmConstArgs& a, int slot_bytes, int max_row, cudaStream_t s) {
  if (a.gdof <= 0) return OK;
  if (slot_bytes != 1 && slot_bytes != 2) return fail(ERR_INVALID, "assemble_const: slot_bytes must be 1 or 2");
  switch (TD * 10 + p) {
    case 21: return launch_asm_const<2, 3>(a, slot_bytes, max_row, s);
    case 22: return launch_asm_const<2, 6>(a, slot_bytes, max_row, s);
    case 23: return launch_asm_const<2, 10>(a, slot_bytes, max_row, s);
    case 31: return launch_asm_const<3, 4>(a, slot_bytes, max_row, s);
    case 32: return launch_asm_const<3, 10>(a, slot_bytes, max_row, s);
    case 33: return launch_asm_const<3, 20>(a, slot_bytes, max_row, s);
    default: return fail(ERR_UNSUPPORTED, "assemble_const: unsupported element TD=%d p=%d", TD, p);
  }
}

template <typename SlotT, int ETD, int NC, int U>
static int launch_gather_u(const AsmKeArgs& a, size_t smem, cudaStream_t s) {
  auto k = assemble_from_ke_kernel<SlotT, ETD, NC, U>;
  FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(unsigned)a.nblk, 128, smem, s>>>(a);
  FB2_LAUNCH_CHECK();
  return OK;
}

template <int ETD, int NC>
static int launch_gather(const AsmKeArgs& a, int slot_bytes, size_t smem, cudaStream_t s) {
  // (cell, i) pairs per row: chunk of independent element-row loads in flight per lane group
  static const int forced = [] { const char* e = getenv("FB2_GATHER_U"); return e ? atoi(e) : 0; }();
  const double valence = a.gdof > 0 ? (double)a.ncell * a.L / (double)a.gdof : 8.0;
  if constexpr (NC == 1) {
    const int u = forced ? forced : (valence < 4.0 ? 2 : 8);
    if (u <= 2) return slot_bytes == 1 ? launch_gather_u<uint8_t, ETD, NC, 2>(a, smem, s) : launch_gather_u<uint16_t, ETD, NC, 2>(a, smem, s);
    if (u <= 4) return slot_bytes == 1 ? launch_gather_u<uint8_t, ETD, NC, 4>(a, smem, s) : launch_gather_u<uint16_t, ETD, NC, 4>(a, smem, s);
    return slot_bytes == 1 ? launch_gather_u<uint8_t, ETD, NC, 8>(a, smem, s) : launch_gather_u<uint16_t, ETD, NC, 8>(a, smem, s);
  } else {
    return slot_bytes == 1 ? launch_gather_u<uint8_t, ETD, NC, 4>(a, smem, s) : launch_gather_u<uint16_t, ETD, NC, 4>(a, smem, s);
  }
}

int assemble_from_ke(AsmKeArgs a, int slot_bytes, int max_row, cudaStream_t s) {
  if (a.gdof <= 0) return OK;
  const int max_out_row = max_row * a.ncomp;
  const size_t smem = (size_t)(a.tile + max_out_row) * 8;
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "assemble_from_ke: row tile does not fit shared memory (max_row=%d)", max_row);
  a.slot_stride = slot_stride(a.L, slot_bytes);
  const int64_t nb = a.nblk;
  if (nb <= 0) return OK;
  if (slot_bytes != 1 && slot_bytes != 2) return fail(ERR_INVALID, "assemble_from_ke: slot_bytes must be 1 or 2");
  if (a.L > 32) return fail(ERR_UNSUPPORTED, "assemble_from_ke: ldof=%d exceeds a warp", a.L);
  switch (a.ncomp) {
    case 1: return launch_gather<0, 1>(a, slot_bytes, smem, s);
    case 2: return launch_gather<0, 2>(a, slot_bytes, smem, s);
    case 3: return launch_gather<0, 3>(a, slot_bytes, smem, s);
    default: return fail(ERR_UNSUPPORTED, "assemble_from_ke: ncomp=%d (1..3 supported)", a.ncomp);
  }
}

int assemble_elasticity_p1(int TD, AsmKeArgs a, int slot_bytes, int max_row, cudaStream_t s) {
  if (a.gdof <= 0 || a.nblk <= 0) return OK;
  if (TD != 2 && TD != 3) return fail(ERR_UNSUPPORTED, "assemble_elasticity_p1: TD must be 2 or 3");
  if (slot_bytes != 1 && slot_bytes != 2) return fail(ERR_INVALID, "assemble_elasticity_p1: slot_bytes must be 1 or 2");
  a.L = TD + 1; a.ncomp = TD; a.Ke = nullptr;
  const size_t smem = (size_t)(a.tile + max_row * a.ncomp) * 8;
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "assemble_elasticity_p1: row tile does not fit shared memory (max_row=%d)", max_row);
  a.slot_stride = slot_stride(a.L, slot_bytes);
  return TD == 2 ? launch_gather<2, 2>(a, slot_bytes, smem, s) : launch_gather<3, 3>(a, slot_bytes, smem, s);
}

int expand_pattern(int64_t gdof, int nc, int prio, const int64_t* crow_s, const int* col_s, int64_t* crow_out, int* col_out,
                   cudaStream_t s) {
  if (gdof <= 0) return OK;
  expand_crow_kernel<<<grid_for(gdof * nc + 1), 256, 0, s>>>(gdof, nc, prio, crow_s, crow_out);
  expand_col_kernel<<<grid_for(gdof * nc), 256, 0, s>>>(gdof, nc, prio, crow_s, col_s, crow_out, col_out);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2

// =====================================================================================
// v4: scheduled batches.  The symbolic phase turns every tile (a run of rows holding at most
// `tile` values, owned by ONE warp) into a list of 32-entry batches such that
//   * all entries of a batch share the local index i      -> the table row T[i] is warp-uniform,
//   * all entries of a batch belong to different rows      -> conflict-free accumulation,
//   * per row, entries appear in (i, cell) order           -> fixed summation order, independent of the tiling.
// The numeric kernel is then a flat loop over batches with every lane busy: lane = one
// (row, cell) pair; geometry comes precomputed per cell (H, reduced form: 7 values per tetrahedron), the
// element row is 7 FMAs per column against uniform table operands, the result is added into the warp's private tile.
// =====================================================================================
namespace fb2 {


// one WARP per tile; COUNT pass returns the number of batches, FILL pass writes them.
// Within a tile the entries of local index i (n of them, at most m in any one row) are dealt
// round-robin over B = max(ceil(n/32), m) batches in adjacency order (row-major): entry k goes to
// batch k mod B, position k / B.  A row's entries are consecutive k and there are at most m <= B of
// them, so they land in different batches; every batch gets at most ceil(n/B) <= 32 entries; and B
// is the minimum possible.  Per row the cells stay in ascending order (see the rotation below).
// (The first version closed a batch whenever a row repeated: 65-70 % of the lanes carried work on
// tet P2; this one reaches 81-85 % with the same tile.)
// Lane = row while walking the adjacency (sorted by (i, cell) per row, so the run of local index i
// is found by advancing a per-lane cursor); lane i keeps the totals of local index i.
constexpr int A4_MAXL = 32;
constexpr int A4_BASE_BITS = 12;     // entry word = tile offset of the row (12 bits) | first-touch mask (ldof <= 20 bits)
constexpr int A4_MAXROW = 1024;      // longest row the first-touch bitmap covers (SYM_CAP bounds a row anyway)
__host__ __device__ constexpr int a4_block_words(int slot_nwords) { return 64 + 32 * slot_nwords + 4; }   // + a 16-byte header

template <bool FILL>
__global__ void __launch_bounds__(128) asm4_schedule_kernel(int ntile, const int32_t* __restrict__ blk_row, const int64_t* __restrict__ crow,
                                                            const int64_t* __restrict__ adj_ptr, const int* __restrict__ adj_pair, int L,
                                                            int* __restrict__ nbatch_of_tile, const int64_t* __restrict__ batch_ptr,
                                                            unsigned char* __restrict__ batch_i, uint32_t* __restrict__ blocks,
                                                            const uint32_t* __restrict__ slot_words, int slot_nwords, int slot_bytes) {
  // packed schedule: one contiguous block of BW 32-bit words per batch (a single bulk copy in the numeric kernel):
  //   [0,32) cell id or -1 | [32,64) row offset in the tile | first-touch mask << 12 | [64,64+32W) slot records | header: local index i
  const int64_t BW = a4_block_words(slot_nwords);
  constexpr unsigned FULL = 0xffffffffu;
  const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= ntile) return;
  const FastDiv fd(L);
  auto local_index = [&](int pair) { return pair - fd.wide(pair) * L; };
  const int64_t r0 = blk_row[t], r1 = blk_row[t + 1];
  const int64_t v0 = crow[r0];
  int cnt = 0, mult = 0;                         // lane i: entries of local index i in the tile, longest run inside one row
  for (int64_t rb = r0; rb < r1; rb += 32) {
    const int64_t r = rb + lane;
    int64_t q = 0, qe = 0;
    if (r < r1) { q = adj_ptr[r]; qe = adj_ptr[r + 1]; }
    for (int i = 0; i < L; ++i) {
      int m = 0;
      while (q + m < qe && local_index(adj_pair[q + m]) == i) ++m;
      q += m;
      const int s = __reduce_add_sync(FULL, m), mx = __reduce_max_sync(FULL, m);
      if (lane == i) { cnt += s; mult = max(mult, mx); }
    }
  }
  const int B = cnt == 0 ? 0 : max((cnt + 31) / 32, mult);     // batches of local index `lane`
  int off = B;                                                 // -> first batch (tile-local) of local index `lane`
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, off, o); if (lane >= o) off += v; }
  const int nb = __shfl_sync(FULL, off, 31);
  off -= B;
  if (!FILL) { if (lane == 0) nbatch_of_tile[t] = nb; return; }
  const int64_t b0 = batch_ptr[t];
  for (int i = 0; i < L; ++i) {
    const int Bi = __shfl_sync(FULL, B, i), offi = __shfl_sync(FULL, off, i), ci = __shfl_sync(FULL, cnt, i);
    for (int j = 0; j < Bi; ++j) {
      const int have = (ci - j + Bi - 1) / Bi;                 // entries k = j, j+B, j+2B, ... < cnt
      if (lane == 0) { batch_i[b0 + offi + j] = (unsigned char)i; blocks[(b0 + offi + j) * BW + 64 + 32 * slot_nwords] = (uint32_t)i; }
      if (lane >= have) blocks[(b0 + offi + j) * BW + lane] = 0xffffffffu;
    }
  }
  int dealt = 0;                                               // lane i: entries of local index i dealt so far
  for (int64_t rb = r0; rb < r1; rb += 32) {
    const int64_t r = rb + lane;
    int64_t q = 0, qe = 0;
    uint32_t base = 0;
    if (r < r1) { q = adj_ptr[r]; qe = adj_ptr[r + 1]; base = (uint32_t)(crow[r] - v0); }
    // first-touch flags: the lane walks its row in execution order ((i, cell), see the rotation below), so it knows
    // which (entry, column) is the first contribution to a value -- the numeric kernel stores that one instead of
    // load-add-store, and the tile needs no zero fill
    uint32_t touched[A4_MAXROW / 32];
#pragma unroll
    for (int w = 0; w < A4_MAXROW / 32; ++w) touched[w] = 0;
    for (int i = 0; i < L; ++i) {
      int m = 0;                                 // this row's run of local index i
      while (q + m < qe && local_index(adj_pair[q + m]) == i) ++m;
      int ex = m;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, ex, o); if (lane >= o) ex += v; }
      const int total = __shfl_sync(FULL, ex, 31);
      const int Bi = __shfl_sync(FULL, B, i), offi = __shfl_sync(FULL, off, i);
      const int k0 = __shfl_sync(FULL, dealt, i) + ex - m;     // rows are dealt in order: row-major deal numbers
      if (lane == i) dealt += total;
      if (m > 0) {
        // the run takes the deal numbers k0 .. k0+m-1; when they wrap around the Bi batches the numbers
        // are handed out rotated, so that the row still meets its cells in ascending batch order:
        // every row is summed in (i, cell) order whatever the tiling (single- and multi-GPU bit-identical)
        const int wrap = max(0, k0 % Bi + m - Bi);
        for (int u = 0; u < m; ++u) {
          const int k = u < wrap ? k0 + (m - wrap) + u : k0 + (u - wrap);
          uint32_t* blk = blocks + (b0 + offi + k % Bi) * BW;
          const int pos = k / Bi;
          blk[pos] = (uint32_t)fd.wide(adj_pair[q + u]);
          uint32_t first = 0;
          for (int w = 0; w < slot_nwords; ++w) blk[64 + pos * slot_nwords + w] = slot_words[(q + u) * slot_nwords + w];
          for (int j = 0; j < L; ++j) {
            const uint32_t wd = slot_words[(q + u) * slot_nwords + (slot_bytes == 1 ? (j >> 2) : (j >> 1))];
            const int sl = slot_bytes == 1 ? (wd >> ((j & 3) * 8)) & 0xffu : (wd >> ((j & 1) * 16)) & 0xffffu;
            if (!((touched[sl >> 5] >> (sl & 31)) & 1u)) { first |= 1u << j; touched[sl >> 5] |= 1u << (sl & 31); }
          }
          blk[32 + pos] = base | (first << A4_BASE_BITS);
        }
      }
      q += m;
    }
  }
}

// element tables in the kernel parameter block: operands come through the constant / uniform
// datapath and cost no LSU (shared-memory) wavefronts -- the LSU data pipe is this kernel's limiter.
//
// Reduced geometry.  sum_k grad(lambda_k) = 0, so the (TD+1)(TD+2)/2 products g_kl = |K| grad(lambda_k).grad(lambda_l)
// are linear in the TD(TD+1)/2 with k, l >= 1:  g_0l = -sum_m g_ml,  g_00 = sum_mn g_mn.  Folding that
// into the table (a4_reduced_table) leaves  Ke[i][j] = sum_{1<=m<=n} Tr[i][j][mn] g_mn + Mm[i][j] |K|:
// 7 instead of 11 values per cell and per FMA chain on tetrahedra, 4 instead of 7 on triangles.
template <int TD>
struct A4Geo {
  static constexpr int NV = TD + 1, NG = NV * (NV + 1) / 2;   // full upper triangle (tables arrive in this layout)
  static constexpr int NR = TD * (TD + 1) / 2;                // reduced: 1 <= m <= n <= TD
  static constexpr int NH = NR + 1;                           // + the mass factor
  static constexpr int HS = ((NH + 1) / 2) * 2;               // record length in doubles (16-byte units)
  static constexpr int QP = HS / 2;                           // 16-byte units per record: 4 (tet), 2 (tri)
  __host__ __device__ static constexpr int full(int k, int l) { return k * NV - k * (k - 1) / 2 + (l - k); }   // k <= l
};
template <int L, int NH>
struct A4Tables { double T[L][L][NH]; };     // T[i][j] = (Tr[i][j][0..NR-1], Mm[i][j])

template <int TD, int L>
static void a4_reduced_table(const double* Ms, const double* Mm, A4Tables<L, A4Geo<TD>::NH>& tb) {
  using GEO = A4Geo<TD>;
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      const double* ms = Ms ? Ms + (size_t)(i * L + j) * GEO::NG : nullptr;
      int t = 0;
      for (int m = 1; m <= TD; ++m)
        for (int n = m; n <= TD; ++n, ++t) {
          if (!ms) { tb.T[i][j][t] = 0.0; continue; }
          const double m00 = ms[GEO::full(0, 0)], m0m = ms[GEO::full(0, m)], m0n = ms[GEO::full(0, n)];
          tb.T[i][j][t] = m == n ? (ms[GEO::full(m, m)] + m00) - m0m : (ms[GEO::full(m, n)] + 2.0 * m00) - (m0m + m0n);
        }
      tb.T[i][j][GEO::NR] = Mm ? Mm[i * L + j] : 0.0;
    }
}

#ifndef FB2_ASM4_WARPS
#define FB2_ASM4_WARPS 3       // 3 CTAs x 3 warps per SM (3.65 ms; 4 x 2: 3.72; 2 x 4: 3.68 -- profiles/r02_tune_asm_v6.txt)
#endif
#ifndef FB2_ASM4_MINBLOCKS
#define FB2_ASM4_MINBLOCKS 3
#endif

// ---- v6 numeric kernel: the same schedule, operands through the bulk-copy engine and registers ----------------------
// Its predecessor (git history, "v4") staged both operand streams of a batch through cp.async rings in shared memory; ncu
// (profiles/r02_ncu_asm_experiments.txt): 393 warp instructions per batch, of which ~150 were operand plumbing (six LDGSTS
// per lane, each with its address arithmetic and three hazard NOPs, plus the read-back of both rings) and 10 % of all
// stall samples waited for the batch's local index.  Here
//   * the batch block (656 bytes on tet P2, header word = local index) arrives with ONE cp.async.bulk issued by lane 0
//     into a 4-deep ring, completion through an mbarrier per ring slot (SASS: UBLKCP + SYNCS);
//   * the geometry record of the lane's cell is loaded straight into registers one batch ahead (the cell id of batch
//     b+1 is already in shared memory when batch b starts): no staging ring, no read-back wavefronts;
//   * the old value of a slot seeds the FMA chain (first-touch columns start from 0), so the accumulate is
//     LDS -> 7 DFMA -> STS with no separate add.
__device__ __forceinline__ void ldg_nc_f64x4(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];\n" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__device__ __forceinline__ void ldg_nc_f64x4(const double* p, double& a, double& b, double& c, double& d, uint64_t pol) {
  asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;\n" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol));
}

#ifndef FB2_ASM6_JC
#define FB2_ASM6_JC 10     // all columns of a row in flight: 3.72 ms against 3.99 with 5 (tet P2 128^3; profiles/r02_tune_asm_v6.txt)
#endif
template <int TD, int L, typename SlotT, int I, typename TabT>
__device__ __forceinline__ void a6_row(const TabT& tb, const double (&h)[A4Geo<TD>::HS],
                                       const uint32_t (&sw)[SlotRec<SlotT, L>::WORDS], uint32_t first, double* __restrict__ my) {
  constexpr int NH = A4Geo<TD>::NH;
  using SR = SlotRec<SlotT, L>;
  constexpr int JC = (L % FB2_ASM6_JC == 0) ? FB2_ASM6_JC : ((L % 5 == 0) ? 5 : ((L % 4 == 0) ? 4 : 3));
#pragma unroll
  for (int j0 = 0; j0 < L; j0 += JC) {
    double s[JC];
#pragma unroll
    for (int jj = 0; jj < JC; ++jj) s[jj] = ((first >> (j0 + jj)) & 1u) ? 0.0 : my[SR::get(sw, j0 + jj)];
#pragma unroll
    for (int t = 0; t < NH; ++t)
#pragma unroll
      for (int jj = 0; jj < JC; ++jj) s[jj] = fma(tb.T[I][j0 + jj][t], h[t], s[jj]);
#pragma unroll
    for (int jj = 0; jj < JC; ++jj) my[SR::get(sw, j0 + jj)] = s[jj];
  }
}
template <int TD, int L, typename SlotT, int I, typename TabT>
__device__ __forceinline__ void a6_dispatch(int i, const TabT& tb, const double (&h)[A4Geo<TD>::HS],
                                            const uint32_t (&sw)[SlotRec<SlotT, L>::WORDS], uint32_t first, double* __restrict__ my) {
  if (i == I) a6_row<TD, L, SlotT, I, TabT>(tb, h, sw, first, my);
  else if constexpr (I + 1 < L) a6_dispatch<TD, L, SlotT, I + 1, TabT>(i, tb, h, sw, first, my);
}

#ifndef FB2_ASM6_RING
#define FB2_ASM6_RING 4
#endif
constexpr int A6_RING = FB2_ASM6_RING;        // batch blocks in flight per warp

template <int TD, int L, typename SlotT>
__global__ void __launch_bounds__(FB2_ASM4_WARPS * 32, FB2_ASM4_MINBLOCKS)
assemble_const_v6_kernel(const __grid_constant__ Asm4Args a, const __grid_constant__ A4Tables<L, A4Geo<TD>::NH> tb) {
  using GEO = A4Geo<TD>;
  constexpr int HS = GEO::HS;
  using SR = SlotRec<SlotT, L>;
  constexpr int W = SR::WORDS, BW = a4_block_words(W), BLK_BYTES = BW * 4, NE = A6_RING;
  static_assert(BLK_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  extern __shared__ __align__(16) double sm4[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tile = blockIdx.x * FB2_ASM4_WARPS + wid;
  if (tile >= a.ntile) return;
  const size_t per_warp = (size_t)a.acc_stride + (NE * BLK_BYTES + ((NE * 8 + 15) & ~15)) / 8;     // keeps every warp's ring 16-byte aligned
  double* acc = sm4 + (size_t)wid * per_warp;                                   // this warp's private tile ...
  const uint32_t* ering = reinterpret_cast<const uint32_t*>(acc + a.acc_stride);  // ... its ring of batch blocks ...
  const uint32_t ering_s = smem_u32(ering);
  const uint32_t bar_s = ering_s + NE * BLK_BYTES;                              // ... and one mbarrier per ring slot
  const int64_t r0 = a.blk_row[tile], r1 = a.blk_row[tile + 1];
  const int64_t v0 = a.crow[r0];
  const int nval = (int)(a.crow[r1] - v0);
  const int64_t b0 = a.batch_ptr[tile];
  const int nb = (int)(a.batch_ptr[tile + 1] - b0);
  const unsigned char* gblk = reinterpret_cast<const unsigned char*>(a.blocks) + b0 * BLK_BYTES;

  // L2 policies (profiles/r02_tune_asm_v6.txt, call h25): the geometry records are re-read ~10 times through L2 and
  // stay (evict_last), the schedule blocks are read once (evict_first) and the values are written once (st.global.cs):
  // 3.647 -> 3.567 ms on config 2 with all three (records alone: no change; + streaming stores 3.60).  -DFB2_A6_NO_L2_POLICY
  // restores the default policies.
#ifndef FB2_A6_NO_L2_POLICY
#define FB2_A6_H_KEEP
#define FB2_A6_BLK_FIRST
#define FB2_A6_ST_CS
#endif
#ifdef FB2_A6_H_KEEP
  const uint64_t pol_h = l2_policy_evict_last();
#endif
#ifdef FB2_A6_BLK_FIRST
  const uint64_t pol_b = l2_policy_evict_first();
#define A6_BULK(dst, src, bytes, bar) bulk_g2s((dst), (src), (bytes), (bar), pol_b)
#else
#define A6_BULK(dst, src, bytes, bar) bulk_g2s((dst), (src), (bytes), (bar))
#endif
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NE; ++k) mbar_init(bar_s + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#pragma unroll
    for (int k = 0; k < NE; ++k)
      if (k < nb) {
        mbar_expect_tx(bar_s + 8 * k, BLK_BYTES);
        A6_BULK(ering_s + k * BLK_BYTES, gblk + (size_t)k * BLK_BYTES, BLK_BYTES, bar_s + 8 * k);
      }
  }
  __syncwarp();

  // 256-bit loads (LDG.E.256, new on sm_100): a scattered load costs the L1 data pipe one wavefront per lane and
  // instruction whatever its width -- ncu of the 128-bit version: 108 global wavefronts per batch, the pipe 82 % busy
  auto load_h = [&](int cell, double (&h)[HS]) {
    const double* src = a.H + (int64_t)cell * HS;
    static_assert(HS % 4 == 0, "geometry records are multiples of 32 bytes");
#pragma unroll
    for (int t = 0; t < HS / 4; ++t)
#ifdef FB2_A6_H_KEEP
      ldg_nc_f64x4(src + 4 * t, h[4 * t], h[4 * t + 1], h[4 * t + 2], h[4 * t + 3], pol_h);
#else
      ldg_nc_f64x4(src + 4 * t, h[4 * t], h[4 * t + 1], h[4 * t + 2], h[4 * t + 3]);
#endif
  };

  // (fetching the records TWO batches ahead -- three register buffers in fixed roles, loop unrolled by three -- measured
  // slower: 3.82 against 3.58 ms, 152 registers and three copies of the 10-way row dispatch; profiles/r02_tune_asm_v6.txt)
  int celln = -1;
  double hn[HS];
#pragma unroll
  for (int t = 0; t < HS; ++t) hn[t] = 0.0;
  if (nb > 0) {
    while (!mbar_try_wait(bar_s, 0)) { }
    celln = (int)ering[lane];
    if (celln >= 0) load_h(celln, hn);
  }
  int slot = 0;
  uint32_t parity = 0;
  for (int k = 0; k < nb; ++k) {
    const uint32_t* ent = ering + slot * BW;
    const int cell = celln;
    double h[HS];
#pragma unroll
    for (int t = 0; t < HS; ++t) h[t] = hn[t];
    const uint32_t bw = ent[32 + lane];
    uint32_t sw[W];
#pragma unroll
    for (int w = 0; w < W; ++w) sw[w] = ent[64 + lane * W + w];
    const int i = (int)ent[64 + 32 * W];
    int nslot = slot + 1;
    uint32_t nparity = parity;
    if (nslot == NE) { nslot = 0; nparity ^= 1u; }
    if (k + 1 < nb) {                                           // next batch: its block is (almost always) already here
      while (!mbar_try_wait(bar_s + 8 * nslot, nparity)) { }
      celln = (int)ering[nslot * BW + lane];
      if (celln >= 0) load_h(celln, hn);
    }
    if (cell >= 0) a6_dispatch<TD, L, SlotT, 0>(i, tb, h, sw, bw >> A4_BASE_BITS, acc + (bw & ((1u << A4_BASE_BITS) - 1u)));
    __syncwarp();                                               // every lane is done with the block in `slot` ...
    if (lane == 0 && k + NE < nb) {                             // ... which is refilled with batch k + NE
      mbar_expect_tx(bar_s + 8 * slot, BLK_BYTES);
      A6_BULK(ering_s + slot * BLK_BYTES, gblk + (size_t)(k + NE) * BLK_BYTES, BLK_BYTES, bar_s + 8 * slot);
    }
    slot = nslot; parity = nparity;
  }
  __syncwarp();
  double* out = a.values + v0;
#ifdef FB2_A6_ST_CS
  for (int t = lane; t < nval; t += 32) __stcs(out + t, acc[t]);
#else
  for (int t = lane; t < nval; t += 32) out[t] = acc[t];
#endif
#undef A6_BULK
}

// per-cell record H = (kd * g_mn for 1 <= m <= n <= TD, km * |K|), padded to a multiple of 2 doubles
template <int TD>
__global__ void __launch_bounds__(256) cell_geometry4_kernel(const double* __restrict__ node, const int* __restrict__ cell, int64_t NC,
                                                             double scal_d, const double* __restrict__ coef_d, double scal_m,
                                                             const double* __restrict__ coef_m, double* __restrict__ H) {
  using GEO = A4Geo<TD>;
  constexpr int NV = TD + 1, NG = GEO::NG, NR = GEO::NR, NH = GEO::NH, HS = GEO::HS;
  __shared__ double stage[256 * HS];
  const int64_t c0 = (int64_t)blockIdx.x * 256;
  const int64_t c = c0 + threadIdx.x;
  if (c < NC) {
    int v[NV];
    double x[NV][TD], cm, G[NG];
    load_verts<TD>(cell, c, v);
    load_coords<TD>(node, v, x);
    geo_from_coords(x, cm, G);
    const double kd = scal_d * (coef_d ? coef_d[c] : 1.0);
    const double km = scal_m * (coef_m ? coef_m[c] : 1.0) * cm;
    double* h = stage + threadIdx.x * HS;
    int t = 0;
#pragma unroll
    for (int m = 1; m <= TD; ++m)
#pragma unroll
      for (int n = m; n <= TD; ++n) h[t++] = kd * G[GEO::full(m, n)];
    h[NR] = km;
    if (HS > NH) h[NH] = 0.0;
  }
  __syncthreads();
  const int64_t ncell = (NC - c0) < 256 ? (NC - c0) : 256;
  double* dst = H + c0 * HS;
  for (int64_t t = threadIdx.x; t < ncell * HS; t += 256) dst[t] = stage[t];       // contiguous, fully coalesced
}

// ---- matrix-free product of the same forms: v = A u without A and without K_e (fem/bilinear_form.py:126-158) ----------
// The reference forms gv = einsum('cij,cj->ci', K_e, u[cell2dof]) and index_adds it into v.  Here one thread owns one
// cell: geometry -> reduced record h (registers), u gathered through cell2dof, and
//     w_i = sum_t h_t * (sum_j T[i][j][t] u_j)
// with the folded table T in the kernel parameter block (warp-uniform operands, L*L*NH + L*NH DFMAs per cell, no
// element matrix anywhere).  The (NC, L) block w goes to HBM once (8*L bytes per cell instead of 8*L*L for K_e) and the
// row-owner gather of gather_vector() sums it per dof in the fixed (i, cell) order: no atomics, bit-reproducible.
#ifndef FB2_MF_MINBLOCKS
#define FB2_MF_MINBLOCKS 1
#endif
template <int TD, int L>
__global__ void __launch_bounds__(128, FB2_MF_MINBLOCKS) matfree_cell_kernel(const double* __restrict__ node, const int* __restrict__ cell,
                                                           const int* __restrict__ c2d, int64_t NC, double scal_d,
                                                           const double* __restrict__ coef_d, double scal_m,
                                                           const double* __restrict__ coef_m, const double* __restrict__ u,
                                                           double* __restrict__ w, const int* __restrict__ pair_pos,
                                                           const __grid_constant__ A4Tables<L, A4Geo<TD>::NH> tb) {
  using GEO = A4Geo<TD>;
  constexpr int NV = TD + 1, NG = GEO::NG, NR = GEO::NR, NH = GEO::NH;
  __shared__ double stage[128 * L];
  const int64_t c0 = (int64_t)blockIdx.x * 128;
  const int64_t c = c0 + threadIdx.x;
  if (c < NC) {
    int v[NV];
    double x[NV][TD], cm, G[NG], h[NH], ul[L];
    load_verts<TD>(cell, c, v);
    const int* dofs = c2d + c * L;
#pragma unroll
    for (int j = 0; j < L; ++j) ul[j] = u[dofs[j]];
    load_coords<TD>(node, v, x);
    geo_from_coords(x, cm, G);
    const double kd = scal_d * (coef_d ? coef_d[c] : 1.0);
    int t = 0;
#pragma unroll
    for (int m = 1; m <= TD; ++m)
#pragma unroll
      for (int n = m; n <= TD; ++n) h[t++] = kd * G[GEO::full(m, n)];
    h[NR] = scal_m * (coef_m ? coef_m[c] : 1.0) * cm;
    double* out = stage + threadIdx.x * L;
#ifdef FB2_MF_IUNROLL
    constexpr int IU = FB2_MF_IUNROLL;
#else
    constexpr int IU = L <= 6 ? L : 1;      // 10 rows unrolled: 156-198 registers, 2.11 ms; rolled: 74 registers, 1.80 ms (tet P2 128^3)
#endif
#pragma unroll(IU)
    for (int i = 0; i < L; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int tt = 0; tt < NH; ++tt) {
        double z = 0.0;
#pragma unroll
        for (int j = 0; j < L; ++j) z = fma(tb.T[i][j][tt], ul[j], z);
        acc = fma(h[tt], z, acc);
      }
      out[i] = acc;
    }
  }
  __syncthreads();
  const int64_t ncell = (NC - c0) < 128 ? (NC - c0) : 128;
  if (pair_pos) {
    // adjacency order: entry (c, i) goes to the position of pair c*L+i in its dof's list, so that the per-dof sum reads ONE
    // contiguous run (the pair-ordered block was read back at sector granularity: 4.8 GB for 1.0 GB of payload); the
    // scattered 8-byte stores of one dof's run come from cells that are neighbours in memory and merge in L2
    const int* __restrict__ pos = pair_pos + c0 * L;
    for (int64_t t = threadIdx.x; t < ncell * L; t += 128) w[pos[t]] = stage[t];
    return;
  }
  double* dst = w + c0 * L;
  for (int64_t t = threadIdx.x; t < ncell * L; t += 128) dst[t] = stage[t];       // contiguous, fully coalesced
}

// pair_pos[pair] = position of (cell, i) pair in the adjacency lists (inverse of adj_pair)
__global__ void __launch_bounds__(256) pair_position_kernel(int64_t npos, const int* __restrict__ adj_pair, int* __restrict__ pair_pos) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npos; q += (int64_t)gridDim.x * blockDim.x)
    pair_pos[adj_pair[q]] = (int)q;
}

// F[d] = sum of the contiguous run w[adj_ptr[d] .. adj_ptr[d+1]) in order (the order of gather_vector_kernel)
__global__ void __launch_bounds__(256) segment_sum_kernel(int64_t gdof, const int64_t* __restrict__ adj_ptr, const double* __restrict__ w,
                                                          double* __restrict__ F) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < gdof; d += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t q = adj_ptr[d]; q < adj_ptr[d + 1]; ++q) s += w[q];
    F[d] = s;
  }
}

int pair_positions(int64_t npos, const int* adj_pair, int* pair_pos, cudaStream_t s) {
  if (npos <= 0) return OK;
  pair_position_kernel<<<grid_for(npos), 256, 0, s>>>(npos, adj_pair, pair_pos);
  FB2_LAUNCH_CHECK();
  return OK;
}

int segment_sum(int64_t gdof, const int64_t* adj_ptr, const double* w, double* F, cudaStream_t s) {
  if (gdof <= 0) return OK;
  segment_sum_kernel<<<grid_for(gdof), 256, 0, s>>>(gdof, adj_ptr, w, F);
  FB2_LAUNCH_CHECK();
  return OK;
}

template <int TD, int L>
static int launch_matfree(const MatfreeArgs& a, cudaStream_t s) {
  using GEO = A4Geo<TD>;
  A4Tables<L, GEO::NH> tb;
  a4_reduced_table<TD, L>(a.Ms_host, a.Mm_host, tb);
  matfree_cell_kernel<TD, L><<<(unsigned)ceil_div(a.NC, 128), 128, 0, s>>>(a.node, a.cell, a.c2d, a.NC, a.Ms_host ? a.scal_d : 0.0, a.coef_d,
                                                                           a.Mm_host ? a.scal_m : 0.0, a.coef_m, a.u, a.w, a.pair_pos, tb);
  FB2_LAUNCH_CHECK();
  return OK;
}

int matfree_scalar_const(int TD, int p, const MatfreeArgs& a, cudaStream_t s) {
  if (a.NC <= 0) return OK;
  switch (TD * 10 + p) {
    case 21: return launch_matfree<2, 3>(a, s);
    case 22: return launch_matfree<2, 6>(a, s);
    case 23: return launch_matfree<2, 10>(a, s);
    case 31: return launch_matfree<3, 4>(a, s);
    case 32: return launch_matfree<3, 10>(a, s);
    case 33: return launch_matfree<3, 20>(a, s);
    default: return fail(ERR_UNSUPPORTED, "matfree: unsupported element TD=%d p=%d", TD, p);
  }
}

// dof -> (cell, local index) adjacency alone (the first half of sym_count): what a matrix-free product or a load
// vector needs of the symbolic phase
size_t adjacency_workspace_bytes(int64_t gdof) { return align_up((size_t)gdof * 4) * 2 + scan_workspace_bytes(gdof + 1) + 1024; }

int build_adjacency(const int* c2d, int64_t NC, int L, int64_t gdof, int64_t* adj_ptr, int* adj_pair, void* ws, cudaStream_t s) {
  const int64_t npair = NC * L;
  if (npair >= ((int64_t)1 << 31)) return fail(ERR_UNSUPPORTED, "adjacency: NC*ldof=%lld exceeds int32 pair ids", (long long)npair);
  Carver c(ws);
  int* deg = c.take<int>(gdof);
  int* cursor = c.take<int>(gdof);
  void* scan_ws = c.take<char>(scan_workspace_bytes(gdof + 1));
  FB2_CUDA(cudaMemsetAsync(deg, 0, (size_t)gdof * 4, s));
  FB2_CUDA(cudaMemsetAsync(cursor, 0, (size_t)gdof * 4, s));
  if (npair > 0) deg_kernel<<<grid_for(npair), 256, 0, s>>>(c2d, npair, deg);
  FB2_TRY(exclusive_scan_i32(deg, adj_ptr, gdof, true, scan_ws, s));
  if (npair > 0) {
    adj_fill_kernel<<<grid_for(npair), 256, 0, s>>>(c2d, npair, adj_ptr, cursor, adj_pair);
    adj_sort_kernel<<<grid_for(gdof), 256, 0, s>>>(gdof, L, adj_ptr, adj_pair);
  }
  FB2_LAUNCH_CHECK();
  return OK;
}

size_t asm4_workspace_bytes(int ntile) { return align_up((size_t)(ntile + 1) * 4) + scan_workspace_bytes(ntile + 1) + 1024; }

int asm4_plan_count(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int* adj_pair, int L,
                    int64_t* batch_ptr, int64_t* nbatch_host, void* ws, cudaStream_t s) {
  Carver c(ws);
  int* cnt = c.take<int>(ntile + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(ntile + 1));
  if (ntile > 0)
    asm4_schedule_kernel<false><<<(unsigned)ceil_div((int64_t)ntile * 32, 128), 128, 0, s>>>(ntile, blk_row, crow, adj_ptr, adj_pair, L, cnt, nullptr,
                                                                               nullptr, nullptr, nullptr, 0, 1);
  FB2_LAUNCH_CHECK();
  FB2_TRY(exclusive_scan_i32(cnt, batch_ptr, ntile, true, scan_ws, s));
  FB2_CUDA(cudaMemcpyAsync(nbatch_host, batch_ptr + ntile, 8, cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  return OK;
}

int asm4_plan_fill(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int* adj_pair, int L,
                   const int64_t* batch_ptr, unsigned char* batch_i, uint32_t* blocks, const void* slots, int slot_bytes,
                   cudaStream_t s) {
  if (ntile <= 0) return OK;
  const int nwords = slot_stride(L, slot_bytes) * slot_bytes / 4;
  asm4_schedule_kernel<true><<<(unsigned)ceil_div((int64_t)ntile * 32, 128), 128, 0, s>>>(ntile, blk_row, crow, adj_ptr, adj_pair, L, nullptr, batch_ptr,
                                                                            batch_i, blocks,
                                                                            static_cast<const uint32_t*>(slots), nwords, slot_bytes);
  FB2_LAUNCH_CHECK();
  return OK;
}

template <int TD, int L>
static int launch_asm4(Asm4Args a, int slot_bytes, cudaStream_t s) {
  using GEO = A4Geo<TD>;
  static_assert(sizeof(A4Tables<L, GEO::NH>) + sizeof(Asm4Args) < 32000, "element tables exceed the kernel parameter block");
  static_assert(L <= A4_MAXL, "scheduler arrays too small");
  cell_geometry4_kernel<TD><<<(unsigned)ceil_div(a.NC, 256), 256, 0, s>>>(a.node, a.cell, a.NC, a.Ms ? a.scal_d : 0.0, a.coef_d,
                                                                         a.Mm ? a.scal_m : 0.0, a.coef_m, a.Hbuf);
  a.H = a.Hbuf;
  a.acc_stride = (a.tile + a.max_row + 1) & ~1;
  A4Tables<L, GEO::NH> tb;
  a4_reduced_table<TD, L>(a.Ms_host, a.Mm_host, tb);
  if (a.tile + a.max_row >= (1 << A4_BASE_BITS) || a.max_row > A4_MAXROW)
    return fail(ERR_UNSUPPORTED, "assemble v4: tile offsets exceed %d bits (tile=%d max_row=%d)", A4_BASE_BITS, a.tile, a.max_row);
  const unsigned grid = (unsigned)ceil_div(a.ntile, FB2_ASM4_WARPS);
  if (grid == 0) return OK;
  const int words = slot_stride(L, slot_bytes) * slot_bytes / 4;
  // per warp: accumulator tile + A6_RING batch blocks + their mbarriers
  const size_t per_warp = (size_t)a.acc_stride * 8 + (size_t)A6_RING * a4_block_words(words) * 4 + ((A6_RING * 8 + 15) & ~15);
  const size_t smem = (size_t)FB2_ASM4_WARPS * per_warp;
  if (smem > 220 * 1024) return fail(ERR_UNSUPPORTED, "assemble v4: tiles do not fit shared memory (tile=%d max_row=%d)", a.tile, a.max_row);
#define FB2_A4_LAUNCH(KERN)                                                                          \
  do {                                                                                               \
    auto k = KERN;                                                                                   \
    FB2_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    k<<<grid, FB2_ASM4_WARPS * 32, smem, s>>>(a, tb);                                                \
  } while (0)
  if (slot_bytes == 1) FB2_A4_LAUNCH((assemble_const_v6_kernel<TD, L, uint8_t>));
  else FB2_A4_LAUNCH((assemble_const_v6_kernel<TD, L, uint16_t>));
#undef FB2_A4_LAUNCH
  FB2_LAUNCH_CHECK();
  return OK;
}

int assemble_v4(int TD, int p, const Asm4Args& a, int slot_bytes, cudaStream_t s) {
  if (slot_bytes != 1 && slot_bytes != 2) return fail(ERR_INVALID, "assemble v4: slot_bytes must be 1 or 2");
  switch (TD * 10 + p) {
    case 21: return launch_asm4<2, 3>(a, slot_bytes, s);
    case 22: return launch_asm4<2, 6>(a, slot_bytes, s);
    case 23: return launch_asm4<2, 10>(a, slot_bytes, s);
    case 31: return launch_asm4<3, 4>(a, slot_bytes, s);
    case 32: return launch_asm4<3, 10>(a, slot_bytes, s);
    case 33: return launch_asm4<3, 20>(a, slot_bytes, s);
    default: return fail(ERR_UNSUPPORTED, "assemble v4: unsupported element TD=%d p=%d", TD, p);
  }
}

}  // namespace fb2

