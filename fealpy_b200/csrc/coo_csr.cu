// K2: deterministic COO -> CSR.  Sort by global (row, col) with a stable radix sort, mark
// segment heads, scan, then reduce every segment left-to-right in original COO order --
// the same order FEALPy's numpy backend produces (sparse/coo_tensor.py:184-213 coalesce:
// lexsort + np.add.at; sparse/coo_tensor.py:137-157 tocsr), so values are reproducible
// and the pattern (explicit zeros included) is bit-identical.  No floating-point atomics.
#include "common.cuh"
#include "sort_scan.cuh"
#include "coo_csr.cuh"

namespace fb2 {

// key = row << cbits | col for the e-th COO entry of a block of element matrices:
// e = (c*Lr + i)*Lc + j, row = rdof[c][i], col = cdof[c][j]  (fem/bilinear_form.py:69-72)
__global__ void __launch_bounds__(256) keys_from_c2d_kernel(const int* __restrict__ rdof, const int* __restrict__ cdof, int64_t NC,
                                                            int Lr, int Lc, int cbits, uint64_t* __restrict__ keys) {
  const int64_t n = NC * Lr * Lc;
  const int LL = Lr * Lc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = e / LL;
    const int ij = (int)(e - c * LL);
    const int i = ij / Lc, j = ij - i * Lc;
    const uint64_t r = (uint32_t)rdof[c * Lr + i], cc = (uint32_t)cdof[c * Lc + j];
    keys[e] = (r << cbits) | cc;
  }
}

template <typename IT>
__global__ void __launch_bounds__(256) keys_from_coo_kernel(const IT* __restrict__ row, const IT* __restrict__ col, int64_t n, int cbits,
                                                            uint64_t* __restrict__ keys) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    keys[e] = ((uint64_t)row[e] << cbits) | (uint64_t)col[e];
}

__global__ void __launch_bounds__(256) mark_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, uint8_t* __restrict__ head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// col[slot], seg_start[slot] for every segment head; seg_start[nnz] = n
template <typename CT>
__global__ void __launch_bounds__(256) fill_pattern_kernel(const uint64_t* __restrict__ keys, const uint8_t* __restrict__ head,
                                                           const int64_t* __restrict__ S, int64_t n, int cbits, CT* __restrict__ col,
                                                           int64_t* __restrict__ seg_start) {
  const uint64_t cmask = (1ull << cbits) - 1ull;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i == n) { seg_start[S[n]] = n; continue; }
    if (head[i]) {
      const int64_t s = S[i];
      col[s] = (CT)(keys[i] & cmask);
      seg_start[s] = i;
    }
  }
}

// crow[r] = number of unique keys whose row < r  (empty rows handled by the search)
__global__ void __launch_bounds__(256) crow_kernel(const uint64_t* __restrict__ keys, const int64_t* __restrict__ S, int64_t n, int cbits,
                                                   int64_t nrow, int64_t* __restrict__ crow) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nrow; r += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t target = (uint64_t)r << cbits;
    int64_t lo = 0, hi = n;     // first index with keys[idx] >= target
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    crow[r] = S[lo];
  }
}

// one thread per output slot; strictly sequential sum over the segment (reference order)
__global__ void __launch_bounds__(256) segment_reduce_kernel(const uint32_t* __restrict__ perm, const int64_t* __restrict__ seg_start,
                                                             int64_t nnz, const double* __restrict__ vin, double* __restrict__ vout) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nnz; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = seg_start[s], e = seg_start[s + 1];
    double acc = 0.0;
    for (int64_t k = b; k < e; ++k) acc += vin[perm[k]];
    vout[s] = acc;
  }
}

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = ceil_div(n, threads);
  const int64_t cap = (int64_t)kNumSM * 32;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

int coo_keys_from_c2d(const int* rdof, const int* cdof, int64_t NC, int Lr, int Lc, int cbits, uint64_t* keys, cudaStream_t s) {
  if (NC <= 0) return OK;
  keys_from_c2d_kernel<<<grid_for(NC * Lr * Lc), 256, 0, s>>>(rdof, cdof, NC, Lr, Lc, cbits, keys);
  FB2_LAUNCH_CHECK();
  return OK;
}

int coo_keys_from_coo(const void* row, const void* col, int index_bytes, int64_t n, int cbits, uint64_t* keys, cudaStream_t s) {
  if (n <= 0) return OK;
  if (index_bytes == 4)
    keys_from_coo_kernel<int32_t><<<grid_for(n), 256, 0, s>>>((const int32_t*)row, (const int32_t*)col, n, cbits, keys);
  else if (index_bytes == 8)
    keys_from_coo_kernel<int64_t><<<grid_for(n), 256, 0, s>>>((const int64_t*)row, (const int64_t*)col, n, cbits, keys);
  else
    return fail(ERR_INVALID, "coo_keys_from_coo: index_bytes must be 4 or 8");
  FB2_LAUNCH_CHECK();
  return OK;
}

size_t coo_symbolic_workspace_bytes(int64_t n) {
  return sort_workspace_bytes(n) + align_up((size_t)n) /*head*/ + align_up((size_t)(n + 1) * 8) /*S*/ + scan_workspace_bytes(n) + 1024;
}

// sorts keys (in place, with perm as payload = original position), leaves in the workspace
// the head flags and their exclusive scan; returns nnz through a synchronous copy.
int coo_symbolic(uint64_t* keys, uint32_t* perm, int64_t n, int nbits, void* ws, int64_t* nnz_host, cudaStream_t s) {
  if (n <= 0) { *nnz_host = 0; return OK; }
  Carver c(ws);
  void* sort_ws = c.take<char>(sort_workspace_bytes(n));
  uint8_t* head = c.take<uint8_t>(n);
  int64_t* S = c.take<int64_t>(n + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes(n));
  uint64_t* ks = nullptr;
  uint32_t* ps = nullptr;   // nullptr on entry = identity payload
  FB2_TRY(radix_sort_pairs(keys, perm, n, nbits, sort_ws, s, &ks, &ps));
  if (ks != keys) {
    FB2_CUDA(cudaMemcpyAsync(keys, ks, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    FB2_CUDA(cudaMemcpyAsync(perm, ps, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  }
  mark_heads_kernel<<<grid_for(n), 256, 0, s>>>(keys, n, head);
  FB2_LAUNCH_CHECK();
  FB2_TRY(exclusive_scan_u8(head, S, n, true, scan_ws, s));
  FB2_CUDA(cudaMemcpyAsync(nnz_host, S + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  return OK;
}

int coo_fill(const uint64_t* keys, int64_t n, int cbits, int64_t nrow, void* ws, int64_t* crow, void* col, int col_bytes,
             int64_t* seg_start, cudaStream_t s) {
  if (n <= 0) {
    FB2_CUDA(cudaMemsetAsync(crow, 0, (size_t)(nrow + 1) * 8, s));
    FB2_CUDA(cudaMemsetAsync(seg_start, 0, 8, s));
    return OK;
  }
  Carver c(ws);
  c.take<char>(sort_workspace_bytes(n));
  uint8_t* head = c.take<uint8_t>(n);
  int64_t* S = c.take<int64_t>(n + 1);
  if (col_bytes == 4)
    fill_pattern_kernel<int32_t><<<grid_for(n + 1), 256, 0, s>>>(keys, head, S, n, cbits, (int32_t*)col, seg_start);
  else if (col_bytes == 8)
    fill_pattern_kernel<int64_t><<<grid_for(n + 1), 256, 0, s>>>(keys, head, S, n, cbits, (int64_t*)col, seg_start);
  else
    return fail(ERR_INVALID, "coo_fill: col_bytes must be 4 or 8");
  crow_kernel<<<grid_for(nrow + 1), 256, 0, s>>>(keys, S, n, cbits, nrow, crow);
  FB2_LAUNCH_CHECK();
  return OK;
}

int coo_reduce(const uint32_t* perm, const int64_t* seg_start, int64_t nnz, const double* vin, double* vout, cudaStream_t s) {
  if (nnz <= 0) return OK;
  segment_reduce_kernel<<<grid_for(nnz), 256, 0, s>>>(perm, seg_start, nnz, vin, vout);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2
