#include "sort_scan.cuh"

namespace fb2 {

// =====================================================================================
// exclusive scan (three sweeps: tile sums, scan of tile sums, tile scan + offset)
// =====================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan_i64(int64_t v, int64_t* total, int64_t* smem /*>=33*/) {
  // returns exclusive prefix of v over the block (blockDim.x multiple of 32, <= 1024)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int64_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int64_t w = lane < nw ? smem[lane] : 0;
    int64_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  int64_t res = smem[wid] + incl - v;
  if (total) *total = smem[32];
  __syncthreads();
  return res;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const T* __restrict__ in, int64_t n, int64_t* __restrict__ sums) {
  __shared__ int64_t sm[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t acc = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
    if (i < n) acc += (int64_t)in[i];
  }
  int64_t tot;
  block_exclusive_scan_i64(acc, &tot, sm);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of `sums`
__global__ void __launch_bounds__(1024) scan_sums_inplace(int64_t* sums, int64_t m, int64_t* total_out) {
  __shared__ int64_t sm[33];
  int64_t carry = 0;
  for (int64_t base = 0; base < m; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    int64_t v = i < m ? sums[i] : 0, tot;
    int64_t ex = block_exclusive_scan_i64(v, &tot, sm);
    if (i < m) sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const T* __restrict__ in, int64_t n, const int64_t* __restrict__ sums,
                                                           int64_t* __restrict__ out) {
  __shared__ int64_t sm[33];
  // blocked arrangement through a transposing index so that loads stay coalesced per item slot
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t v[SCAN_ITEMS];
  int64_t acc = 0;
  // thread t owns elements [t*ITEMS, (t+1)*ITEMS) of the tile
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + k;
    v[k] = i < n ? (int64_t)in[i] : 0;
    acc += v[k];
  }
  int64_t ex = block_exclusive_scan_i64(acc, nullptr, sm) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

size_t scan_workspace_bytes(int64_t n) { return align_up((size_t)(ceil_div(n, SCAN_TILE) + 1) * sizeof(int64_t)); }

template <typename T>
static int exclusive_scan_impl(const T* in, int64_t* out, int64_t n, bool with_total, void* ws, cudaStream_t s) {
  if (n <= 0) {
    if (with_total) FB2_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), s));
    return OK;
  }
  int64_t* sums = static_cast<int64_t*>(ws);
  const int64_t nb = ceil_div(n, SCAN_TILE);
  scan_tile_sums<T><<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, sums);
  scan_sums_inplace<<<1, 1024, 0, s>>>(sums, nb, with_total ? out + n : nullptr);
  scan_tiles<T><<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, sums, out);
  FB2_LAUNCH_CHECK();
  return OK;
}
int exclusive_scan_u8(const uint8_t* in, int64_t* out, int64_t n, bool t, void* ws, cudaStream_t s) { return exclusive_scan_impl(in, out, n, t, ws, s); }
int exclusive_scan_u32(const uint32_t* in, int64_t* out, int64_t n, bool t, void* ws, cudaStream_t s) { return exclusive_scan_impl(in, out, n, t, ws, s); }
int exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n, bool t, void* ws, cudaStream_t s) { return exclusive_scan_impl(in, out, n, t, ws, s); }

// =====================================================================================
// radix sort: 8-bit digits, tile = 256 threads x 16 items, stable
// =====================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096
constexpr int RS_WARP_ITEMS = 32 * RS_ITEMS;     // 512 contiguous items per warp
constexpr int RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t mask,
                                                           uint32_t* __restrict__ hist, int64_t nb) {
  __shared__ uint32_t h[RADIX];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    int64_t i = base + (int64_t)k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & mask], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];   // digit-major for the global scan
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                         uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                         int shift, uint32_t mask, const int64_t* __restrict__ ghist, int64_t nb) {
  __shared__ uint32_t cnt[RS_WARPS][RADIX];     // per-warp running digit counts -> exclusive warp bases
  __shared__ uint32_t dig_base[RADIX];          // exclusive scan over digits of the block totals
  __shared__ int64_t gbase[RADIX];              // global base of (digit, this block) minus dig_base
  __shared__ int64_t scan_sm[33];
  extern __shared__ __align__(16) unsigned char rs_dyn[];          // RS_TILE * 12 bytes
  uint64_t* skey = reinterpret_cast<uint64_t*>(rs_dyn);
  uint32_t* sval = reinterpret_cast<uint32_t*>(rs_dyn + (size_t)RS_TILE * 8);

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();

  const int64_t tile0 = (int64_t)blockIdx.x * RS_TILE;
  const int64_t warp0 = tile0 + (int64_t)wid * RS_WARP_ITEMS;
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  uint16_t rank[RS_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = warp0 + r * 32 + lane;
    const bool ok = i < n;
    key[r] = ok ? keys_in[i] : ~0ull;
    val[r] = ok ? (vals_in ? vals_in[i] : (uint32_t)i) : 0u;
    // out-of-range items get digit `mask` handled via `ok` below: they must not be counted
    const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
    const uint32_t act = __ballot_sync(0xffffffffu, ok);
    uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : 0xffffffffu);
    peers &= act;
    uint32_t before = 0;
    if (ok) before = cnt[wid][d];
    __syncwarp();
    if (ok) {
      rank[r] = (uint16_t)(before + __popc(peers & lt));
      if ((peers & lt) == 0) cnt[wid][d] = before + __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // thread d: exclusive scan over warps for digit d, block total
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = cnt[w][d];
      cnt[w][d] = run;
      run += c;
    }
    int64_t ex = block_exclusive_scan_i64((int64_t)run, nullptr, scan_sm);
    dig_base[d] = (uint32_t)ex;
    gbase[d] = ghist[(int64_t)d * nb + blockIdx.x] - ex;
  }
  __syncthreads();
  // local reorder by digit (stable)
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = warp0 + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
      const uint32_t p = dig_base[d] + cnt[wid][d] + rank[r];
      skey[p] = key[r];
      sval[p] = val[r];
    }
  }
  __syncthreads();
  const int64_t rem = n - tile0;
  const int cntTile = rem < RS_TILE ? (int)rem : RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int p = k * RS_THREADS + threadIdx.x;
    if (p < cntTile) {
      const uint64_t kk = skey[p];
      const uint32_t d = (uint32_t)(kk >> shift) & mask;
      const int64_t g = gbase[d] + p;
      keys_out[g] = kk;
      vals_out[g] = sval[p];
    }
  }
}

size_t sort_workspace_bytes(int64_t n) {
  const int64_t nb = ceil_div(n > 0 ? n : 1, RS_TILE);
  size_t b = 0;
  b += align_up((size_t)n * 8);                 // alt keys
  b += align_up((size_t)n * 4);                 // alt vals
  b += align_up((size_t)RADIX * nb * 4);        // hist
  b += align_up((size_t)(RADIX * nb + 1) * 8);  // scanned hist
  b += scan_workspace_bytes((int64_t)RADIX * nb);
  return b + 1024;
}

int radix_sort_pairs(uint64_t* keys, uint32_t* vals, int64_t n, int nbits, void* ws, cudaStream_t s,
                     uint64_t** keys_out, uint32_t** vals_out) {
  // `vals` must be a valid buffer of n u32.  Entry convention: *vals_out == nullptr means the
  // payload is the original position (synthesised in the first pass, `vals` content ignored).
  if (n >= ((int64_t)1 << 32)) return fail(ERR_UNSUPPORTED, "radix_sort_pairs: n=%lld exceeds u32 payload range", (long long)n);
  const int64_t nb = ceil_div(n > 0 ? n : 1, RS_TILE);
  Carver c(ws);
  uint64_t* kalt = c.take<uint64_t>(n);
  uint32_t* valt = c.take<uint32_t>(n);
  uint32_t* hist = c.take<uint32_t>((size_t)RADIX * nb);
  int64_t* ghist = c.take<int64_t>((size_t)RADIX * nb + 1);
  void* scan_ws = c.take<char>(scan_workspace_bytes((int64_t)RADIX * nb));
  uint64_t *kin = keys, *kout = kalt;
  uint32_t *vin = vals, *vout = valt;
  const bool identity = (*vals_out == nullptr);  // entry flag: payload = original position
  bool first = true;
  static bool attr_set = false;
  if (!attr_set) {
    FB2_CUDA(cudaFuncSetAttribute(rs_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_TILE * 12));
    attr_set = true;
  }
  if (nbits <= 0) return fail(ERR_INVALID, "radix_sort_pairs: nbits must be > 0");
  if (n > 0) {
    for (int shift = 0; shift < nbits; shift += 8) {
      const int bits = (nbits - shift) < 8 ? (nbits - shift) : 8;
      const uint32_t mask = (1u << bits) - 1u;
      rs_histogram<<<(unsigned)nb, RS_THREADS, 0, s>>>(kin, n, shift, mask, hist, nb);
      FB2_TRY(exclusive_scan_u32(hist, ghist, (int64_t)RADIX * nb, false, scan_ws, s));
      rs_scatter<<<(unsigned)nb, RS_THREADS, RS_TILE * 12, s>>>(kin, (first && identity) ? nullptr : vin, kout, vout, n, shift, mask, ghist, nb);
      FB2_LAUNCH_CHECK();
      first = false;
      uint64_t* tk = kin; kin = kout; kout = tk;
      uint32_t* tv = vin; vin = vout; vout = tv;
    }
  }
  *keys_out = kin;
  *vals_out = vin;
  return OK;
}

}  // namespace fb2
