// K3/K4: CSR SpMV with a fused, deterministic p.Ap reduction and the fused CG vector
// updates; the whole iteration is three kernels with all scalars resident on the device,
// captured into a CUDA graph in chunks so that the host only polls a convergence flag.
//
// Reference semantics reproduced (solver/cg.py:76-123): zero-rhs early return, stop test on
// sqrt(r.z) with '<' (atol first, then rtol*|b|, then maxit), x of the stopping iteration is
// returned, z = M r with M a diagonal (Jacobi) preconditioner or identity.
// SpMV semantics: sparse/csr_tensor.py:411-452 -> backend csr_spmm (numpy_backend.py:180-199).
#include "cg.cuh"

#include <climits>
#include <cstdlib>
#include <cmath>

namespace fb2 {

constexpr int CG_THREADS = 256;

// ---- deterministic grid reduction ("last block finishes", fixed summation order) ---------
template <typename Finish>
__device__ __forceinline__ void grid_reduce(double v, double* partials, unsigned int* counter, Finish finish) {
  __shared__ double red[CG_THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    const unsigned int ticket = atomicInc(counter, gridDim.x - 1);   // wraps back to 0 for the next use
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += (int)blockDim.x) t += __ldcg(partials + i);
  t = warp_sum(t);
  __syncthreads();
  if (lane == 0) red[wid] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    finish(tot);
  }
}

// ---- SpMV: T lanes per row --------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(CG_THREADS) spmv_kernel(int64_t n, const int64_t* __restrict__ crow, const int32_t* __restrict__ col,
                                                          const double* __restrict__ val, const double* __restrict__ x,
                                                          double* __restrict__ y, const double* __restrict__ b, int mode,
                                                          double* dot_out, double* partials, unsigned int* counter,
                                                          const CgScalars* sc, OwnRange own) {
  if (sc && sc->done) return;
  constexpr int RPB = CG_THREADS / T;
  const int sub = threadIdx.x % T, rib = threadIdx.x / T;
  double dsum = 0.0;
  const int64_t ntile = (n + RPB - 1) / RPB;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t r = tile * RPB + rib;
    double acc = 0.0;
    if (r < n) {
      const int64_t s = crow[r], e = crow[r + 1];
      for (int64_t k = s + sub; k < e; k += T) acc += ld_stream(val + k) * x[ld_stream(col + k)];
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (r < n && sub == 0) {
      const double yv = mode ? b[r] - acc : acc;
      y[r] = yv;
      if (dot_out && own.has(r)) dsum += x[r] * yv;
    }
  }
  if (dot_out) grid_reduce(dsum, partials, counter, [=](double tot) { *dot_out = tot; });
}

template <int T, int NB>
__global__ void __launch_bounds__(CG_THREADS) spmm_kernel(int64_t n, const int64_t* __restrict__ crow, const int32_t* __restrict__ col,
                                                          const double* __restrict__ val, const double* __restrict__ X,
                                                          double* __restrict__ Y, int nb, int b0) {
  constexpr int RPB = CG_THREADS / T;
  const int sub = threadIdx.x % T, rib = threadIdx.x / T;
  const int64_t ntile = (n + RPB - 1) / RPB;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t r = tile * RPB + rib;
    double acc[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) acc[k] = 0.0;
    if (r < n) {
      const int64_t s = crow[r], e = crow[r + 1];
      for (int64_t k = s + sub; k < e; k += T) {
        const double v = val[k];
        const double* xr = X + (int64_t)col[k] * nb + b0;
#pragma unroll
        for (int q = 0; q < NB; ++q)
          if (b0 + q < nb) acc[q] += v * xr[q];
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int o = T / 2; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    if (r < n && sub == 0)
#pragma unroll
      for (int q = 0; q < NB; ++q)
        if (b0 + q < nb) Y[r * nb + b0 + q] = acc[q];
  }
}

// ---- SpMV, "row-aligned nnz tiles": CTA b owns the rows whose first entry index lies in
// [b*tile, (b+1)*tile).  Phase 1 streams the tile's (val, col) with perfectly coalesced, deeply
// unrolled loads (memory-level parallelism independent of the row lengths), gathers x and
// parks the products in shared memory; phase 2 reduces each row with G lanes.  Row sums and
// the fused p.Ap partials are formed in a fixed order -> bitwise reproducible.
// Round-2 experiments on this kernel, all measured SLOWER on tet P2 128^3 and removed again (profiles/r02_tune_spmv.txt;
// the kernel is bound by the latency / L2 sector traffic of the x gather -- DRAM 58 % busy, issue slots 23 %):
//   * 16-bit window-relative column stream (10 instead of 12 bytes per nonzero): 1.61 vs 1.66 ms per CG iteration (-3 %);
//   * per-tile list of distinct columns, x staged in shared memory, 16-bit positions: 2.37 ms (a third dependent phase);
//   * (val, col) tiles fetched by cp.async.bulk + mbarrier into a double buffer, products in place: 1.95 - 3.4 ms
//     depending on threads / tile (fewer gathers in flight than 8 resident CTAs x 256 threads of this kernel);
//   * SM-chunked tile map (co-resident CTAs sweep adjacent tiles, for L1 reuse of x): 1.73 vs 1.67 ms;
//   * the per-tile chain of dependent loads (blk_row -> crow[r0] -> stream -> gather -> barrier -> crow[r] -> x[r]) cut to
//     stream -> gather by issuing the other loads one phase early: with plain loads 1.616 vs 1.612 ms (no change: eight
//     resident CTAs already hide the chain), with 4/8-byte cp.async into shared memory 2.56 ms (small LDGSTS that miss to
//     DRAM park in the LSU queue: mio_throttle 15.8 per issue);
//   * bound experiment (tools/gpu_spmv_bound.py, SpMV alone): real columns 1.356 ms, a row's columns consecutive 1.249 ms,
//     every column = 0 (free gather) 1.163 ms, dofs renumbered along a Morton curve 1.380 / 1.432 ms -- the gather costs
//     14 % of the kernel, no DOF reordering can recover more than that, and the Z-curve is worse than the lexicographic
//     numbering of from_box.  The remaining gap to a plain copy (0.92 ms for the same bytes) is the stream -> shared memory
//     -> row-sum structure itself.
#ifndef FB2_ST_UNROLL
#define FB2_ST_UNROLL 4
#endif
constexpr int ST_UNROLL = FB2_ST_UNROLL;
#ifndef FB2_ST_THREADS
#define FB2_ST_THREADS 256
#endif
constexpr int ST_THREADS = FB2_ST_THREADS;     // threads per CTA of the streaming SpMV kernel

// (2048 / ST_THREADS resident CTAs: the kernel lives on occupancy -- at 40 registers (6 CTAs) it ran 2.17 instead of 1.65 ms)
template <int G>
__global__ void __launch_bounds__(ST_THREADS, 2048 / ST_THREADS) spmv_stream_kernel(int64_t n, const int64_t* __restrict__ crow,
                                                                 const int32_t* __restrict__ col, const double* __restrict__ val,
                                                                 const double* __restrict__ x, double* __restrict__ y,
                                                                 const double* __restrict__ b, int mode,
                                                                 const int32_t* __restrict__ blk_row, int nblk, double* dot_out,
                                                                 double* partials, unsigned int* counter, const CgScalars* sc,
                                                                 OwnRange own, const int32_t* __restrict__ blk_end, HaloWait hw) {
  if (sc && sc->done) return;
  extern __shared__ __align__(16) double prod[];
  const int tid = threadIdx.x;
  const int g = tid % G, grp = tid / G;
  constexpr int NGRP = ST_THREADS / G;
  double dsum = 0.0;
  bool waited = hw.nnb == 0;
  // L2 policies (profiles/r02_tune_spmv.txt, call h24): gathering x with evict_last keeps it in L2 while 5.8 GB of matrix
  // stream past (config 4: 0.861 -> 0.837 ms per CG iteration, config 2: 1.604 -> 1.591, config 3 unchanged); marking the
  // (val, col) stream evict_first on top of that, or alone, measured no better (knob kept for tuning builds)
#ifdef FB2_STREAM_L2_EVICT_FIRST
  const uint64_t pol_s = l2_policy_evict_first();
#define FB2_LDS(p) ld_stream((p), pol_s)
#else
#define FB2_LDS(p) ld_stream(p)
#endif
#ifndef FB2_GATHER_L2_DEFAULT
  const uint64_t pol_x = l2_policy_evict_last();
#define FB2_LDX(p) ld_hint((p), pol_x)
#else
#define FB2_LDX(p) (*(p))
#endif
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    if (!waited && blk >= hw.first_tile) {
      // tiles from hw.first_tile on hold boundary rows: they read halo entries of x that the neighbours push over NVLink
      // (csrc/peer.cu).  The grid is persistent and fully resident, the flags are raised by OTHER GPUs: spinning is safe.
      if (tid == 0) {
        const unsigned long long seq = *hw.epoch + (unsigned long long)sc->niter;
        for (int k = 0; k < hw.nnb; ++k) {
          const volatile unsigned long long* f = hw.flags + (k == 0 ? hw.nb0 : hw.nb1);
          while (*f < seq) { }
        }
        __threadfence_system();
      }
      __syncthreads();
      waited = true;
    }
    const int r0 = blk_row[blk], r1 = blk_end ? blk_end[blk] : blk_row[blk + 1];
    if (r0 == r1) continue;
    const int64_t v0 = crow[r0];
    const int nval = (int)(crow[r1] - v0);
    const double* __restrict__ vp = val + v0;
    const int32_t* __restrict__ cp = col + v0;
    {
      int k = tid;
      for (; k + (ST_UNROLL - 1) * ST_THREADS < nval; k += ST_UNROLL * ST_THREADS) {
        double vv[ST_UNROLL];
        int cc[ST_UNROLL];
#pragma unroll
        for (int u = 0; u < ST_UNROLL; ++u) { vv[u] = FB2_LDS(vp + k + u * ST_THREADS); cc[u] = FB2_LDS(cp + k + u * ST_THREADS); }
#pragma unroll
        for (int u = 0; u < ST_UNROLL; ++u) prod[k + u * ST_THREADS] = vv[u] * FB2_LDX(x + cc[u]);
      }
      for (; k < nval; k += ST_THREADS) prod[k] = FB2_LDS(vp + k) * FB2_LDX(x + FB2_LDS(cp + k));
    }
    __syncthreads();
    for (int base = r0; base < r1; base += NGRP) {
      const int r = base + grp;
      double acc = 0.0;
      if (r < r1) {
        const int s = (int)(crow[r] - v0), e = (int)(crow[r + 1] - v0);
        for (int q = s + g; q < e; q += G) acc += prod[q];
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (r < r1 && g == 0) {
        const double yv = mode ? b[r] - acc : acc;
        y[r] = yv;
        if (dot_out && own.has(r)) dsum += x[r] * yv;
      }
    }
    __syncthreads();
  }
  if (dot_out) grid_reduce(dsum, partials, counter, [=](double tot) { *dot_out = tot; });
}

__global__ void __launch_bounds__(256) partition_rows_kernel(const int64_t* __restrict__ crow, int64_t n, int tile, int nblk,
                                                             int32_t* __restrict__ blk_row) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nblk) return;
  if (b == nblk) { blk_row[b] = (int32_t)n; return; }
  const int64_t target = crow[0] + (int64_t)b * tile;        // crow may point into the middle of a matrix (a row range)
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (crow[mid] < target) lo = mid + 1; else hi = mid;
  }
  blk_row[b] = (int32_t)lo;
}

__global__ void __launch_bounds__(256) max_row_kernel(const int64_t* __restrict__ crow, int64_t n, int* out) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (int)(crow[i + 1] - crow[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void __launch_bounds__(CG_THREADS) dot_kernel(int64_t n, const double* __restrict__ a, const double* __restrict__ b,
                                                         double* out, double* partials, unsigned int* counter) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += a[i] * b[i];
  grid_reduce(acc, partials, counter, [=](double tot) { *out = tot; });
}

// p = z = M r ; rTr = r.z
__global__ void __launch_bounds__(CG_THREADS) cg_start_kernel(int64_t n, const double* __restrict__ r, const double* __restrict__ minv,
                                                              double* __restrict__ p, CgScalars* sc, double* partials, OwnRange own) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i], zi = minv ? minv[i] * ri : ri;
    p[i] = zi;
    if (own.has(i)) acc += ri * zi;
  }
  grid_reduce(acc, partials, &sc->counter[1], [=](double tot) { sc->rTr = tot; });
}

__device__ __forceinline__ void cg_finalize_dev(CgScalars* sc) {
  // solver/cg.py:97-121
  const double rn = sqrt(sc->rTr_new);
  sc->rnorm = rn;
  const int it = sc->niter + 1;
  sc->niter = it;
  if (rn < sc->atol || rn < sc->rtol * sc->bnorm || it >= sc->maxit) {
    sc->done = 1;
  } else {
    sc->beta = sc->rTr_new / sc->rTr;
    sc->rTr = sc->rTr_new;
  }
}

__global__ void __launch_bounds__(CG_THREADS) cg_update_xr_kernel(int64_t n, double* __restrict__ x, double* __restrict__ r,
                                                                  const double* __restrict__ p, const double* __restrict__ Ap,
                                                                  const double* __restrict__ minv, CgScalars* sc, double* partials,
                                                                  int fuse_finalize, OwnRange own) {
  if (sc->done) return;
  const double alpha = sc->rTr / sc->pAp;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = x[i] + alpha * p[i];
    const double ri = r[i] - alpha * Ap[i];
    r[i] = ri;
    if (own.has(i)) acc += ri * (minv ? minv[i] * ri : ri);
  }
  grid_reduce(acc, partials, &sc->counter[1], [=](double tot) {
    sc->rTr_new = tot;
    sc->alpha = alpha;
    if (fuse_finalize) cg_finalize_dev(sc);
  });
}

__global__ void cg_finalize_kernel(CgScalars* sc) {
  if (sc->done) return;
  cg_finalize_dev(sc);
}

__global__ void __launch_bounds__(CG_THREADS) cg_update_p_kernel(int64_t n, double* __restrict__ p, const double* __restrict__ r,
                                                                 const double* __restrict__ minv, const CgScalars* sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i];
    p[i] = (minv ? minv[i] * ri : ri) + beta * p[i];
  }
}

__global__ void cg_init_scalars_kernel(CgScalars* sc, double atol, double rtol, int maxit) {
  sc->rTr = sc->pAp = sc->rTr_new = sc->rnorm = sc->alpha = sc->beta = 0.0;
  sc->niter = 0;
  sc->done = 0;
  sc->maxit = maxit;
  sc->atol = atol;
  sc->rtol = rtol;
  for (int i = 0; i < 4; ++i) sc->counter[i] = 0;
}

// ---- host side --------------------------------------------------------------------------
static inline int vec_grid(int64_t n) {
  int64_t b = ceil_div(n, CG_THREADS * 4);
  const int64_t cap = kNumSM * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int pick_T(int64_t n, int64_t nnz) {
  const double avg = n > 0 ? (double)nnz / (double)n : 1.0;
  if (avg <= 3.0) return 2;
  if (avg <= 6.0) return 4;
  if (avg <= 24.0) return 8;
  if (avg <= 96.0) return 16;
  return 32;
}

template <int T>
static void launch_spmv(int64_t n, const int64_t* crow, const int32_t* col, const double* val, const double* x, double* y,
                        const double* b, int mode, double* dot_out, double* partials, unsigned int* counter, const CgScalars* sc,
                        cudaStream_t s, OwnRange own) {
  constexpr int RPB = CG_THREADS / T;
  int64_t nbk = ceil_div(n, RPB);
  const int64_t cap = dot_out ? (int64_t)CG_PARTIALS : (int64_t)kNumSM * 64;
  if (nbk > cap) nbk = cap;
  if (nbk > kNumSM * 16 && dot_out) nbk = kNumSM * 16;
  spmv_kernel<T><<<(unsigned)(nbk < 1 ? 1 : nbk), CG_THREADS, 0, s>>>(n, crow, col, val, x, y, b, mode, dot_out, partials, counter, sc, own);
}

int spmv_plan_blocks(int64_t nnz, int tile) { return (int)ceil_div(nnz > 0 ? nnz : 1, tile); }

// blk_row (nblk+1 int32) and the longest row; *max_row_dev is a device int (zeroed here)
int spmv_plan_build(int64_t n, const int64_t* crow, int tile, int nblk, int32_t* blk_row, int* max_row_dev, cudaStream_t s) {
  if (n >= ((int64_t)1 << 31)) return fail(ERR_UNSUPPORTED, "spmv_plan: more than 2^31 rows");
  FB2_CUDA(cudaMemsetAsync(max_row_dev, 0, sizeof(int), s));
  partition_rows_kernel<<<(unsigned)ceil_div(nblk + 1, 256), 256, 0, s>>>(crow, n, tile, nblk, blk_row);
  if (n > 0) max_row_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSM * 8), 256, 0, s>>>(crow, n, max_row_dev);
  FB2_LAUNCH_CHECK();
  return OK;
}

static int spmv_stream_launch(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* val, const double* x,
                              double* y, const double* b, int mode, const SpmvPlan& plan, double* dot_out, double* partials,
                              unsigned int* counter, const CgScalars* sc, cudaStream_t s, OwnRange own) {
  const size_t smem = (size_t)(plan.tile + plan.max_row) * sizeof(double);
  const double avg = n > 0 ? (double)nnz / (double)n : 1.0;
  static const int per_sm_cap = [] { const char* e = getenv("FB2_SPMV_PERSM"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 2048 / ST_THREADS; }();
  const int per_sm = (int)std::min<size_t>(per_sm_cap, (200 * 1024) / (smem + 1024));
  int grid = std::min(plan.nblk, kNumSM * std::max(per_sm, 1));
  if (grid > CG_PARTIALS) grid = CG_PARTIALS;
  if (grid < 1) grid = 1;
#define FB2_ST(GV)                                                                                             \
  do {                                                                                                         \
    auto kern = spmv_stream_kernel<GV>;                                                                        \
    if (smem > 48 * 1024) FB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, ST_THREADS, smem, s>>>(n, crow, col, val, x, y, b, mode, plan.blk_row, plan.nblk, dot_out, partials, counter, sc, own, plan.blk_end, plan.halo); \
  } while (0)
  // lanes per row in the reduce phase: ONE (a thread sums its row sequentially out of shared memory) measured
  // best by a wide margin -- 1.60 ms/iteration against 1.92 with 8 lanes + shuffles (profiles/r01_tune_spmv.txt)
#ifdef FB2_SPMV_G
  FB2_ST(FB2_SPMV_G);
#else
  if (avg <= 160.0) FB2_ST(1);
  else FB2_ST(4);
#endif
#undef FB2_ST
  FB2_LAUNCH_CHECK();
  return OK;
}

static int spmv_impl(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* val, const double* x, double* y,
                     const double* b, int mode, double* dot_out, double* partials, unsigned int* counter, const CgScalars* sc,
                     cudaStream_t s, const SpmvPlan* plan = nullptr, OwnRange own = OwnRange{}) {
  if (n <= 0) return OK;
  if (own.hi0 == 0 && own.hi1 == 0) own.hi0 = n;      // default: every row contributes to the dot
  if (plan && plan->blk_row && (size_t)(plan->tile + plan->max_row) * 8 <= 200 * 1024)
    return spmv_stream_launch(n, nnz, crow, col, val, x, y, b, mode, *plan, dot_out, partials, counter, sc, s, own);
  switch (pick_T(n, nnz)) {
    case 2: launch_spmv<2>(n, crow, col, val, x, y, b, mode, dot_out, partials, counter, sc, s, own); break;
    case 4: launch_spmv<4>(n, crow, col, val, x, y, b, mode, dot_out, partials, counter, sc, s, own); break;
    case 8: launch_spmv<8>(n, crow, col, val, x, y, b, mode, dot_out, partials, counter, sc, s, own); break;
    case 16: launch_spmv<16>(n, crow, col, val, x, y, b, mode, dot_out, partials, counter, sc, s, own); break;
    default: launch_spmv<32>(n, crow, col, val, x, y, b, mode, dot_out, partials, counter, sc, s, own); break;
  }
  FB2_LAUNCH_CHECK();
  return OK;
}

// workspace layout shared by all entry points: [partials | counters | ... ]
struct PartialWs {
  double* partials;
  unsigned int* counter;
  explicit PartialWs(void* ws) {
    partials = static_cast<double*>(ws);
    counter = reinterpret_cast<unsigned int*>(partials + CG_PARTIALS);
  }
  static size_t bytes() { return align_up(CG_PARTIALS * sizeof(double) + 64); }
};

// values per SpMV tile inside fb2_cg (FB2_SPMV_TILE overrides, for tuning)
static int cg_tile() {
  static int t = [] { const char* e = getenv("FB2_SPMV_TILE"); const int v = e ? atoi(e) : 0; return v >= 256 ? v : 2560; }();
  return t;
}
#define CG_TILE cg_tile()

size_t cg_workspace_bytes(int64_t n, int64_t nnz) {
  return PartialWs::bytes() + align_up(sizeof(CgScalars)) + 3 * align_up((size_t)n * sizeof(double)) +
         align_up((size_t)(spmv_plan_blocks(nnz, CG_TILE) + 2) * sizeof(int32_t)) + 1024;
}

int spmv(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* val, const double* x, double* y,
         const double* b, int mode, double* dot_out, void* partial_ws, cudaStream_t s, const SpmvPlan* plan, OwnRange own,
         const CgScalars* sc) {
  if (dot_out && !partial_ws) return fail(ERR_INVALID, "spmv: fused dot needs the (zero-initialised) partial workspace");
  PartialWs pw(partial_ws);
  return spmv_impl(n, nnz, crow, col, val, x, y, b, mode, dot_out, partial_ws ? pw.partials : nullptr,
                   partial_ws ? pw.counter : nullptr, sc, s, plan, own);
}
size_t partial_workspace_bytes() { return PartialWs::bytes(); }

int spmm(int64_t n, const int64_t* crow, const int32_t* col, const double* val, const double* X, double* Y, int nb, cudaStream_t s) {
  if (n <= 0 || nb <= 0) return OK;
  int64_t nbk = ceil_div(n, CG_THREADS / 8);
  if (nbk > (int64_t)kNumSM * 64) nbk = (int64_t)kNumSM * 64;
  // A is streamed once per group of right-hand sides: 8 columns per pass (one 64-byte run of X per nonzero), 4 when nb <= 4
  if (nb <= 4) spmm_kernel<8, 4><<<(unsigned)nbk, CG_THREADS, 0, s>>>(n, crow, col, val, X, Y, nb, 0);
  else
    for (int b0 = 0; b0 < nb; b0 += 8) spmm_kernel<8, 8><<<(unsigned)nbk, CG_THREADS, 0, s>>>(n, crow, col, val, X, Y, nb, b0);
  FB2_LAUNCH_CHECK();
  return OK;
}

int dot(int64_t n, const double* a, const double* b, double* out, void* partial_ws, cudaStream_t s) {
  PartialWs pw(partial_ws);
  dot_kernel<<<vec_grid(n), CG_THREADS, 0, s>>>(n, a, b, out, pw.partials, pw.counter);
  FB2_LAUNCH_CHECK();
  return OK;
}

int cg_init_scalars(CgScalars* sc, double atol, double rtol, int maxit, cudaStream_t s) {
  cg_init_scalars_kernel<<<1, 1, 0, s>>>(sc, atol, rtol, maxit);
  FB2_LAUNCH_CHECK();
  return OK;
}
int cg_start(int64_t n, const double* r, const double* minv, double* p, CgScalars* sc, void* partial_ws, OwnRange own,
             cudaStream_t s) {
  PartialWs pw(partial_ws);
  if (own.hi0 == 0 && own.hi1 == 0) own.hi0 = n;
  cg_start_kernel<<<vec_grid(n), CG_THREADS, 0, s>>>(n, r, minv, p, sc, pw.partials, own);
  FB2_LAUNCH_CHECK();
  return OK;
}
int cg_update_xr(int64_t n, double* x, double* r, const double* p, const double* Ap, const double* minv, CgScalars* sc,
                 void* partial_ws, int fuse_finalize, cudaStream_t s, OwnRange own) {
  PartialWs pw(partial_ws);
  if (own.hi0 == 0 && own.hi1 == 0) own.hi0 = n;
  cg_update_xr_kernel<<<vec_grid(n), CG_THREADS, 0, s>>>(n, x, r, p, Ap, minv, sc, pw.partials, fuse_finalize, own);
  FB2_LAUNCH_CHECK();
  return OK;
}
int cg_finalize(CgScalars* sc, cudaStream_t s) {
  cg_finalize_kernel<<<1, 1, 0, s>>>(sc);
  FB2_LAUNCH_CHECK();
  return OK;
}
int cg_update_p(int64_t n, double* p, const double* r, const double* minv, CgScalars* sc, cudaStream_t s) {
  cg_update_p_kernel<<<vec_grid(n), CG_THREADS, 0, s>>>(n, p, r, minv, sc);
  FB2_LAUNCH_CHECK();
  return OK;
}

int cg_solve(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* val, const double* b, double* x,
             const double* minv, double atol, double rtol, int maxit, int chunk, void* ws, int* niter_out, double* resid_out,
             cudaStream_t user_stream, const SpmvPlan* prebuilt) {
  *niter_out = 0;
  *resid_out = 0.0;
  if (n <= 0) return OK;
  // graph capture is illegal on the legacy default stream: borrow a private stream, ordered after it
  cudaStream_t s = user_stream;
  if (user_stream == nullptr || user_stream == cudaStreamLegacy || user_stream == cudaStreamPerThread) {
    // one private stream + event per (thread, device): a stream is bound to the device that was current at creation
    constexpr int kMaxDev = 64;
    static thread_local cudaStream_t priv_of[kMaxDev] = {};
    static thread_local cudaEvent_t ev_of[kMaxDev] = {};
    int dev = 0;
    FB2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return fail(ERR_UNSUPPORTED, "cg: device ordinal %d out of range", dev);
    if (!priv_of[dev]) {
      FB2_CUDA(cudaStreamCreateWithFlags(&priv_of[dev], cudaStreamNonBlocking));
      FB2_CUDA(cudaEventCreateWithFlags(&ev_of[dev], cudaEventDisableTiming));
    }
    FB2_CUDA(cudaEventRecord(ev_of[dev], user_stream));
    FB2_CUDA(cudaStreamWaitEvent(priv_of[dev], ev_of[dev], 0));
    s = priv_of[dev];
  }
  Carver c(ws);
  char* pws = c.take<char>(PartialWs::bytes());
  CgScalars* sc = c.take<CgScalars>(1);
  double* r = c.take<double>(n);
  double* p = c.take<double>(n);
  double* Ap = c.take<double>(n);
  PartialWs pw(pws);
  SpmvPlan plan{};
  plan.tile = CG_TILE;
  plan.nblk = spmv_plan_blocks(nnz, CG_TILE);
  int32_t* blk_row = c.take<int32_t>(plan.nblk + 2);
  plan.blk_row = blk_row;
  FB2_TRY(cg_init_scalars(sc, atol, rtol, maxit < 0 ? INT_MAX : maxit, s));
  if (prebuilt && prebuilt->blk_row) {
    plan = *prebuilt;                       // the caller's plan of this matrix (possibly with compressed columns)
  } else {
    FB2_TRY(spmv_plan_build(n, crow, CG_TILE, plan.nblk, blk_row, &sc->pad, s));
    FB2_CUDA(cudaMemcpyAsync(&plan.max_row, &sc->pad, sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  // |b| and the zero-rhs early return (solver/cg.py:79-80)
  dot_kernel<<<vec_grid(n), CG_THREADS, 0, s>>>(n, b, b, &sc->dot_tmp, pw.partials, &sc->counter[0]);
  double bb = 0.0;
  FB2_CUDA(cudaMemcpyAsync(&bb, &sc->dot_tmp, sizeof(double), cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  const double bnorm = std::sqrt(bb);
  if (bnorm < 1e-15) {
    FB2_CUDA(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), s));
    FB2_CUDA(cudaStreamSynchronize(s));      // s may be the private stream: x must be zero before the caller's stream reads it
    return OK;
  }
  FB2_CUDA(cudaMemcpyAsync(&sc->bnorm, &bnorm, sizeof(double), cudaMemcpyHostToDevice, s));
  // r = b - A x0 ; p = z = M r ; rTr = r.z
  FB2_TRY(spmv_impl(n, nnz, crow, col, val, x, r, b, 1, nullptr, nullptr, nullptr, nullptr, s, &plan));
  cg_start_kernel<<<vec_grid(n), CG_THREADS, 0, s>>>(n, r, minv, p, sc, pw.partials, OwnRange{0, n, 0, 0});
  FB2_LAUNCH_CHECK();

  if (chunk <= 0) {
    // keep the host poll interval around a few hundred microseconds of GPU work
    const double bytes_it = 12.0 * (double)nnz + 112.0 * (double)n;
    const double t_it = bytes_it / 5.0e12 + 8e-6;
    chunk = (int)(4e-4 / t_it);
    if (chunk < 2) chunk = 2;
    if (chunk > 64) chunk = 64;
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  FB2_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int rc = OK;
  for (int k = 0; k < chunk && rc == OK; ++k) {
    rc = spmv_impl(n, nnz, crow, col, val, p, Ap, nullptr, 0, &sc->pAp, pw.partials, &sc->counter[0], sc, s, &plan);
    if (rc == OK) rc = cg_update_xr(n, x, r, p, Ap, minv, sc, pws, 1, s, OwnRange{0, n, 0, 0});
    if (rc == OK) rc = cg_update_p(n, p, r, minv, sc, s);
  }
  cudaError_t ce = cudaStreamEndCapture(s, &graph);
  if (rc != OK) { if (graph) cudaGraphDestroy(graph); return rc; }
  FB2_CUDA(ce);
  FB2_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  struct { int niter, done; } st = {0, 0};
  int guard = 0;
  while (true) {
    ce = cudaGraphLaunch(exec, s);
    if (ce != cudaSuccess) break;
    ce = cudaMemcpyAsync(&st, &sc->niter, sizeof(st), cudaMemcpyDeviceToHost, s);
    if (ce != cudaSuccess) break;
    ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) break;
    if (st.done) break;
    if (++guard > (1 << 28)) break;
  }
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  FB2_CUDA(ce);
  double rn = 0.0;
  FB2_CUDA(cudaMemcpyAsync(&rn, &sc->rnorm, sizeof(double), cudaMemcpyDeviceToHost, s));
  FB2_CUDA(cudaStreamSynchronize(s));
  *niter_out = st.niter;
  *resid_out = rn;
  return OK;
}

}  // namespace fb2

// =====================================================================================
// batched right-hand sides (b of shape (n, B), row-major): per-column alpha / beta, joint stopping
// test on sqrt(sum_k r_k.z_k)  (solver/cg.py:88-121).  Building blocks driven from the host.
// =====================================================================================
namespace fb2 {

constexpr int BCG_MAXB = 32;

// out[k] = sum_i a[i,k]*b[i,k]  (deterministic: per-CTA partials, last CTA sums in order)
__global__ void __launch_bounds__(CG_THREADS) bcg_dots_kernel(int64_t n, int B, const double* __restrict__ a, const double* __restrict__ b,
                                                              double* __restrict__ out, double* __restrict__ partials,
                                                              unsigned int* counter) {
  __shared__ double red[CG_THREADS / 32][BCG_MAXB];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < B; ++k) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += a[i * B + k] * b[i * B + k];
    acc = warp_sum(acc);
    if (lane == 0) red[wid][k] = acc;
  }
  __syncthreads();
  if (threadIdx.x < B) {
    double t = 0.0;
    for (int w = 0; w < CG_THREADS / 32; ++w) t += red[w][threadIdx.x];
    partials[(int64_t)blockIdx.x * B + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicInc(counter, gridDim.x - 1);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < B) {
    double t = 0.0;
    for (int g = 0; g < (int)gridDim.x; ++g) t += __ldcg(partials + (int64_t)g * B + threadIdx.x);
    out[threadIdx.x] = t;
  }
}

// x += alpha_k p ; r -= alpha_k Ap   with alpha_k = rTr[k] / pAp[k]
// `state` (device, 4 doubles: done flag, iteration count, residual norm, pad; may be null): once the joint stopping test
// (bcg_check_kernel) has fired, the update kernels do nothing -- the host can run several iterations per poll and x stays
// the iterate of the stopping iteration, as in the reference (solver/cg.py:97-121)
__global__ void __launch_bounds__(CG_THREADS) bcg_update_xr_kernel(int64_t n, int B, double* __restrict__ x, double* __restrict__ r,
                                                                   const double* __restrict__ p, const double* __restrict__ Ap,
                                                                   const double* __restrict__ rTr, const double* __restrict__ pAp,
                                                                   const double* __restrict__ state) {
  if (state && state[0] != 0.0) return;
  const int64_t tot = n * B;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % B);
    const double alpha = rTr[k] / pAp[k];
    x[t] += alpha * p[t];
    r[t] -= alpha * Ap[t];
  }
}

// p = z + beta_k p  with beta_k = rTr_new[k] / rTr[k],  z = minv .* r (or r)
__global__ void __launch_bounds__(CG_THREADS) bcg_update_p_kernel(int64_t n, int B, double* __restrict__ p, const double* __restrict__ r,
                                                                  const double* __restrict__ minv, const double* __restrict__ rTr_new,
                                                                  const double* __restrict__ rTr, const double* __restrict__ state) {
  if (state && state[0] != 0.0) return;
  const int64_t tot = n * B;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % B);
    const double z = minv ? minv[t / B] * r[t] : r[t];
    p[t] = z + (rTr_new[k] / rTr[k]) * p[t];
  }
}

// joint stopping test of the batched solve on the device: r_norm = sqrt(sum_k rTr_new[k]); atol, then rtol * |b|, then maxit
__global__ void bcg_check_kernel(int B, const double* __restrict__ rTr_new, double atol, double rtol_bnorm, int maxit, double* state) {
  if (state[0] != 0.0) return;
  double t = 0.0;
  for (int k = 0; k < B; ++k) t += rTr_new[k];
  const double rn = sqrt(t);
  const double it = state[1] + 1.0;
  state[1] = it;
  state[2] = rn;
  if (rn < atol || rn < rtol_bnorm || (maxit >= 0 && it >= (double)maxit)) state[0] = 1.0;
}

int bcg_check(int B, const double* rTr_new, double atol, double rtol_bnorm, int maxit, double* state, cudaStream_t s) {
  bcg_check_kernel<<<1, 1, 0, s>>>(B, rTr_new, atol, rtol_bnorm, maxit, state);
  FB2_LAUNCH_CHECK();
  return OK;
}

int bcg_dots(int64_t n, int B, const double* a, const double* b, double* out, void* partial_ws, cudaStream_t s) {
  if (B < 1 || B > BCG_MAXB) return fail(ERR_UNSUPPORTED, "batched cg: 1 <= batch <= %d", BCG_MAXB);
  PartialWs pw(partial_ws);
  int grid = vec_grid(n);
  if ((int64_t)grid * B > CG_PARTIALS) grid = CG_PARTIALS / B;
  bcg_dots_kernel<<<grid, CG_THREADS, 0, s>>>(n, B, a, b, out, pw.partials, pw.counter);
  FB2_LAUNCH_CHECK();
  return OK;
}
int bcg_update_xr(int64_t n, int B, double* x, double* r, const double* p, const double* Ap, const double* rTr, const double* pAp,
                  const double* state, cudaStream_t s) {
  bcg_update_xr_kernel<<<vec_grid(n * B), CG_THREADS, 0, s>>>(n, B, x, r, p, Ap, rTr, pAp, state);
  FB2_LAUNCH_CHECK();
  return OK;
}
int bcg_update_p(int64_t n, int B, double* p, const double* r, const double* minv, const double* rTr_new, const double* rTr,
                 const double* state, cudaStream_t s) {
  bcg_update_p_kernel<<<vec_grid(n * B), CG_THREADS, 0, s>>>(n, B, p, r, minv, rTr_new, rTr, state);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2
