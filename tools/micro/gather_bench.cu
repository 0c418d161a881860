// Micro-benchmark (tuning evidence, not product code): how fast can one warp stream 64-byte per-cell geometry records,
// selected by a dependent index, into its shared memory?  (a) cooperative 16-byte cp.async (LDGSTS) -- what
// assemble_const_v4_kernel does; (b) one cp.async.bulk (UBLKCP) per lane with an mbarrier; (c) plain LDG.128 into registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gather_bench tools/micro/gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int WARPS = 4, DEPTH = 3;

__device__ __forceinline__ void cp16(void* d, const void* s) {
  unsigned a = (unsigned)__cvta_generic_to_shared(d);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(s) : "memory");
}

// mode 0: cooperative LDGSTS; mode 1: per-lane UBLKCP; mode 2: LDG.128 x4 per lane (registers)
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 2) gather_kernel(const double* __restrict__ H, const int* __restrict__ idx, int64_t nbatch,
                                                               double* __restrict__ out) {
  __shared__ __align__(128) double st[WARPS][DEPTH][32 * 8];
  __shared__ __align__(8) unsigned long long bars[WARPS][DEPTH];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t nw = (int64_t)gridDim.x * WARPS, w = (int64_t)blockIdx.x * WARPS + wid;
  const int64_t per = (nbatch + nw - 1) / nw, b0 = w * per, b1 = (b0 + per < nbatch) ? b0 + per : nbatch;
  double acc = 0.0;
  if (MODE == 1) {
    if (lane == 0)
      for (int s = 0; s < DEPTH; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bars[wid][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  auto issue = [&](int64_t b, int s) {
    const int c = idx[b * 32 + lane];
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = r * 32 + lane, rec = k >> 2, part = k & 3;
        const int cc = __shfl_sync(0xffffffffu, c, rec);
        cp16(&st[wid][s][rec * 8 + 2 * ((part + (rec >> 1)) & 3)], H + (int64_t)cc * 8 + 2 * part);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (MODE == 1) {
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[wid][s]);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(32 * 64) : "memory");
      __syncwarp();
      const unsigned dst = (unsigned)__cvta_generic_to_shared(&st[wid][s][lane * 8]);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];" ::"r"(dst),
                   "l"(H + (int64_t)c * 8), "r"(bar) : "memory");
    }
  };
  if (MODE != 2)
    for (int k = 0; k < DEPTH - 1; ++k)
      if (b0 + k < b1) issue(b0 + k, k);
  int s = 0;
  unsigned phase = 0;
  for (int64_t b = b0; b < b1; ++b) {
    if (MODE == 2) {
      const int c = idx[b * 32 + lane];
      const double2* p = reinterpret_cast<const double2*>(H + (int64_t)c * 8);
#pragma unroll
      for (int t = 0; t < 4; ++t) { const double2 v = p[t]; acc += v.x + v.y; }
      continue;
    }
    int sn = s + DEPTH - 1; if (sn >= DEPTH) sn -= DEPTH;
    if (b + DEPTH - 1 < b1) issue(b + DEPTH - 1, sn);
    else if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
    if (MODE == 0) {
      asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double2 v = *reinterpret_cast<const double2*>(&st[wid][s][lane * 8 + 2 * ((t + (lane >> 1)) & 3)]);
        acc += v.x + v.y;
      }
    } else {
      const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[wid][s]);
      unsigned ok = 0;
      while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double2 v = *reinterpret_cast<const double2*>(&st[wid][s][lane * 8 + 2 * t]);
        acc += v.x + v.y;
      }
    }
    __syncwarp();
    if (++s == DEPTH) { s = 0; phase ^= 1; }
  }
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int64_t NC = 12582912, nbatch = 4600000;
  std::vector<int> h((size_t)nbatch * 32);
  uint64_t rng = 12345;
  for (int64_t b = 0; b < nbatch; ++b) {       // a batch touches cells clustered around a moving centre (like a row tile does)
    const int64_t centre = (int64_t)((double)b / nbatch * (NC - 2000));
    for (int l = 0; l < 32; ++l) { rng = rng * 6364136223846793005ull + 1442695040888963407ull; h[b * 32 + l] = (int)(centre + (rng >> 33) % 2000); }
  }
  double *H, *out; int* idx;
  CK(cudaMalloc(&H, NC * 64)); CK(cudaMemset(H, 0, NC * 64));
  CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const int grid = 148 * 2;
  CK(cudaMalloc(&out, (size_t)grid * WARPS * 32 * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) gather_kernel<0><<<grid, WARPS * 32>>>(H, idx, nbatch, out);
      else if (mode == 1) gather_kernel<1><<<grid, WARPS * 32>>>(H, idx, nbatch, out);
      else gather_kernel<2><<<grid, WARPS * 32>>>(H, idx, nbatch, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
    }
    printf("mode %d (%s): %.3f ms for %lld batches of 32 x 64 B records = %.1f G records/s\n", mode,
           mode == 0 ? "cooperative cp.async 16B" : mode == 1 ? "per-lane cp.async.bulk 64B" : "LDG.128 x4 to registers", best,
           (long long)nbatch, nbatch * 32 / best * 1e-6);
  }
  return 0;
}
