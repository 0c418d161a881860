// Multi-GPU CG over NVLink peer memory: the per-iteration halo exchange and the two scalar all-reduces are done by the
// kernels of the iteration themselves, with plain stores into the neighbours' (symmetric, peer-mapped) memory followed by
// a sequence flag -- no NCCL call, no host round trip, inside the iteration.  (replaces the
// per-iteration `batch_isend_irecv` + two 1-element `all_reduce`s of parallel/dist_cg.py, which cost ~100 us per iteration
// at 8 GPUs.)
//
// Protocol.  Every rank owns a PeerCtrl block at the start of its symmetric buffer, followed by its search direction p.
//   sequence numbers: seq = *epoch + (number of completed iterations) (+ 1); *epoch grows by 2^32 per solve, so flags
//   of earlier solves and iterations always compare "older".
//   all-reduce (kind 0: p.Ap, kind 1: r.z):  every rank stores its partial into red[kind][iteration parity][rank] of
//   EVERY rank, fences (system scope) and raises rflag[kind][rank] there; then waits until all `world` flags in its own
//   block have reached seq and sums the partials in rank order -- the same order on every rank, so every rank computes the
//   bit-identical sum (and therefore takes the identical convergence decision, without another exchange).  A slot of
//   parity q is rewritten two iterations later; by then every rank has read it: a rank cannot be two reductions ahead
//   of another one, because each reduction needs everybody's contribution.
//   halo: the p-update kernel also stores the owned boundary slices of the new p into the neighbours' p vectors (the
//   neighbour's halo slots), each thread fences its stores, the last CTA raises hflag[rank] on the neighbours.  The
//   neighbour's boundary-row SpMV of the next iteration waits for that flag; its interior-row SpMV does not, so the transfer
//   overlaps the interior product.  The neighbour cannot still be reading the old halo values: its previous boundary SpMV
//   precedes (in stream order) its contribution to the r.z reduction that this p update has already consumed.
#include "cg.cuh"

namespace fb2 {

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_vol(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_vol(double* p, double v) { asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

__device__ __forceinline__ void cg_finalize_peer(CgScalars* sc) {       // solver/cg.py:97-121 (same as cg_finalize_dev in cg.cu)
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

__global__ void __launch_bounds__(32) peer_allreduce_kernel(PeerCtrl* mine, const unsigned long long* __restrict__ peer_base, int world,
                                                            int rank, int kind, const double* src0, const double* src1, double* dst,
                                                            CgScalars* sc, int finalize, const unsigned long long* epoch) {
  if (sc->done) return;                                   // identical on every rank (see above): nobody pushes, nobody waits
  const int lane = threadIdx.x;
  const int it = sc->niter;
  const unsigned long long seq = *epoch + (unsigned long long)it + 1ull;
  const int par = it & 1;
  const double mine_v = *src0 + (src1 ? *src1 : 0.0);
  double pv = 0.0;
  if (lane < world) {
    PeerCtrl* peer = reinterpret_cast<PeerCtrl*>(peer_base[lane]);
    st_vol(&peer->red[kind][par][rank], mine_v);
    __threadfence_system();
    st_flag(&peer->rflag[kind][rank], seq);
    while (ld_flag(&mine->rflag[kind][lane]) < seq) { }
    __threadfence_system();
    pv = ld_vol(&mine->red[kind][par][lane]);
  }
  double tot = 0.0;
  for (int r = 0; r < world; ++r) tot += __shfl_sync(0xffffffffu, pv, r);      // rank order: the same sum on every rank
  if (lane == 0) {
    *dst = tot;
    if (finalize) cg_finalize_peer(sc);
  }
}

__global__ void __launch_bounds__(32) peer_wait_halo_kernel(PeerCtrl* mine, int nnb, int nb0, int nb1, const CgScalars* sc,
                                                            const unsigned long long* epoch) {
  if (sc->done) return;
  const unsigned long long seq = *epoch + (unsigned long long)sc->niter;       // pushed at the end of the previous iteration
  const int lane = threadIdx.x;
  if (lane < nnb) {
    const int nb = lane == 0 ? nb0 : nb1;
    while (ld_flag(&mine->hflag[nb]) < seq) { }
    __threadfence_system();
  }
}

// p = z + beta p on the OWNED rows only (halo entries of p belong to the neighbours' pushes), then the boundary slices go
// to the neighbours
__global__ void __launch_bounds__(256) cg_update_p_push_kernel(OwnRange own, double* __restrict__ p, const double* __restrict__ r,
                                                               const double* __restrict__ minv, const CgScalars* sc, PeerPush push) {
  if (sc->done) return;
  const double beta = sc->beta;
  const int64_t len0 = own.hi0 - own.lo0, tot = len0 + (own.hi1 - own.lo1);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t < len0 ? own.lo0 + t : own.lo1 + (t - len0);
    const double ri = r[i];
    const double v = (minv ? minv[i] * ri : ri) + beta * p[i];
    p[i] = v;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < push.nslice && i >= push.slice[k].lo && i < push.slice[k].hi)
        push.slice[k].peer_p[push.slice[k].peer_lo + (i - push.slice[k].lo)] = v;
  }
  __threadfence_system();                                 // my stores (also the remote ones) before my ticket
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicInc(push.counter, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x < push.nnb) {
    __threadfence_system();
    const unsigned long long seq = *push.epoch + (unsigned long long)sc->niter;     // niter already counts this iteration
    st_flag(&push.nb_ctrl[threadIdx.x]->hflag[push.rank], seq);
  }
}

// pack kernel of the general (Morton) partition: out[k] = v[idx[k]] -- the owned values one neighbour needs, contiguous
__global__ void __launch_bounds__(256) gather_f64_kernel(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ v,
                                                         double* __restrict__ out) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) out[k] = v[idx[k]];
}
int gather_f64(int64_t n, const int64_t* idx, const double* v, double* out, cudaStream_t s) {
  if (n <= 0) return OK;
  int64_t b = ceil_div(n, 256);
  const int64_t cap = kNumSM * 16;
  gather_f64_kernel<<<(unsigned)(b > cap ? cap : b), 256, 0, s>>>(n, idx, v, out);
  FB2_LAUNCH_CHECK();
  return OK;
}

int peer_allreduce(PeerCtrl* mine, const unsigned long long* peer_base, int world, int rank, int kind, const double* src0,
                   const double* src1, double* dst, CgScalars* sc, int finalize, const unsigned long long* epoch, cudaStream_t s) {
  if (world < 1 || world > PEER_MAXW || rank < 0 || rank >= world || kind < 0 || kind > 1)
    return fail(ERR_INVALID, "peer_allreduce: world=%d rank=%d kind=%d", world, rank, kind);
  peer_allreduce_kernel<<<1, 32, 0, s>>>(mine, peer_base, world, rank, kind, src0, src1, dst, sc, finalize, epoch);
  FB2_LAUNCH_CHECK();
  return OK;
}

int peer_wait_halo(PeerCtrl* mine, int nnb, const int* nb_rank, const CgScalars* sc, const unsigned long long* epoch, cudaStream_t s) {
  if (nnb < 0 || nnb > 2) return fail(ERR_INVALID, "peer_wait_halo: at most two neighbours (slab partition)");
  if (nnb == 0) return OK;
  peer_wait_halo_kernel<<<1, 32, 0, s>>>(mine, nnb, nb_rank[0], nnb > 1 ? nb_rank[1] : 0, sc, epoch);
  FB2_LAUNCH_CHECK();
  return OK;
}

int cg_update_p_push(OwnRange own, double* p, const double* r, const double* minv, const CgScalars* sc, const PeerPush& push,
                     cudaStream_t s) {
  const int64_t tot = (own.hi0 - own.lo0) + (own.hi1 - own.lo1);
  if (tot <= 0) return OK;
  int64_t b = ceil_div(tot, 256 * 4);
  const int64_t cap = kNumSM * 8;
  const int grid = (int)(b < 1 ? 1 : (b > cap ? cap : b));
  cg_update_p_push_kernel<<<grid, 256, 0, s>>>(own, p, r, minv, sc, push);
  FB2_LAUNCH_CHECK();
  return OK;
}

}  // namespace fb2
