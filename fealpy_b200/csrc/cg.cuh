#pragma once
#include "common.cuh"

namespace fb2 {

// device-resident CG scalars (one struct per solve; lives in a caller-provided buffer)
struct CgScalars {
  double rTr, pAp, rTr_new, bnorm, rnorm, alpha, beta, dot_tmp;
  int niter, done, maxit, pad;
  double atol, rtol;
  unsigned int counter[4];   // last-block-done tickets for the three reductions
};

constexpr int CG_PARTIALS = 4096;

// rows that contribute to dot products (the owned rows of a rank): [lo0,hi0) U [lo1,hi1)
struct OwnRange {
  int64_t lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
  __host__ __device__ bool has(int64_t r) const { return (r >= lo0 && r < hi0) || (r >= lo1 && r < hi1); }
};

// multi-GPU: tiles >= first_tile read halo entries that the (<= 2) neighbours push over NVLink; the CTA that reaches such a
// tile first waits until flags[nb] >= *epoch + sc->niter (csrc/peer.cu).  nnb == 0: no waiting (single GPU).
struct HaloWait {
  const unsigned long long* flags = nullptr;   // hflag array of this rank's PeerCtrl
  const unsigned long long* epoch = nullptr;
  int nnb = 0, nb0 = 0, nb1 = 0, first_tile = 0;
};

// row-aligned nnz tiling of a CSR matrix (built once per matrix, see spmv_plan_build)
struct SpmvPlan {
  const int32_t* blk_row = nullptr;   // (nblk+1) first row of every tile
  int nblk = 0, tile = 0, max_row = 0;
  const int32_t* blk_end = nullptr;   // optional (nblk): one-past-last row of every tile -- tiles of SEVERAL row ranges in one plan
  HaloWait halo;
};   // max blocks contributing partial sums per reduction

size_t cg_workspace_bytes(int64_t n, int64_t nnz);
int spmv_plan_blocks(int64_t nnz, int tile);
int spmv_plan_build(int64_t n, const int64_t* crow, int tile, int nblk, int32_t* blk_row, int* max_row_dev, cudaStream_t s);

// y = A x  (mode 0),  y = b - A x (mode 1); optional fused dot  sum_r x[r]*y[r] -> *dot_out (deterministic)
size_t partial_workspace_bytes();
int spmv(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* val, const double* x, double* y,
         const double* b, int mode, double* dot_out, void* partial_ws, cudaStream_t s, const SpmvPlan* plan = nullptr,
         OwnRange own = OwnRange{}, const CgScalars* sc = nullptr);
// SpMM with a row-major (n, nb) dense block: Y = A X
int spmm(int64_t n, const int64_t* crow, const int32_t* col, const double* val, const double* X, double* Y, int nb, cudaStream_t s);
// deterministic dot product
int dot(int64_t n, const double* a, const double* b, double* out, void* partial_ws, cudaStream_t s);

// full solve: reference recurrence (solver/cg.py:76-123); x holds x0 on entry, the solution on exit
int cg_solve(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* val, const double* b, double* x,
             const double* minv_diag, double atol, double rtol, int maxit, int chunk, void* ws, int* niter_out,
             double* resid_out, cudaStream_t s, const SpmvPlan* prebuilt = nullptr);

// building blocks for the distributed (multi-GPU) driver
int cg_init_scalars(CgScalars* sc, double atol, double rtol, int maxit, cudaStream_t s);
int cg_start(int64_t n, const double* r, const double* minv, double* p, CgScalars* sc, void* partial_ws, OwnRange own, cudaStream_t s);
int cg_update_xr(int64_t n, double* x, double* r, const double* p, const double* Ap, const double* minv, CgScalars* sc,
                 void* partial_ws, int fuse_finalize, cudaStream_t s, OwnRange own = OwnRange{});
int cg_finalize(CgScalars* sc, cudaStream_t s);
int cg_update_p(int64_t n, double* p, const double* r, const double* minv, CgScalars* sc, cudaStream_t s);

// ---- multi-GPU CG over NVLink peer memory (csrc/peer.cu): the halo exchange and the scalar all-reduces of the CG iteration
// are done by the CG kernels themselves with stores into the neighbours' memory + sequence flags -- no NCCL call per iteration.
constexpr int PEER_MAXW = 16;
struct PeerCtrl {                         // one per rank, at the start of the rank's symmetric buffer (same layout everywhere)
  double red[2][2][PEER_MAXW];            // [kind: 0 = p.Ap, 1 = r.z][parity of the iteration][source rank] partial sums
  unsigned long long rflag[2][PEER_MAXW]; // [kind][source rank] sequence number of the newest partial
  unsigned long long hflag[PEER_MAXW];    // [source rank] sequence number of the newest halo pushed by that neighbour
};
struct PeerSlice { int64_t lo, hi, peer_lo; double* peer_p; };     // my owned p[lo, hi) -> the neighbour's p[peer_lo, ...)
struct PeerPush {
  int nslice, nnb, rank;
  PeerSlice slice[4];
  PeerCtrl* nb_ctrl[2];                   // control blocks of the (<= 2) neighbours
  unsigned int* counter;                  // local last-block ticket
  const unsigned long long* epoch;        // device scalar: base of this solve's sequence numbers
};
int peer_allreduce(PeerCtrl* mine, const unsigned long long* peer_base, int world, int rank, int kind, const double* src0,
                   const double* src1, double* dst, CgScalars* sc, int finalize, const unsigned long long* epoch, cudaStream_t s);
int peer_wait_halo(PeerCtrl* mine, int nnb, const int* nb_rank, const CgScalars* sc, const unsigned long long* epoch, cudaStream_t s);
int gather_f64(int64_t n, const int64_t* idx, const double* v, double* out, cudaStream_t s);
int cg_update_p_push(OwnRange own, double* p, const double* r, const double* minv, const CgScalars* sc, const PeerPush& push,
                     cudaStream_t s);

// batched right-hand sides (row-major (n, B))
int bcg_dots(int64_t n, int B, const double* a, const double* b, double* out, void* partial_ws, cudaStream_t s);
int bcg_update_xr(int64_t n, int B, double* x, double* r, const double* p, const double* Ap, const double* rTr, const double* pAp,
                  const double* state, cudaStream_t s);
int bcg_update_p(int64_t n, int B, double* p, const double* r, const double* minv, const double* rTr_new, const double* rTr,
                 const double* state, cudaStream_t s);
int bcg_check(int B, const double* rTr_new, double atol, double rtol_bnorm, int maxit, double* state, cudaStream_t s);
}  // namespace fb2
