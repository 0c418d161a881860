// extern "C" boundary (include/fealpy_b200.h): thin argument marshalling onto the kernels.
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/fealpy_b200.h"
#include "assemble.cuh"
#include "bc_source.cuh"
#include "cg.cuh"
#include "common.cuh"
#include "coo_csr.cuh"
#include "elem.cuh"
#include "sort_scan.cuh"
#include "topo.cuh"

namespace fb2 {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
}  // namespace fb2

using namespace fb2;
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* fb2_last_error(void) { return g_last_error.c_str(); }
int fb2_version(void) { return 100; }

// ---- topology ---------------------------------------------------------------------------
int fb2_tri_from_box(const double box[4], int nx, int ny, double* node, int32_t* cell, void* stream) {
  if (nx < 1 || ny < 1) return fail(ERR_INVALID, "tri_from_box: nx, ny must be >= 1");
  return tri_from_box(box, nx, ny, node, cell, S(stream));
}
int fb2_tet_from_box(const double box[6], int nx, int ny, int nz, double* node, int32_t* cell, void* stream) {
  if (nx < 1 || ny < 1 || nz < 1) return fail(ERR_INVALID, "tet_from_box: nx, ny, nz must be >= 1");
  return tet_from_box(box, nx, ny, nz, node, cell, S(stream));
}
size_t fb2_entity_workspace_bytes(int64_t NC, int per_cell) { return entity_workspace_bytes(NC, per_cell); }
int fb2_build_entities(const int32_t* cell, int64_t NC, int TD, int kind, int64_t NN, int32_t* cell2ent, int64_t* count_host,
                       void* ws, void* stream) {
  if ((TD != 2 && TD != 3) || (kind != 1 && kind != 2) || (kind == 2 && TD != 3))
    return fail(ERR_INVALID, "build_entities: TD=%d kind=%d not supported", TD, kind);
  return build_entities(cell, NC, TD, kind, NN, cell2ent, count_host, ws, S(stream));
}
int fb2_entities_emit(const int32_t* cell, int64_t NC, int TD, int kind, int32_t* cell2ent, int32_t* ent, void* ws, void* stream) {
  return entities_emit(cell, NC, TD, kind, cell2ent, ent, ws, S(stream));
}
int fb2_cell_to_dof(const int32_t* cell, const int32_t* cell2edge, const int32_t* edge, const int32_t* cell2face, int64_t NC, int TD,
                    int p, int64_t NN, int64_t NE, int64_t NF, const unsigned char* mi, int ldof, int32_t* c2d, void* stream) {
  return cell_to_dof(cell, cell2edge, edge, cell2face, NC, TD, p, NN, NE, NF, mi, ldof, c2d, S(stream));
}
int fb2_tensor_cell_to_dof(const int32_t* c2d, int64_t NC, int ldof, int GD, int64_t gdof, int prio, int32_t* out, void* stream) {
  return tensor_cell_to_dof(c2d, NC, ldof, GD, gdof, prio, out, S(stream));
}

// ---- K1 -----------------------------------------------------------------------------------
int fb2_elem_scalar_const(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const double* Ms, const double* Mm,
                          double scal_d, const double* coef_d, double scal_m, const double* coef_m, double* out, void* stream) {
  if (!Ms && !Mm) return fail(ERR_INVALID, "elem_scalar_const: need a diffusion and/or a mass table");
  ElemConstArgs a{};
  a.node = node; a.cell = cell; a.NC = NC;
  a.has_diff = Ms != nullptr; a.has_mass = Mm != nullptr;
  a.Ms = Ms; a.Mm = Mm; a.scal_d = scal_d; a.scal_m = scal_m; a.coef_d = coef_d; a.coef_m = coef_m; a.out = out;
  return elem_const(TD, p, a, S(stream));
}
int fb2_elem_scalar_const_acc(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const double* Ms, const double* Mm,
                              double scal_d, const double* coef_d, double scal_m, const double* coef_m, double* out, int accumulate,
                              void* stream) {
  if (!Ms && !Mm) return fail(ERR_INVALID, "elem_scalar_const: need a diffusion and/or a mass table");
  ElemConstArgs a{};
  a.node = node; a.cell = cell; a.NC = NC;
  a.has_diff = Ms != nullptr; a.has_mass = Mm != nullptr;
  a.Ms = Ms; a.Mm = Mm; a.scal_d = scal_d; a.scal_m = scal_m; a.coef_d = coef_d; a.coef_m = coef_m; a.out = out;
  a.accumulate = accumulate ? 1 : 0;
  return elem_const(TD, p, a, S(stream));
}
int fb2_elem_scalar_quad_fused(int TD, int p, int64_t NC, const double* node, const int32_t* cell, int is_mass, int NQ,
                               const double* ws, const double* table, int coef_kind, const double* coef, const double* Ms_const,
                               const double* Mm_const, double scal_d, const double* coef_d_cell, double scal_m,
                               const double* coef_m_cell, double* out, int accumulate, void* stream) {
  if (coef_kind != 2 && coef_kind != 3) return fail(ERR_INVALID, "elem_scalar_quad: coef_kind must be 2 (NC,NQ) or 3 (NC,NQ,GD,GD)");
  if (is_mass && coef_kind == 3) return fail(ERR_INVALID, "elem_scalar_quad: matrix coefficients apply to diffusion only");
  ElemQuadArgs a{};
  a.node = node; a.cell = cell; a.NC = NC; a.is_mass = is_mass; a.NQ = NQ; a.ws = ws; a.tab = table;
  a.coef_kind = coef_kind; a.coef = coef; a.out = out; a.accumulate = accumulate ? 1 : 0;
  a.xMs = Ms_const; a.xMm = Mm_const; a.x_scal_d = scal_d; a.x_scal_m = scal_m; a.x_coef_d = coef_d_cell; a.x_coef_m = coef_m_cell;
  return elem_quad(TD, p, a, S(stream));
}
int fb2_elem_scalar_quad(int TD, int p, int64_t NC, const double* node, const int32_t* cell, int is_mass, int NQ, const double* ws,
                         const double* table, int coef_kind, const double* coef, double* out, void* stream) {
  return fb2_elem_scalar_quad_fused(TD, p, NC, node, cell, is_mass, NQ, ws, table, coef_kind, coef, nullptr, nullptr, 0.0, nullptr, 0.0,
                                    nullptr, out, 0, stream);
}
int fb2_elem_elasticity(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const double* M4, double d_diag,
                        double d_lam, double d_shear, int dof_priority, double* out, void* stream) {
  ElemElasticityArgs a{};
  a.node = node; a.cell = cell; a.NC = NC; a.M4 = M4; a.d_diag = d_diag; a.d_lam = d_lam; a.d_shear = d_shear;
  a.dof_priority = dof_priority; a.out = out;
  return elem_elasticity(TD, p, a, S(stream));
}

// ---- public geometry / basis (rows a7, a9) ----------------------------------------------------
int fb2_cell_gradients(int TD, int64_t NC, const double* node, const int32_t* cell, double* out, void* stream) {
  return cell_gradients(TD, NC, node, cell, out, S(stream));
}
int fb2_grad_basis(int TD, int64_t NC, int NQ, int ldof, const double* cell_gradient_records, const double* R, double* out,
                   void* stream) {
  return grad_basis(TD, NC, NQ, ldof, cell_gradient_records, R, out, S(stream));
}

// ---- K2 -----------------------------------------------------------------------------------
int fb2_coo_keys_from_c2d(const int32_t* rdof, const int32_t* cdof, int64_t NC, int lr, int lc, int col_bits, uint64_t* keys,
                          void* stream) {
  return coo_keys_from_c2d(rdof, cdof, NC, lr, lc, col_bits, keys, S(stream));
}
int fb2_coo_keys_from_coo(const void* row, const void* col, int index_bytes, int64_t n, int col_bits, uint64_t* keys, void* stream) {
  return coo_keys_from_coo(row, col, index_bytes, n, col_bits, keys, S(stream));
}
size_t fb2_coo_workspace_bytes(int64_t n) { return coo_symbolic_workspace_bytes(n); }
int fb2_coo_symbolic(uint64_t* keys, uint32_t* perm, int64_t n, int key_bits, void* ws, int64_t* nnz_host, void* stream) {
  return coo_symbolic(keys, perm, n, key_bits, ws, nnz_host, S(stream));
}
int fb2_coo_fill(const uint64_t* keys, int64_t n, int col_bits, int64_t nrow, void* ws, int64_t* crow, void* col, int col_bytes,
                 int64_t* seg_start, void* stream) {
  return coo_fill(keys, n, col_bits, nrow, ws, crow, col, col_bytes, seg_start, S(stream));
}
int fb2_coo_reduce(const uint32_t* perm, const int64_t* seg_start, int64_t nnz, const double* vin, double* vout, void* stream) {
  return coo_reduce(perm, seg_start, nnz, vin, vout, S(stream));
}

// ---- symbolic + fused numeric ---------------------------------------------------------------
size_t fb2_sym_workspace_bytes(int64_t NC, int ldof, int64_t gdof) { return sym_workspace_bytes(NC, ldof, gdof); }
int fb2_slot_stride(int ldof, int slot_bytes) { return slot_stride(ldof, slot_bytes); }
int fb2_sym_count(const int32_t* c2d, int64_t NC, int ldof, int64_t gdof, int64_t* adj_ptr, int32_t* adj_pair, int64_t* crow,
                  int64_t* nnz_host, int32_t* max_row_host, uint16_t* stash, void* ws, void* stream) {
  return sym_count(c2d, NC, ldof, gdof, adj_ptr, adj_pair, crow, nnz_host, max_row_host, stash, ws, S(stream));
}
int fb2_sym_fill(const int32_t* c2d, int64_t NC, int ldof, int64_t gdof, const int64_t* adj_ptr, const int32_t* adj_pair,
                 const int64_t* crow, int32_t* col, void* slots, int slot_bytes, const uint16_t* stash, void* stream) {
  return sym_fill(c2d, NC, ldof, gdof, adj_ptr, adj_pair, crow, col, slots, slot_bytes, stash, S(stream));
}
int fb2_assemble_scalar_const(int TD, int p, int64_t NC, int64_t gdof, const double* node, const int32_t* cell,
                              const int64_t* adj_ptr, const int32_t* adj_pair, const void* slots, int slot_bytes,
                              const int64_t* crow, int32_t max_row, const int32_t* blk_row, int nblk, int tile,
                              const double* Ms, const double* Mm, double scal_d, const double* coef_d, double scal_m,
                              const double* coef_m, double* values, void* stream) {
  if (!Ms && !Mm) return fail(ERR_INVALID, "assemble_scalar_const: need a diffusion and/or a mass table");
  AsmConstArgs a{};
  a.node = node; a.cell = cell; a.gdof = gdof;
  a.adj_ptr = adj_ptr; a.adj_pair = adj_pair; a.slots = slots; a.crow = crow;
  a.has_diff = Ms != nullptr; a.has_mass = Mm != nullptr; a.Ms = Ms; a.Mm = Mm;
  a.scal_d = scal_d; a.scal_m = scal_m; a.coef_d = coef_d; a.coef_m = coef_m; a.values = values;
  a.nnz = 0;
  a.blk_row = blk_row; a.nblk = nblk; a.tile = tile; a.threads = 0;
  a.NC = NC;
  return assemble_const(TD, p, a, slot_bytes, max_row, S(stream));
}
int fb2_assemble_from_ke(int64_t NC, int ldof, int ncomp, int dof_priority, int64_t gdof_scalar, const double* Ke,
                         const int64_t* adj_ptr, const int32_t* adj_pair, const void* slots, int slot_bytes,
                         const int64_t* crow_scalar, int32_t max_row, const int64_t* crow_out, const int32_t* blk_row, int nblk,
                         int tile, double* values, void* stream) {
  AsmKeArgs a{};
  a.gdof = gdof_scalar; a.ncell = NC; a.L = ldof; a.ncomp = ncomp; a.dof_priority = dof_priority; a.Ke = Ke;
  a.adj_ptr = adj_ptr; a.adj_pair = adj_pair; a.slots = slots; a.crow_s = crow_scalar;
  a.crow_out = crow_out ? crow_out : crow_scalar; a.values = values;
  a.nnz_out = 0;
  a.blk_row = blk_row; a.nblk = nblk; a.tile = tile;
  return assemble_from_ke(a, slot_bytes, max_row, S(stream));
}
int fb2_assemble_elasticity_p1(int TD, int64_t NC, const double* node, const int32_t* cell, int dof_priority, int64_t gdof_scalar,
                               double d_diag, double d_lam, double d_shear, double wsum, const int64_t* adj_ptr,
                               const int32_t* adj_pair, const void* slots, int slot_bytes, const int64_t* crow_scalar, int32_t max_row,
                               const int64_t* crow_out, const int32_t* blk_row, int nblk, int tile, double* geo_ws, double* values,
                               void* stream) {
  FB2_TRY(cell_gradients(TD, NC, node, cell, geo_ws, S(stream)));
  AsmKeArgs a{};
  a.gdof = gdof_scalar; a.ncell = NC; a.dof_priority = dof_priority; a.adj_ptr = adj_ptr; a.adj_pair = adj_pair; a.slots = slots;
  a.crow_s = crow_scalar; a.crow_out = crow_out; a.values = values; a.blk_row = blk_row; a.nblk = nblk; a.tile = tile;
  a.geo = geo_ws; a.d_diag = d_diag; a.d_lam = d_lam; a.d_shear = d_shear; a.wsum = wsum;
  return assemble_elasticity_p1(TD, a, slot_bytes, max_row, S(stream));
}
size_t fb2_asm4_workspace_bytes(int ntile) { return asm4_workspace_bytes(ntile); }
int fb2_asm4_plan_count(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int32_t* adj_pair,
                        int ldof, int64_t* batch_ptr, int64_t* nbatch_host, void* ws, void* stream) {
  return asm4_plan_count(ntile, blk_row, crow, adj_ptr, adj_pair, ldof, batch_ptr, nbatch_host, ws, S(stream));
}
int fb2_asm4_plan_fill(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int32_t* adj_pair,
                       int ldof, const int64_t* batch_ptr, uint8_t* batch_i, uint32_t* blocks, const void* slots, int slot_bytes,
                       void* stream) {
  return asm4_plan_fill(ntile, blk_row, crow, adj_ptr, adj_pair, ldof, batch_ptr, batch_i, blocks, slots, slot_bytes, S(stream));
}
int fb2_assemble_scalar_const_v4(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const int64_t* crow,
                                 const int32_t* blk_row, int ntile, int tile, int32_t max_row, const int64_t* batch_ptr,
                                 const uint8_t* batch_i, const uint32_t* blocks, int slot_bytes, const double* Ms_host, const double* Mm_host, double scal_d, const double* coef_d,
                                 double scal_m, const double* coef_m, double* geom_ws, double* values, void* stream) {
  const double *Ms = Ms_host, *Mm = Mm_host;      // only their presence matters on the device side
  if (!Ms && !Mm) return fail(ERR_INVALID, "assemble_scalar_const_v4: need a diffusion and/or a mass table");
  Asm4Args a{};
  a.node = node; a.cell = cell; a.NC = NC; a.crow = crow; a.blk_row = blk_row; a.ntile = ntile; a.tile = tile; a.max_row = max_row;
  a.batch_ptr = batch_ptr; a.batch_i = batch_i; a.blocks = blocks;
  a.Ms = Ms; a.Mm = Mm; a.Ms_host = Ms_host; a.Mm_host = Mm_host;
  a.scal_d = scal_d; a.scal_m = scal_m; a.coef_d = coef_d; a.coef_m = coef_m; a.Hbuf = geom_ws; a.values = values;
  return assemble_v4(TD, p, a, slot_bytes, S(stream));
}
int fb2_expand_pattern(int64_t gdof_scalar, int ncomp, int dof_priority, const int64_t* crow_scalar, const int32_t* col_scalar,
                       int64_t* crow_out, int32_t* col_out, void* stream) {
  return expand_pattern(gdof_scalar, ncomp, dof_priority, crow_scalar, col_scalar, crow_out, col_out, S(stream));
}

// ---- K3/K4 ----------------------------------------------------------------------------------
size_t fb2_partial_workspace_bytes(void) { return partial_workspace_bytes(); }
int fb2_spmv_plan_blocks(int64_t nnz, int tile) { return spmv_plan_blocks(nnz, tile); }
int fb2_spmv_plan_build(int64_t n, const int64_t* crow, int tile, int32_t* blk_row, int64_t nnz, int32_t* max_row_host, void* stream) {
  const int nblk = spmv_plan_blocks(nnz, tile);
  int* dmax = blk_row + nblk + 1;        // the caller allocates nblk + 2 entries; the last one is scratch
  FB2_TRY(spmv_plan_build(n, crow, tile, nblk, blk_row, dmax, S(stream)));
  FB2_CUDA(cudaMemcpyAsync(max_row_host, dmax, sizeof(int), cudaMemcpyDeviceToHost, S(stream)));
  FB2_CUDA(cudaStreamSynchronize(S(stream)));
  return OK;
}
static SpmvPlan make_plan(const int32_t* blk_row, int64_t nnz, int tile, int max_row) {
  SpmvPlan pl{};
  pl.blk_row = blk_row; pl.tile = tile; pl.max_row = max_row;
  pl.nblk = blk_row ? spmv_plan_blocks(nnz, tile) : 0;
  return pl;
}
int fb2_csr_spmv(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* x, double* y,
                 const int32_t* blk_row, int tile, int32_t max_row, void* stream) {
  const SpmvPlan pl = make_plan(blk_row, nnz, tile, max_row);
  return spmv(n, nnz, crow, col, values, x, y, nullptr, 0, nullptr, nullptr, S(stream), blk_row ? &pl : nullptr);
}
int fb2_csr_spmm(int64_t n, const int64_t* crow, const int32_t* col, const double* values, const double* X, double* Y, int nb,
                 void* stream) {
  return spmm(n, crow, col, values, X, Y, nb, S(stream));
}
int fb2_dot(int64_t n, const double* a, const double* b, double* out_dev, void* partial_ws, void* stream) {
  return dot(n, a, b, out_dev, partial_ws, S(stream));
}
size_t fb2_cg_workspace_bytes(int64_t n, int64_t nnz) { return cg_workspace_bytes(n, nnz); }
int fb2_cg(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* b, double* x,
           const double* minv_diag, double atol, double rtol, int maxit, int chunk, const int32_t* blk_row,
           int tile, int32_t max_row, void* ws, int* niter_host, double* residual_host, void* stream) {
  const SpmvPlan pl = make_plan(blk_row, nnz, tile, max_row);
  return cg_solve(n, nnz, crow, col, values, b, x, minv_diag, atol, rtol, maxit, chunk, ws, niter_host, residual_host, S(stream),
                  blk_row ? &pl : nullptr);
}
int fb2_cg_init(void* scalars, double atol, double rtol, int maxit, double bnorm, double rTr, void* stream) {
  CgScalars* sc = static_cast<CgScalars*>(scalars);
  FB2_TRY(cg_init_scalars(sc, atol, rtol, maxit < 0 ? 2147483647 : maxit, S(stream)));
  FB2_CUDA(cudaMemcpyAsync(&sc->bnorm, &bnorm, 8, cudaMemcpyHostToDevice, S(stream)));
  FB2_CUDA(cudaMemcpyAsync(&sc->rTr, &rTr, 8, cudaMemcpyHostToDevice, S(stream)));
  FB2_CUDA(cudaStreamSynchronize(S(stream)));   // bnorm / rTr live on the caller's stack
  return OK;
}
static OwnRange make_own(const int64_t* own) {
  OwnRange o{};
  if (own) { o.lo0 = own[0]; o.hi0 = own[1]; o.lo1 = own[2]; o.hi1 = own[3]; }
  return o;
}
int fb2_cg_residual(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* x,
                    const double* b, double* r, const int32_t* blk_row, int tile, int32_t max_row,
                    void* stream) {
  const SpmvPlan pl = make_plan(blk_row, nnz, tile, max_row);
  return spmv(n, nnz, crow, col, values, x, r, b, 1, nullptr, nullptr, S(stream), blk_row ? &pl : nullptr);
}
int fb2_cg_start(int64_t n, const double* r, const double* minv_diag, double* p, void* scalars, void* partial_ws,
                 const int64_t own[4], void* stream) {
  return cg_start(n, r, minv_diag, p, static_cast<CgScalars*>(scalars), partial_ws, make_own(own), S(stream));
}
int fb2_cg_spmv_dot(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* p,
                    double* Ap, const int32_t* blk_row, int tile, int32_t max_row, void* scalars,
                    void* partial_ws, const int64_t own[4], void* stream) {
  CgScalars* sc = static_cast<CgScalars*>(scalars);
  const SpmvPlan pl = make_plan(blk_row, nnz, tile, max_row);
  return spmv(n, nnz, crow, col, values, p, Ap, nullptr, 0, &sc->pAp, partial_ws, S(stream), blk_row ? &pl : nullptr, make_own(own), sc);
}
int fb2_cg_update_xr(int64_t n, double* x, double* r, const double* p, const double* Ap, const double* minv_diag, void* scalars,
                     void* partial_ws, int fuse_finalize, const int64_t own[4], void* stream) {
  return cg_update_xr(n, x, r, p, Ap, minv_diag, static_cast<CgScalars*>(scalars), partial_ws, fuse_finalize, S(stream), make_own(own));
}
// ---- multi-GPU CG over NVLink peer memory (csrc/peer.cu) -----------------------------------------------------------
int fb2_gather_f64(int64_t n, const int64_t* idx, const double* v, double* out, void* stream) { return gather_f64(n, idx, v, out, S(stream)); }
int fb2_peer_ctrl_bytes(void) { return (int)align_up(sizeof(PeerCtrl), 4096); }
int fb2_cg_spmv_dot_ranges(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* p,
                           double* Ap, const int32_t* blk_lo, const int32_t* blk_hi, int nblk, int first_halo_tile, int tile,
                           int32_t max_row, double* dot_out_dev, void* scalars, void* partial_ws, const int64_t own[4],
                           const void* ctrl_mine, int nnb, const int32_t* nb_rank_host, const uint64_t* epoch_dev, void* stream) {
  if (nblk <= 0) {                                  // no rows on this rank: the partial dot is 0
    FB2_CUDA(cudaMemsetAsync(dot_out_dev, 0, sizeof(double), S(stream)));
    return OK;
  }
  if (nnb < 0 || nnb > 2) return fail(ERR_INVALID, "cg_spmv_dot_ranges: at most two halo neighbours");
  SpmvPlan pl{};
  pl.blk_row = blk_lo; pl.blk_end = blk_hi; pl.nblk = nblk; pl.tile = tile; pl.max_row = max_row;
  if (nnb > 0) {
    pl.halo.flags = static_cast<const PeerCtrl*>(ctrl_mine)->hflag;     // (address arithmetic only: ctrl_mine is a device pointer)
    pl.halo.epoch = reinterpret_cast<const unsigned long long*>(epoch_dev);
    pl.halo.nnb = nnb; pl.halo.nb0 = nb_rank_host[0]; pl.halo.nb1 = nnb > 1 ? nb_rank_host[1] : 0;
    pl.halo.first_tile = first_halo_tile;
  }
  return spmv(n, nnz, crow, col, values, p, Ap, nullptr, 0, dot_out_dev, partial_ws, S(stream), &pl, make_own(own),
              static_cast<CgScalars*>(scalars));
}
int fb2_peer_allreduce(void* ctrl_mine, const uint64_t* peer_base_dev, int world, int rank, int kind, const double* src0,
                       const double* src1, double* dst, void* scalars, int finalize, const uint64_t* epoch_dev, void* stream) {
  return peer_allreduce(static_cast<PeerCtrl*>(ctrl_mine), reinterpret_cast<const unsigned long long*>(peer_base_dev), world, rank, kind,
                        src0, src1, dst, static_cast<CgScalars*>(scalars), finalize,
                        reinterpret_cast<const unsigned long long*>(epoch_dev), S(stream));
}
int fb2_peer_wait_halo(void* ctrl_mine, int nnb, const int32_t* nb_rank_host, const void* scalars, const uint64_t* epoch_dev,
                       void* stream) {
  return peer_wait_halo(static_cast<PeerCtrl*>(ctrl_mine), nnb, nb_rank_host, static_cast<const CgScalars*>(scalars),
                        reinterpret_cast<const unsigned long long*>(epoch_dev), S(stream));
}
int fb2_cg_update_p_push(const int64_t own[4], double* p, const double* r, const double* minv_diag, const void* scalars, int nslice,
                         const int64_t* lo, const int64_t* hi, const int64_t* peer_lo, void* const* peer_p, int nnb,
                         void* const* nb_ctrl, int rank, uint32_t* counter_dev, const uint64_t* epoch_dev, void* stream) {
  if (nslice < 0 || nslice > 4 || nnb < 0 || nnb > 2) return fail(ERR_INVALID, "cg_update_p_push: nslice=%d nnb=%d", nslice, nnb);
  PeerPush push{};
  push.nslice = nslice; push.nnb = nnb; push.rank = rank;
  for (int k = 0; k < nslice; ++k) push.slice[k] = PeerSlice{lo[k], hi[k], peer_lo[k], static_cast<double*>(peer_p[k])};
  for (int k = 0; k < nnb; ++k) push.nb_ctrl[k] = static_cast<PeerCtrl*>(nb_ctrl[k]);
  push.counter = counter_dev;
  push.epoch = reinterpret_cast<const unsigned long long*>(epoch_dev);
  return cg_update_p_push(make_own(own), p, r, minv_diag, static_cast<const CgScalars*>(scalars), push, S(stream));
}

int64_t fb2_box_edges_before(int nx, int ny, int nz, int i, int j, int k) { return box_edges_before_host(nx, ny, nz, i, j, k); }
int fb2_tet_box_slab(const double box[6], int nx, int ny, int nz, int cube_layer_lo, int cube_layer_hi, int p, double* node,
                     int32_t* cell, int32_t* cell2dof, void* stream) {
  return tet_box_slab(box, nx, ny, nz, cube_layer_lo, cube_layer_hi, p, node, cell, cell2dof, S(stream));
}
int fb2_cg_finalize(void* scalars, void* stream) { return cg_finalize(static_cast<CgScalars*>(scalars), S(stream)); }
int fb2_cg_update_p(int64_t n, double* p, const double* r, const double* minv_diag, void* scalars, void* stream) {
  return cg_update_p(n, p, r, minv_diag, static_cast<CgScalars*>(scalars), S(stream));
}

int fb2_bcg_dots(int64_t n, int nb, const double* a, const double* b, double* out_dev, void* partial_ws, void* stream) {
  return bcg_dots(n, nb, a, b, out_dev, partial_ws, S(stream));
}
int fb2_bcg_update_xr(int64_t n, int nb, double* x, double* r, const double* p, const double* Ap, const double* rTr, const double* pAp,
                      const double* state, void* stream) {
  return bcg_update_xr(n, nb, x, r, p, Ap, rTr, pAp, state, S(stream));
}
int fb2_bcg_update_p(int64_t n, int nb, double* p, const double* r, const double* minv_diag, const double* rTr_new, const double* rTr,
                     const double* state, void* stream) {
  return bcg_update_p(n, nb, p, r, minv_diag, rTr_new, rTr, state, S(stream));
}
int fb2_bcg_check(int nb, const double* rTr_new, double atol, double rtol_bnorm, int maxit, double* state, void* stream) {
  if (!state) return fail(ERR_INVALID, "bcg_check: state is required");
  return bcg_check(nb, rTr_new, atol, rtol_bnorm, maxit, state, S(stream));
}

// ---- next rows: source vector, Dirichlet -----------------------------------------------------
int fb2_elem_source(int TD, int64_t NC, int ldof, int NQ, const double* node, const int32_t* cell, const double* phiw, int kind,
                    double scal, const double* f, double* out, void* stream) {
  if (kind < 0 || kind > 2) return fail(ERR_INVALID, "elem_source: kind must be 0 (scalar), 1 (NC,) or 2 (NC,NQ)");
  return elem_source(TD, NC, ldof, NQ, node, cell, phiw, kind, scal, f, out, S(stream));
}
int fb2_bc_to_points(int TD, int64_t NC, int NQ, const double* node, const int32_t* cell, const double* bcs, double* out, void* stream) {
  return bc_to_points(TD, NC, NQ, node, cell, bcs, out, S(stream));
}
int fb2_gather_vector(int64_t gdof, const int64_t* adj_ptr, const int32_t* adj_pair, const double* fe, double* F, void* stream) {
  return gather_vector(gdof, adj_ptr, adj_pair, fe, F, S(stream));
}
int fb2_matfree_apply(int64_t gdof, int ldof, const int64_t* adj_ptr, const int32_t* adj_pair, const int32_t* cell2dof,
                      const double* Ke, const double* u, double* v, void* stream) {
  return matfree_apply(gdof, ldof, adj_ptr, adj_pair, cell2dof, Ke, u, v, S(stream));
}
size_t fb2_adjacency_workspace_bytes(int64_t gdof) { return adjacency_workspace_bytes(gdof); }
int fb2_adjacency(const int32_t* c2d, int64_t NC, int ldof, int64_t gdof, int64_t* adj_ptr, int32_t* adj_pair, void* ws, void* stream) {
  return build_adjacency(c2d, NC, ldof, gdof, adj_ptr, adj_pair, ws, S(stream));
}
int fb2_pair_positions(int64_t npos, const int32_t* adj_pair, int32_t* pair_pos, void* stream) {
  return pair_positions(npos, adj_pair, pair_pos, S(stream));
}
int fb2_matfree_scalar_const(int TD, int p, int64_t NC, int64_t gdof, const double* node, const int32_t* cell, const int32_t* cell2dof,
                             const int64_t* adj_ptr, const int32_t* adj_pair, const int32_t* pair_pos, const double* Ms_host,
                             const double* Mm_host, double scal_d, const double* coef_d, double scal_m, const double* coef_m,
                             const double* u, double* cell_ws, double* v, void* stream) {
  if (!Ms_host && !Mm_host) return fail(ERR_INVALID, "matfree_scalar_const: need a diffusion and/or a mass table");
  MatfreeArgs a{};
  a.node = node; a.cell = cell; a.c2d = cell2dof; a.NC = NC; a.Ms_host = Ms_host; a.Mm_host = Mm_host;
  a.scal_d = scal_d; a.scal_m = scal_m; a.coef_d = coef_d; a.coef_m = coef_m; a.u = u; a.w = cell_ws; a.pair_pos = pair_pos;
  FB2_TRY(matfree_scalar_const(TD, p, a, S(stream)));
  if (pair_pos) return segment_sum(gdof, adj_ptr, cell_ws, v, S(stream));
  return gather_vector(gdof, adj_ptr, adj_pair, cell_ws, v, S(stream));
}
size_t fb2_bc_workspace_bytes(int64_t n) { return bc_workspace_bytes(n); }
int fb2_bc_matrix_count(int64_t n, const int64_t* crow, const int32_t* col, const uint8_t* isbd, int64_t* crow_new,
                        int64_t* nnz_host, void* ws, void* stream) {
  return bc_matrix_count(n, crow, col, isbd, crow_new, nnz_host, ws, S(stream));
}
int fb2_bc_matrix_fill(int64_t n, const int64_t* crow, const int32_t* col, const double* values, const uint8_t* isbd,
                       const int64_t* crow_new, int32_t* col_new, double* values_new, void* stream) {
  return bc_matrix_fill(n, crow, col, values, isbd, crow_new, col_new, values_new, S(stream));
}
int fb2_bc_vector(int64_t n, const uint8_t* isbd, const double* uh, double* f, void* stream) {
  return bc_vector(n, isbd, uh, f, S(stream));
}

// ---- raw primitives -------------------------------------------------------------------------
size_t fb2_sort_workspace_bytes(int64_t n) { return sort_workspace_bytes(n); }
int fb2_sort_pairs(uint64_t* keys, uint32_t* vals, int identity_payload, int64_t n, int key_bits, void* ws, void* stream) {
  uint64_t* ko = nullptr;
  uint32_t* vo = identity_payload ? nullptr : vals;
  FB2_TRY(radix_sort_pairs(keys, vals, n, key_bits, ws, S(stream), &ko, &vo));
  if (n > 0 && ko != keys) {
    FB2_CUDA(cudaMemcpyAsync(keys, ko, (size_t)n * 8, cudaMemcpyDeviceToDevice, S(stream)));
    FB2_CUDA(cudaMemcpyAsync(vals, vo, (size_t)n * 4, cudaMemcpyDeviceToDevice, S(stream)));
  }
  return OK;
}
size_t fb2_scan_workspace_bytes(int64_t n) { return scan_workspace_bytes(n); }
int fb2_exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n, void* ws, void* stream) {
  return exclusive_scan_i32(in, out, n, true, ws, S(stream));
}

}  // extern "C"
