/* fealpy_b200 -- C ABI of the B200-native (sm_100a) Lagrange-FEM assembly + CG hot path.
 *
 * This is the drop-in boundary for FEALPy's path
 *     BilinearForm(space).add_integrator(...).assembly() -> CSRTensor -> fealpy.solver.cg
 * Every entry point takes plain device pointers (from tensor.data_ptr()), sizes and a
 * cudaStream_t passed as void*; no torch types.  All functions return 0 on success or a
 * non-zero FB2_ERR_* code; fb2_last_error() returns the message of the calling thread's last
 * failure.  Paths cited below are relative to the reference tree (FEALPy 3.4.0).
 *
 * Memory protocol: outputs and workspaces are caller-allocated (the Python shim allocates
 * torch tensors).  Where an output size is data dependent (nnz, number of edges) the call is
 * split in a "symbolic/count" step that returns the size through a host pointer (it
 * synchronises the stream) and a "fill" step.
 */
#ifndef FEALPY_B200_H
#define FEALPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FB2_OK = 0, FB2_ERR_INVALID = 1, FB2_ERR_CUDA = 2, FB2_ERR_UNSUPPORTED = 3, FB2_ERR_WORKSPACE = 4 };

const char* fb2_last_error(void);
int fb2_version(void);

/* ---- mesh generation + topology -> DOF numbering --------------------------------------
 * replaces TriangleMesh.from_box (mesh/triangle_mesh.py:1386-1435), TetrahedronMesh.from_box
 * (mesh/tetrahedron_mesh.py:1016-1086), MeshDS.construct (mesh/mesh_data_structure.py:428-464),
 * cell_to_ipoint (mesh/triangle_mesh.py:218-270, mesh/tetrahedron_mesh.py:388-441) and
 * to_tensor_dof (functionspace/utils.py:83-95). */
int fb2_tri_from_box(const double box[4], int nx, int ny, double* node, int32_t* cell, void* stream);
int fb2_tet_from_box(const double box[6], int nx, int ny, int nz, double* node, int32_t* cell, void* stream);
size_t fb2_entity_workspace_bytes(int64_t NC, int per_cell);
/* kind: 1 = edges, 2 = faces (tets).  Step 1 fills cell2ent (NC x per_cell) and *count. */
int fb2_build_entities(const int32_t* cell, int64_t NC, int TD, int kind, int64_t NN, int32_t* cell2ent, int64_t* count_host,
                       void* ws, void* stream);
/* Step 2 (same ws): ent (count x nve) = first-occurrence vertex tuples. */
int fb2_entities_emit(const int32_t* cell, int64_t NC, int TD, int kind, int32_t* cell2ent, int32_t* ent, void* ws, void* stream);
int fb2_cell_to_dof(const int32_t* cell, const int32_t* cell2edge, const int32_t* edge, const int32_t* cell2face, int64_t NC, int TD,
                    int p, int64_t NN, int64_t NE, int64_t NF, const unsigned char* multi_index_host, int ldof, int32_t* cell2dof,
                    void* stream);
int fb2_tensor_cell_to_dof(const int32_t* cell2dof, int64_t NC, int ldof, int GD, int64_t gdof, int dof_priority, int32_t* out,
                           void* stream);
/* x-slab of TetrahedronMesh.from_box for the multi-GPU row partition: cube layers
 * [cube_layer_lo, cube_layer_hi), node planes [lo, hi]; node coordinates and cell2dof follow the
 * GLOBAL from_box numbering in closed form (no sort), expressed in the slab's window:
 * nodes -> [0, NNw), edges whose smaller node lies in the window -> NNw + (global edge id -
 * fb2_box_edges_before(nx,ny,nz,lo,0,0)).  p = 1 or 2.  node / cell may be NULL (only cell2dof is wanted: the whole box,
 * layers [0, nx), gives the global P2 numbering of a from_box mesh without building its edges). */
int64_t fb2_box_edges_before(int nx, int ny, int nz, int i, int j, int k);
int fb2_tet_box_slab(const double box[6], int nx, int ny, int nz, int cube_layer_lo, int cube_layer_hi, int p, double* node,
                     int32_t* cell, int32_t* cell2dof, void* stream);

/* ---- K1: element matrices ---------------------------------------------------------------
 * replaces ScalarDiffusionIntegrator.assembly / ScalarMassIntegrator.assembly /
 * LinearElasticityIntegrator.assembly (fem/scalar_diffusion_integrator.py:54-79,
 * fem/scalar_mass_integrator.py:47-54, fem/linear_elasticity_integrator.py:60-181) and
 * bilinear_integral (functional.py:68-106).  out is (NC, l, l) C-order float64.
 *
 * constant / per-cell coefficients:  out = sd*cd[c] * (Ms : G_c) + sm*cm[c] * vol_c * Mm
 *   Ms (l,l,NG) device table, NG = (TD+1)(TD+2)/2;  Mm (l,l) device table; either may be NULL. */
int fb2_elem_scalar_const(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const double* Ms, const double* Mm,
                          double scal_d, const double* coef_d_cell, double scal_m, const double* coef_m_cell, double* out,
                          void* stream);
/* quadrature loop: is_mass 0 -> table = R (NQ,l,TD+1), 1 -> table = phi (NQ,l);
 * coef_kind 2 -> coef (NC,NQ), 3 -> coef (NC,NQ,GD,GD) (diffusion only). */
int fb2_elem_scalar_quad(int TD, int p, int64_t NC, const double* node, const int32_t* cell, int is_mass, int NQ, const double* ws,
                         const double* table, int coef_kind, const double* coef, double* out, void* stream);
/* The same kernels for a SUM of integrators written into one block (GroupIntegrator / several groups on the K_e gather path):
 * `accumulate` != 0 adds to `out` instead of overwriting it, and the quadrature kernel can fold constant / per-cell coefficient
 * terms (the K of fb2_elem_scalar_const: Ms_const / Mm_const device tables or NULL) into its own final write -- a variable
 * diffusion plus a constant mass term then writes K_e once. */
int fb2_elem_scalar_const_acc(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const double* Ms, const double* Mm,
                              double scal_d, const double* coef_d_cell, double scal_m, const double* coef_m_cell, double* out,
                              int accumulate, void* stream);
int fb2_elem_scalar_quad_fused(int TD, int p, int64_t NC, const double* node, const int32_t* cell, int is_mass, int NQ,
                               const double* ws, const double* table, int coef_kind, const double* coef, const double* Ms_const,
                               const double* Mm_const, double scal_d, const double* coef_d_cell, double scal_m,
                               const double* coef_m_cell, double* out, int accumulate, void* stream);
/* isotropic linear elasticity; M4 (l,l,TD+1,TD+1) device table; out (NC, GD*l, GD*l). */
int fb2_elem_elasticity(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const double* M4, double d_diag,
                        double d_lam, double d_shear, int dof_priority, double* out, void* stream);

/* ---- cell geometry and basis gradients as public arrays ----------------------------------------
 * replaces mesh.entity_measure('cell') / cell_volume / grad_lambda (mesh/triangle_mesh.py:45-70,115-129,
 * mesh/tetrahedron_mesh.py:144-175,208-219 -> backend simplex_measure / triangle_grad_lambda_2d /
 * tetrahedron_grad_lambda_3d, backend/numpy_backend.py:413-421,586-629) and
 * mesh.grad_shape_function(variables='x') = LagrangeFESpace.grad_basis (mesh/mesh_base.py:713-749,
 * functionspace/lagrange_fe_space.py:146-154).  The same device functions feed the element kernels.
 * out (NC, (TD+1)*TD + 1): grad lambda_k[x] (k-major), then the signed measure. */
int fb2_cell_gradients(int TD, int64_t NC, const double* node, const int32_t* cell, double* out, void* stream);
/* out (NC, NQ, ldof, TD) = sum_b R[q][i][b] * Dlambda[c][b][m]; R (NQ, ldof, TD+1) device table (dphi/dlambda) */
int fb2_grad_basis(int TD, int64_t NC, int NQ, int ldof, const double* cell_gradient_records, const double* R, double* out,
                   void* stream);

/* ---- K2: deterministic COO -> CSR ---------------------------------------------------------
 * replaces BilinearForm._scalar_assembly index build (fem/bilinear_form.py:46-75),
 * COOTensor.coalesce (sparse/coo_tensor.py:184-213) and COOTensor.tocsr (:137-157). */
int fb2_coo_keys_from_c2d(const int32_t* row_dof, const int32_t* col_dof, int64_t NC, int lr, int lc, int col_bits, uint64_t* keys,
                          void* stream);
int fb2_coo_keys_from_coo(const void* row, const void* col, int index_bytes, int64_t n, int col_bits, uint64_t* keys, void* stream);
size_t fb2_coo_workspace_bytes(int64_t n);
/* sorts keys in place (perm = original positions), returns the number of distinct keys */
int fb2_coo_symbolic(uint64_t* keys, uint32_t* perm, int64_t n, int key_bits, void* ws, int64_t* nnz_host, void* stream);
/* crow (nrow+1, int64), col (nnz, int32 or int64 by col_bytes), seg_start (nnz+1, int64) */
int fb2_coo_fill(const uint64_t* keys, int64_t n, int col_bits, int64_t nrow, void* ws, int64_t* crow, void* col, int col_bytes,
                 int64_t* seg_start, void* stream);
/* values_out[s] = sum (left to right, original COO order) of values_in[perm[k]], k in segment s */
int fb2_coo_reduce(const uint32_t* perm, const int64_t* seg_start, int64_t nnz, const double* values_in, double* values_out,
                   void* stream);

/* ---- symbolic + fused numeric assembly (pattern cached per space) -------------------------
 * Same result as K1 + K2, without materialising the COO: the CSR pattern is a pure function
 * of cell2dof (fem/bilinear_form.py:69-72), so it is built once (symbolic) and every row then
 * sums its incident cells in ascending cell order (numeric). */
size_t fb2_sym_workspace_bytes(int64_t NC, int ldof, int64_t gdof);
/* step 1: dof -> (cell, local index) adjacency, row lengths; returns nnz and max row length.
 * stash (optional, NC*ldof*ldof uint16 of scratch, may be NULL): the count pass has to rank every
 * row's candidate columns anyway; with a stash it keeps (rank << 1 | representative) per candidate and
 * fb2_sym_fill given the same stash replays it instead of ranking a second time. */
int fb2_sym_count(const int32_t* cell2dof, int64_t NC, int ldof, int64_t gdof, int64_t* adj_ptr, int32_t* adj_pair, int64_t* crow,
                  int64_t* nnz_host, int32_t* max_row_host, uint16_t* stash, void* ws, void* stream);
/* step 2: col (nnz) and slot map: NC*ldof records of fb2_slot_stride(ldof, slot_bytes) elements
 * (uint8 when max_row<=255 else uint16; records padded to 4-byte multiples) */
int fb2_slot_stride(int ldof, int slot_bytes);
int fb2_sym_fill(const int32_t* cell2dof, int64_t NC, int ldof, int64_t gdof, const int64_t* adj_ptr, const int32_t* adj_pair,
                 const int64_t* crow, int32_t* col, void* slots, int slot_bytes, const uint16_t* stash, void* stream);
/* numeric, constant / per-cell coefficient scalar forms (diffusion and/or mass fused).
 * blk_row/nblk/tile: row tiling of crow from fb2_spmv_plan_build (one CTA per tile).  This is the
 * "v2" kernel (lane = row, tables in shared memory); it also serves rows too long for the 12-bit tile
 * offsets of the v4 kernel below. */
int fb2_assemble_scalar_const(int TD, int p, int64_t NC, int64_t gdof, const double* node, const int32_t* cell,
                              const int64_t* adj_ptr, const int32_t* adj_pair, const void* slots, int slot_bytes,
                              const int64_t* crow, int32_t max_row, const int32_t* blk_row, int nblk, int tile,
                              const double* Ms, const double* Mm, double scal_d, const double* coef_d_cell, double scal_m,
                              const double* coef_m_cell, double* values, void* stream);
/* numeric, generic: gathers rows of a precomputed element-matrix block Ke (NC, lt, lt);
 * ncomp > 1 = tensor space over the scalar pattern (interleaved or dof-priority layout) */
int fb2_assemble_from_ke(int64_t NC, int ldof, int ncomp, int dof_priority, int64_t gdof_scalar, const double* Ke,
                         const int64_t* adj_ptr, const int32_t* adj_pair, const void* slots, int slot_bytes,
                         const int64_t* crow_scalar, int32_t max_row, const int64_t* crow_out, const int32_t* blk_row, int nblk,
                         int tile, double* values, void* stream);
/* LinearElasticityIntegrator on a P1 tensor space, fused: grad(phi_i) = grad(lambda_i) is constant per cell, so the
 * element-matrix entry is a closed form of a (TD+1)*TD+1-double per-cell record (geo_ws: that many doubles per
 * cell of scratch) and the gather never materialises K_e (fem/linear_elasticity_integrator.py:159-179);
 * d_diag / d_lam / d_shear as in fb2_elem_elasticity, wsum = sum of the quadrature weights (M4[0,0,0,0]);
 * pattern arguments as in fb2_assemble_from_ke with ncomp = TD */
int fb2_assemble_elasticity_p1(int TD, int64_t NC, const double* node, const int32_t* cell, int dof_priority, int64_t gdof_scalar,
                               double d_diag, double d_lam, double d_shear, double wsum, const int64_t* adj_ptr,
                               const int32_t* adj_pair, const void* slots, int slot_bytes, const int64_t* crow_scalar, int32_t max_row,
                               const int64_t* crow_out, const int32_t* blk_row, int nblk, int tile, double* geo_ws, double* values,
                               void* stream);
/* "v4" numeric kernel for the same forms: the symbolic phase additionally turns every warp tile
 * (rows holding <= tile values; fb2_spmv_plan_build) into 32-entry batches that share the local
 * index (warp-uniform element-table row) and touch 32 different rows (conflict-free adds);
 * geom_ws = NC * 2*ceil((TD*(TD+1)/2+1)/2) doubles of scratch for the per-cell geometry (8 per
 * tetrahedron, 4 per triangle).  The element tables are HOST pointers here: they travel in the kernel
 * parameter block (constant bank), which costs no load/store-unit bandwidth.
 * The schedule is ONE packed array `blocks`: per batch BW = 64 + 32*W + 4 uint32 words (W = words of a slot record =
 * fb2_slot_stride(ldof, slot_bytes) * slot_bytes / 4), fetched by the numeric kernel with a single bulk copy:
 *   words [0,32)        cell id of lane's entry, 0xffffffff = pad lane
 *   words [32,64)       offset of the entry's row inside its tile (low 12 bits: tile + max_row < 4096) | mask of the columns
 *                       whose value this entry is the FIRST to touch (bit j + 12): those are stored, the others
 *                       load-add-stored, so the accumulator tile needs no zero fill
 *   words [64,64+32W)   the 32 slot records (position of the cell's j-th dof inside the row)
 *   word  64+32W        the batch's local index i (also in batch_i[b]); 3 pad words
 * `blocks` must be zero-initialised by the caller before fb2_asm4_plan_fill. */
size_t fb2_asm4_workspace_bytes(int ntile);
int fb2_asm4_plan_count(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int32_t* adj_pair,
                        int ldof, int64_t* batch_ptr, int64_t* nbatch_host, void* ws, void* stream);
int fb2_asm4_plan_fill(int ntile, const int32_t* blk_row, const int64_t* crow, const int64_t* adj_ptr, const int32_t* adj_pair,
                       int ldof, const int64_t* batch_ptr, uint8_t* batch_i, uint32_t* blocks, const void* slots, int slot_bytes,
                       void* stream);
int fb2_assemble_scalar_const_v4(int TD, int p, int64_t NC, const double* node, const int32_t* cell, const int64_t* crow,
                                 const int32_t* blk_row, int ntile, int tile, int32_t max_row, const int64_t* batch_ptr,
                                 const uint8_t* batch_i, const uint32_t* blocks, int slot_bytes, const double* Ms_host, const double* Mm_host, double scal_d,
                                 const double* coef_d_cell, double scal_m, const double* coef_m_cell, double* geom_ws, double* values,
                                 void* stream);
/* expands the scalar pattern to the tensor-space pattern */
int fb2_expand_pattern(int64_t gdof_scalar, int ncomp, int dof_priority, const int64_t* crow_scalar, const int32_t* col_scalar,
                       int64_t* crow_out, int32_t* col_out, void* stream);

/* ---- K3/K4: SpMV + CG ----------------------------------------------------------------------
 * replaces CSRTensor.matmul -> bm.csr_spmm (sparse/csr_tensor.py:411-452,
 * backend/numpy_backend.py:180-199) and fealpy.solver.cg (solver/cg.py:14-123). */
size_t fb2_partial_workspace_bytes(void);           /* zero-initialise once */
/* SpMV plan (built once per matrix): row-aligned tiles of `tile` stored values.
 * blk_row has fb2_spmv_plan_blocks(nnz, tile) + 1 int32 entries; *max_row_host = longest row. */
int fb2_spmv_plan_blocks(int64_t nnz, int tile);
int fb2_spmv_plan_build(int64_t n, const int64_t* crow, int tile, int32_t* blk_row, int64_t nnz, int32_t* max_row_host, void* stream);
/* y = A x.  blk_row may be NULL (row-per-lane-group kernel); with a plan the streaming kernel runs. */
int fb2_csr_spmv(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* x,
                 double* y, const int32_t* blk_row, int tile, int32_t max_row, void* stream);
int fb2_csr_spmm(int64_t n, const int64_t* crow, const int32_t* col, const double* values, const double* X, double* Y, int nb,
                 void* stream);
int fb2_dot(int64_t n, const double* a, const double* b, double* out_dev, void* partial_ws, void* stream);
size_t fb2_cg_workspace_bytes(int64_t n, int64_t nnz);
/* x: x0 on entry, solution on exit.  minv_diag: NULL or the diagonal of M (z = M r).
 * maxit < 0 means "no limit" (reference maxit=None).  chunk <= 0: automatic.
 * blk_row / tile / max_row: a prebuilt SpMV plan of the matrix (fb2_spmv_plan_build); blk_row NULL: the solver builds one in its
 * workspace. */
int fb2_cg(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* b, double* x,
           const double* minv_diag, double atol, double rtol, int maxit, int chunk, const int32_t* blk_row,
           int tile, int32_t max_row, void* ws, int* niter_host, double* residual_host, void* stream);
/* building blocks of the distributed driver (device-resident scalars in `scalars`, 256 bytes) */
int fb2_cg_init(void* scalars, double atol, double rtol, int maxit, double bnorm, double rTr, void* stream);
/* own[4] = {lo0, hi0, lo1, hi1}: rows (owned dofs of this rank) that contribute to the dot
 * products; NULL = all rows.  The caller all-reduces scalars[1] (p.Ap) / scalars[2] (r.z). */
int fb2_cg_residual(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* x,
                    const double* b, double* r, const int32_t* blk_row, int tile, int32_t max_row,
                    void* stream);
int fb2_cg_start(int64_t n, const double* r, const double* minv_diag, double* p, void* scalars, void* partial_ws,
                 const int64_t own[4], void* stream);
int fb2_cg_spmv_dot(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* p,
                    double* Ap, const int32_t* blk_row, int tile, int32_t max_row, void* scalars,
                    void* partial_ws, const int64_t own[4], void* stream);
int fb2_cg_update_xr(int64_t n, double* x, double* r, const double* p, const double* Ap, const double* minv_diag, void* scalars,
                     void* partial_ws, int fuse_finalize, const int64_t own[4], void* stream);
int fb2_cg_finalize(void* scalars, void* stream);
int fb2_cg_update_p(int64_t n, double* p, const double* r, const double* minv_diag, void* scalars, void* stream);

/* ---- multi-GPU CG over NVLink peer memory ----------------------------------------------------
 * The halo exchange and the two scalar all-reduces of a CG iteration, done by the iteration's own kernels with stores
 * into the neighbours' peer-mapped (symmetric) memory + sequence flags: no NCCL call inside the iteration.  Every rank owns a
 * control block of fb2_peer_ctrl_bytes() bytes at the start of its symmetric buffer, followed by its vector p; peer_base_dev
 * is a device array of the `world` base addresses.  *epoch_dev (device) is the base of the solve's sequence numbers (the host
 * raises it by 2^32 per solve).  See csrc/peer.cu for the protocol and its ordering argument.
 *   fb2_cg_spmv_dot_ranges : Ap = A p on the rows of a multi-range tile plan (blk_lo / blk_hi per tile), fused partial p.Ap;
 *                            tiles >= first_halo_tile (the boundary rows, placed last) wait inside the kernel until the
 *                            neighbours' halo pushes of the previous iteration have landed -- the interior tiles overlap them
 *   fb2_peer_wait_halo     : (1 warp) the same wait as a stand-alone kernel
 *   fb2_peer_allreduce     : (1 warp) *dst = sum over ranks (in rank order, bit-identical everywhere) of *src0 (+ *src1);
 *                            finalize != 0 applies the stopping rules of solver/cg.py:97-121 to the reduced r.z
 *   fb2_cg_update_p_push   : p = z + beta p on the owned rows, owned boundary slices stored into the neighbours' p, flags raised */
/* pack kernel of the general-mesh partition (parallel/mesh_partition.py): out[k] = v[idx[k]], idx int64 local ids */
int fb2_gather_f64(int64_t n, const int64_t* idx, const double* v, double* out, void* stream);
int fb2_peer_ctrl_bytes(void);
int fb2_cg_spmv_dot_ranges(int64_t n, int64_t nnz, const int64_t* crow, const int32_t* col, const double* values, const double* p,
                           double* Ap, const int32_t* blk_lo, const int32_t* blk_hi, int nblk, int first_halo_tile, int tile,
                           int32_t max_row, double* dot_out_dev, void* scalars, void* partial_ws, const int64_t own[4],
                           const void* ctrl_mine, int nnb, const int32_t* nb_rank_host, const uint64_t* epoch_dev, void* stream);
int fb2_peer_allreduce(void* ctrl_mine, const uint64_t* peer_base_dev, int world, int rank, int kind, const double* src0,
                       const double* src1, double* dst, void* scalars, int finalize, const uint64_t* epoch_dev, void* stream);
int fb2_peer_wait_halo(void* ctrl_mine, int nnb, const int32_t* nb_rank_host, const void* scalars, const uint64_t* epoch_dev,
                       void* stream);
int fb2_cg_update_p_push(const int64_t own[4], double* p, const double* r, const double* minv_diag, const void* scalars, int nslice,
                         const int64_t* lo, const int64_t* hi, const int64_t* peer_lo, void* const* peer_p, int nnb,
                         void* const* nb_ctrl, int rank, uint32_t* counter_dev, const uint64_t* epoch_dev, void* stream);

/* batched right-hand sides, b of shape (n, nb) row-major (solver/cg.py:88-121): per-column dots,
 * x/r update with alpha_k = rTr[k]/pAp[k], p update with beta_k = rTr_new[k]/rTr[k]; the host drives
 * the loop with fb2_csr_spmm.  All scalar arrays live on the device. */
int fb2_bcg_dots(int64_t n, int nb, const double* a, const double* b, double* out_dev, void* partial_ws, void* stream);
int fb2_bcg_update_xr(int64_t n, int nb, double* x, double* r, const double* p, const double* Ap, const double* rTr, const double* pAp,
                      const double* state, void* stream);
int fb2_bcg_update_p(int64_t n, int nb, double* p, const double* r, const double* minv_diag, const double* rTr_new, const double* rTr,
                     const double* state, void* stream);
/* the joint stopping test on the device (solver/cg.py:97-121: r_norm = sqrt(sum_k rTr_new[k]) against atol, then rtol*|b|, then
 * maxit; maxit < 0 = none).  `state`: 4 device doubles, zero-initialised by the caller = (done flag, iterations, r_norm, pad);
 * once done is raised the update kernels that receive `state` do nothing, so the host may launch several iterations per poll
 * and x remains the iterate of the stopping iteration. */
int fb2_bcg_check(int nb, const double* rTr_new, double atol, double rtol_bnorm, int maxit, double* state, void* stream);

/* ---- next rows: right-hand side and Dirichlet step ------------------------
 * replaces ScalarSourceIntegrator.assembly + LinearForm.assembly (fem/scalar_source_integrator.py:13-57,
 * fem/linear_form.py:36-86) and DirichletBC.apply (fem/dirichlet_bc.py:101-235). */
/* F_e (NC,l) = vol_c * scal * sum_q phiw[q][i] f_cq; kind 0: f = 1, 1: f (NC,), 2: f (NC,NQ) */
int fb2_elem_source(int TD, int64_t NC, int ldof, int NQ, const double* node, const int32_t* cell, const double* phiw, int kind,
                    double scal, const double* f, double* out, void* stream);
/* physical points of NQ barycentric points (bcs: NQ x (TD+1), device) in every cell, out (NC, NQ, TD):
 * mesh.bc_to_point (mesh/mesh_base.py:454-478 -> bm.bc_to_points, backend/numpy_backend.py:401-407), where callable
 * coefficients / sources are evaluated */
int fb2_bc_to_points(int TD, int64_t NC, int NQ, const double* node, const int32_t* cell, const double* bcs, double* out, void* stream);
/* F[d] = sum of F_e over the (cell, i) pairs of dof d in ascending order (adjacency of fb2_sym_count) */
int fb2_gather_vector(int64_t gdof, const int64_t* adj_ptr, const int32_t* adj_pair, const double* fe, double* F, void* stream);
/* matrix-free v = A u from the element matrices Ke (NC,l,l) (BilinearForm.__matmul__ before assembly,
 * fem/bilinear_form.py:126-158): row-owner gather over the adjacency of fb2_sym_count, no atomics */
int fb2_matfree_apply(int64_t gdof, int ldof, const int64_t* adj_ptr, const int32_t* adj_pair, const int32_t* cell2dof,
                      const double* Ke, const double* u, double* v, void* stream);
/* dof -> (cell, local index) adjacency alone (first half of fb2_sym_count): all a matrix-free product or a load vector
 * needs of the symbolic phase.  ws: fb2_adjacency_workspace_bytes(gdof) */
size_t fb2_adjacency_workspace_bytes(int64_t gdof);
int fb2_adjacency(const int32_t* cell2dof, int64_t NC, int ldof, int64_t gdof, int64_t* adj_ptr, int32_t* adj_pair, void* ws,
                  void* stream);
/* fused matrix-free v = A u for scalar diffusion + mass forms with constant / per-cell coefficients (the forms
 * fb2_assemble_scalar_const_v4 assembles): neither A nor K_e is formed.  One thread per cell recomputes the geometry,
 * gathers u through cell2dof and writes w_c = K_c u_c (NC, ldof) to `cell_ws`; the row-owner gather over the adjacency
 * sums it per dof in fixed order (no atomics).  Replaces einsum('cij,cj->ci') + index_add of fem/bilinear_form.py:126-158.
 * Ms_host / Mm_host: host tables as for fb2_assemble_scalar_const_v4 (null = term absent). */
int fb2_matfree_scalar_const(int TD, int p, int64_t NC, int64_t gdof, const double* node, const int32_t* cell,
                             const int32_t* cell2dof, const int64_t* adj_ptr, const int32_t* adj_pair, const int32_t* pair_pos,
                             const double* Ms_host, const double* Mm_host, double scal_d, const double* coef_d_cell, double scal_m,
                             const double* coef_m_cell, const double* u, double* cell_ws, double* v, void* stream);
/* pair_pos (NC*ldof int32, optional argument of fb2_matfree_scalar_const): the position of every (cell, i) pair in the
 * adjacency lists.  With it the cell kernel writes its products in adjacency order and the per-dof sum reads one contiguous
 * run (same summation order, bit-identical result; NULL = pair order + gather). */
int fb2_pair_positions(int64_t npos, const int32_t* adj_pair, int32_t* pair_pos, void* stream);
size_t fb2_bc_workspace_bytes(int64_t n);
/* canonical CSR of the constrained matrix: boundary rows/columns removed, unit diagonal on boundary rows */
int fb2_bc_matrix_count(int64_t n, const int64_t* crow, const int32_t* col, const uint8_t* isbd, int64_t* crow_new,
                        int64_t* nnz_host, void* ws, void* stream);
int fb2_bc_matrix_fill(int64_t n, const int64_t* crow, const int32_t* col, const double* values, const uint8_t* isbd,
                       const int64_t* crow_new, int32_t* col_new, double* values_new, void* stream);
/* f[r] = uh[r] on boundary dofs (f must already hold f - A uh, e.g. from fb2_cg_residual) */
int fb2_bc_vector(int64_t n, const uint8_t* isbd, const double* uh, double* f, void* stream);

/* ---- raw primitives (exported for tests) ---------------------------------------------------*/
size_t fb2_sort_workspace_bytes(int64_t n);
int fb2_sort_pairs(uint64_t* keys, uint32_t* vals, int identity_payload, int64_t n, int key_bits, void* ws, void* stream);
size_t fb2_scan_workspace_bytes(int64_t n);
int fb2_exclusive_scan_i32(const int32_t* in, int64_t* out, int64_t n, void* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEALPY_B200_H */
