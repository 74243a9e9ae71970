/*
 * spis_b200.h -- C ABI of the B200-native conservative-FGMRES ("CGMRES") hot path.
 *
 * This is the drop-in boundary below the Python `solvers` module.  The reference
 * (JamesJackaman/StructurePreservingIterativeSolvers) has no native layer: every O(n)
 * operation of solvers.py is a numpy/scipy call.  Each entry point below names the
 * reference lines (solvers.py:<line>) whose arithmetic it replaces.  The host side
 * (structurepreservingiterativesolvers_b200/solvers.py) binds these with ctypes.
 *
 * Conventions
 *   - every pointer argument is HOST memory owned by the caller (plain or pinned);
 *     the library never keeps a host pointer after the call returns;
 *   - all functions return 0 on success, a negative SPIS_E_* code on failure, and
 *     never throw; spis_last_error(ctx) gives the message of the last failure;
 *   - a context owns one CUDA stream; calls on one context are synchronous with
 *     respect to each other (the *_launch/*_wait pairs expose the one place where the
 *     device runs ahead of the host);
 *   - doubles are IEEE fp64, CSR indices are int32 (PETSc getValuesCSR ->
 *     scipy.sparse.csr_matrix, lkdv/lkdv.py:109-110).
 *
 * There is deliberately no CPU fallback: if no sm_100 device is present
 * spis_ctx_create fails with SPIS_E_CUDA.
 */
#ifndef SPIS_B200_H
#define SPIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define SPIS_ABI_VERSION 5

/* error codes */
#define SPIS_OK            0
#define SPIS_E_INVALID    -1   /* bad argument / call order */
#define SPIS_E_CUDA       -2   /* CUDA runtime error (message has cudaGetErrorString) */
#define SPIS_E_NOMEM      -3   /* device or pinned allocation failed */
#define SPIS_E_UNSUPPORTED -4

/* matrix slots: 0 = system matrix A, 1 = sparse preconditioner P (z = P q),
 * 2.. = constraint matrices M_c (slot 2+c).                                  */
#define SPIS_SLOT_A      0
#define SPIS_SLOT_PRE    1
#define SPIS_SLOT_CON0   2
#define SPIS_MAX_SLOTS   18

/* vector ids for spis_upload_vec / spis_download_vec */
#define SPIS_VEC_B        0   /* right-hand side b                       (solvers.py:167) */
#define SPIS_VEC_X0       1   /* initial guess x0                        (solvers.py:167) */
#define SPIS_VEC_R0       2   /* r0 = b - A x0 (download only; dict['x'][0], quirk Q1, solvers.py:169) */
#define SPIS_VEC_Q        3   /* Arnoldi basis vector q[j]               (solvers.py:172) */
#define SPIS_VEC_Z        4   /* preconditioned vector z[j]              (solvers.py:173) */
#define SPIS_VEC_X        5   /* most recent iterate x_j                 (solvers.py:287) */
#define SPIS_VEC_PRE_DIAG 6   /* Jacobi preconditioner: the diagonal d, z = d (.) q */
#define SPIS_VEC_W        7   /* work vector (download only, for tests)  */

/* preconditioner kinds (solvers.py:149-161: identity / pre.solve / pre @ vec) */
#define SPIS_PRE_NONE     0   /* identity: z[j] aliases q[j], no copy                  */
#define SPIS_PRE_JACOBI   1   /* z = d (.) q, d uploaded as SPIS_VEC_PRE_DIAG           */
#define SPIS_PRE_CSR      2   /* z = P q with sparse P in slot SPIS_SLOT_PRE            */
#define SPIS_PRE_BLOCK    3   /* z = blockdiag(B_i) q, dense b x b blocks (spis_upload_blocks) */
#define SPIS_PRE_HOST     4   /* host callback bridge: spis_host_pre_get / _put         */

/* orthogonalisation variants for spis_set_option("orth", ...) */
#define SPIS_ORTH_CGS2    0   /* classical Gram-Schmidt with reorthogonalisation (default) */
#define SPIS_ORTH_CGS1    1   /* single classical pass                                      */
#define SPIS_ORTH_MGS     2   /* modified Gram-Schmidt, the reference's loop solvers.py:193-195 */

/* SpMV storage formats for spis_set_option("spmv_format", ...) */
#define SPIS_FMT_AUTO     0   /* row patterns if the matrix has few distinct stencils, else SELL-32, CSR if padding > 25 % */
#define SPIS_FMT_SELL     1
#define SPIS_FMT_CSR      2
#define SPIS_FMT_SELL2     3   /* SELL-32 with the entries of a row packed in pairs (128-bit value loads) */
#define SPIS_FMT_SELLD     5   /* SELL-32 with 8-bit dictionary codes for the values (<= 256 distinct doubles) */
#define SPIS_FMT_PATTERN   4   /* 16-bit stencil id per row + stencil table (matrices with <= 4096 distinct rows) */

/* timer classes returned by spis_get_profile (CUDA-event time, algorithmic bytes, launches) */
#define SPIS_PROF_SPMV     0   /* system matrix A: A z_j, b - A x0, ||A x_j - b||          */
#define SPIS_PROF_MDOT     1   /* tall-skinny V^T w                 */
#define SPIS_PROF_LINCOMB  2   /* w -= V h  and  x = x0 + Z y       */
#define SPIS_PROF_SCALE    3
#define SPIS_PROF_PRECOND  4
#define SPIS_PROF_OTHER    5
#define SPIS_PROF_ORTHMID  6   /* fused  w -= V h1 ; h2 = V^T w  (middle of CGS2, one pass over V) */
#define SPIS_PROF_SPMV_AUX 7   /* constraint matrices M_c z_j (solvers.py:33)                   */
#define SPIS_PROF_CLASSES  8

typedef struct spis_ctx spis_ctx;

/* ---- library / context ------------------------------------------------------------ */
int         spis_abi_version(void);
int         spis_device_count(int* count_out);
/* n = number of unknowns owned by this context, n_halo = extra ghost entries appended to
 * every SpMV input vector (0 on a single GPU), k_max = maximum Krylov dimension.
 * stream = a cudaStream_t to launch on (e.g. torch's current stream) or NULL to create one. */
int         spis_ctx_create(int device, int64_t n, int64_t n_halo, int k_max, void* stream, spis_ctx** ctx_out);
int         spis_ctx_destroy(spis_ctx* ctx);
const char* spis_last_error(const spis_ctx* ctx);
const char* spis_last_global_error(void);            /* for failures of spis_ctx_create itself */
/* keys: "orth", "spmv_format", "profile", "x0_is_zero", "fuse_jacobi", "fuse_iterate", "orth_fused", "force_nonsymmetric",
 *       "spmv_dual" (A q and ||A x - b|| from one pass), "spmv_multi" (grouped constraint SpMVs), "spmv_variant",
 *       "pinned_scan_dma", and the tuning knobs "ctas_per_sm", "spmv_ctas_per_sm", "spmv_pipe_ctas_per_sm",
 *       "spmv_dual_ctas_per_sm", "mdot_variant", "mdot_reg_auto", "mdot_reg_ctas_per_sm", "lincomb_variant",
 *       "lincomb2_ctas_per_sm", "mdotm_ctas_per_sm", "orth_mid_max_stages", "auto_pattern", "auto_dict", "auto_sell2" */
int         spis_set_option(spis_ctx* ctx, const char* key, int64_t value);
int         spis_get_info(const spis_ctx* ctx, const char* key, int64_t* value_out);

/* Page-locked host buffers from a process-wide pool (results of spis_download_vec land at full
 * PCIe speed when the destination comes from here; pageable destinations run at ~4 GB/s).     */
int         spis_pinned_alloc(size_t bytes, void** ptr_out);
int         spis_pinned_free(void* ptr);
int         spis_pinned_trim(void);            /* release every unused pooled buffer */
/* Device buffers of destroyed contexts are cached for the next context (exact-size reuse); this
 * returns all cached device memory to the driver.                                              */
int         spis_device_trim(void);
/* host utility: *out = 1 if any of the n doubles is non-zero (multi-threaded scan; used to
 * recognise the explicit-zero constraint matrix `0*A` of lkdv/LinearSolver.py:30)             */
int         spis_host_any_nonzero(const double* data, size_t n, int* out);
/* bytes the upload entry points of this process have actually sent to devices (a matrix stored by row patterns found on
 * the host crosses as 2 bytes per row); reset != 0 also clears the counter                                          */
long long   spis_h2d_bytes(int reset);
/* Row stencils of a host CSR matrix, found by host threads (what spis_upload_csr does before it uploads anything when
 * option host_pattern is set -- off by default, see DESIGN -- : a
 * matrix assembled on a uniform mesh with constant coefficients -- lkdv/lkdv.py:109-111 exports such a one -- crosses
 * PCIe as a 16-bit id per row and a table instead of 12 bytes per entry).  Two rows share a stencil iff their
 * (column - row, value bits) lists are identical.  pid_out[nrows]; rep_out[<= 4096] = one representative row per stencil;
 * *npat_out = 0: more than 4096 stencils or a row longer than 64 entries.  Columns >= n_local move by col_shift
 * (ghost columns of a row-sharded strip).  No device is touched.                                                   */
int         spis_host_find_patterns(const int32_t* indptr, const int32_t* indices, const double* data, int64_t nrows,
                                    int64_t n_local, int64_t col_shift, int nthreads, uint16_t* pid_out, int32_t* rep_out,
                                    int* npat_out, int* maxlen_out, int64_t* changes_out);
/* same question, answered where it is cheapest: page-locked host buffers are pulled through a device
 * scratch block by the copy engine and tested there (no host CPU time; runs on the calling thread's
 * upload stream), pageable ones are scanned by host threads                                        */
int         spis_any_nonzero(spis_ctx* ctx, const double* data, size_t n, int* out);

/* ---- uploads ------------------------------------------------------------------------ */
/* CSR matrix -> device (and SELL-32 conversion on device).  Replaces holding `A`, `pre`
 * and `const.M` as scipy objects (solvers.py:131,150,33).  ncols may exceed nrows by
 * n_halo (ghost columns).                                                               */
int spis_upload_csr(spis_ctx* ctx, int slot, int64_t nrows, int64_t ncols, int64_t nnz,
                    const int32_t* indptr, const int32_t* indices, const double* data);
int spis_upload_vec(spis_ctx* ctx, int which, const double* host, int64_t n);
/* dense blocks for SPIS_PRE_BLOCK: nblk blocks of bs x bs (row-major, block-major on the
 * host); element f of block i is global index i*idx_stride_block + f*idx_stride_field
 * (contiguous blocks: bs,1; field-blocked [u;v;w] ordering of lkdv/refd.py:17: 1,nblk). */
int spis_upload_blocks(spis_ctx* ctx, int bs, int64_t nblk, int64_t idx_stride_block,
                       int64_t idx_stride_field, const double* blocks);
int spis_set_precond(spis_ctx* ctx, int kind);
/* After spis_thread_use_aux_stream(ctx, 1) the upload entry points above and spis_constraint_define,
 * when called BY THE SAME HOST THREAD, run on the context's auxiliary CUDA stream: a helper thread can
 * stage the constraint matrices (only needed at the first constrained step, solvers.py:242-247) while
 * the main thread already drives the Krylov loop.  (ctx, 0) waits for that stream and switches back.
 * The helper must finish (join) before the first spis_constraint_terms call.                      */
int spis_thread_use_aux_stream(spis_ctx* ctx, int on);

/* ---- Krylov loop -------------------------------------------------------------------- */
/* r0 = b - A x0, beta = ||r0||, q[0] = r0/beta                     (solvers.py:167-177) */
int spis_solve_begin(spis_ctx* ctx, double* beta_out);
/* one Arnoldi step j: z[j] = P q[j]; w = A z[j]; orthogonalise against q[0..j];
 * h[j+1,j] = ||w||; q[j+1] = w/h[j+1,j]                            (solvers.py:190-198).
 * hcol_out receives h[0..j+1, j] (j+2 doubles).  launch/wait split lets the host run
 * the small minimisation of step j-1 while the device does step j.                      */
int spis_arnoldi_launch(spis_ctx* ctx, int j);
int spis_arnoldi_wait(spis_ctx* ctx, int j, double* hcol_out);
int spis_arnoldi_step(spis_ctx* ctx, int j, double* hcol_out);
/* spis_arnoldi_launch in two halves.  _begin queues z_j, w = A z_j and the first 1.5 projections; _finish
 * queues the last projection, the normalisation and the copy of the Hessenberg column.  With m_it > 0 the
 * last projection ALSO forms the iterate of the previous step, x = x0 + Z[:, :m_it] y_it (solvers.py:287),
 * from the same sweep over the basis (CGS2, no preconditioner; spis_get_info "can_fuse_iterate"): follow it
 * with spis_residual_launch (= the second half of spis_iterate_residual_launch) and _wait.              */
int spis_arnoldi_begin(spis_ctx* ctx, int j);
int spis_arnoldi_finish(spis_ctx* ctx, int j, int m_it, const double* y_it);
int spis_residual_launch(spis_ctx* ctx);
/* spis_arnoldi_begin(j) whose SpMV ALSO measures ||A x - b|| of the iterate in the X buffer (formed by the
 * preceding spis_arnoldi_finish(j-1, m_it, y_it)): w = A z_j (solvers.py:191) and the residual of the previous
 * step's iterate (solvers.py:290) share one pass over the matrix.  Replaces the pair spis_residual_launch +
 * spis_arnoldi_begin(j); collect the norm with spis_iterate_residual_wait.                                   */
int spis_arnoldi_begin_residual(spis_ctx* ctx, int j);
/* x_j = x0 + Z[:, :m] y ; resnorm = ||A x_j - b||                  (solvers.py:287,290) */
int spis_iterate_residual(spis_ctx* ctx, int m, const double* y, double* resnorm_out);
/* the same in two halves: after _launch the host may fetch the next Hessenberg column and queue the
 * Arnoldi step after it; _wait blocks only until the residual norm of THIS pair has arrived           */
int spis_iterate_residual_launch(spis_ctx* ctx, int m, const double* y);
int spis_iterate_residual_wait(spis_ctx* ctx, double* resnorm_out);
/* _launch for an iterate the caller expects to be the LAST one (the loop ends on its residual, solvers.py:296-297, and
 * the solver returns it, :313): x_j is formed in `chunks` row blocks and every block is streamed to host_dst (n doubles
 * of page-locked memory) as soon as it is complete, so the device-to-host copy of the result overlaps its formation
 * and its residual check.  *started_out = 0: host_dst is not page-locked, plain launch.  spis_download_join blocks
 * until the copy is complete (host_dst must stay alive until then); an iterate that turns out not to be the last
 * one is joined and dropped.                                                                                        */
int spis_iterate_residual_launch_dl(spis_ctx* ctx, int m, const double* y, double* host_dst, int chunks, int* started_out);
int spis_download_join(spis_ctx* ctx);
/* only x = x0 + Z[:, :m] y (lazy re-materialisation of dict['x'][j], solvers.py:318)    */
int spis_form_iterate(spis_ctx* ctx, int m, const double* y);

/* ---- pipelined loop: whole Arnoldi steps queued ahead, Givens update + y on the device --------------------------
 * Replaces, for the unconstrained phase (solvers.py:230-235), the per-iteration host round trip "Hessenberg column
 * down, coefficients up": a one-warp kernel does the Givens update of solvers.py:113 and the back substitution, the
 * last Gram-Schmidt sweep of the NEXT step forms x_j = x0 + Z y_j (solvers.py:287) from device-resident y_j, and the
 * SpMV of the step after measures ||A x_j - b|| (solvers.py:290).  The device also takes the phase decision of
 * solvers.py:230 (`residual[-1] > contol*tol`) for the steps that are already queued: once a measured residual is
 * <= thr (or a pivot of the Hessenberg QR vanishes) it stops forming iterates and the host takes over.  h[j+1,j]
 * comes out of the second Gram-Schmidt reduction (|w'|^2 - |h2|^2), so q[j+1] is written normalised by the last
 * sweep: no scale pass, no reduction in it.  Records reach the host through mapped page-locked memory.
 * Needs CGS2, a device preconditioner (or none) and the peer-memory transport (or one GPU): spis_get_info
 * "device_pipeline".  At most 8 steps / 8 residual measurements may be outstanding.                          */
int spis_pipe_begin(spis_ctx* ctx, double thr, int phase0);
/* flags: 1 = the SpMV also measures ||A x - b|| of the iterate in the X buffer (*ticket_out: see spis_resid_wait);
 *        2 = form an iterate with the device's y if the phase word allows: x_{j-1} from the same sweep over the basis
 *            (no preconditioner), x_j from a sweep over Z (preconditioned)                                       */
int spis_step_enqueue(spis_ctx* ctx, int j, int flags, int64_t* ticket_out);
/* col_out: h[0..j+1, j]; y_out: argmin_y |beta e1 - H_j y| (j+1); info_out[8]: [1] y is valid, [2] the minimum,
 * [3] h[j+1,j]^2, [4] |w'|^2, [5] |h2|^2, [6] phase word seen by the step's last sweep                           */
int spis_step_wait(spis_ctx* ctx, int j, double* col_out, double* y_out, double* info_out);
/* *res2_out = ||A x - b||^2 as reduced on the device (the number compared with thr^2); *go_out = 1 while the phase
 * word still says "unconstrained"                                                                                  */
int spis_resid_wait(spis_ctx* ctx, int64_t ticket, double* res2_out, int* go_out);

/* ---- constraint stage (solvers.py:21-36, constraint_container.__init__) ------------- */
/* class-form constraint c: 1/2 x^T M x + v^T x + cc.  mat_slot < 0: M is identically zero
 * (lkdv/LinearSolver.py:30 `0*A`); v may be NULL.  On a single GPU an all-zero v is recognised and dropped;
 * a row-sharded context (halo / collectives configured) keeps whatever it is given, because "v is zero" is a
 * global property that the caller decides collectively (a rank with an all-zero SLICE still joins the sums). */
int spis_constraint_define(spis_ctx* ctx, int c, int mat_slot, const double* v, double cc);
/* term0 (scalar), term1 (m), term2 (m x m row-major) for Z = z[:m].T, incremental in m:
 * MZ = M@Z (:33), term0 (:34), term1 = v@Z + x0@MZ (:35), term2 = 1/2 Z.T@MZ (:36)      */
int spis_constraint_terms(spis_ctx* ctx, int c, int m, double* term0, double* term1, double* term2);
/* The same for nc constraints cs[0..nc) at once: term0[nc], term1[nc][m], term2[nc][m][m].  When every constraint is one
 * column behind m (a constrained iteration, solvers.py:242-247; every iteration of cgmres_p, :398-407) the quadratic ones
 * share ONE pass over Z (their M z_col as up to four right-hand sides), one cross-rank reduction and one
 * synchronisation; otherwise this is spis_constraint_terms per constraint.                                            */
int spis_constraint_terms_batch(spis_ctx* ctx, int nc, const int32_t* cs, int m, double* term0, double* term1, double* term2);
/* replaces the scalar of a defined constraint (time loops: same M and v, new invariant values per step) */
int spis_constraint_set_constant(spis_ctx* ctx, int c, double cc);
/* replaces the linear term v of a defined constraint (NULL: no linear term); the matrix stays where it is */
int spis_constraint_set_vector(spis_ctx* ctx, int c, const double* v);
/* Class-form constraint c staged by a native helper thread on the auxiliary stream while the caller runs the
 * Krylov loop: zero test of M's values (-> mat_slot < 0), spis_upload_csr into slot SPIS_SLOT_CON0 + c,
 * spis_constraint_define.  Exception to the pointer rule above: the host arrays must stay valid until
 * spis_constraint_setup_wait returns (it joins all helpers and reports the first failure); call it before the
 * first spis_constraint_terms.  The constraint data is first needed at the first constrained step
 * (solvers.py:242-247), many iterations after the solve has started.                                        */
int spis_constraint_setup_async(spis_ctx* ctx, int c, int64_t nrows, int64_t ncols, int64_t nnz,
                                const int32_t* indptr, const int32_t* indices, const double* data,
                                const double* v, double cc);
/* the same with the answer to "is M identically zero?" supplied (m_is_zero = 0 / 1; -1 = scan): row-sharded
 * sessions take that decision collectively, once, for all ranks                                              */
int spis_constraint_setup_async2(spis_ctx* ctx, int c, int64_t nrows, int64_t ncols, int64_t nnz,
                                 const int32_t* indptr, const int32_t* indices, const double* data,
                                 const double* v, double cc, int m_is_zero);
int spis_constraint_setup_wait(spis_ctx* ctx);

/* ---- the k-dimensional constrained minimisation (HOST code: solvers.py:251-255 keeps it off the device) ------
 * min_y |beta e1 - H y|^2 subject to term0[c] + term1[c].y + y'term2[c] y = 0, c < nc (the reduced quadratic
 * invariants of spis_constraint_terms); H is (m+1) x m row-major with row stride ldh.  Newton on the KKT system in
 * QR coordinates (the caller settles the signs for the reference's signed 1e-12 acceptance test).  *handled = 0: not the
 * plain converged case -- use the general (Python) solver; y_out is then untouched.                            */
int spis_small_kkt(int m, int ldh, const double* H, double beta, int nc, const double* term0, const double* term1,
                   const double* term2, double* y_out, double* fval_out, int* nit_out, int* handled);

/* The sign settling that follows (smallsolve._settle_signs; solvers.py:14-18,266 accept a constrained step only if
 * max_c g_c(y) <= 1e-12, signed): y moves by minimum-norm corrections of a few ulps until every g_c(y) <= 0 in this
 * routine's arithmetic (at most `tries` times).  The caller checks the signs once more with its own evaluation.      */
int spis_small_settle(int m, int nc, const double* term0, const double* term1, const double* term2, double* y, int tries,
                      int* moved_out);

/* ---- downloads / host bridges ------------------------------------------------------- */
int spis_download_vec(spis_ctx* ctx, int which, int j, double* host, int64_t n);
/* rows z[j0..j1) as a (j1-j0) x n row-major block (dict-form constraints need Z on the
 * host: lkdvRK/LinearSolver.py:29-67)                                                   */
int spis_download_Z(spis_ctx* ctx, int j0, int j1, double* host);
/* SPIS_PRE_HOST bridge: fetch q[j], store z[j] = pre.solve(q[j])   (solvers.py:152-154) */
int spis_host_pre_get(spis_ctx* ctx, int j, double* q_host);
int spis_host_pre_put(spis_ctx* ctx, int j, const double* z_host);

/* ---- multi-GPU hooks (row-block sharding) -------------------------------- */
/* Reductions: after every local reduction the library calls allreduce(user, device_ptr, count)
 * with a DEVICE pointer to `count` doubles that must be summed in place over all ranks, ordered
 * on the context's stream.
 * Halo: spis_halo_set_plan uploads the local indices this rank must send (concatenated per
 * destination rank).  Before every SpMV the library gathers those entries of the input vector
 * into a contiguous device buffer (halo_pack_kernel) and calls halo(user, send_dev, recv_dev):
 * the callback must deliver the peers' packed entries into recv_dev = &vec[hoff], n_halo
 * doubles ordered by source rank.  The host language supplies both callbacks
 * (torch.distributed over NCCL in distributed.py).                                        */
typedef int (*spis_allreduce_fn)(void* user, void* device_ptr, int64_t count);
typedef int (*spis_halo_fn)(void* user, void* send_device_ptr, void* recv_device_ptr);
int spis_set_collectives(spis_ctx* ctx, spis_allreduce_fn allreduce, spis_halo_fn halo, void* user);
int spis_halo_set_plan(spis_ctx* ctx, const int32_t* send_idx, int64_t n_send);
/* NVLink peer-memory communicator (one process per GPU on one node).  Each rank creates a comm
 * buffer and exports it as a CUDA IPC handle (64 bytes); the host language all-gathers the
 * handles and hands the concatenation (world x 64 bytes, rank order) to spis_xcomm_connect.
 * From then on the all-reduces are FUSED into the tail of the producing kernels (the last CTA
 * stores its partial sums into every peer, flags, waits, sums in rank order) and the halo
 * exchange is a push kernel writing straight into the neighbours' buffers: no callbacks, no NCCL
 * launches on the hot path.  spis_xcomm_set_halo gives, for each entry of the send plan, the
 * destination rank and its position in that rank's ghost ordering, plus 0/1 masks of the ranks
 * this rank sends to / receives from.                                                        */
int spis_xcomm_create(spis_ctx* ctx, int rank, int world, int64_t halo_cap, void* handle_out, int64_t handle_capacity);
int spis_xcomm_connect(spis_ctx* ctx, const void* handles);
int spis_xcomm_set_halo(spis_ctx* ctx, const int32_t* dest_rank, const int32_t* dest_off,
                        const int32_t* send_to, const int32_t* recv_from);
/* The same communicator as an object that OUTLIVES the contexts (one per process): created and connected once
 * (handles all-gathered by the host language as above), then attached to every context of a solver call with
 * spis_ctx_attach_comm -- nothing is allocated, exchanged, IPC-opened or barriered on the per-call path.  red_cap:
 * doubles per fused reduction (>= k_max + 5); halo_cap: ghost entries per vector.  spis_comm_allreduce sums `count`
 * host doubles over all ranks in place (the collective yes/no decisions of a session set-up).                   */
typedef struct spis_comm spis_comm;
int spis_comm_create(int device, int rank, int world, int64_t red_cap, int64_t halo_cap, void* handle_out,
                     int64_t handle_capacity, spis_comm** comm_out);
int spis_comm_connect(spis_comm* comm, const void* handles);
int spis_comm_destroy(spis_comm* comm);
int spis_comm_capacity(const spis_comm* comm, int64_t* red_cap_out, int64_t* halo_cap_out);
int spis_comm_allreduce(spis_comm* comm, double* vals, int count);
int spis_ctx_attach_comm(spis_ctx* ctx, spis_comm* comm);
/* out[4]: SM cycles this rank spent waiting for its peers inside fused reductions, number of fused reductions, the same
 * two numbers for halo exchanges, since the last call (cleared by the call).  Evidence for where a sharded solve's time
 * goes: streaming versus waiting on NVLink flags.                                                                     */
int spis_xcomm_stats(spis_ctx* ctx, uint64_t* out);
int spis_sync(spis_ctx* ctx);

/* ---- measurement --------------------------------------------------------------------- */
/* per class: ms_out[c] CUDA-event milliseconds, bytes_out[c] algorithmic bytes,
 * launches_out[c] kernel launches since the last spis_reset_profile.                    */
int spis_get_profile(spis_ctx* ctx, double* ms_out, double* bytes_out, int64_t* launches_out);
int spis_reset_profile(spis_ctx* ctx);
/* moved_out[c]: bytes the launches of class c move with the storage format they ran on (row-pattern ids, value codes,
 * one pass over the matrix for two products); equals bytes_out of spis_get_profile for everything but SpMV          */
int spis_get_profile_moved(spis_ctx* ctx, double* moved_out);
/* profile mode: gaps_out[c] = milliseconds the device was idle right before the launches of class c (from the end of the
 * previous profiled kernel): where a solve waits for the host, for a copy or for another stream                       */
int spis_get_profile_gaps(spis_ctx* ctx, double* gaps_out);
/* profile mode: the launches resolved by the last spis_get_profile* call, in launch order -- kernel class, start (ms
 * after the first of them) and duration; *n_out = how many there are (at most cap are written).  A diagnostic for
 * the reference's `timing=True` question (solvers.py:300-312): where on the device timeline a solve waits for the host */
int spis_get_profile_trace(spis_ctx* ctx, int32_t* cls_out, double* start_ms_out, double* dur_ms_out, int64_t cap, int64_t* n_out);
/* CUDA-event stopwatch on the context's stream: device time between the two calls.      */
int spis_timer_start(spis_ctx* ctx);
int spis_timer_stop(spis_ctx* ctx, double* ms_out);

/* ---- single-kernel entry points (parity tests and tuning; host buffers in and out) --- */
int spis_op_spmv(spis_ctx* ctx, int slot, const double* x, double* y);                 /* y = M_slot x */
int spis_op_mdot(spis_ctx* ctx, int m, const double* V /* m x n */, const double* w, double* out /* m+1: V w, w.w */);
int spis_op_lincomb(spis_ctx* ctx, int m, const double* V, const double* base, const double* coef,
                    double sign, double* out, double* sumsq_out);
int spis_op_precond(spis_ctx* ctx, const double* q, double* z);                        /* z = P q */
/* fused middle of CGS2 (solvers.py:193-195 applied twice): w_out = w - V^T coef, dots_out[i] = V_i . w_out */
int spis_op_orth_mid(spis_ctx* ctx, int m, const double* V, const double* w, const double* coef,
                     double* w_out, double* dots_out);
/* time `reps` back-to-back launches of one kernel class on resident random data and
 * return the mean milliseconds per launch (tuning / roofline sweeps).                   */
int spis_bench_kernel(spis_ctx* ctx, int prof_class, int m, int reps, double* ms_out, double* bytes_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SPIS_B200_H */
