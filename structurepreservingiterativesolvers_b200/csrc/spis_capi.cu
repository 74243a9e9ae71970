// spis_capi.cu -- context object + extern "C" entry points declared in include/spis_b200.h.
//
// Host-side orchestration of one Krylov context on one B200: device buffers, the launch
// sequences for an Arnoldi step / iterate+residual / constraint terms, CUDA-event profiling
// and the hooks through which the host language supplies halo exchange and all-reduce for
// row-sharded multi-GPU runs.  No torch types, no exceptions across the ABI, no CPU path.
#include "../../include/spis_b200.h"
#include "spis_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <atomic>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

using namespace spis;

namespace {

thread_local char g_global_err[512] = "";

struct Matrix {
  bool present = false;
  int fmt = SPIS_FMT_SELL;
  int64_t nrows = 0, ncols = 0, nnz = 0, nnz_padded = 0;
  int32_t* indptr = nullptr;   // CSR (kept only for SPIS_FMT_CSR)
  int32_t* cols = nullptr;
  double* vals = nullptr;
  int64_t* slice_off = nullptr;  // SELL-32
  uint8_t* rowperm = nullptr;    // SELL-C-sigma (sigma = 256): position inside its window of the row stored at a slot; null = unsorted
  int32_t* scols = nullptr;
  double* svals = nullptr;
  int csr_lanes = 8;
  // row-pattern storage (SPIS_FMT_PATTERN)
  uint16_t* pid = nullptr; int32_t* tab_len = nullptr; int32_t* tab_off = nullptr; double* tab_val = nullptr;
  int npat = 0, patW = 0;
  // SELLW: per-tile x windows + 16-bit window-local columns (spmv_sellw_kernel), built next to SELL / SELLD storage
  SwTile* sw_tiles = nullptr; uint16_t* sw_lcol = nullptr; int sw_cap = 0;
  // field-window twin of the row-pattern storage (spmv_fw_kernel): F fields of N nodes, node shifts |d| <= D
  FwEntry* fw_tab = nullptr; int fwF = 0, fwN = 0, fwD = 0;
  // dictionary-coded values (SPIS_FMT_SELLD): scols as in SELL, one code per entry, table of <= 256 doubles
  uint32_t* codes = nullptr; int64_t* code_off = nullptr; double* dict = nullptr; int ndict = 0;
};

struct Constraint {
  bool defined = false;
  int slot = -1;              // matrix slot, <0: M == 0
  double* v = nullptr;        // device, ld doubles, or null when v == 0
  double cc = 0.0;
  double* MZ = nullptr;       // device, kmax x ld, lazily allocated (non-symmetric M only)
  int symmetric = -1;         // -1 not tested yet, 1: M == M^T to round-off (probabilistic device test)
  int cols_done = 0;
  bool term0_done = false;
  double term0 = 0.0;
  std::vector<double> T1, T2; // host mirrors, kmax and kmax*kmax
};

struct ProfRec { int cls; double bytes; cudaEvent_t e0, e1; };
struct AsyncJob { std::thread th; int rc = 0; std::string err; };   // spis_constraint_setup_async
struct DevBlock { void* p; size_t bytes; int device; };

}  // namespace

extern "C" int spis_device_trim(void);

// NVLink peer-memory communicator that OUTLIVES the contexts (one per process and device): comm buffer, IPC mappings of
// the peers' buffers and the sequence counters of the flag protocol.  A solver call creates a context and attaches it
// (spis_ctx_attach_comm): no allocation, handle exchange, cudaIpcOpenMemHandle or barrier on the per-call path.
struct spis_comm {
  int device = 0;
  double* xbuf = nullptr; void* xpeer[kMaxRanks] = {nullptr};
  XView xv;
  unsigned long long xseq = 1, hseq = 1;
  cudaStream_t stream = nullptr;         // for spis_comm_allreduce
  double* d_tmp = nullptr; double* h_tmp = nullptr;
  bool connected = false;
  char err[256] = "";
};

struct spis_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t aux = nullptr;   // uploads issued by a helper host thread (spis_thread_use_aux_stream)
  std::mutex mu;                // guards `owned` and the constraint table against that helper thread
  int nsm = 148;
  int64_t n = 0, n_halo = 0, hoff = 0, ld = 0;
  int kmax = 0, K = 0;          // K = kmax + 4 (small-array stride)
  // options
  int orth = SPIS_ORTH_CGS2;
  int fmt_pref = SPIS_FMT_AUTO;
  int profile = 0;
  int ctas_per_sm = 4;
  int spmv_ctas_per_sm = 8;
  int spmv_variant = 1;         // 0: first-generation SpMV kernels; 1+: prefetching / software-pipelined ones (see launch_spmv_mode)
  int spmv_pipe_ctas_per_sm = 0;   // 0 = the kernel's own default
  int mdot_reg_auto = 1;        // mdot_variant 0 (auto) may pick the register-accumulator kernel
  int mdot_reg_ctas_per_sm = 0;
  int pinned_scan_dma = 0;      // spis_any_nonzero on page-locked memory: 1 = copy engine + kernel, 0 = host threads
  int sell_sigma = 0;           // SELL / SELLD: sort rows by length inside windows of 256 when that removes >= 5 % of the padding.
                                // Off by default: on swe it removes the 18.7 % padding (148.4 M -> 126.5 M entries) and the kernels that
                                // only READ rows gain (mode 2: 211 -> 203 us, pipelined SELL 308 -> 288 us), but every kernel that
                                // STORES y now scatters 8-byte stores over its 2 KB window (SELLD mode 0: 227 -> 278 us, the grouped
                                // constraint SpMV 1.98 -> 2.55 ms per solve): 22.6 -> 23.3 ms per swe solve.  Un-permuting y through
                                // shared memory (one contiguous 2 KB store per CTA; built, 342 GPU tests green) did not rescue it
                                // either: 23.7 ms -- rows sorted by length sit next to rows of the same TYPE from a 1.7x wider span,
                                // and the x gathers of a slice touch more lines; on this L1-bound kernel that outweighs 18.7 %
                                // fewer entries.
  int spmv_multi = 1;           // constraint stage: M z_j for a group of 2 / 4 Krylov columns from one pass over M
  int spmv_dual = 1;            // A q_{j+2} and ||A x_j - b|| from one pass over A (spis_arnoldi_begin_residual)
  int spmv_dual_ctas_per_sm = 0;
  int spmv_sellw = 0;           // SELL / SELLD with x windows staged in shared memory and 16-bit columns (spmv_sellw_kernel).
                                // Bit mask like spmv_fw: 1 = dual product, 2 = single products, 4 = grouped constraint products.
                                // OFF by default: measured on the 1e7 operators (tools/tune_sellw.py, profiles/tune_sellw_r2.json) the staged
                                // kernels LOSE to the L1-gather ones although they move 2 bytes less per entry -- swe (SELLD) 7.08 ms of SpMV
                                // per solve against 6.63, lkdv forced to SELL 5.78 against 5.16: an LDS crosses the same data pipe as an
                                // L1 hit, the TMA writes of the windows add to it, and two CTAs of 8 consumer warps hide the latency of
                                // the matrix stream worse than four to five CTAs of the plain kernels.
  int batch_terms = 1;          // spis_constraint_terms_batch: one pass over Z for all quadratic constraints when one column is new
  int sweep_reverse = 3;        // pipelined loop: bit 0 = the first projection, bit 1 = the last sweep walk their tiles back to front (L2 reuse)
  int hess_async = 1;           // pipelined loop: the Givens / least-squares kernel runs beside the normalising sweep
  // Look for row patterns in the caller's CSR arrays with host threads before anything is uploaded.  OFF by default:
  // on the bench box (16 hardware threads, PCIe 5) the detection alone runs at 94 GB/s (8.1 ms for the 760 MB lkdv
  // operator against 14.4 ms of PCIe), but inside an end-to-end solve it took 12.3 ms + 4 ms for the ids to arrive
  // (36.7-37.8 ms per solve against 34.6).  It cuts the bytes sent per solve from 1.28 GB to 0.28 GB: worth it where
  // PCIe is the scarcer resource (slower links, more host cores).
  int host_pattern = 0;
  int64_t host_pattern_min_nnz = 1 << 20;
  int host_threads = 0;         // 0: hardware threads, at most 16 (half of that on a helper thread)
  int spmv_fw_rows = 8;         // spmv_fw_kernel: rows per thread and tile (4: narrow tiles, three CTAs per SM; 8: wide tiles, two)
  int spmv_fw = 1;              // row patterns on field-blocked systems, x windows staged in shared memory by TMA (spmv_fw_kernel).
                                // Bit mask: 1 = the dual product of an Arnoldi step, 2 = single products, 4 = grouped constraint products.
                                // Measured on the 1e7 lkdv operator (tools/tune_fw.py): dual 97 us against 109 us for the L1-gather kernel
                                // (each x entry travels from HBM once instead of once per field block), single products 65 / 80 / 70 us
                                // against 63 / 79 / 65 us, four columns per pass slower (the shared-memory data pipe, which every
                                // gather crosses as an LDS, is as busy as the L1 pipe was) -- so only the dual product takes it by default.
  int mdot_variant = 0, lincomb_variant = 4;   // mdot 0 = auto (tools/tune.py sweep, profiles/tune_r1.md)
  int x0_is_zero = 0;
  int fuse_jacobi = 1;
  int lincomb2_ctas_per_sm = 4;
  int fuse_iterate = 1;         // form x_j inside the last projection pass of Arnoldi step j+1 (no preconditioner)
  int arnoldi_part1 = -1;       // step whose first half (SpMV, first projection, middle pass) is queued
  int mdotm_ctas_per_sm = 4;       // tools/tune_mdotm.py: 5.1-5.5 TB/s at 4, 3.5-4.8 at 2, 3.8-4.3 at 8
  int bench_mdotm_nw = 0;       // tuning: spis_bench_kernel(SPIS_PROF_MDOT) times mdotm_kernel<nw> instead
  int auto_dict = 1;            // spmv_format=auto codes the values of a SELL matrix with <= 256 distinct ones in 8 bits
  int auto_pattern = 1;         // spmv_format=auto first tries the row-pattern storage (few distinct stencils)
  int auto_sell2 = 0;           // spmv_format=auto picks the pair-packed SELL layout
  int force_nonsymmetric = 0;   // tests: take the general (stored M Z) path of the constraint stage
  int orth_fused = 1;           // CGS2: fuse (w -= V h1) with (h2 = V^T w) through TMA-staged tiles
  int orth_mid_max_stages = 8;
  int orth_mid_probe = 0;       // tuning: stream the tiles through shared memory without consuming them
  int orth_mid_force_e = 0;     // tuning: 1 or 2 doubles per thread per tile (0 = choose)
  // vectors
  double *V = nullptr, *Z = nullptr, *W = nullptr, *T = nullptr, *R0 = nullptr, *B = nullptr, *X0 = nullptr, *X = nullptr;
  double* G = nullptr;          // 4 x ld group buffer of the constraint stage (lazy)
  std::vector<int> piggy;       // linear constraints whose v.Z rides in the next one-pass reduction (spis_constraint_terms)
  double* GW = nullptr; int gw_cols = 0;   // M Z[c0..m) of the one-pass constraint reduction (gram_kernel), gw_cols x ld (lazy)
  double* d_gpartial = nullptr; unsigned int* d_gcounter = nullptr;
  int gram = 1;                 // constraint stage: 8 or more new columns of a symmetric M go through gram_kernel
  double* pre_diag = nullptr;
  double* pre_blocks = nullptr; int pre_bs = 0; int64_t pre_nblk = 0, pre_sb = 0, pre_sf = 0;
  int pre_kind = SPIS_PRE_NONE;
  bool z_ready_next = false;    // fused Jacobi already produced z[j] for the coming step
  int z_ready_index = -1;
  // small device arrays
  double* d_small = nullptr;    // [h1 (K) | h2 (K) | scal (8)]
  double* d_y = nullptr;        // K
  double* d_cout = nullptr;     // kmax * 2K
  double* d_partial = nullptr; unsigned* d_counter = nullptr; int pstride = 0; int max_grid = 0;
  double* h_small = nullptr;    // pinned mirror of d_small
  double* h_y = nullptr;        // pinned K
  double* h_cout = nullptr;     // pinned kmax*2K
  cudaEvent_t ev_arnoldi = nullptr;
  cudaStream_t dstream = nullptr; cudaEvent_t ev_dl = nullptr, ev_chunk = nullptr; bool dl_inflight = false;   // early download of a final-candidate iterate
  cudaStream_t hstream = nullptr; cudaEvent_t ev_orth = nullptr, ev_hess = nullptr;   // hess_kernel runs beside the normalising sweep
  cudaEvent_t ev_resid = nullptr; double* h_resid = nullptr; bool resid_inflight = false;   // pinned residual slot of its own
  // pipelined loop (spis_pipe_begin / spis_step_enqueue): Givens state, least-squares coefficients and the phase word on
  // the device, per-step and per-residual records in mapped page-locked memory (rings of kRecSlots)
  double *hs_cs = nullptr, *hs_sn = nullptr, *hs_gv = nullptr, *hs_R = nullptr; int* hs_tracking = nullptr;
  int* d_phase = nullptr; double* d_ydev = nullptr; double* d_rpartial = nullptr;
  double* h_rec = nullptr; double* d_rec = nullptr; int rec_stride = 0;
  unsigned long long rec_counter = 1;
  unsigned long long step_seq[8] = {0}, res_seq[8] = {0};
  long long res_tickets = 0;
  double pipe_thr2 = 0.0; bool pipe_ready = false;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  int arnoldi_inflight = -1;
  bool began = false;
  Matrix mats[SPIS_MAX_SLOTS];
  Constraint cons[SPIS_MAX_SLOTS];
  // collectives
  spis_allreduce_fn allreduce = nullptr; spis_halo_fn halo = nullptr; void* cuser = nullptr;
  int32_t* d_send_idx = nullptr; double* d_send = nullptr; int64_t n_send = 0;
  bool defer_allreduce = false;   // constraint stage: one all-reduce for a whole batch of dot blocks
  // NVLink peer-memory collectives (spis_xcomm_*)
  bool xactive = false;
  double* xbuf = nullptr; void* xpeer[kMaxRanks] = {nullptr};
  XView xv;
  unsigned long long xseq_own = 1, hseq_own = 1;
  unsigned long long* xseq_p = &xseq_own; unsigned long long* hseq_p = &hseq_own;   // the attached communicator's counters, if any
  spis_comm* comm = nullptr;              // attached (not owned) communicator
  int32_t *d_dest_rank = nullptr, *d_dest_off = nullptr, *d_send_to = nullptr, *d_recv_from = nullptr;
  // profiling
  std::vector<ProfRec> recs; std::vector<cudaEvent_t> evpool;
  struct TraceRec { int cls; double start_ms, dur_ms; };
  std::vector<TraceRec> trace;   // profile mode: per-launch timeline of the batch resolved last (spis_get_profile_trace)
  std::vector<DevBlock> owned;   // device blocks currently held by this context
  std::vector<AsyncJob*> jobs;   // native helper threads staging constraint data (joined by spis_constraint_setup_wait)
  double prof_ms[SPIS_PROF_CLASSES] = {0}; double prof_bytes[SPIS_PROF_CLASSES] = {0}; int64_t prof_launch[SPIS_PROF_CLASSES] = {0};
  double prof_gap_ms[SPIS_PROF_CLASSES] = {0};  // profile mode: device idle time BEFORE the launches of each class (end of the previous kernel -> start)
  double prof_moved[SPIS_PROF_CLASSES] = {0};   // bytes the launches move with the storage format they actually run on (== prof_bytes except SpMV)
  char err[512] = "";
};

namespace {

int fail(spis_ctx* c, int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  if (c) vsnprintf(c->err, sizeof(c->err), fmt, ap); else vsnprintf(g_global_err, sizeof(g_global_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      return fail(ctx, e_ == cudaErrorMemoryAllocation ? SPIS_E_NOMEM : SPIS_E_CUDA,           \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define TRY(call) do { int r_ = (call); if (r_ != SPIS_OK) return r_; } while (0)
#define REQUIRE(cond, ...) do { if (!(cond)) return fail(ctx, SPIS_E_INVALID, __VA_ARGS__); } while (0)

inline int64_t roundup(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// SPIS_TRACE=1: wall-clock of the phases of the heavier entry points, on stderr
struct PhaseTrace {
  bool on; std::chrono::steady_clock::time_point t;
  PhaseTrace() : on(getenv("SPIS_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
  void mark(const char* what) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[spis trace]   native: %-30s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// Device memory comes from a process-wide cache of exact-size blocks.  The reference's call pattern
// is one solver call per time step (lkdv/Evolve.py:39-56), i.e. one context per solve with the SAME
// buffer sizes every time; cudaMalloc/cudaFree of ~10 GB per solve costs 100+ ms (measured), and the
// stream-ordered pool (cudaMallocAsync) stalled for up to a second when it had to grow or could not
// coalesce.  Freed blocks are kept (spis_device_trim releases them) and handed out again on an
// exact size match, so from the second solve on allocation is free.
std::atomic<long long> g_h2d_bytes(0);     // bytes the upload entry points actually sent to a device (spis_h2d_bytes)
std::mutex g_dev_mu;
std::vector<DevBlock> g_dev_free;
std::atomic<long long> g_dev_hits(0), g_dev_misses(0), g_dev_miss_bytes(0);

// Upload entry points (spis_upload_csr / _vec / _blocks, spis_constraint_define) called by a host
// thread that has switched this on run on the context's AUXILIARY stream, so that a helper thread can
// stage the constraint data while the main thread drives the Krylov loop on the main stream.
thread_local bool tl_use_aux = false;
inline cudaStream_t up_stream(spis_ctx* ctx) { return (tl_use_aux && ctx->aux) ? ctx->aux : ctx->stream; }

template <class Tp> int dalloc(spis_ctx* ctx, Tp** p, size_t count, bool zero = true) {
  *p = nullptr;
  if (count == 0) count = 1;
  const size_t bytes = (count * sizeof(Tp) + 255) / 256 * 256;
  void* q = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    for (size_t i = 0; i < g_dev_free.size(); ++i)
      if (g_dev_free[i].bytes == bytes && g_dev_free[i].device == ctx->device) {
        q = g_dev_free[i].p;
        g_dev_free[i] = g_dev_free.back();
        g_dev_free.pop_back();
        break;
      }
  }
  if (q) g_dev_hits++;
  if (!q) {
    g_dev_misses++; g_dev_miss_bytes += (long long)bytes;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e == cudaErrorMemoryAllocation) {          // give cached blocks back to the driver and retry
      cudaGetLastError();
      spis_device_trim();
      e = cudaMalloc(&q, bytes);
    }
    if (e != cudaSuccess) return fail(ctx, e == cudaErrorMemoryAllocation ? SPIS_E_NOMEM : SPIS_E_CUDA,
                                      "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  *p = static_cast<Tp*>(q);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->owned.push_back({q, bytes, ctx->device});
  }
  if (zero) CU(cudaMemsetAsync(q, 0, bytes, up_stream(ctx)));
  return SPIS_OK;
}

// Return a block to the cache.  The caller guarantees that the stream it works on has drained every
// kernel that touched the block (all call sites follow a stream synchronisation, or synchronise here).
template <class Tp> void dfree(spis_ctx* ctx, Tp*& p) {
  if (!p) return;
  cudaStreamSynchronize(up_stream(ctx));
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (size_t i = 0; i < ctx->owned.size(); ++i)
      if (ctx->owned[i].p == (void*)p) {
        std::lock_guard<std::mutex> lk2(g_dev_mu);
        g_dev_free.push_back(ctx->owned[i]);
        ctx->owned[i] = ctx->owned.back();
        ctx->owned.pop_back();
        break;
      }
  }
  p = nullptr;
}

int h2d(spis_ctx* ctx, void* dst, const void* src, size_t bytes) {
  g_h2d_bytes += (long long)bytes;
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, up_stream(ctx)));
  CU(cudaStreamSynchronize(up_stream(ctx)));
  return SPIS_OK;
}
int d2h(spis_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SPIS_OK;
}

// ---- profiling ------------------------------------------------------------------------
int prof_begin(spis_ctx* ctx, int cls, double bytes, double moved = -1.0) {
  ctx->prof_launch[cls] += 1;
  ctx->prof_bytes[cls] += bytes;
  ctx->prof_moved[cls] += moved >= 0.0 ? moved : bytes;
  if (!ctx->profile) return SPIS_OK;
  ProfRec r; r.cls = cls; r.bytes = bytes;
  for (cudaEvent_t* e : {&r.e0, &r.e1}) {
    if (!ctx->evpool.empty()) { *e = ctx->evpool.back(); ctx->evpool.pop_back(); }
    else CU(cudaEventCreate(e));
  }
  CU(cudaEventRecord(r.e0, ctx->stream));
  ctx->recs.push_back(r);
  return SPIS_OK;
}
int prof_end(spis_ctx* ctx) {
  if (!ctx->profile) return SPIS_OK;
  CU(cudaEventRecord(ctx->recs.back().e1, ctx->stream));
  return SPIS_OK;
}
int prof_resolve(spis_ctx* ctx) {
  if (ctx->recs.empty()) return SPIS_OK;
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->trace.clear();
  for (size_t i = 0; i < ctx->recs.size(); ++i) {
    auto& r = ctx->recs[i];
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r.e0, r.e1));
    ctx->prof_ms[r.cls] += ms;
    if (ctx->recs.size() <= 100000) {
      float t0 = 0.f;
      if (i > 0 && cudaEventElapsedTime(&t0, ctx->recs[0].e0, r.e0) != cudaSuccess) { cudaGetLastError(); t0 = 0.f; }
      ctx->trace.push_back({r.cls, (double)t0, (double)ms});
    }
    if (i > 0) {
      float gap = 0.f;
      if (cudaEventElapsedTime(&gap, ctx->recs[i - 1].e1, r.e0) == cudaSuccess && gap > 0.f) ctx->prof_gap_ms[r.cls] += gap;
      else cudaGetLastError();
    }
  }
  for (auto& r : ctx->recs) { ctx->evpool.push_back(r.e0); ctx->evpool.push_back(r.e1); }
  ctx->recs.clear();
  return SPIS_OK;
}

// An early download (spis_iterate_residual_launch_dl) may still be reading X on the copy stream: whatever writes X next
// on the main stream waits for it (device-side; the host does not block).
int dl_fence(spis_ctx* ctx) {
  if (ctx->dl_inflight) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_dl, 0));
  return SPIS_OK;
}

// ---- kernel launchers -----------------------------------------------------------------
int grid_for(spis_ctx* ctx, int64_t work_items, int per_sm) {
  int64_t g = (int64_t)ctx->nsm * per_sm;
  if (g > ctx->max_grid) g = ctx->max_grid;
  if (work_items < g) g = work_items;
  return (int)(g < 1 ? 1 : g);
}

// view + sequence number handed to a reducing kernel: cross-GPU part fused into its tail unless the
// reduction is being batched (constraint stage)
inline XView fused_view(spis_ctx* ctx) { return (ctx->xactive && !ctx->defer_allreduce) ? ctx->xv : XView(); }
inline unsigned long long fused_seq(spis_ctx* ctx) {
  if (ctx->xactive && !ctx->defer_allreduce) return (*ctx->xseq_p)++;
  return 0;
}

int do_allreduce(spis_ctx* ctx, double* dev, int64_t count, bool force = false) {
  if (ctx->xactive) {
    if (!force) return SPIS_OK;                    // already reduced inside the producing kernel
    const unsigned long long chunks = (unsigned long long)((count + ctx->xv.red_cap - 1) / ctx->xv.red_cap);
    xreduce_kernel<<<1, kThreads, 0, ctx->stream>>>(dev, count, ctx->xv, *ctx->xseq_p);
    CU(cudaGetLastError());
    *ctx->xseq_p += chunks;
    ctx->prof_launch[SPIS_PROF_OTHER] += 1;
    return SPIS_OK;
  }
  if (!ctx->allreduce || (ctx->defer_allreduce && !force)) return SPIS_OK;
  int r = ctx->allreduce(ctx->cuser, dev, count);
  if (r != 0) return fail(ctx, SPIS_E_INVALID, "allreduce callback failed (%d)", r);
  return SPIS_OK;
}
int do_halo2(spis_ctx* ctx, double* a, double* b);
int do_halo(spis_ctx* ctx, double* vec) {
  if (ctx->xactive) return do_halo2(ctx, vec, nullptr);
  if (!ctx->halo || (ctx->n_halo == 0 && ctx->n_send == 0)) return SPIS_OK;
  if (ctx->n_send > 0) {
    const int grid = (int)((ctx->n_send + 255) / 256 < (int64_t)ctx->nsm * 8 ? (ctx->n_send + 255) / 256 : (int64_t)ctx->nsm * 8);
    halo_pack_kernel<<<grid, 256, 0, ctx->stream>>>(vec, ctx->d_send_idx, ctx->n_send, ctx->d_send);
    CU(cudaGetLastError());
    ctx->prof_launch[SPIS_PROF_OTHER] += 1;
  }
  int r = ctx->halo(ctx->cuser, ctx->d_send, vec + ctx->hoff);
  if (r != 0) return fail(ctx, SPIS_E_INVALID, "halo callback failed (%d)", r);
  return SPIS_OK;
}

// ghosts of up to two vectors from ONE exchange (peer-memory transport only; b may be null)
int do_halo2(spis_ctx* ctx, double* a, double* b) {
  if (!ctx->xactive) {
    TRY(do_halo(ctx, a));
    return b ? do_halo(ctx, b) : SPIS_OK;
  }
  if (ctx->n_halo == 0 && ctx->n_send == 0) return SPIS_OK;
  REQUIRE(ctx->xv.halo_cap / 4 >= ctx->n_halo, "comm buffer holds %lld ghost entries per vector, %lld needed", (long long)(ctx->xv.halo_cap / 4), (long long)ctx->n_halo);
  const unsigned long long seq = (*ctx->hseq_p)++;
  halo_xchg_kernel<<<1, 1024, 0, ctx->stream>>>(a, b, ctx->hoff, ctx->n_halo, ctx->d_send_idx, ctx->d_dest_rank, ctx->d_dest_off,
                                                 ctx->n_send, ctx->xv, seq);
  CU(cudaGetLastError());
  ctx->prof_launch[SPIS_PROF_OTHER] += 1;
  return SPIS_OK;
}

// out[0..m) = V_i.w, then extra.w (if extra), then w.w (if with_sumsq); all-reduced over ranks
int launch_mdot(spis_ctx* ctx, const double* V, int m, const double* extra, int with_sumsq,
                const double* w, double* out, const TailExtra* ride = nullptr) {
  const int nrows = m + (extra ? 1 : 0) + (with_sumsq ? 1 : 0);
  if (nrows == 0) return SPIS_OK;
  const TailExtra tx = ride ? *ride : TailExtra{};
  const int nred = nrows + (tx.rpart ? 1 : 0);
  REQUIRE(nrows <= ctx->pstride, "mdot: %d rows exceed workspace %d", nrows, ctx->pstride);
  const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
  // measured on B200 (n = 1e7): few rows want more loads per thread on fewer CTAs, many rows the opposite
  // (tools/tune_mdot_reg.py, n = 1e7: register sums win by 3-8 % for 5..32 rows and by 20-35 % for 1-2 rows, lose 2-5 % at 3-4 and beyond 34)
  if ((ctx->mdot_variant == 1 || (ctx->mdot_variant == 0 && ctx->mdot_reg_auto && (nrows <= 2 || (nrows >= 5 && nrows <= 32)))) && nrows <= 40) {
    // partial sums in registers, one reduction at the end (mdot_reg_kernel)
    const int MB = (nrows + 7) / 8 * 8;
    const int per = ctx->mdot_reg_ctas_per_sm > 0 ? ctx->mdot_reg_ctas_per_sm : 2;
    const int grid = grid_for(ctx, ntiles, per);
    TRY(prof_begin(ctx, SPIS_PROF_MDOT, (double)(m + (extra ? 1 : 0) + 1) * 8.0 * (double)ctx->n));
    const XView xv = fused_view(ctx);
    const unsigned long long seq = fused_seq(ctx);
#define SPIS_MDR(MBB) case MBB: mdot_reg_kernel<MBB><<<grid, kThreads, 0, ctx->stream>>>(V, ctx->ld, m, extra, with_sumsq, w, ctx->n, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq, tx); break;
    switch (MB) { SPIS_MDR(8) SPIS_MDR(16) SPIS_MDR(24) SPIS_MDR(32) default: SPIS_MDR(40) }
#undef SPIS_MDR
    CU(cudaGetLastError());
    TRY(prof_end(ctx));
    return do_allreduce(ctx, out, nred);
  }
  int variant = ctx->mdot_variant, per_sm = ctx->ctas_per_sm;
  if (variant == 0) { variant = nrows <= 6 ? 4 : 2; per_sm = nrows <= 6 ? (ctx->ctas_per_sm > 2 ? 2 : ctx->ctas_per_sm) : ctx->ctas_per_sm; }
  const int grid = grid_for(ctx, ntiles, per_sm);
  const size_t smem = (size_t)(kWarps * nrows + kWarps * 32) * sizeof(double);
  TRY(prof_begin(ctx, SPIS_PROF_MDOT, (double)(m + (extra ? 1 : 0) + 1) * 8.0 * (double)ctx->n));
  const XView xv = fused_view(ctx);
  const unsigned long long seq = fused_seq(ctx);
  if (variant == 8)
    mdot_kernel<8><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, extra, with_sumsq, w, ctx->n, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq, tx);
  else if (variant == 2)
    mdot_kernel<2><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, extra, with_sumsq, w, ctx->n, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq, tx);
  else
    mdot_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, extra, with_sumsq, w, ctx->n, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq, tx);
  CU(cudaGetLastError());
  TRY(prof_end(ctx));
  return do_allreduce(ctx, out, nred);
}

// out[c*nrows + i] = row_i . W_c  for nw (2 or 4) vectors W_c = W + c*wstride; rows = V[0..m) (+ extra)
int launch_mdotm(spis_ctx* ctx, int nw, const double* V, int m, const double* extra, const double* W,
                 int64_t wstride, double* out) {
  const int nrows = m + (extra ? 1 : 0);
  if (nrows == 0) return SPIS_OK;
  const int nout = nw * nrows;
  REQUIRE(nout <= ctx->pstride, "mdotm: %d outputs exceed workspace %d", nout, ctx->pstride);
  const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
  const int grid = grid_for(ctx, ntiles, ctx->mdotm_ctas_per_sm);
  const size_t smem = (size_t)(kWarps * nout + kWarps * 32) * sizeof(double);
  REQUIRE(smem <= 227 * 1024, "mdotm: %zu bytes of shared memory needed", smem);
  TRY(prof_begin(ctx, SPIS_PROF_MDOT, (double)(nrows + nw) * 8.0 * (double)ctx->n));
  const XView xv = fused_view(ctx);
  const unsigned long long seq = fused_seq(ctx);
  if (nw == 4)
    mdotm_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, extra, W, wstride, ctx->n, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq);
  else
    mdotm_kernel<2><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, extra, W, wstride, ctx->n, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq);
  CU(cudaGetLastError());
  TRY(prof_end(ctx));
  return do_allreduce(ctx, out, nout);
}

// out[j * ra + i] = row_i . B_j for rows = A[0..ma) (+ extra) and ALL mb columns B_j = B + j*ldb, one pass over A per
// 24 columns (gram_kernel); tri: only entries with i <= c0 + j (and the extra row) are needed.  No cross-rank
// reduction here: the caller batches it (spis_constraint_terms).
bool gram_applicable(spis_ctx* ctx, int ma, bool extra) { return ctx->gram && (ma + (extra ? 1 : 0) + 7) / 8 <= kGramMaxIB; }
int launch_gram(spis_ctx* ctx, const double* A, int ma, const double* extra, const double* B, int mb, int c0, int tri,
                double* out, int ra, const double* const* xcols = nullptr, int nxcols = 0) {
  const int nib_all = (ma + (extra ? 1 : 0) + 7) / 8;
  REQUIRE(nib_all >= 1 && nib_all <= kGramMaxIB && ra >= nib_all * 8 - 7, "gram: %d rows not supported", ma);
  REQUIRE(nxcols >= 0 && nxcols <= 4, "gram: at most 4 extra columns");
  const size_t pmax = (size_t)2 * ctx->nsm * kGramMaxIB * kGramJB * 64;
  if (!ctx->d_gpartial) { TRY(dalloc(ctx, &ctx->d_gpartial, pmax, false)); TRY(dalloc(ctx, &ctx->d_gcounter, 64)); }
  const int ncols = mb + nxcols;               // columns j < mb: B + j*ld; then the extra columns (every row of A is needed for those)
  for (int j0 = 0; j0 < ncols; j0 += 8 * kGramJB) {
    const int ncp = std::min(ncols - j0, 8 * kGramJB);
    const int mbp = std::max(0, std::min(mb - j0, ncp));
    const int nxp = ncp - mbp;
    // rows this launch needs: i <= c0 + j0 + mbp - 1 (the extra row and the extra columns keep every tile row alive)
    int nib = nib_all;
    int rows_a = ma;
    if (tri && !extra && nxp == 0) { rows_a = std::min(ma, c0 + j0 + mbp); nib = (rows_a + 7) / 8; }
    GramArgs g{A, ctx->ld, rows_a, extra, B + (size_t)j0 * ctx->ld, ctx->ld, mbp, {nullptr, nullptr, nullptr, nullptr}, nxp,
               c0 + j0, tri, ctx->n, ctx->d_gpartial, ctx->d_gcounter, out + (size_t)j0 * ra, ra};
    for (int x = 0; x < nxp; ++x) g.bx[x] = xcols[j0 + mbp + x - mb];
    const int grid = ctx->nsm * (nib <= 3 ? 2 : 1);
    const size_t smem = (size_t)nib * kGramJB * 64 * sizeof(double);
    TRY(prof_begin(ctx, SPIS_PROF_MDOT, (double)(rows_a + (extra ? 1 : 0) + ncp) * 8.0 * (double)ctx->n));
    switch (nib) {
      case 1: gram_kernel<1><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
      case 2: gram_kernel<2><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
      case 3: gram_kernel<3><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
      case 4: gram_kernel<4><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
      case 5: gram_kernel<5><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
      case 6: gram_kernel<6><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
      default: gram_kernel<7><<<grid, kGramThreads, smem, ctx->stream>>>(g); break;
    }
    CU(cudaGetLastError());
    TRY(prof_end(ctx));
  }
  return SPIS_OK;
}

int launch_lincomb(spis_ctx* ctx, const double* V, int m, const double* coef, const double* coef2,
                   double sign, const double* base, double* out, int with_sumsq, double* sumsq_out) {
  const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
  const int grid = grid_for(ctx, ntiles, ctx->ctas_per_sm);
  const size_t smem = (size_t)(m + 2 + kWarps * 32) * sizeof(double);
  TRY(prof_begin(ctx, SPIS_PROF_LINCOMB, (double)(m + (base ? 1 : 0) + 1) * 8.0 * (double)ctx->n));
  const XView xv = with_sumsq ? fused_view(ctx) : XView();
  const unsigned long long seq = with_sumsq ? fused_seq(ctx) : 0;
  if (ctx->lincomb_variant == 8)
    lincomb_kernel<8><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, coef, coef2, sign, base, out, ctx->n, with_sumsq, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq);
  else if (ctx->lincomb_variant == 2)
    lincomb_kernel<2><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, coef, coef2, sign, base, out, ctx->n, with_sumsq, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq);
  else
    lincomb_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, coef, coef2, sign, base, out, ctx->n, with_sumsq, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq);
  CU(cudaGetLastError());
  TRY(prof_end(ctx));
  if (with_sumsq) return do_allreduce(ctx, sumsq_out, 1);
  return SPIS_OK;
}

// outA = baseA - V[0..m) coefA (+ ||outA||^2 -> sumsq_out), outB = baseB + V[0..mB) coefB, one sweep over V
int launch_lincomb2(spis_ctx* ctx, const double* V, int m, const double* coefA, const double* coefB, int mB,
                    const double* baseA, const double* baseB, double* outA, double* outB, double* sumsq_out) {
  const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
  const int grid = grid_for(ctx, ntiles, ctx->lincomb2_ctas_per_sm);
  const size_t smem = (size_t)(2 * (m + 2) + kWarps * 32) * sizeof(double);
  TRY(prof_begin(ctx, SPIS_PROF_LINCOMB, (double)(m + 3 + (baseB ? 1 : 0)) * 8.0 * (double)ctx->n));
  const XView xv = fused_view(ctx);
  const unsigned long long seq = fused_seq(ctx);
  lincomb2_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(V, ctx->ld, m, coefA, coefB, mB, baseA, baseB, outA, outB, ctx->n, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq);
  CU(cudaGetLastError());
  TRY(prof_end(ctx));
  return do_allreduce(ctx, sumsq_out, 1);
}

// Fused middle of CGS2 (orth_mid_kernel): w <- w - V coef, out[i] = V_i . w(new).  Picks the widest
// tile that still leaves a ring of staging slots in shared memory; returns SPIS_E_UNSUPPORTED (and
// launches nothing) when m is too large for one slot pair, so the caller takes the two-kernel path.
constexpr size_t kOrthMidSmemBudget = 225 * 1024;
template <int MB>
int launch_orth_mid_mb(spis_ctx* ctx, int E, int grid, size_t smem, const double* V, int m, const double* coef,
                       double* w, int nstages, double* out, const XView& xv, unsigned long long seq, int with_norm) {
  if (E == 2)
    orth_mid_kernel<MB, 2><<<grid, kOrthMidThreads, smem, ctx->stream>>>(V, ctx->ld, m, coef, w, ctx->hoff, nstages, ctx->orth_mid_probe, with_norm, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq);
  else
    orth_mid_kernel<MB, 1><<<grid, kOrthMidThreads, smem, ctx->stream>>>(V, ctx->ld, m, coef, w, ctx->hoff, nstages, ctx->orth_mid_probe, with_norm, ctx->d_partial, ctx->pstride, ctx->d_counter, out, xv, seq);
  CU(cudaGetLastError());
  return SPIS_OK;
}

bool orth_mid_plan(const spis_ctx* ctx, int m, int* E_out, int* stages_out, int* MB_out, size_t* smem_out) {
  if (m < 1 || m > 64) return false;
  const int MB = (m + 7) / 8 * 8;
  for (int E = 2; E >= 1; --E) {
    if (ctx->orth_mid_force_e && E != ctx->orth_mid_force_e) continue;
    const size_t fixed = orth_mid_smem(MB, E, m, 0);
    const size_t stage = (size_t)(m + 1) * kThreads * E * sizeof(double) + 2 * sizeof(uint64_t);
    if (fixed + stage > kOrthMidSmemBudget) continue;
    int stages = (int)((kOrthMidSmemBudget - fixed) / stage);
    if (stages > ctx->orth_mid_max_stages) stages = ctx->orth_mid_max_stages;
    if (stages < 2) continue;
    *E_out = E; *stages_out = stages; *MB_out = MB; *smem_out = orth_mid_smem(MB, E, m, stages);
    return true;
  }
  return false;
}

int launch_orth_mid(spis_ctx* ctx, const double* V, int m, const double* coef, double* w, double* out, int with_norm = 0) {
  int E = 0, stages = 0, MB = 0; size_t smem = 0;
  if (!orth_mid_plan(ctx, m, &E, &stages, &MB, &smem)) return fail(ctx, SPIS_E_UNSUPPORTED, "orth_mid: m=%d does not fit shared memory", m);
  const int T = kThreads * E;
  const int64_t ntiles = (ctx->hoff + T - 1) / T;
  const int grid = grid_for(ctx, ntiles, 1);
  TRY(prof_begin(ctx, SPIS_PROF_ORTHMID, (double)(m + 2) * 8.0 * (double)ctx->n));
  const XView xv = fused_view(ctx);
  const unsigned long long seq = fused_seq(ctx);
  int rc;
  switch (MB) {
    case 8: rc = launch_orth_mid_mb<8>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    case 16: rc = launch_orth_mid_mb<16>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    case 24: rc = launch_orth_mid_mb<24>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    case 32: rc = launch_orth_mid_mb<32>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    case 40: rc = launch_orth_mid_mb<40>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    case 48: rc = launch_orth_mid_mb<48>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    case 56: rc = launch_orth_mid_mb<56>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
    default: rc = launch_orth_mid_mb<64>(ctx, E, grid, smem, V, m, coef, w, stages, out, xv, seq, with_norm); break;
  }
  TRY(rc);
  TRY(prof_end(ctx));
  return do_allreduce(ctx, out, m + (with_norm ? 1 : 0));
}

// bytes one pass over the STORED matrix moves (format-specific; the CSR model of SURVEY 8d is 12 nnz + 4 (n + 1))
double matrix_bytes(const Matrix& M) {
  const double nsl = (double)((M.nrows + 31) / 32);
  switch (M.fmt) {
    case SPIS_FMT_PATTERN: return 2.0 * (double)M.nrows;                                   // 16-bit stencil id per row (the table stays in L1)
    case SPIS_FMT_SELLD: return (M.sw_lcol ? 3.0 : 5.0) * (double)M.nnz_padded + 16.0 * nsl;    // 32- (16-) bit column + 8-bit value code
    case SPIS_FMT_SELL: return (M.sw_lcol ? 10.0 : 12.0) * (double)M.nnz_padded + 8.0 * nsl;
    case SPIS_FMT_SELL2: return 12.0 * (double)M.nnz_padded + 8.0 * nsl;
    default: return 12.0 * (double)M.nnz + 4.0 * (double)(M.nrows + 1);
  }
}

// Field-window SpMV (spmv_fw_kernel).  Tile width T and window stride WS from the shared-memory budget; *grid_out CTAs
// each leave one partial sum (KIND != 0).  Returns false when the matrix has no field-window table or nothing fits.
constexpr size_t kFwSmemBudget = 110 * 1024;       // two CTAs per SM
bool fw_plan(const spis_ctx* ctx, const Matrix& M, int NV, int* T_out, int* WS_out, size_t* smem_out, int use = 7) {
  if (!M.fw_tab || !(ctx->spmv_fw & use) || M.fmt != SPIS_FMT_PATTERN || M.patW > 16 || M.fwF * M.npat > kFwMaxTable / 4) return false;
  const int rows = (ctx->spmv_fw_rows == 4 && M.fwF <= 4) ? 4 : 8;
  for (int T = 8 * kFwThreads; T >= kFwThreads; T /= 2) {
    if (M.fwF * (T / kFwThreads) > rows) continue;
    const int WS = (T + 2 * M.fwD + 2 + 7) / 8 * 8;
    const size_t smem = fw_smem_bytes(NV, M.npat, M.patW, M.fwF, WS);
    if (smem > kFwSmemBudget) continue;
    *T_out = T; *WS_out = WS; *smem_out = smem;
    return true;
  }
  return false;
}

template <int NV, int KIND>
int launch_fw(spis_ctx* ctx, const Matrix& M, const FwVecs& vv, const double* b, double* partial, int* grid_out) {
  int T = 0, WS = 0; size_t smem = 0;
  if (!fw_plan(ctx, M, NV, &T, &WS, &smem)) return fail(ctx, SPIS_E_UNSUPPORTED, "field-window SpMV does not apply");   // (callers checked their bit)
  FwArgs P;
  P.pid = M.pid; P.tab_len = M.tab_len; P.tab = M.fw_tab;
  P.npat = M.npat; P.W = M.patW; P.F = M.fwF; P.N = M.fwN; P.D = M.fwD; P.T = T; P.WS = WS; P.ld = ctx->ld;
  const int ntiles = (M.fwN + T - 1) / T;
  const int per_sm = (M.fwF * (T / kFwThreads) <= 4 && ctx->spmv_fw_rows == 4) ? 3 : 2;
  const int grid = ntiles < per_sm * ctx->nsm ? ntiles : per_sm * ctx->nsm;
  static bool attr_done = false;     // per template instance
  if (!attr_done) {
#define SPIS_FW_ATTR(NCHV, RV) CU(cudaFuncSetAttribute(spmv_fw_kernel<NV, KIND, NCHV, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwSmemBudget));
    SPIS_FW_ATTR(1, 4) SPIS_FW_ATTR(2, 4) SPIS_FW_ATTR(3, 4) SPIS_FW_ATTR(4, 4) SPIS_FW_ATTR(1, 8) SPIS_FW_ATTR(2, 8) SPIS_FW_ATTR(3, 8) SPIS_FW_ATTR(4, 8)
#undef SPIS_FW_ATTR
    attr_done = true;
  }
  const bool r4 = M.fwF * (T / kFwThreads) <= 4 && ctx->spmv_fw_rows == 4;
#define SPIS_FW_LAUNCH(NCHV) if (r4) spmv_fw_kernel<NV, KIND, NCHV, 4><<<grid, kFwThreads, smem, ctx->stream>>>(P, vv, b, partial); \
                             else spmv_fw_kernel<NV, KIND, NCHV, 8><<<grid, kFwThreads, smem, ctx->stream>>>(P, vv, b, partial);
  switch (M.patW / 4) {
    case 1: SPIS_FW_LAUNCH(1) break;
    case 2: SPIS_FW_LAUNCH(2) break;
    case 3: SPIS_FW_LAUNCH(3) break;
    default: SPIS_FW_LAUNCH(4) break;
  }
#undef SPIS_FW_LAUNCH
  CU(cudaGetLastError());
  if (grid_out) *grid_out = grid;
  return SPIS_OK;
}

// SELLW SpMV (spmv_sellw_kernel): SELL / SELLD storage with staged x windows and 16-bit columns
constexpr size_t kSwSmemBudget = 110 * 1024;       // two CTAs per SM
inline size_t sw_smem_bytes(int NV, int cap) { return 2112 + (size_t)2 * NV * cap * sizeof(double); }
bool sw_plan(const spis_ctx* ctx, const Matrix& M, int NV, int use) {
  return M.sw_lcol && (ctx->spmv_sellw & use) && (M.fmt == SPIS_FMT_SELL || M.fmt == SPIS_FMT_SELLD) && !M.rowperm &&
         sw_smem_bytes(NV, M.sw_cap) <= kSwSmemBudget;
}

template <int NV, int KIND>
int launch_sellw(spis_ctx* ctx, const Matrix& M, const FwVecs& vv, const double* b, double* partial, int* grid_out) {
  const int64_t nslices = (M.nrows + 31) / 32;
  const int64_t ntiles = (nslices + kSwSlices - 1) / kSwSlices;
  const int grid = (int)(ntiles < 2 * (int64_t)ctx->nsm ? ntiles : 2 * (int64_t)ctx->nsm);
  const size_t smem = sw_smem_bytes(NV, M.sw_cap);
  static bool attr_done = false;     // per template instance
  if (!attr_done) {
    CU(cudaFuncSetAttribute(spmv_sellw_kernel<NV, KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmemBudget));
    CU(cudaFuncSetAttribute(spmv_sellw_kernel<NV, KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSwSmemBudget));
    attr_done = true;
  }
  if (M.fmt == SPIS_FMT_SELLD)
    spmv_sellw_kernel<NV, KIND, true><<<grid, kSwThreads, smem, ctx->stream>>>(M.slice_off, M.code_off, M.sw_lcol, nullptr, M.codes, M.dict, M.sw_tiles, M.nrows, M.sw_cap, vv, b, partial);
  else
    spmv_sellw_kernel<NV, KIND, false><<<grid, kSwThreads, smem, ctx->stream>>>(M.slice_off, nullptr, M.sw_lcol, M.svals, nullptr, nullptr, M.sw_tiles, M.nrows, M.sw_cap, vv, b, partial);
  CU(cudaGetLastError());
  if (grid_out) *grid_out = grid;
  return SPIS_OK;
}

template <int MODE>
int launch_spmv_mode(spis_ctx* ctx, const Matrix& M, const double* x, const double* b, double* y, double* sumsq_out) {
  const XView xv = MODE != 0 ? fused_view(ctx) : XView();
  const unsigned long long seq = MODE != 0 ? fused_seq(ctx) : 0;
  if (sw_plan(ctx, M, 1, 2)) {
    FwVecs vv{}; vv.x[0] = x; vv.y[0] = y;
    int grid = 1;
    TRY((launch_sellw<1, MODE>(ctx, M, vv, b, ctx->d_partial, &grid)));
    if (MODE != 0) reduce_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partial, grid, sumsq_out, xv, seq);
  } else if (M.fmt == SPIS_FMT_SELL && ctx->spmv_variant > 0) {
    // software-pipelined kernel (4 CTAs per SM by its register budget)
    const int64_t nslices = (M.nrows + 31) / 32;
    const int per_sm = ctx->spmv_pipe_ctas_per_sm > 0 ? ctx->spmv_pipe_ctas_per_sm : 4;
    const int grid = grid_for(ctx, (nslices + kWarps - 1) / kWarps, per_sm);
    spmv_sellp_kernel<MODE, false><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, M.rowperm, nullptr, M.scols, M.svals, nullptr, nullptr, M.nrows, x, b, y, ctx->d_partial);
    if (MODE != 0) reduce_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partial, grid, sumsq_out, xv, seq);
  } else if (M.fmt == SPIS_FMT_SELL) {
    const int64_t nslices = (M.nrows + 31) / 32;
    const int grid = grid_for(ctx, (nslices + kWarps - 1) / kWarps, (MODE != 0 && ctx->spmv_ctas_per_sm > 6) ? 6 : ctx->spmv_ctas_per_sm);
    spmv_sell_kernel<MODE><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, M.rowperm, M.scols, M.svals, M.nrows, x, b, y, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq);
  } else if (M.fmt == SPIS_FMT_SELLD) {
    const int64_t nslices = (M.nrows + 31) / 32;
    const int grid = grid_for(ctx, (nslices + kWarps - 1) / kWarps, (MODE != 0 && ctx->spmv_ctas_per_sm > 6) ? 6 : ctx->spmv_ctas_per_sm);
    spmv_selld_kernel<MODE><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, M.rowperm, M.code_off, M.scols, M.codes, M.dict, M.nrows, x, b, y, ctx->d_partial);
    if (MODE != 0) reduce_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partial, grid, sumsq_out, xv, seq);
  } else if (int T_ = 0, WS_ = 0; M.fmt == SPIS_FMT_PATTERN && [&] { size_t sm_ = 0; return fw_plan(ctx, M, 1, &T_, &WS_, &sm_, 2); }()) {
    // field-blocked system: x windows staged in shared memory, every field block of a node range by one CTA
    FwVecs vv{}; vv.x[0] = x; vv.y[0] = y;
    int grid = 1;
    TRY((launch_fw<1, MODE>(ctx, M, vv, b, ctx->d_partial, &grid)));
    if (MODE != 0) reduce_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partial, grid, sumsq_out, xv, seq);
  } else if (M.fmt == SPIS_FMT_PATTERN) {
    // variant 0: first-generation kernel; 1+: the id of the next round's row is prefetched
    const bool pf = ctx->spmv_variant > 0;
    int grid = grid_for(ctx, (M.nrows + kThreads - 1) / kThreads, ctx->spmv_ctas_per_sm);
    while (grid > 1 && (int64_t)M.nrows + (int64_t)grid * kThreads >= (int64_t)INT32_MAX) grid /= 2;   // 32-bit row arithmetic of the prefetch
#define SPIS_PAT_LAUNCH(NC, PF) spmv_pattern_kernel<MODE, NC, PF><<<grid, kThreads, 0, ctx->stream>>>(M.pid, M.patW, M.tab_off, M.tab_val, (int)M.nrows, x, b, y, ctx->d_partial)
#define SPIS_PAT_CASE(NC) case NC: if (pf) SPIS_PAT_LAUNCH(NC, true); else SPIS_PAT_LAUNCH(NC, false); break;
    switch (M.patW / 4) { SPIS_PAT_CASE(1) SPIS_PAT_CASE(2) SPIS_PAT_CASE(3) SPIS_PAT_CASE(4)
      default: if (pf) SPIS_PAT_LAUNCH(0, true); else SPIS_PAT_LAUNCH(0, false); }
#undef SPIS_PAT_CASE
#undef SPIS_PAT_LAUNCH
    if (MODE != 0) reduce_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partial, grid, sumsq_out, xv, seq);
  } else if (M.fmt == SPIS_FMT_SELL2) {
    const int64_t nslices = (M.nrows + 31) / 32;
    const int grid = grid_for(ctx, (nslices + kWarps - 1) / kWarps, (MODE != 0 && ctx->spmv_ctas_per_sm > 6) ? 6 : ctx->spmv_ctas_per_sm);
    spmv_sell2_kernel<MODE><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, reinterpret_cast<const int2*>(M.scols), reinterpret_cast<const double2*>(M.svals), M.nrows, x, b, y, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq);
  } else {
    const int T = M.csr_lanes;
    const int64_t nblocks = (M.nrows + (kThreads / T) - 1) / (kThreads / T);
    const int grid = grid_for(ctx, nblocks, ctx->spmv_ctas_per_sm);
#define SPIS_CSR_CASE(TT) case TT: spmv_csr_kernel<TT, MODE><<<grid, kThreads, 0, ctx->stream>>>(M.indptr, M.cols, M.vals, M.nrows, x, b, y, ctx->d_partial, ctx->d_counter, sumsq_out, xv, seq); break;
    switch (T) { SPIS_CSR_CASE(2) SPIS_CSR_CASE(4) SPIS_CSR_CASE(8) SPIS_CSR_CASE(16) default: SPIS_CSR_CASE(32) }
#undef SPIS_CSR_CASE
  }
  CU(cudaGetLastError());
  return SPIS_OK;
}

// mode 0: y = Mx ; 1: y = b - Mx, sumsq ; 2: sumsq = ||Mx - b||^2.  x must have its halo filled.
int launch_spmv(spis_ctx* ctx, int slot, int mode, const double* x, const double* b, double* y, double* sumsq_out) {
  const Matrix& M = ctx->mats[slot];
  REQUIRE(M.present, "matrix slot %d has not been uploaded", slot);
  const double bytes = 12.0 * (double)M.nnz + 4.0 * (double)(M.nrows + 1) + (mode == 1 ? 24.0 : 16.0) * (double)M.nrows;
  const double moved = matrix_bytes(M) + (mode == 1 ? 24.0 : 16.0) * (double)M.nrows;
  TRY(prof_begin(ctx, slot == SPIS_SLOT_A ? SPIS_PROF_SPMV : SPIS_PROF_SPMV_AUX, bytes, moved));
  if (mode == 0) TRY(launch_spmv_mode<0>(ctx, M, x, b, y, sumsq_out));
  else if (mode == 1) TRY(launch_spmv_mode<1>(ctx, M, x, b, y, sumsq_out));
  else TRY(launch_spmv_mode<2>(ctx, M, x, b, y, sumsq_out));
  TRY(prof_end(ctx));
  if (mode != 0) return do_allreduce(ctx, sumsq_out, 1);
  return SPIS_OK;
}

// y1 = A x1 and sumsq = ||A x2 - b||^2 from one pass over the system matrix (formats without a dual kernel: two launches)
// ride_partial != null: the per-CTA partial sums of the norm are left there (*ride_parts of them) for the tail of the
// next reducing kernel to finish (TailExtra) instead of a reduce_partials launch; formats without a dual kernel finish
// the norm themselves and report *ride_parts = 0.
int launch_spmv_dual(spis_ctx* ctx, const double* x1, double* y1, const double* x2, const double* b, double* sumsq_out,
                     double* ride_partial = nullptr, int* ride_parts = nullptr) {
  const Matrix& M = ctx->mats[SPIS_SLOT_A];
  REQUIRE(M.present, "matrix slot %d has not been uploaded", SPIS_SLOT_A);
  const bool fused = ctx->spmv_dual && (M.fmt == SPIS_FMT_PATTERN || M.fmt == SPIS_FMT_SELL || M.fmt == SPIS_FMT_SELLD);
  if (ride_parts) *ride_parts = 0;
  if (!fused) {
    TRY(launch_spmv(ctx, SPIS_SLOT_A, 0, x1, nullptr, y1, nullptr));
    return launch_spmv(ctx, SPIS_SLOT_A, 2, x2, b, nullptr, sumsq_out);
  }
  double* part = ride_partial ? ride_partial : ctx->d_partial;
  const double bytes = 2.0 * (12.0 * (double)M.nnz + 4.0 * (double)(M.nrows + 1) + 16.0 * (double)M.nrows);   // two SpMVs' worth
  const double moved = matrix_bytes(M) + 32.0 * (double)M.nrows;      // the matrix once, x1, x2 and b read, y1 written
  TRY(prof_begin(ctx, SPIS_PROF_SPMV, bytes, moved));
  int grid = 1;
  int fwT = 0, fwWS = 0; size_t fwsm = 0;
  if (fw_plan(ctx, M, 2, &fwT, &fwWS, &fwsm, 1)) {
    FwVecs vv{}; vv.x[0] = x1; vv.x[1] = x2; vv.y[0] = y1;
    TRY((launch_fw<2, 3>(ctx, M, vv, b, part, &grid)));
  } else if (sw_plan(ctx, M, 2, 1)) {
    FwVecs vv{}; vv.x[0] = x1; vv.x[1] = x2; vv.y[0] = y1;
    TRY((launch_sellw<2, 3>(ctx, M, vv, b, part, &grid)));
  } else if (M.fmt == SPIS_FMT_PATTERN) {
    grid = grid_for(ctx, (M.nrows + kThreads - 1) / kThreads, ctx->spmv_dual_ctas_per_sm > 0 ? ctx->spmv_dual_ctas_per_sm : 6);
    while (grid > 1 && (int64_t)M.nrows + (int64_t)grid * kThreads >= (int64_t)INT32_MAX) grid /= 2;
#define SPIS_PATD_CASE(NC) case NC: spmv_pattern_dual_kernel<NC><<<grid, kThreads, 0, ctx->stream>>>(M.pid, M.patW, M.tab_off, M.tab_val, (int)M.nrows, x1, y1, x2, b, part); break;
    switch (M.patW / 4) { SPIS_PATD_CASE(1) SPIS_PATD_CASE(2) SPIS_PATD_CASE(3) SPIS_PATD_CASE(4)
      default: spmv_pattern_dual_kernel<0><<<grid, kThreads, 0, ctx->stream>>>(M.pid, M.patW, M.tab_off, M.tab_val, (int)M.nrows, x1, y1, x2, b, part); }
#undef SPIS_PATD_CASE
  } else {
    const int64_t nslices = (M.nrows + 31) / 32;
    const bool coded = M.fmt == SPIS_FMT_SELLD;
    grid = grid_for(ctx, (nslices + kWarps - 1) / kWarps, ctx->spmv_dual_ctas_per_sm > 0 ? ctx->spmv_dual_ctas_per_sm : (coded ? 5 : 4));
    if (coded) spmv_sell_dual_kernel<true><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, M.rowperm, M.code_off, M.scols, nullptr, M.codes, M.dict, M.nrows, x1, y1, x2, b, part);
    else spmv_sell_dual_kernel<false><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, M.rowperm, nullptr, M.scols, M.svals, nullptr, nullptr, M.nrows, x1, y1, x2, b, part);
  }
  if (ride_partial) {
    CU(cudaGetLastError());
    if (ride_parts) *ride_parts = grid;
    return prof_end(ctx);
  }
  const XView xv = fused_view(ctx);
  const unsigned long long seq = fused_seq(ctx);
  reduce_partials_kernel<<<1, kThreads, 0, ctx->stream>>>(ctx->d_partial, grid, sumsq_out, xv, seq);
  CU(cudaGetLastError());
  TRY(prof_end(ctx));
  return do_allreduce(ctx, sumsq_out, 1);
}

// y_c = M x_c, c < nv (nv = 2 or 4), vectors `xstride` / `ystride` doubles apart: one pass over the matrix in `slot`
// for the whole group (formats without a multi-vector kernel: nv launches)
int launch_spmv_multi(spis_ctx* ctx, int slot, int nv, const double* x, int64_t xstride, double* y, int64_t ystride) {
  const Matrix& M = ctx->mats[slot];
  REQUIRE(M.present, "matrix slot %d has not been uploaded", slot);
  const bool fused = ctx->spmv_multi && (nv == 2 || nv == 4) &&
                     (M.fmt == SPIS_FMT_PATTERN || M.fmt == SPIS_FMT_SELL || M.fmt == SPIS_FMT_SELLD);
  if (!fused) {
    for (int c = 0; c < nv; ++c) TRY(launch_spmv(ctx, slot, 0, x + (size_t)c * xstride, nullptr, y + (size_t)c * ystride, nullptr));
    return SPIS_OK;
  }
  const double bytes = (double)nv * (12.0 * (double)M.nnz + 4.0 * (double)(M.nrows + 1) + 16.0 * (double)M.nrows);   // nv SpMVs' worth
  const double moved = matrix_bytes(M) + (double)nv * 16.0 * (double)M.nrows;
  TRY(prof_begin(ctx, slot == SPIS_SLOT_A ? SPIS_PROF_SPMV : SPIS_PROF_SPMV_AUX, bytes, moved));
  int fwT = 0, fwWS = 0; size_t fwsm = 0;
  if (fw_plan(ctx, M, nv, &fwT, &fwWS, &fwsm, 4)) {
    FwVecs vv{};
    for (int c = 0; c < nv; ++c) { vv.x[c] = x + (size_t)c * xstride; vv.y[c] = y + (size_t)c * ystride; }
    if (nv == 2) TRY((launch_fw<2, 0>(ctx, M, vv, nullptr, nullptr, nullptr)));
    else TRY((launch_fw<4, 0>(ctx, M, vv, nullptr, nullptr, nullptr)));
  } else if (sw_plan(ctx, M, nv, 4)) {
    FwVecs vv{};
    for (int c = 0; c < nv; ++c) { vv.x[c] = x + (size_t)c * xstride; vv.y[c] = y + (size_t)c * ystride; }
    if (nv == 2) TRY((launch_sellw<2, 0>(ctx, M, vv, nullptr, nullptr, nullptr)));
    else TRY((launch_sellw<4, 0>(ctx, M, vv, nullptr, nullptr, nullptr)));
  } else if (M.fmt == SPIS_FMT_PATTERN) {
    int grid = grid_for(ctx, (M.nrows + kThreads - 1) / kThreads, nv == 2 ? 6 : 4);
    while (grid > 1 && (int64_t)M.nrows + (int64_t)grid * kThreads >= (int64_t)INT32_MAX) grid /= 2;
#define SPIS_PATM(NC, NVV) spmv_pattern_multi_kernel<NC, NVV><<<grid, kThreads, 0, ctx->stream>>>(M.pid, M.patW, M.tab_off, M.tab_val, (int)M.nrows, x, xstride, y, ystride)
#define SPIS_PATM_CASE(NC) case NC: if (nv == 2) SPIS_PATM(NC, 2); else SPIS_PATM(NC, 4); break;
    switch (M.patW / 4) { SPIS_PATM_CASE(1) SPIS_PATM_CASE(2) SPIS_PATM_CASE(3) SPIS_PATM_CASE(4)
      default: if (nv == 2) SPIS_PATM(0, 2); else SPIS_PATM(0, 4); }
#undef SPIS_PATM_CASE
#undef SPIS_PATM
  } else {
    const int64_t nslices = (M.nrows + 31) / 32;
    const bool coded = M.fmt == SPIS_FMT_SELLD;
    const int grid = grid_for(ctx, (nslices + kWarps - 1) / kWarps, nv == 2 ? 5 : 4);
#define SPIS_SELLM(CD, NVV) spmv_sell_multi_kernel<CD, NVV><<<grid, kThreads, 0, ctx->stream>>>(M.slice_off, M.rowperm, M.code_off, M.scols, M.svals, M.codes, M.dict, M.nrows, x, xstride, y, ystride)
    if (coded) { if (nv == 2) SPIS_SELLM(true, 2); else SPIS_SELLM(true, 4); }
    else { if (nv == 2) SPIS_SELLM(false, 2); else SPIS_SELLM(false, 4); }
#undef SPIS_SELLM
  }
  CU(cudaGetLastError());
  return prof_end(ctx);
}

int launch_scale(spis_ctx* ctx, double* v, const double* sumsq, const double* jac, double* znext) {
  const int grid = grid_for(ctx, (ctx->n + 2 * kThreads - 1) / (2 * kThreads), 8);
  TRY(prof_begin(ctx, SPIS_PROF_SCALE, (jac ? 32.0 : 16.0) * (double)ctx->n));
  scale_kernel<<<grid, kThreads, 0, ctx->stream>>>(v, sumsq, ctx->n, jac, znext);
  CU(cudaGetLastError());
  return prof_end(ctx);
}

int launch_precond(spis_ctx* ctx, const double* q, double* z) {
  switch (ctx->pre_kind) {
    case SPIS_PRE_JACOBI: {
      REQUIRE(ctx->pre_diag, "Jacobi preconditioner: diagonal not uploaded");
      const int grid = grid_for(ctx, (ctx->n + 2 * kThreads - 1) / (2 * kThreads), 8);
      TRY(prof_begin(ctx, SPIS_PROF_PRECOND, 24.0 * (double)ctx->n));
      jacobi_kernel<<<grid, kThreads, 0, ctx->stream>>>(ctx->pre_diag, q, z, ctx->n);
      CU(cudaGetLastError());
      return prof_end(ctx);
    }
    case SPIS_PRE_CSR: {
      // the sparse preconditioner sees q through the same ghost layout as A
      TRY(do_halo(ctx, const_cast<double*>(q)));
      const Matrix& M = ctx->mats[SPIS_SLOT_PRE];
      REQUIRE(M.present, "sparse preconditioner not uploaded");
      const double bytes = 12.0 * (double)M.nnz + 4.0 * (double)(M.nrows + 1) + 16.0 * (double)M.nrows;
      TRY(prof_begin(ctx, SPIS_PROF_PRECOND, bytes));
      TRY(launch_spmv_mode<0>(ctx, M, q, nullptr, z, nullptr));
      return prof_end(ctx);
    }
    case SPIS_PRE_BLOCK: {
      REQUIRE(ctx->pre_blocks, "block preconditioner not uploaded");
      const int grid = grid_for(ctx, (ctx->pre_nblk + kThreads - 1) / kThreads, 8);
      const int bs = ctx->pre_bs;
      TRY(prof_begin(ctx, SPIS_PROF_PRECOND, (8.0 * bs * bs + 16.0 * bs) * (double)ctx->pre_nblk));
#define SPIS_BLK_CASE(BB) case BB: blockdiag_kernel<BB><<<grid, kThreads, 0, ctx->stream>>>(ctx->pre_blocks, ctx->pre_nblk, ctx->pre_sb, ctx->pre_sf, q, z); break;
      switch (bs) { SPIS_BLK_CASE(1) SPIS_BLK_CASE(2) SPIS_BLK_CASE(3) SPIS_BLK_CASE(4) SPIS_BLK_CASE(5) SPIS_BLK_CASE(6) SPIS_BLK_CASE(7) SPIS_BLK_CASE(8)
        default: return fail(ctx, SPIS_E_UNSUPPORTED, "block size %d not supported (1..8)", bs); }
#undef SPIS_BLK_CASE
      CU(cudaGetLastError());
      return prof_end(ctx);
    }
    default:
      return fail(ctx, SPIS_E_INVALID, "launch_precond called for kind %d", ctx->pre_kind);
  }
}

void free_matrix(spis_ctx* ctx, Matrix& M) {
  dfree(ctx, M.indptr); dfree(ctx, M.cols); dfree(ctx, M.vals); dfree(ctx, M.fw_tab); dfree(ctx, M.sw_tiles); dfree(ctx, M.sw_lcol);
  dfree(ctx, M.slice_off); dfree(ctx, M.scols); dfree(ctx, M.svals); dfree(ctx, M.rowperm);
  dfree(ctx, M.pid); dfree(ctx, M.tab_len); dfree(ctx, M.tab_off); dfree(ctx, M.tab_val);
  dfree(ctx, M.codes); dfree(ctx, M.code_off); dfree(ctx, M.dict);
  M = Matrix();
}

inline double* zbase(spis_ctx* ctx) { return ctx->pre_kind == SPIS_PRE_NONE ? ctx->V : ctx->Z; }

int ensure_Z(spis_ctx* ctx) {
  if (ctx->pre_kind != SPIS_PRE_NONE && !ctx->Z) {
    TRY(dalloc(ctx, &ctx->Z, (size_t)ctx->kmax * ctx->ld, false));
    if (ctx->ld > ctx->n) {
      zero_pads_kernel<<<ctx->kmax, 64, 0, ctx->stream>>>(ctx->Z, ctx->n, ctx->ld);
      CU(cudaGetLastError());
    }
  }
  return SPIS_OK;
}

}  // namespace

// ---- process-wide pool of page-locked host buffers -----------------------------------------
// cudaMallocHost costs ~0.3 ms/MB; results (the 8n-byte solution vector) and the small staging
// buffers are recycled across contexts instead.
namespace {
struct PinnedBlock { void* p; size_t cap; bool used; };
std::mutex g_pin_mu;
std::vector<PinnedBlock> g_pin;
}  // namespace

extern "C" {

int spis_pinned_alloc(size_t bytes, void** out) {
  if (!out) return SPIS_E_INVALID;
  *out = nullptr;
  if (bytes == 0) bytes = 8;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  PinnedBlock* best = nullptr;
  for (auto& b : g_pin)
    if (!b.used && b.cap >= bytes && b.cap <= 2 * bytes + 4096 && (!best || b.cap < best->cap)) best = &b;
  if (best) { best->used = true; *out = best->p; return SPIS_OK; }
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e != cudaSuccess) { spis_ctx* ctx = nullptr; return fail(ctx, SPIS_E_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); }
  g_pin.push_back({p, bytes, true});
  *out = p;
  return SPIS_OK;
}

int spis_pinned_free(void* p) {
  if (!p) return SPIS_OK;
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (auto& b : g_pin)
    if (b.p == p) { b.used = false; return SPIS_OK; }
  return SPIS_E_INVALID;
}

int spis_device_trim(void) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  for (auto& b : g_dev_free) { cudaSetDevice(b.device); cudaFree(b.p); }
  g_dev_free.clear();
  cudaSetDevice(dev);
  return SPIS_OK;
}

int spis_pinned_trim(void) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  std::vector<PinnedBlock> keep;
  for (auto& b : g_pin) { if (b.used) keep.push_back(b); else cudaFreeHost(b.p); }
  g_pin.swap(keep);
  return SPIS_OK;
}

// Multi-threaded "does this host buffer hold any non-zero double?" (-0.0 counts as zero, NaN as
// non-zero).  The reference's mass constraint is `0*A`, a CSR with A's pattern and all-zero data
// (lkdv/LinearSolver.py:30); recognising it (to skip its SpMM) means scanning
// 8*nnz bytes, which numpy's any() does at ~8 GB/s on one core.
int spis_host_any_nonzero(const double* p, size_t n, int* out) {
  if (!out || (!p && n)) return SPIS_E_INVALID;
  {
    // Data that is not zero almost always says so in its first entries: look at 64 K of them on the calling
    // thread before paying for a team of threads (which, next to another scan, waited milliseconds for cores).
    const uint64_t* q0 = reinterpret_cast<const uint64_t*>(p);
    const size_t probe = n < ((size_t)1 << 16) ? n : ((size_t)1 << 16);
    uint64_t acc = 0;
    for (size_t k = 0; k < probe; ++k) acc |= q0[k];
    if (acc & 0x7fffffffffffffffull) {
      for (size_t k = 0; k < probe; ++k)
        if (q0[k] & 0x7fffffffffffffffull) { *out = 1; return SPIS_OK; }
    }
    if (probe == n) { *out = 0; return SPIS_OK; }
  }
  std::atomic<int> found(0);
  auto scan = [&](size_t lo, size_t hi) {
    const uint64_t* q = reinterpret_cast<const uint64_t*>(p);
    const size_t blk = 8192;
    for (size_t i = lo; i < hi && !found.load(std::memory_order_relaxed); i += blk) {
      const size_t e = i + blk < hi ? i + blk : hi;
      uint64_t acc = 0;
      for (size_t k = i; k < e; ++k) acc |= q[k];
      if (acc & 0x7fffffffffffffffull) {
        for (size_t k = i; k < e; ++k)
          if (q[k] & 0x7fffffffffffffffull) { found.store(1); break; }
      }
    }
  };
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 16) nt = 16;
  if (tl_use_aux) {                   // helper thread: leave cores to the thread that feeds the GPU
    const char* env = getenv("SPIS_HELPER_SCAN_THREADS");
    unsigned cap = env ? (unsigned)atoi(env) : (nt > 4 ? nt / 2 : nt);
    if (cap >= 1 && nt > cap) nt = cap;
  }
  if (n < (size_t)1 << 20) nt = 1;
  if (nt == 1) scan(0, n);
  else {
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
      const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
      if (lo < hi) th.emplace_back(scan, lo, hi);
    }
    for (auto& t : th) t.join();
  }
  *out = found.load();
  return SPIS_OK;
}

}  // extern "C" (templates below)

// ---- row patterns found on the HOST ------------------------------------------------------------------------------
// try_pattern_storage (below) finds the row stencils of a CSR matrix on the device, AFTER all of it has crossed PCIe:
// 760 MB and 14.4 ms for the 1e7 lkdv operator, of which 20 MB survive (a 16-bit id per row and a table).  The same
// question is answered here by a team of host threads reading the caller's arrays in place: consecutive rows almost
// always share their stencil, so a row is first compared, entry by entry, with the stencil of the row before it
// (column - row and the value's bits: lossless, exactly the device's criterion), and only a row that differs is
// hashed and looked up in the thread's own small dictionary.  The threads' dictionaries are merged afterwards.  A
// matrix that is not of this kind (more than kMaxPatterns distinct rows: unstructured, variable coefficients,
// assembly round-off) makes every thread give up within a few thousand rows, and the device path takes over.
// col_shift: column indices >= n_local are ghosts and move by this much (remap_cols_kernel does it on the device).
namespace {

struct HostPatterns {
  int npat = 0, maxlen = 0;
  int64_t changes = 0;                       // rows whose stencil differs from the row before (pattern_assign_kernel's info[5])
  std::vector<int32_t> rep;                  // representative row of every pattern
  uint16_t* pid = nullptr;                   // IN: the caller's buffer of nrows entries
};

inline unsigned long long hmix64(unsigned long long h, unsigned long long e) {
  h = (h ^ e) * 0xFF51AFD7ED558CCDull;
  return h ^ (h >> 32);
}

inline bool same_stencil(const int32_t* indptr, const int32_t* cols, const double* vals, int64_t r, int64_t q,
                         int32_t n_local, int32_t col_shift) {
  const int32_t p0 = indptr[r], q0 = indptr[q];
  const int32_t len = indptr[r + 1] - p0;
  if (len != indptr[q + 1] - q0) return false;
  const int32_t dr = (int32_t)(r - q);
  const uint64_t* vr = reinterpret_cast<const uint64_t*>(vals + p0);
  const uint64_t* vq = reinterpret_cast<const uint64_t*>(vals + q0);
  unsigned diff = 0;
  for (int32_t k = 0; k < len; ++k) {
    int32_t cr = cols[p0 + k], cq = cols[q0 + k];
    if (cr >= n_local) cr += col_shift;
    if (cq >= n_local) cq += col_shift;
    diff |= (unsigned)((cr - cq) != dr) | (unsigned)(vr[k] != vq[k]);
  }
  return diff == 0;
}

// Row r against the row right before it, whose entries sit right before its own in the CSR arrays: same stencil iff
// every column is the previous row's + 1 and every value has the same bits (equality with the previous row is equality
// with its representative).  Fixed lengths are unrolled and vectorised by the compiler; no ghost-column shift here.
template <int L>
inline bool same_as_previous_fixed(const int32_t* c, const uint64_t* v) {
  uint64_t diff = 0;
#pragma unroll
  for (int k = 0; k < L; ++k) diff |= (uint64_t)(uint32_t)(c[k] - c[k - L] - 1) | (v[k] ^ v[k - L]);
  return diff == 0;
}
inline bool same_as_previous(const int32_t* c, const uint64_t* v, int len) {
  switch (len) {
    case 1: return same_as_previous_fixed<1>(c, v);   case 2: return same_as_previous_fixed<2>(c, v);
    case 3: return same_as_previous_fixed<3>(c, v);   case 4: return same_as_previous_fixed<4>(c, v);
    case 5: return same_as_previous_fixed<5>(c, v);   case 6: return same_as_previous_fixed<6>(c, v);
    case 7: return same_as_previous_fixed<7>(c, v);   case 8: return same_as_previous_fixed<8>(c, v);
    case 9: return same_as_previous_fixed<9>(c, v);   case 10: return same_as_previous_fixed<10>(c, v);
    case 11: return same_as_previous_fixed<11>(c, v); case 12: return same_as_previous_fixed<12>(c, v);
    case 13: return same_as_previous_fixed<13>(c, v); case 14: return same_as_previous_fixed<14>(c, v);
    case 15: return same_as_previous_fixed<15>(c, v); case 16: return same_as_previous_fixed<16>(c, v);
    default: {
      uint64_t diff = 0;
      for (int k = 0; k < len; ++k) diff |= (uint64_t)(uint32_t)(c[k] - c[k - len] - 1) | (v[k] ^ v[k - len]);
      return diff == 0;
    }
  }
}

// returns false: not a pattern matrix (or not worth it) -- nothing is kept
bool host_find_patterns(const int32_t* indptr, const int32_t* cols, const double* vals, int64_t nrows, int32_t n_local,
                        int32_t col_shift, unsigned nthreads, HostPatterns& out, bool early_reject = true) {
  if (nrows <= 0 || !out.pid) return false;
  uint16_t* pid = out.pid;                   // the caller's buffer, nrows entries
  struct Local { std::vector<unsigned long long> key; std::vector<int32_t> rep; std::vector<int32_t> slot; int count = 0; int64_t changes = 0; bool failed = false; };
  constexpr int kSlots = 16384;
  if (nthreads < 1) nthreads = 1;
  if ((int64_t)nthreads > nrows / 4096 + 1) nthreads = (unsigned)(nrows / 4096 + 1);
  std::vector<Local> loc(nthreads);
  std::atomic<int> give_up(0);
  const int64_t per = (nrows + nthreads - 1) / nthreads;
  auto work = [&](unsigned t) {
    Local& L = loc[t];
    L.key.assign(kSlots, 0ull); L.rep.assign(kSlots, -1);
    const int64_t lo = (int64_t)t * per, hi = std::min(nrows, lo + per);
    int64_t last_rep = -1; int last_slot = -1;
    const uint64_t* vbits = reinterpret_cast<const uint64_t*>(vals);
    int32_t prevlen = -1;
    for (int64_t r = lo; r < hi; ++r) {
      if ((r & 1023) == 0) {
        if (give_up.load(std::memory_order_relaxed)) { L.failed = true; return; }
        // storage by patterns needs neighbouring rows to share their stencil (at most one change per eight rows over
        // the matrix, try_pattern_storage): a range where more than every fourth row changes so far ends the attempt
        // (swe: ten row types in sequence; a matrix with assembly round-off: every row its own stencil)
        if (early_reject && r - lo >= 1024 && L.changes * 4 > r - lo) { L.failed = true; give_up.store(1); return; }
      }
      const int32_t p0 = indptr[r], p1 = indptr[r + 1];
      const int32_t len = p1 - p0;
      if (last_rep >= 0 && len == prevlen) {
        const bool same = len == 0 ? true
                          : col_shift == 0 ? same_as_previous(cols + p0, vbits + p0, len)
                                           : same_stencil(indptr, cols, vals, r, last_rep, n_local, col_shift);
        if (same) { pid[r] = (uint16_t)last_slot; continue; }
      }
      prevlen = len;
      if (p1 - p0 > kMaxPatternWidth) { L.failed = true; give_up.store(1); return; }
      unsigned long long h = hmix64(0x9E3779B97F4A7C15ull, (unsigned long long)(p1 - p0));
      for (int32_t p = p0; p < p1; ++p) {
        int32_t c = cols[p]; if (c >= n_local) c += col_shift;
        unsigned long long bits; memcpy(&bits, vals + p, 8);
        h = hmix64(h, (unsigned long long)(unsigned)(c - (int32_t)r) * 0xD6E8FEB86659FD93ull ^ bits);
      }
      h |= 1ull;
      unsigned slot = (unsigned)(h >> 17) % kSlots;
      for (int probes = 0;; ++probes) {
        if (L.key[slot] == 0ull) {
          if (L.count >= kMaxPatterns) { L.failed = true; give_up.store(1); return; }
          L.key[slot] = h; L.rep[slot] = (int32_t)r; ++L.count;
          break;
        }
        if (L.key[slot] == h && same_stencil(indptr, cols, vals, r, L.rep[slot], n_local, col_shift)) break;
        if (probes > 512) { L.failed = true; give_up.store(1); return; }
        slot = (slot + 1) % kSlots;
      }
      if (r > lo) ++L.changes;
      pid[r] = (uint16_t)slot; last_slot = (int)slot; last_rep = L.rep[slot];
    }
  };
  if (nthreads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto& x : th) x.join();
  }
  bool failed = give_up.load() != 0;
  for (auto& L : loc) failed = failed || L.failed;
  // merge: global patterns in order of (thread, slot); a thread's stencil equal to an earlier one takes its id
  std::vector<int32_t> grep_;                            // representative rows of the global patterns
  std::vector<std::vector<int>> map(nthreads);
  if (!failed) {
    std::vector<unsigned long long> gkey;
    for (unsigned t = 0; t < nthreads && !failed; ++t) {
      map[t].assign(kSlots, -1);
      for (int sl = 0; sl < kSlots && !failed; ++sl) {
        if (!loc[t].key[sl]) continue;
        int g = -1;
        for (size_t i = 0; i < gkey.size(); ++i)
          if (gkey[i] == loc[t].key[sl] && same_stencil(indptr, cols, vals, loc[t].rep[sl], grep_[i], n_local, col_shift)) { g = (int)i; break; }
        if (g < 0) {
          if ((int)gkey.size() >= kMaxPatterns) { failed = true; break; }
          g = (int)gkey.size(); gkey.push_back(loc[t].key[sl]); grep_.push_back(loc[t].rep[sl]);
        }
        map[t][sl] = g;
      }
    }
  }
  if (failed) return false;
  // local slot numbers -> global ids; stencil changes across the seams between the threads' ranges
  int64_t changes = 0;
  auto remap = [&](unsigned t) {
    const int64_t lo = (int64_t)t * per, hi = std::min(nrows, lo + per);
    const std::vector<int>& m = map[t];
    for (int64_t r = lo; r < hi; ++r) pid[r] = (uint16_t)m[pid[r]];
  };
  if (nthreads == 1) remap(0);
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; ++t) th.emplace_back(remap, t);
    for (auto& x : th) x.join();
  }
  for (unsigned t = 0; t < nthreads; ++t) {
    changes += loc[t].changes;
    const int64_t lo = (int64_t)t * per;
    if (t > 0 && lo < nrows && pid[lo] != pid[lo - 1]) ++changes;
  }
  out.npat = (int)grep_.size();
  out.rep = grep_;
  out.maxlen = 0;
  for (int32_t r : grep_) out.maxlen = std::max(out.maxlen, (int)(indptr[r + 1] - indptr[r]));
  out.changes = changes;
  return out.npat >= 1;
}

}  // namespace

extern "C" {

static bool is_pinned_host(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// any non-zero among n doubles at `p` (device memory or page-locked host memory), on the calling
// thread's upload stream
static int device_any_nonzero(spis_ctx* ctx, const double* p, size_t n, int* out) {
  cudaStream_t s = up_stream(ctx);
  int* flag = nullptr;                         // per call: several helper threads may be scanning at once
  TRY(dalloc(ctx, &flag, 1));
  const int64_t want = (int64_t)((n / 2 + 255) / 256);
  const int grid = (int)(want < 1 ? 1 : want > ctx->nsm ? ctx->nsm : want);
  any_nonzero_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(p), (int64_t)n, flag);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  dfree(ctx, flag);
  if (e != cudaSuccess) return fail(ctx, SPIS_E_CUDA, "zero test failed: %s", cudaGetErrorString(e));
  return SPIS_OK;
}

// Page-locked host buffer: pulled through a 64 MB device scratch block by the COPY ENGINE and tested
// there.  (A kernel reading the host buffer directly over PCIe also works, but its CTAs stay resident
// for milliseconds, and the Krylov kernels are persistent grids that fill every SM: one foreign CTA per
// SM pushes them into a second wave and doubles their run time -- measured.)
static int pinned_any_nonzero(spis_ctx* ctx, const double* p, size_t n, int* out) {
  cudaStream_t s = up_stream(ctx);
  const size_t chunk = (size_t)8 << 20;                       // doubles per chunk
  double* scratch = nullptr;
  TRY(dalloc(ctx, &scratch, chunk < n ? chunk : n, false));
  int* flag = nullptr;
  TRY(dalloc(ctx, &flag, 1));
  cudaError_t e = cudaSuccess;
  size_t done = 0;
  int found = 0, k = 0;
  while (e == cudaSuccess && done < n && !found) {
    const size_t cnt = (n - done) < chunk ? (n - done) : chunk;
    e = cudaMemcpyAsync(scratch, p + done, cnt * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) break;
    const int64_t want = (int64_t)((cnt / 2 + 255) / 256);
    const int grid = (int)(want < 1 ? 1 : want > ctx->nsm ? ctx->nsm : want);
    any_nonzero_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const unsigned long long*>(scratch), (int64_t)cnt, flag);
    e = cudaGetLastError();
    done += cnt;
    ++k;
    if (e == cudaSuccess && (k == 1 || (k & 3) == 0 || done == n)) {     // look at the flag now and then: non-zero data ends the scan
      e = cudaMemcpyAsync(&found, flag, sizeof(int), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
  }
  dfree(ctx, scratch);
  dfree(ctx, flag);
  if (e != cudaSuccess) return fail(ctx, SPIS_E_CUDA, "zero scan failed: %s", cudaGetErrorString(e));
  *out = found;
  return SPIS_OK;
}

int spis_any_nonzero(spis_ctx* ctx, const double* p, size_t n, int* out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(out && (p || n == 0), "null argument");
  *out = 0;
  if (n == 0) return SPIS_OK;
  CU(cudaSetDevice(ctx->device));
  // Page-locked buffers CAN be tested by the copy engine + a kernel (pinned_any_nonzero: no host CPU time), but
  // measured on the GPU box the host threads scan 480 MB in 1.4-2.5 ms against 9.1 ms through PCIe, and in an
  // end-to-end solve PCIe is the scarce resource (1.3 GB of operands to upload): option pinned_scan_dma = 0.
  if (ctx->pinned_scan_dma && n >= ((size_t)1 << 16) && is_pinned_host(p)) return pinned_any_nonzero(ctx, p, n, out);
  return spis_host_any_nonzero(p, n, out);
}

// ---- host-side small solve ---------------------------------------------------------------
// Equality-constrained least squares  min_y |beta e1 - H y|^2  s.t.  g_c(y) = t0_c + t1_c.y + y'T2_c y = 0,
// the k-dimensional problem that stays on the HOST (solvers.py:251-255).  Native twin of smallsolve.kkt (Python):
// Householder QR of H, Newton on the KKT conditions in u = R y - Q'(beta e1) (objective Hessian 2 I), damped on
// the constraint residual (the sign settling of smallsolve._settle_signs follows in Python).  It sits on the critical path of
// every constrained iteration with the device idle; the numpy version costs 0.4-1.2 ms, this one ~20 us.
// Anything that is not the plain converged case (rank-deficient H, singular KKT matrix, exhausted line search, no
// convergence from the least-squares start, a strongly active constraint that asks for the second start) returns
// *handled = 0 and the caller takes the Python route, so behaviour in the hard cases is unchanged.
namespace {

// in-place LU with partial pivoting; returns false if a pivot vanishes
bool lu_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  for (int k = 0; k < n; ++k) {
    int p = k; double best = std::fabs(A[(size_t)k * n + k]);
    for (int i = k + 1; i < n; ++i) { const double v = std::fabs(A[(size_t)i * n + k]); if (v > best) { best = v; p = i; } }
    if (!(best > 0.0) || !std::isfinite(best)) return false;
    if (p != k) { for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]); std::swap(b[k], b[p]); }
    const double piv = A[(size_t)k * n + k];
    for (int i = k + 1; i < n; ++i) {
      const double f = A[(size_t)i * n + k] / piv;
      if (f != 0.0) { for (int j = k + 1; j < n; ++j) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j]; b[i] -= f * b[k]; }
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = b[i];
    for (int j = i + 1; j < n; ++j) t -= A[(size_t)i * n + j] * b[j];
    b[i] = t / A[(size_t)i * n + i];
  }
  return true;
}

struct KktProblem {
  int m, nc;
  const double *t0, *t1, *t2;            // nc, nc x m, nc x m x m
  std::vector<double> S, c;              // R^{-1} (m x m, row-major, upper triangular), Q'(beta e1)
  std::vector<char> linear;              // T2_c == 0 (mass-type invariants): no curvature, no quadratic form
  mutable std::vector<double> cu, Jy, yT;
  // y = S (c + u); g_c(y); Ju = (t1_c + 2 y'T2_c) S.  Every sum runs over its index in ascending order (the loops
  // are arranged as row updates so that the compiler vectorises them; exact zeros of the triangle are skipped).
  void evaluate(const std::vector<double>& u, std::vector<double>& y, std::vector<double>& g, std::vector<double>& Ju) const {
    cu.resize(m); Jy.resize(m); yT.resize(m);
    for (int i = 0; i < m; ++i) cu[i] = c[i] + u[i];
    for (int i = 0; i < m; ++i) { double t = 0.0; const double* Si = &S[(size_t)i * m]; for (int j = i; j < m; ++j) t += Si[j] * cu[j]; y[i] = t; }
    for (int q = 0; q < nc; ++q) {
      const double* T1 = t1 + (size_t)q * m; const double* T2 = t2 + (size_t)q * m * m;
      double lin = 0.0, quad = 0.0;
      std::fill(yT.begin(), yT.end(), 0.0);
      if (!linear[q])
        for (int i = 0; i < m; ++i) { const double yi = y[i]; const double* Ti = T2 + (size_t)i * m; for (int j = 0; j < m; ++j) yT[j] += yi * Ti[j]; }
      for (int j = 0; j < m; ++j) {
        Jy[j] = T1[j] + 2.0 * yT[j];
        quad += yT[j] * y[j];
        lin += T1[j] * y[j];
      }
      g[q] = (t0[q] + lin) + quad;
      double* Jq = &Ju[(size_t)q * m];
      std::fill(Jq, Jq + m, 0.0);
      for (int i = 0; i < m; ++i) { const double a = Jy[i]; const double* Si = &S[(size_t)i * m]; for (int j = i; j < m; ++j) Jq[j] += a * Si[j]; }
    }
  }
};

}  // namespace

// The sign settling of smallsolve._settle_signs (solvers.py:14-18,266: the reference accepts a constrained step only
// if max_c g_c(y) <= 1e-12, SIGNED): y moves by the minimum-norm correction dy = J'(J J')^{-1}(target - g) that puts
// every g_c a few ulps below zero.  The caller re-evaluates g_c(y) with its own arithmetic afterwards (that evaluation
// is the one the acceptance test uses) and repeats in Python if a sign still disagrees.  *moved_out = corrections made.
int spis_small_settle(int m, int nc, const double* term0, const double* term1, const double* term2, double* y, int tries,
                      int* moved_out) {
  if (!term0 || !term1 || !term2 || !y || m < 1 || nc < 1) return SPIS_E_INVALID;
  if (moved_out) *moved_out = 0;
  const double eps = 2.220446049250313e-16;
  std::vector<double> g(nc), J((size_t)nc * m), yT(m), G((size_t)nc * nc), rhs(nc), ynew(m);
  for (int k = 0; k < tries; ++k) {
    double gmax = -1e300; bool finite = true;
    for (int q = 0; q < nc; ++q) {
      const double* T1 = term1 + (size_t)q * m; const double* T2 = term2 + (size_t)q * m * m;
      std::fill(yT.begin(), yT.end(), 0.0);
      for (int i = 0; i < m; ++i) { const double yi = y[i]; const double* Ti = T2 + (size_t)i * m; for (int j = 0; j < m; ++j) yT[j] += yi * Ti[j]; }
      double lin = 0.0, quad = 0.0;
      for (int j = 0; j < m; ++j) { J[(size_t)q * m + j] = T1[j] + 2.0 * yT[j]; lin += T1[j] * y[j]; quad += yT[j] * y[j]; }
      g[q] = (term0[q] + lin) + quad;
      finite = finite && std::isfinite(g[q]);
      gmax = std::max(gmax, g[q]);
    }
    // "safely negative": the caller's evaluation of g_c differs from this one by an ulp or two of |term0|, so a value
    // within one ulp below zero still counts as positive here and the target sits four ulps below (the caller's own
    // loop aims two ulps below)
    bool settled = finite;
    for (int q = 0; q < nc && settled; ++q) settled = g[q] <= -eps * std::max(std::fabs(term0[q]), std::fabs(g[q]));
    (void)gmax;
    if (!finite || settled) return SPIS_OK;
    for (int q = 0; q < nc; ++q)
      rhs[q] = -std::ldexp(4.0 * eps * std::max(std::fabs(term0[q]), std::fabs(g[q])), k) - g[q];
    for (int a = 0; a < nc; ++a) for (int b = 0; b < nc; ++b) {
      double t = 0.0; for (int j = 0; j < m; ++j) t += J[(size_t)a * m + j] * J[(size_t)b * m + j];
      G[(size_t)a * nc + b] = t;
    }
    if (!lu_solve(G, nc, rhs)) return SPIS_OK;       // dependent gradients: the caller's lstsq handles it
    bool ok = true;
    for (int j = 0; j < m; ++j) {
      double t = 0.0; for (int q = 0; q < nc; ++q) t += J[(size_t)q * m + j] * rhs[q];
      ynew[j] = y[j] + t; ok = ok && std::isfinite(ynew[j]);
    }
    if (!ok) return SPIS_OK;
    for (int j = 0; j < m; ++j) y[j] = ynew[j];
    if (moved_out) *moved_out += 1;
  }
  return SPIS_OK;
}

int spis_small_kkt(int m, int ldh, const double* H, double beta, int nc, const double* term0, const double* term1,
                   const double* term2, double* y_out, double* fval_out, int* nit_out, int* handled) {
  if (!H || !term0 || !term1 || !term2 || !y_out || !handled || m < 1 || nc < 1 || ldh < m) return SPIS_E_INVALID;
  *handled = 0;
  const int rows = m + 1;
  // Householder QR of the (m+1) x m matrix; c = first m entries of Q'(beta e1)
  std::vector<double> A((size_t)rows * m), rhs(rows, 0.0);
  for (int i = 0; i < rows; ++i) for (int j = 0; j < m; ++j) A[(size_t)i * m + j] = H[(size_t)i * ldh + j];
  rhs[0] = beta;
  for (int k = 0; k < m; ++k) {
    double nrm = 0.0;
    for (int i = k; i < rows; ++i) nrm += A[(size_t)i * m + k] * A[(size_t)i * m + k];
    nrm = std::sqrt(nrm);
    if (!(nrm > 0.0)) return SPIS_OK;                       // rank-deficient: Python route
    const double akk = A[(size_t)k * m + k];
    const double alpha = akk > 0.0 ? -nrm : nrm;
    std::vector<double> v(rows - k);
    v[0] = akk - alpha;
    for (int i = k + 1; i < rows; ++i) v[i - k] = A[(size_t)i * m + k];
    double vv = 0.0; for (double t : v) vv += t * t;
    if (!(vv > 0.0)) continue;
    for (int j = k; j < m; ++j) {
      double d = 0.0; for (int i = k; i < rows; ++i) d += v[i - k] * A[(size_t)i * m + j];
      const double f = 2.0 * d / vv;
      for (int i = k; i < rows; ++i) A[(size_t)i * m + j] -= f * v[i - k];
    }
    double d = 0.0; for (int i = k; i < rows; ++i) d += v[i - k] * rhs[i];
    const double f = 2.0 * d / vv;
    for (int i = k; i < rows; ++i) rhs[i] -= f * v[i - k];
  }
  double dmin = 1e300;
  for (int k = 0; k < m; ++k) dmin = std::min(dmin, std::fabs(A[(size_t)k * m + k]));
  if (!(dmin > 1e-300)) return SPIS_OK;
  KktProblem P; P.m = m; P.nc = nc; P.t0 = term0; P.t1 = term1; P.t2 = term2;
  P.c.assign(rhs.begin(), rhs.begin() + m);
  P.S.assign((size_t)m * m, 0.0);
  for (int col = 0; col < m; ++col) {                       // S = R^{-1} by back substitution, column by column
    for (int i = col; i >= 0; --i) {
      double t = (i == col) ? 1.0 : 0.0;
      for (int j = i + 1; j <= col; ++j) t -= A[(size_t)i * m + j] * P.S[(size_t)j * m + col];
      P.S[(size_t)i * m + col] = t / A[(size_t)i * m + i];
    }
  }
  // curvature S'(T2 + T2')S per constraint (S is upper triangular; sums over l ascending, as row updates)
  std::vector<double> curv((size_t)nc * m * m, 0.0), tmp((size_t)m * m);
  P.linear.assign(nc, 1);
  for (int q = 0; q < nc; ++q) {
    const double* T2 = term2 + (size_t)q * m * m;
    for (size_t e = 0; e < (size_t)m * m; ++e) if (T2[e] != 0.0) { P.linear[q] = 0; break; }
    if (P.linear[q]) continue;
    std::fill(tmp.begin(), tmp.end(), 0.0);
    for (int i = 0; i < m; ++i) {
      double* ti = &tmp[(size_t)i * m];
      for (int l = 0; l < m; ++l) {
        const double a = T2[(size_t)i * m + l] + T2[(size_t)l * m + i];
        const double* Sl = &P.S[(size_t)l * m];
        for (int j = l; j < m; ++j) ti[j] += a * Sl[j];
      }
    }
    double* Cq = &curv[(size_t)q * m * m];
    for (int l = 0; l < m; ++l) {
      const double* tl = &tmp[(size_t)l * m];
      for (int i = l; i < m; ++i) {
        const double sli = P.S[(size_t)l * m + i];
        double* ci = Cq + (size_t)i * m;
        for (int j = 0; j < m; ++j) ci[j] += sli * tl[j];
      }
    }
  }
  auto nrm2 = [](const std::vector<double>& a) { double t = 0.0; for (double x : a) t += x * x; return std::sqrt(t); };
  std::vector<double> u(m, 0.0), lam(nc, 0.0), y(m), g(nc), Ju((size_t)nc * m), scale(nc);
  P.evaluate(u, y, g, Ju);
  for (int q = 0; q < nc; ++q) scale[q] = std::max(std::fabs(term0[q]), 1e-300);
  const double cnorm = nrm2(P.c);
  bool converged = false; int nit = 0;
  const int N = m + nc;
  std::vector<double> K((size_t)N * N), r(N), un(m), yn(m), gn(nc), Jun((size_t)nc * m), du(m), JtL(m), st(m);
  for (nit = 1; nit <= 40; ++nit) {
    std::fill(K.begin(), K.end(), 0.0);
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) {
      double t = (i == j) ? 2.0 : 0.0;
      for (int q = 0; q < nc; ++q) t += lam[q] * curv[((size_t)q * m + i) * m + j];
      K[(size_t)i * N + j] = t;
    }
    for (int q = 0; q < nc; ++q) for (int j = 0; j < m; ++j) { K[(size_t)j * N + m + q] = Ju[(size_t)q * m + j]; K[(size_t)(m + q) * N + j] = Ju[(size_t)q * m + j]; }
    for (int j = 0; j < m; ++j) { double t = 2.0 * u[j]; for (int q = 0; q < nc; ++q) t += Ju[(size_t)q * m + j] * lam[q]; r[j] = -t; }
    for (int q = 0; q < nc; ++q) r[m + q] = -g[q];
    if (!lu_solve(K, N, r)) return SPIS_OK;                 // singular KKT matrix: Python route (lstsq there)
    bool finite = true; for (double t : r) finite = finite && std::isfinite(t);
    if (!finite) break;
    for (int j = 0; j < m; ++j) du[j] = r[j];
    double gnorm = 0.0; for (int q = 0; q < nc; ++q) gnorm = std::max(gnorm, std::fabs(g[q]) / scale[q]);
    double t = 1.0; bool accepted = false;
    for (int tries = 0; tries < 12; ++tries) {
      for (int j = 0; j < m; ++j) un[j] = u[j] + t * du[j];
      P.evaluate(un, yn, gn, Jun);
      double gm = 0.0; for (int q = 0; q < nc; ++q) gm = std::max(gm, std::fabs(gn[q]) / scale[q]);
      if (gm <= std::max(gnorm, 1e-15) * (1.0 + 1e-3) || gnorm < 1e-13) { accepted = true; break; }
      t *= 0.5;
    }
    if (!accepted) return SPIS_OK;                          // exhausted line search: Python route
    u = un; y = yn; g = gn; Ju = Jun;
    for (int q = 0; q < nc; ++q) lam[q] += t * r[m + q];
    for (int j = 0; j < m; ++j) st[j] = t * du[j];
    const bool small_step = nrm2(st) <= 1e-15 * std::max(std::max(cnorm, nrm2(u)), 1e-300);
    double feas = 0.0; for (int q = 0; q < nc; ++q) feas = std::max(feas, std::fabs(g[q]) / scale[q]);
    for (int j = 0; j < m; ++j) { double a = 0.0; for (int q = 0; q < nc; ++q) a += Ju[(size_t)q * m + j] * lam[q]; JtL[j] = a; st[j] = 2.0 * u[j] + a; }
    const bool stationary = nrm2(st) <= 1e-13 * std::max(nrm2(JtL), 1e-300);
    if (small_step || (feas <= 4e-16 && stationary)) { converged = true; break; }
  }
  if (!converged) return SPIS_OK;
  double fval = 0.0; for (double t : u) fval += t * t;
  if (fval > 1e-4 * cnorm * cnorm) return SPIS_OK;          // strongly active constraints: the caller also tries its warm start
  // (the final word on the signs stays with the caller, smallsolve._settle_signs: it must see the constraint values
  //  exactly as the acceptance test of solvers.py:266 evaluates them -- one ulp of a 1e4-sized invariant is 1.8e-12)
  for (int j = 0; j < m; ++j) y_out[j] = y[j];
  spis_small_settle(m, nc, term0, term1, term2, y_out, 8, nullptr);   // (the caller verifies the signs with its own evaluation)
  if (fval_out) *fval_out = fval;
  if (nit_out) *nit_out = nit;
  *handled = 1;
  return SPIS_OK;
}

// bytes the upload entry points (matrices, vectors, preconditioner data) have sent to devices since the last reset
long long spis_h2d_bytes(int reset) {
  return reset ? g_h2d_bytes.exchange(0) : g_h2d_bytes.load();
}

int spis_abi_version(void) { return SPIS_ABI_VERSION; }

const char* spis_last_error(const spis_ctx* ctx) { return ctx ? ctx->err : g_global_err; }
const char* spis_last_global_error(void) { return g_global_err; }

int spis_device_count(int* count_out) {
  spis_ctx* ctx = nullptr;
  if (!count_out) return fail(ctx, SPIS_E_INVALID, "count_out is null");
  CU(cudaGetDeviceCount(count_out));
  return SPIS_OK;
}

int spis_ctx_create(int device, int64_t n, int64_t n_halo, int k_max, void* stream, spis_ctx** ctx_out) {
  spis_ctx* ctx = nullptr;
  if (!ctx_out) return fail(ctx, SPIS_E_INVALID, "ctx_out is null");
  *ctx_out = nullptr;
  if (n <= 0 || n_halo < 0 || k_max <= 0 || k_max > 2048) return fail(ctx, SPIS_E_INVALID, "bad sizes n=%lld n_halo=%lld k_max=%d", (long long)n, (long long)n_halo, k_max);
  if (n + n_halo + 64 >= (int64_t)INT32_MAX) return fail(ctx, SPIS_E_UNSUPPORTED, "n + n_halo must fit int32 column indices");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(ctx, SPIS_E_CUDA, "no CUDA device available (%s); this library has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(ctx, SPIS_E_INVALID, "device %d out of range (%d devices)", device, ndev);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(ctx, SPIS_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  // cudaGetDeviceProperties takes 5-100 ms (it queries every attribute, some through the driver's
  // global lock); the two attributes needed here cost microseconds
  int cc_major = 0, cc_minor = 0, n_sm = 0;
  e = cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&cc_minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return fail(ctx, SPIS_E_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (cc_major < 10) return fail(ctx, SPIS_E_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", cc_major, cc_minor);

  PhaseTrace pt;
  spis_ctx* c = new spis_ctx();
  ctx = c;
  c->device = device;
  c->nsm = n_sm;
  c->n = n; c->n_halo = n_halo;
  c->hoff = roundup(n, 16);
  c->ld = roundup(c->hoff + n_halo, 16);
  c->kmax = k_max; c->K = k_max + 4;
  c->pstride = 4 * c->K;        // mdotm_kernel<4> reduces 4 x (m+1) sums per launch
  c->max_grid = c->nsm * 16;
  auto bail = [&](int code) { std::string msg = c->err; spis_ctx_destroy(c); snprintf(g_global_err, sizeof(g_global_err), "%s", msg.c_str()); return code; };
#define CTRY(call) do { int r_ = (call); if (r_ != SPIS_OK) return bail(r_); } while (0)
#define CCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, e_ == cudaErrorMemoryAllocation ? SPIS_E_NOMEM : SPIS_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); return bail(e_ == cudaErrorMemoryAllocation ? SPIS_E_NOMEM : SPIS_E_CUDA); } } while (0)
  if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
  else { CCU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  CCU(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
  CCU(cudaEventCreateWithFlags(&c->ev_arnoldi, cudaEventDisableTiming));
  CCU(cudaEventCreateWithFlags(&c->ev_resid, cudaEventDisableTiming));
  { int lo = 0, hi = 0; CCU(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CCU(cudaStreamCreateWithPriority(&c->hstream, cudaStreamNonBlocking, hi)); }
  CCU(cudaStreamCreateWithFlags(&c->dstream, cudaStreamNonBlocking));
  CCU(cudaEventCreateWithFlags(&c->ev_dl, cudaEventDisableTiming));
  CCU(cudaEventCreateWithFlags(&c->ev_chunk, cudaEventDisableTiming));
  CCU(cudaEventCreateWithFlags(&c->ev_orth, cudaEventDisableTiming));
  CCU(cudaEventCreateWithFlags(&c->ev_hess, cudaEventDisableTiming));
  CCU(cudaEventCreate(&c->ev_t0));
  CCU(cudaEventCreate(&c->ev_t1));
  pt.mark("streams + events");
  const size_t ld = (size_t)c->ld;
  // The basis is written before it is read except for the pad entries [n, ld) of each row, which every
  // kernel relies on being zero: clear those (one strided memset) instead of all (k+1) n doubles.
  CTRY(dalloc(c, &c->V, (size_t)(k_max + 1) * ld, false));
  if (c->ld > c->n) {
    zero_pads_kernel<<<k_max + 1, 64, 0, c->stream>>>(c->V, c->n, c->ld);
    CCU(cudaGetLastError());
  }
  CTRY(dalloc(c, &c->W, ld));
  CTRY(dalloc(c, &c->T, ld));
  CTRY(dalloc(c, &c->R0, ld));
  CTRY(dalloc(c, &c->B, ld));
  CTRY(dalloc(c, &c->X0, ld));
  CTRY(dalloc(c, &c->X, ld));
  CTRY(dalloc(c, &c->d_small, (size_t)2 * c->K + 8));
  CTRY(dalloc(c, &c->d_y, (size_t)c->K));
  CTRY(dalloc(c, &c->d_cout, (size_t)k_max * 2 * c->K));
  CTRY(dalloc(c, &c->d_partial, (size_t)c->max_grid * c->pstride));
  CTRY(dalloc(c, &c->d_counter, 4));
  pt.mark("device blocks");
  if (spis_pinned_alloc(((size_t)2 * c->K + 8) * sizeof(double), (void**)&c->h_small) != SPIS_OK ||
      spis_pinned_alloc((size_t)c->K * sizeof(double), (void**)&c->h_y) != SPIS_OK ||
      spis_pinned_alloc(64, (void**)&c->h_resid) != SPIS_OK ||
      spis_pinned_alloc((size_t)k_max * 2 * c->K * sizeof(double), (void**)&c->h_cout) != SPIS_OK) {
    fail(c, SPIS_E_NOMEM, "pinned host allocation failed: %s", g_global_err);
    return bail(SPIS_E_NOMEM);
  }
  pt.mark("pinned mirrors");
  // mdot needs up to (8*K+256)*8 bytes of dynamic shared memory, mdotm<4> four times the K part
  {
    const int need4 = (kWarps * 4 * c->K + kWarps * 32) * (int)sizeof(double);
    if (need4 > 48 * 1024 && need4 <= 227 * 1024) {
      CCU(cudaFuncSetAttribute(mdotm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, need4));
      CCU(cudaFuncSetAttribute(mdotm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, need4));
    }
  }
  const int smem_need = (kWarps * c->K + kWarps * 32) * (int)sizeof(double);
  if (smem_need > 48 * 1024) {
    CCU(cudaFuncSetAttribute(mdot_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_need));
    CCU(cudaFuncSetAttribute(mdot_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_need));
    CCU(cudaFuncSetAttribute(mdot_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_need));
  }
#define SPIS_OM_ATTR(MBV) \
  CCU(cudaFuncSetAttribute(orth_mid_kernel<MBV, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOrthMidSmemBudget)); \
  CCU(cudaFuncSetAttribute(orth_mid_kernel<MBV, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOrthMidSmemBudget));
  SPIS_OM_ATTR(8) SPIS_OM_ATTR(16) SPIS_OM_ATTR(24) SPIS_OM_ATTR(32) SPIS_OM_ATTR(40) SPIS_OM_ATTR(48) SPIS_OM_ATTR(56) SPIS_OM_ATTR(64)
#undef SPIS_OM_ATTR
  pt.mark("kernel attributes");
  CCU(cudaStreamSynchronize(c->stream));
  pt.mark("stream drain");
#undef CTRY
#undef CCU
  *ctx_out = c;
  return SPIS_OK;
}

int spis_ctx_destroy(spis_ctx* ctx) {
  if (!ctx) return SPIS_OK;
  cudaSetDevice(ctx->device);
  spis_constraint_setup_wait(ctx);
  if (ctx->aux) cudaStreamSynchronize(ctx->aux);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto& e : ctx->evpool) cudaEventDestroy(e);
  if (ctx->stream) {
    for (auto& M : ctx->mats) free_matrix(ctx, M);
    for (auto& c : ctx->cons) { dfree(ctx, c.v); dfree(ctx, c.MZ); }
    dfree(ctx, ctx->V); dfree(ctx, ctx->Z); dfree(ctx, ctx->W); dfree(ctx, ctx->T); dfree(ctx, ctx->R0);
    dfree(ctx, ctx->B); dfree(ctx, ctx->X0); dfree(ctx, ctx->X); dfree(ctx, ctx->G); dfree(ctx, ctx->GW); dfree(ctx, ctx->d_gpartial); dfree(ctx, ctx->d_gcounter); dfree(ctx, ctx->pre_diag); dfree(ctx, ctx->pre_blocks);
    dfree(ctx, ctx->d_send_idx); dfree(ctx, ctx->d_send);
    dfree(ctx, ctx->d_dest_rank); dfree(ctx, ctx->d_dest_off); dfree(ctx, ctx->d_send_to); dfree(ctx, ctx->d_recv_from);
    dfree(ctx, ctx->d_small); dfree(ctx, ctx->d_y); dfree(ctx, ctx->d_cout); dfree(ctx, ctx->d_partial); dfree(ctx, ctx->d_counter);
    dfree(ctx, ctx->hs_cs); dfree(ctx, ctx->hs_sn); dfree(ctx, ctx->hs_gv); dfree(ctx, ctx->hs_R); dfree(ctx, ctx->hs_tracking);
    dfree(ctx, ctx->d_phase); dfree(ctx, ctx->d_ydev); dfree(ctx, ctx->d_rpartial);
    cudaStreamSynchronize(ctx->stream);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    for (auto& b : ctx->owned) g_dev_free.push_back(b);
    ctx->owned.clear();
  }
  if (!ctx->comm) {
    for (int r = 0; r < kMaxRanks; ++r)
      if (ctx->xpeer[r]) cudaIpcCloseMemHandle(ctx->xpeer[r]);
    if (ctx->xbuf) cudaFree(ctx->xbuf);
  }
  spis_pinned_free(ctx->h_small); spis_pinned_free(ctx->h_y); spis_pinned_free(ctx->h_cout); spis_pinned_free(ctx->h_resid);
  spis_pinned_free(ctx->h_rec);
  if (ctx->ev_arnoldi) cudaEventDestroy(ctx->ev_arnoldi);
  if (ctx->ev_resid) cudaEventDestroy(ctx->ev_resid);
  if (ctx->dstream) { cudaStreamSynchronize(ctx->dstream); cudaStreamDestroy(ctx->dstream); }
  if (ctx->ev_dl) cudaEventDestroy(ctx->ev_dl);
  if (ctx->ev_chunk) cudaEventDestroy(ctx->ev_chunk);
  if (ctx->ev_orth) cudaEventDestroy(ctx->ev_orth);
  if (ctx->ev_hess) cudaEventDestroy(ctx->ev_hess);
  if (ctx->hstream) cudaStreamDestroy(ctx->hstream);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SPIS_OK;
}

int spis_set_option(spis_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return SPIS_E_INVALID;
  std::string k(key);
  if (k == "orth") { REQUIRE(value >= 0 && value <= 2, "orth must be 0..2"); ctx->orth = (int)value; }
  else if (k == "spmv_format") { REQUIRE(value >= 0 && value <= 5, "spmv_format must be 0..5"); ctx->fmt_pref = (int)value; }
  else if (k == "auto_sell2") { ctx->auto_sell2 = value ? 1 : 0; }
  else if (k == "auto_pattern") { ctx->auto_pattern = value ? 1 : 0; }
  else if (k == "fuse_iterate") { ctx->fuse_iterate = value ? 1 : 0; }
  else if (k == "lincomb2_ctas_per_sm") { REQUIRE(value >= 1 && value <= 8, "lincomb2_ctas_per_sm must be 1..8"); ctx->lincomb2_ctas_per_sm = (int)value; }
  else if (k == "auto_dict") { ctx->auto_dict = value ? 1 : 0; }
  else if (k == "mdotm_ctas_per_sm") { REQUIRE(value >= 1 && value <= 8, "mdotm_ctas_per_sm must be 1..8"); ctx->mdotm_ctas_per_sm = (int)value; }
  else if (k == "bench_mdotm_nw") { REQUIRE(value == 0 || value == 2 || value == 4, "bench_mdotm_nw must be 0, 2 or 4"); ctx->bench_mdotm_nw = (int)value; }
  else if (k == "profile") { ctx->profile = value ? 1 : 0; }
  else if (k == "ctas_per_sm") { REQUIRE(value >= 1 && value <= 16, "ctas_per_sm must be 1..16"); ctx->ctas_per_sm = (int)value; }
  else if (k == "spmv_ctas_per_sm") { REQUIRE(value >= 1 && value <= 16, "spmv_ctas_per_sm must be 1..16"); ctx->spmv_ctas_per_sm = (int)value; }
  else if (k == "spmv_variant") { REQUIRE(value >= 0 && value <= 1, "spmv_variant must be 0 or 1"); ctx->spmv_variant = (int)value; }
  else if (k == "spmv_pipe_ctas_per_sm") { REQUIRE(value >= 0 && value <= 16, "spmv_pipe_ctas_per_sm must be 0..16"); ctx->spmv_pipe_ctas_per_sm = (int)value; }
  else if (k == "pinned_scan_dma") { ctx->pinned_scan_dma = value ? 1 : 0; }
  else if (k == "sell_sigma") { ctx->sell_sigma = value ? 1 : 0; }
  else if (k == "spmv_multi") { ctx->spmv_multi = value ? 1 : 0; }
  else if (k == "spmv_dual") { ctx->spmv_dual = value ? 1 : 0; }
  else if (k == "hess_async") ctx->hess_async = value != 0;
  else if (k == "sweep_reverse") ctx->sweep_reverse = (int)value & 3;
  else if (k == "batch_terms") ctx->batch_terms = value != 0;
  else if (k == "host_pattern") ctx->host_pattern = value != 0;
  else if (k == "host_pattern_min_nnz") ctx->host_pattern_min_nnz = value;
  else if (k == "host_threads") { REQUIRE(value >= 0 && value <= 256, "host_threads out of range"); ctx->host_threads = (int)value; }
  else if (k == "gram") ctx->gram = value != 0;
  else if (k == "spmv_fw_rows") { REQUIRE(value == 4 || value == 8, "spmv_fw_rows must be 4 or 8"); ctx->spmv_fw_rows = (int)value; }
  else if (k == "spmv_sellw") { REQUIRE(value >= 0 && value <= 7, "spmv_sellw is a bit mask 0..7"); ctx->spmv_sellw = (int)value; }
  else if (k == "spmv_fw") { REQUIRE(value >= 0 && value <= 7, "spmv_fw is a bit mask 0..7"); ctx->spmv_fw = (int)value; }
  else if (k == "spmv_dual_ctas_per_sm") { REQUIRE(value >= 0 && value <= 16, "spmv_dual_ctas_per_sm must be 0..16"); ctx->spmv_dual_ctas_per_sm = (int)value; }
  else if (k == "mdot_variant") { REQUIRE(value == 0 || value == 1 || value == 2 || value == 4 || value == 8, "mdot_variant must be 0 (auto), 1 (register sums), 2, 4 or 8"); ctx->mdot_variant = (int)value; }
  else if (k == "mdot_reg_auto") { ctx->mdot_reg_auto = value ? 1 : 0; }
  else if (k == "mdot_reg_ctas_per_sm") { REQUIRE(value >= 0 && value <= 8, "mdot_reg_ctas_per_sm must be 0..8"); ctx->mdot_reg_ctas_per_sm = (int)value; }
  else if (k == "lincomb_variant") { REQUIRE(value == 2 || value == 4 || value == 8, "lincomb_variant must be 2, 4 or 8"); ctx->lincomb_variant = (int)value; }
  else if (k == "x0_is_zero") { ctx->x0_is_zero = value ? 1 : 0; }
  else if (k == "fuse_jacobi") { ctx->fuse_jacobi = value ? 1 : 0; }
  else if (k == "force_nonsymmetric") { ctx->force_nonsymmetric = value ? 1 : 0; }
  else if (k == "orth_fused") { ctx->orth_fused = value ? 1 : 0; }
  else if (k == "orth_mid_probe") { ctx->orth_mid_probe = value ? 1 : 0; }
  else if (k == "orth_mid_force_e") { REQUIRE(value >= 0 && value <= 2, "orth_mid_force_e must be 0..2"); ctx->orth_mid_force_e = (int)value; }
  else if (k == "orth_mid_max_stages") { REQUIRE(value >= 2 && value <= 64, "orth_mid_max_stages must be 2..64"); ctx->orth_mid_max_stages = (int)value; }
  else return fail(ctx, SPIS_E_INVALID, "unknown option '%s'", key);
  return SPIS_OK;
}

int spis_get_info(const spis_ctx* cctx, const char* key, int64_t* value_out) {
  spis_ctx* ctx = const_cast<spis_ctx*>(cctx);
  if (!ctx || !key || !value_out) return SPIS_E_INVALID;
  std::string k(key);
  if (k == "n") *value_out = ctx->n;
  else if (k == "ld") *value_out = ctx->ld;
  else if (k == "hoff") *value_out = ctx->hoff;
  else if (k == "n_halo") *value_out = ctx->n_halo;
  else if (k == "k_max") *value_out = ctx->kmax;
  else if (k == "num_sms") *value_out = ctx->nsm;
  else if (k == "device") *value_out = ctx->device;
  else if (k.rfind("fmt:", 0) == 0 || k.rfind("nnz_padded:", 0) == 0 || k.rfind("nnz:", 0) == 0) {
    const int slot = atoi(k.c_str() + k.find(':') + 1);
    REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS && ctx->mats[slot].present, "slot %d not uploaded", slot);
    if (k[0] == 'f') *value_out = ctx->mats[slot].fmt;
    else if (k.rfind("nnz_padded:", 0) == 0) *value_out = ctx->mats[slot].nnz_padded;
    else *value_out = ctx->mats[slot].nnz;
  }
  else if (k.rfind("ndict:", 0) == 0) {
    const int slot = atoi(k.c_str() + 6);
    REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS && ctx->mats[slot].present, "slot %d not uploaded", slot);
    *value_out = ctx->mats[slot].ndict;
  }
  else if (k.rfind("sellw_cap:", 0) == 0) {
    const int slot = atoi(k.c_str() + 10);
    REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS && ctx->mats[slot].present, "slot %d not uploaded", slot);
    *value_out = (ctx->mats[slot].sw_lcol && ctx->spmv_sellw) ? ctx->mats[slot].sw_cap : 0;
  }
  else if (k.rfind("fw_fields:", 0) == 0) {
    const int slot = atoi(k.c_str() + 10);
    REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS && ctx->mats[slot].present, "slot %d not uploaded", slot);
    *value_out = (ctx->mats[slot].fw_tab && ctx->spmv_fw) ? ctx->mats[slot].fwF : 0;
  }
  else if (k.rfind("npat:", 0) == 0) {
    const int slot = atoi(k.c_str() + 5);
    REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS && ctx->mats[slot].present, "slot %d not uploaded", slot);
    *value_out = ctx->mats[slot].npat;
  }
  else if (k == "alloc_hits") *value_out = g_dev_hits.load();
  else if (k == "alloc_misses") *value_out = g_dev_misses.load();
  else if (k == "alloc_miss_bytes") *value_out = g_dev_miss_bytes.load();
  else if (k == "can_fuse_iterate") *value_out = (ctx->fuse_iterate && ctx->orth == SPIS_ORTH_CGS2 && ctx->pre_kind == SPIS_PRE_NONE) ? 1 : 0;
  else if (k == "pipe_lag") *value_out = ctx->pre_kind == SPIS_PRE_NONE ? 1 : 0;
  else if (k == "sharded") *value_out = (ctx->xactive || ctx->allreduce != nullptr || ctx->n_halo > 0) ? 1 : 0;
  else if (k == "device_pipeline") *value_out = (ctx->orth == SPIS_ORTH_CGS2 && ctx->pre_kind != SPIS_PRE_HOST && !ctx->allreduce && !ctx->halo &&
                                                 (ctx->fuse_iterate || ctx->pre_kind != SPIS_PRE_NONE)) ? 1 : 0;
  else if (k == "device_ptr:small") *value_out = (int64_t)(intptr_t)ctx->d_small;
  else if (k == "stream") *value_out = (int64_t)(intptr_t)ctx->stream;
  else if (k == "n_send") *value_out = ctx->n_send;
  else return fail(ctx, SPIS_E_INVALID, "unknown info key '%s'", key);
  return SPIS_OK;
}

// ---- uploads --------------------------------------------------------------------------
// SELL matrix on the device -> dictionary-coded values if it holds at most 256 distinct doubles.
static int try_value_dictionary(spis_ctx* ctx, Matrix& M, cudaStream_t s, const std::vector<int64_t>& off, int* ok_out) {
  *ok_out = 0;
  if (M.nnz_padded <= 0) return SPIS_OK;
  unsigned long long* keys = nullptr; int *dense = nullptr, *info = nullptr;
  TRY(dalloc(ctx, &keys, (size_t)kDictSlots));
  TRY(dalloc(ctx, &dense, (size_t)kDictSlots, false));
  TRY(dalloc(ctx, &info, 8));
  TRY(dalloc(ctx, &M.dict, 256));
  auto cleanup = [&]() { dfree(ctx, keys); dfree(ctx, dense); dfree(ctx, info); };
  int h_info[8] = {0};
  const int grid = ctx->nsm * 8;
  dict_insert_kernel<<<grid, 256, 0, s>>>(M.svals, M.nnz_padded, keys, info);
  dict_number_kernel<<<1, 32, 0, s>>>(keys, dense, M.dict, info);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_info, info, sizeof(h_info), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { cleanup(); dfree(ctx, M.dict); return fail(ctx, SPIS_E_CUDA, "value dictionary failed: %s", cudaGetErrorString(e)); }
  if (h_info[1] || h_info[2] < 1 || h_info[2] > 256) { cleanup(); dfree(ctx, M.dict); return SPIS_OK; }
  // word offsets of the packed codes: ceil(width / 4) words per lane and slice
  const int64_t nslices = (int64_t)off.size() - 1;
  std::vector<int64_t> coff((size_t)nslices + 1);
  coff[0] = 0;
  for (int64_t q = 0; q < nslices; ++q) coff[q + 1] = coff[q] + (((off[q + 1] - off[q]) / 32 + 3) / 4) * 32;
  int rc = dalloc(ctx, &M.codes, (size_t)coff[nslices], false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.code_off, (size_t)nslices + 1, false);
  if (rc != SPIS_OK) { cleanup(); dfree(ctx, M.codes); dfree(ctx, M.dict); return rc; }
  e = cudaMemcpyAsync(M.code_off, coff.data(), ((size_t)nslices + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s);
  const int cgrid = (int)((nslices * 32 + 255) / 256);
  dict_encode_kernel<<<cgrid, 256, 0, s>>>(M.svals, M.slice_off, M.code_off, nslices, keys, dense, M.codes);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);       // coff (pageable) must outlive the copy
  cleanup();
  if (e != cudaSuccess) { dfree(ctx, M.codes); dfree(ctx, M.code_off); dfree(ctx, M.dict); return fail(ctx, SPIS_E_CUDA, "value encoding failed: %s", cudaGetErrorString(e)); }
  dfree(ctx, M.svals);
  M.ndict = h_info[2];
  M.fmt = SPIS_FMT_SELLD;
  *ok_out = 1;
  return SPIS_OK;
}

// Try to store the CSR matrix (already on the device) as row patterns.  *ok_out = 1 on success.
static int try_pattern_storage(spis_ctx* ctx, Matrix& M, cudaStream_t s, int* ok_out) {
  *ok_out = 0;
  const int64_t nrows = M.nrows;
  if (nrows <= 0 || M.nnz <= 0) return SPIS_OK;
  unsigned long long* hash = nullptr; unsigned long long* keys = nullptr;
  int *rep = nullptr, *dense = nullptr, *info = nullptr; int32_t* slot_of_row = nullptr;
  TRY(dalloc(ctx, &hash, (size_t)nrows, false));
  TRY(dalloc(ctx, &slot_of_row, (size_t)nrows, false));
  TRY(dalloc(ctx, &keys, (size_t)kPatternSlots));
  TRY(dalloc(ctx, &rep, (size_t)kPatternSlots, false));
  TRY(dalloc(ctx, &dense, (size_t)kPatternSlots, false));
  TRY(dalloc(ctx, &info, 8));
  auto cleanup = [&]() { dfree(ctx, hash); dfree(ctx, slot_of_row); dfree(ctx, keys); dfree(ctx, rep); dfree(ctx, dense); dfree(ctx, info); };
  auto drop = [&]() { dfree(ctx, M.pid); dfree(ctx, M.tab_len); dfree(ctx, M.tab_off); dfree(ctx, M.tab_val); };
  int rc = SPIS_OK;
  int h_info[8] = {0};
  const int grid = ctx->nsm * 8;
  cudaError_t e = cudaMemsetAsync(rep, 0x7f, (size_t)kPatternSlots * sizeof(int), s);      // "infinity" for atomicMin
  pattern_hash_kernel<<<grid, 256, 0, s>>>(M.indptr, M.cols, M.vals, nrows, hash);
  pattern_insert_kernel<<<grid, 256, 0, s>>>(hash, nrows, keys, rep, slot_of_row, info);
  pattern_number_kernel<<<1, 32, 0, s>>>(keys, dense, info);
  pattern_maxlen_kernel<<<kPatternSlots / 256, 256, 0, s>>>(M.indptr, rep, keys, info);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_info, info, sizeof(h_info), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { cleanup(); return fail(ctx, SPIS_E_CUDA, "pattern detection failed: %s", cudaGetErrorString(e)); }
  const int npat = h_info[2], maxlen = h_info[3];
  if (h_info[1] || npat < 1 || npat > kMaxPatterns || maxlen > kMaxPatternWidth) { cleanup(); return SPIS_OK; }
  const int W = maxlen < 4 ? 4 : (maxlen + 3) / 4 * 4;
  // worth it only if the stencil table is small next to the matrix it replaces (and so stays in L1)
  if ((double)npat * W * 8.0 > (double)M.nnz && ctx->fmt_pref != SPIS_FMT_PATTERN) { cleanup(); return SPIS_OK; }
  rc = dalloc(ctx, &M.pid, (size_t)nrows, false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.tab_len, (size_t)npat, false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.tab_off, (size_t)npat * W, false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.tab_val, (size_t)npat * W, false);
  if (rc != SPIS_OK) { cleanup(); drop(); return rc; }
  pattern_table_kernel<<<kPatternSlots / 256, 256, 0, s>>>(M.indptr, M.cols, M.vals, rep, dense, W, M.tab_len, M.tab_off, M.tab_val);
  pattern_assign_kernel<<<grid, 256, 0, s>>>(M.indptr, M.cols, M.vals, nrows, slot_of_row, dense, W, M.tab_len, M.tab_off, M.tab_val, M.pid, info);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_info, info, sizeof(h_info), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cleanup();
  if (e != cudaSuccess) { drop(); return fail(ctx, SPIS_E_CUDA, "pattern assignment failed: %s", cudaGetErrorString(e)); }
  if (h_info[1]) { drop(); return SPIS_OK; }              // a hash collision: keep the general storage
  // neighbouring rows must mostly share their stencil (see pattern_assign_kernel); measured: the swe energy
  // matrix (10 row types in sequence) runs at 310 us as patterns, 262 us as SELL, 180 us as SELLD
  if ((double)h_info[5] * 8.0 > (double)nrows && ctx->fmt_pref != SPIS_FMT_PATTERN) { drop(); return SPIS_OK; }
  M.npat = npat; M.patW = W;
  *ok_out = 1;
  return SPIS_OK;
}

// SELL / SELLD matrix on the device -> per-tile x windows and 16-bit window-local columns (spmv_sellw_kernel), if every
// tile of 256 rows decomposes into at most 12 windows of at most kSwCapMax doubles together.
static int try_sellw(spis_ctx* ctx, Matrix& M, cudaStream_t s) {
  const int64_t nslices = (M.nrows + 31) / 32;
  const int64_t ntiles = (nslices + kSwSlices - 1) / kSwSlices;
  if (ntiles < 1 || ntiles > 0x7fffffff) return SPIS_OK;
  int* info = nullptr;
  TRY(dalloc(ctx, &info, 4));
  TRY(dalloc(ctx, &M.sw_tiles, (size_t)ntiles, false));
  TRY(dalloc(ctx, &M.sw_lcol, (size_t)M.nnz_padded, false));
  sellw_analyse_kernel<<<(unsigned)ntiles, 256, 0, s>>>(M.slice_off, M.scols, M.nrows, M.sw_tiles, M.sw_lcol, kSwCapMax, info);
  int h[4] = {0};
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  dfree(ctx, info);
  if (e != cudaSuccess) { dfree(ctx, M.sw_tiles); dfree(ctx, M.sw_lcol); return fail(ctx, SPIS_E_CUDA, "window analysis failed: %s", cudaGetErrorString(e)); }
  if (h[0] || h[1] < 2) { dfree(ctx, M.sw_tiles); dfree(ctx, M.sw_lcol); return SPIS_OK; }     // some tile does not decompose: plain kernels
  M.sw_cap = (h[1] + 15) / 16 * 16;
  return SPIS_OK;
}

// Row-pattern matrix on the device -> field-window table for spmv_fw_kernel, if the unknowns are F fields of N nodes
// (n = F N) and (nearly) every stencil offset is q N + d with a small node shift d: the reference's field-blocked
// systems [u; v; w] (lkdv/refd.py:17), the stage-block systems of lkdvRK, and any banded single-field matrix (F = 1).
static int try_field_windows(spis_ctx* ctx, Matrix& M, cudaStream_t s) {
  if (!ctx->spmv_fw || M.npat < 1 || M.npat > kFwMaxPat || M.npat * M.patW > kFwMaxTable || M.nrows < 32) return SPIS_OK;
  const int npat = M.npat, W = M.patW;
  std::vector<int32_t> len(npat), off((size_t)npat * W);
  std::vector<double> val((size_t)npat * W);
  std::vector<unsigned long long> cnt(npat, 0ull);
  unsigned long long* d_hist = nullptr;
  TRY(dalloc(ctx, &d_hist, (size_t)npat));
  pid_hist_kernel<<<ctx->nsm * 4, 256, 0, s>>>(M.pid, M.nrows, npat, d_hist);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt.data(), d_hist, (size_t)npat * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(len.data(), M.tab_len, (size_t)npat * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(off.data(), M.tab_off, off.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(val.data(), M.tab_val, val.size() * sizeof(double), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  dfree(ctx, d_hist);
  if (e != cudaSuccess) return fail(ctx, SPIS_E_CUDA, "field-window analysis failed: %s", cudaGetErrorString(e));
  constexpr int kDmax = 64;
  const long long n = M.nrows;
  int bestF = 0, bestD = 0;
  for (int F = 1; F <= 8 && !bestF; ++F) {
    if (n % F) continue;
    const long long N = n / F;
    if (N < 8) break;
    unsigned long long covered = 0;
    int Dm = 0;
    for (int p = 0; p < npat; ++p) {
      bool regular = true;
      int dm = 0;
      for (int k = 0; k < len[p] && regular; ++k) {
        const long long o = off[(size_t)p * W + k];
        const long long q = (o >= 0 ? o + N / 2 : o - N / 2) / N;
        const long long d = o - q * N;
        if (d < -kDmax || d > kDmax || q < -63 || q > 63) regular = false;
        else dm = std::max<int>(dm, (int)std::llabs(d));
      }
      if (regular) { covered += cnt[p]; Dm = std::max(Dm, dm); }
    }
    if ((double)covered >= 0.98 * (double)n) { bestF = F; bestD = Dm; }
  }
  if (!bestF) return SPIS_OK;
  const long long N = n / bestF;
  const int D = std::max(2, (bestD + 1) / 2 * 2);
  std::vector<FwEntry> tab((size_t)npat * W);
  for (int p = 0; p < npat; ++p)
    for (int k = 0; k < W; ++k) {
      FwEntry t; t.val = k < len[p] ? val[(size_t)p * W + k] : 0.0;
      const long long o = k < len[p] ? off[(size_t)p * W + k] : 0;
      const long long q = (o >= 0 ? o + N / 2 : o - N / 2) / N;
      const long long d = o - q * N;
      if (d >= -D && d <= D && q >= -63 && q <= 63) { t.reg = (int)(2 * (q + 64) + 1); t.soff = (int)(d + D); }
      else { t.reg = 0; t.soff = (int)o; }
      tab[(size_t)p * W + k] = t;
    }
  TRY(dalloc(ctx, &M.fw_tab, tab.size(), false));
  TRY(h2d(ctx, M.fw_tab, tab.data(), tab.size() * sizeof(FwEntry)));
  M.fwF = bestF; M.fwN = (int)N; M.fwD = D;
  return SPIS_OK;
}

// The host-side detection alone (no device involved): pid_out[nrows], rep_out[<= 4096] representative rows.  *npat_out = 0:
// the matrix is not a pattern matrix.  For tests and for callers that want to know before they build a context.
int spis_host_find_patterns(const int32_t* indptr, const int32_t* indices, const double* data, int64_t nrows, int64_t n_local,
                            int64_t col_shift, int nthreads, uint16_t* pid_out, int32_t* rep_out, int* npat_out,
                            int* maxlen_out, int64_t* changes_out) {
  if (!indptr || !npat_out || nrows < 0 || (nrows > 0 && indptr[nrows] > 0 && (!indices || !data))) return SPIS_E_INVALID;
  *npat_out = 0;
  HostPatterns hp;
  std::vector<uint16_t> tmp;
  if (!pid_out) { tmp.resize((size_t)nrows); pid_out = tmp.data(); }
  hp.pid = pid_out;
  if (!host_find_patterns(indptr, indices, data, nrows, (int32_t)n_local, (int32_t)col_shift, nthreads > 0 ? (unsigned)nthreads : 1u, hp)) return SPIS_OK;
  if (rep_out) for (int p = 0; p < hp.npat; ++p) rep_out[p] = hp.rep[p];
  *npat_out = hp.npat;
  if (maxlen_out) *maxlen_out = hp.maxlen;
  if (changes_out) *changes_out = hp.changes;
  return SPIS_OK;
}

// Row patterns of a HOST CSR matrix (host_find_patterns) -> pattern storage on the device, under the acceptance rules of
// try_pattern_storage.  *ok_out = 0: not applicable, nothing allocated, the caller uploads the CSR arrays.
static int host_pattern_upload(spis_ctx* ctx, Matrix& M, cudaStream_t s, const int32_t* indptr, const int32_t* cols,
                               const double* vals, int* ok_out) {
  *ok_out = 0;
  PhaseTrace pt;
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 16) nt = 16;
  if (ctx->host_threads > 0) nt = (unsigned)ctx->host_threads;
  else if (tl_use_aux && nt > 4) nt /= 2;          // helper thread: leave cores to the thread that feeds the GPU
  const int32_t shift = (ctx->n_halo > 0 && ctx->hoff != ctx->n) ? (int32_t)(ctx->hoff - ctx->n) : 0;
  HostPatterns hp;
  void* pinned = nullptr;
  if (spis_pinned_alloc((size_t)M.nrows * sizeof(uint16_t), &pinned) != SPIS_OK) { ctx->err[0] = 0; return SPIS_OK; }
  hp.pid = static_cast<uint16_t*>(pinned);
  pt.mark("host patterns: id buffer");
  if (!host_find_patterns(indptr, cols, vals, M.nrows, (int32_t)ctx->n, shift, nt, hp, ctx->fmt_pref != SPIS_FMT_PATTERN)) {
    spis_pinned_free(pinned); pt.mark("host patterns: not applicable"); return SPIS_OK;
  }
  pt.mark("host patterns: found");
  const int npat = hp.npat, maxlen = hp.maxlen;
  const int W = maxlen < 4 ? 4 : (maxlen + 3) / 4 * 4;
  const bool forced = ctx->fmt_pref == SPIS_FMT_PATTERN;
  if (npat > kMaxPatterns || maxlen > kMaxPatternWidth ||
      (!forced && ((double)npat * W * 8.0 > (double)M.nnz || (double)hp.changes * 8.0 > (double)M.nrows))) {
    spis_pinned_free(hp.pid);
    return SPIS_OK;
  }
  std::vector<int32_t> len(npat), off((size_t)npat * W, 0);
  std::vector<double> val((size_t)npat * W, 0.0);
  for (int p = 0; p < npat; ++p) {
    const int32_t r = hp.rep[p], p0 = indptr[r];
    len[p] = indptr[r + 1] - p0;
    for (int k = 0; k < len[p]; ++k) {
      int32_t c = cols[p0 + k]; if (c >= (int32_t)ctx->n) c += shift;
      off[(size_t)p * W + k] = c - r;
      val[(size_t)p * W + k] = vals[p0 + k];
    }
  }
  auto drop = [&]() { dfree(ctx, M.pid); dfree(ctx, M.tab_len); dfree(ctx, M.tab_off); dfree(ctx, M.tab_val); spis_pinned_free(hp.pid); };
  int rc = dalloc(ctx, &M.pid, (size_t)M.nrows, false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.tab_len, (size_t)npat, false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.tab_off, (size_t)npat * W, false);
  if (rc == SPIS_OK) rc = dalloc(ctx, &M.tab_val, (size_t)npat * W, false);
  if (rc != SPIS_OK) { drop(); return rc; }
  pt.mark("host patterns: device blocks");
  cudaError_t e = cudaMemcpyAsync(M.pid, hp.pid, (size_t)M.nrows * sizeof(uint16_t), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(M.tab_len, len.data(), (size_t)npat * sizeof(int32_t), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(M.tab_off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(M.tab_val, val.data(), val.size() * sizeof(double), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);           // the staging arrays must outlive the copies
  if (e != cudaSuccess) { drop(); return fail(ctx, SPIS_E_CUDA, "pattern upload failed: %s", cudaGetErrorString(e)); }
  spis_pinned_free(hp.pid);
  g_h2d_bytes += (long long)((size_t)M.nrows * sizeof(uint16_t) + (size_t)npat * (4 + (size_t)W * 12));
  M.npat = npat; M.patW = W;
  pt.mark("host patterns: uploaded");
  *ok_out = 1;
  return SPIS_OK;
}

int spis_upload_csr(spis_ctx* ctx, int slot, int64_t nrows, int64_t ncols, int64_t nnz,
                    const int32_t* indptr, const int32_t* indices, const double* data) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS, "slot %d out of range", slot);
  REQUIRE(nrows == ctx->n, "matrix has %lld rows, context owns %lld", (long long)nrows, (long long)ctx->n);
  REQUIRE(ncols >= 0 && ncols <= ctx->n + ctx->n_halo, "matrix has %lld columns, context allows %lld", (long long)ncols, (long long)(ctx->n + ctx->n_halo));
  REQUIRE(nnz >= 0 && nnz < (int64_t)INT32_MAX, "nnz %lld must fit int32 row pointers", (long long)nnz);
  REQUIRE(indptr && (nnz == 0 || (indices && data)), "null CSR arrays");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = up_stream(ctx);
  Matrix& M = ctx->mats[slot];
  free_matrix(ctx, M);
  M.nrows = nrows; M.ncols = ncols; M.nnz = nnz;
  // Row patterns found by host threads in the caller's arrays: 2 bytes per row cross PCIe instead of 12 per entry
  if (ctx->host_pattern && nnz >= ctx->host_pattern_min_nnz &&
      (ctx->fmt_pref == SPIS_FMT_PATTERN || (ctx->fmt_pref == SPIS_FMT_AUTO && ctx->auto_pattern))) {
    int ok = 0;
    TRY(host_pattern_upload(ctx, M, s, indptr, indices, data, &ok));
    if (ok) {
      M.fmt = SPIS_FMT_PATTERN;
      M.nnz_padded = nnz;
      TRY(try_field_windows(ctx, M, s));
      M.present = true;
      return SPIS_OK;
    }
  }
  TRY(dalloc(ctx, &M.indptr, (size_t)nrows + 1, false));
  TRY(dalloc(ctx, &M.cols, (size_t)nnz, false));
  TRY(dalloc(ctx, &M.vals, (size_t)nnz, false));
  CU(cudaMemcpyAsync(M.indptr, indptr, ((size_t)nrows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  g_h2d_bytes += (long long)(((size_t)nrows + 1) * sizeof(int32_t) + (size_t)nnz * 12);
  if (nnz) {
    CU(cudaMemcpyAsync(M.cols, indices, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(M.vals, data, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  if (ctx->n_halo > 0 && ctx->hoff != ctx->n && nnz)
    remap_cols_kernel<<<ctx->nsm * 8, 256, 0, s>>>(M.cols, nnz, (int32_t)ctx->n, (int32_t)(ctx->hoff - ctx->n));
  if (ctx->fmt_pref == SPIS_FMT_PATTERN || (ctx->fmt_pref == SPIS_FMT_AUTO && ctx->auto_pattern)) {
    int ok = 0;
    TRY(try_pattern_storage(ctx, M, s, &ok));
    if (ok) {
      dfree(ctx, M.indptr); dfree(ctx, M.cols); dfree(ctx, M.vals);
      M.fmt = SPIS_FMT_PATTERN;
      M.nnz_padded = nnz;
      TRY(try_field_windows(ctx, M, s));
      M.present = true;
      return SPIS_OK;
    }
    REQUIRE(ctx->fmt_pref != SPIS_FMT_PATTERN, "matrix in slot %d has too many distinct row patterns for spmv_format=pattern", slot);
  }
  // SELL-32 slice widths -> offsets (tiny scan on the host)
  const int64_t nslices = (nrows + 31) / 32;
  int32_t* d_width = nullptr;
  TRY(dalloc(ctx, &d_width, (size_t)nslices, false));
  const int cgrid = (int)((nslices * 32 + 255) / 256);
  sell_width_kernel<<<cgrid, 256, 0, s>>>(M.indptr, nrows, d_width);
  CU(cudaGetLastError());
  std::vector<int32_t> width((size_t)nslices);
  CU(cudaMemcpyAsync(width.data(), d_width, (size_t)nslices * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  // SELL-C-sigma: rows sorted by length inside windows of 256 -- kept only if it removes >= 5 % of the stored entries
  if (ctx->sell_sigma && nnz > 0 && ctx->fmt_pref != SPIS_FMT_SELL2 && !(ctx->fmt_pref == SPIS_FMT_AUTO && ctx->auto_sell2) && ctx->fmt_pref != SPIS_FMT_CSR) {
    const int64_t nwin = (nrows + kSigma - 1) / kSigma;
    uint8_t* perm = nullptr;
    TRY(dalloc(ctx, &perm, (size_t)nwin * kSigma, false));
    sell_sigma_kernel<<<(unsigned)nwin, kSigma, 0, s>>>(M.indptr, nrows, perm, d_width);
    CU(cudaGetLastError());
    std::vector<int32_t> wsorted((size_t)nslices);
    CU(cudaMemcpyAsync(wsorted.data(), d_width, (size_t)nslices * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    int64_t tot = 0, tots = 0;
    for (int64_t q = 0; q < nslices; ++q) { tot += width[q]; tots += wsorted[q]; }
    if ((double)tots <= 0.95 * (double)tot) { width.swap(wsorted); M.rowperm = perm; }
    else dfree(ctx, perm);
  }
  dfree(ctx, d_width);
  std::vector<int64_t> off((size_t)nslices + 1);
  off[0] = 0;
  int fmt = ctx->fmt_pref;
  const bool want2 = fmt == SPIS_FMT_SELL2 || (fmt == SPIS_FMT_AUTO && ctx->auto_sell2);
  for (int64_t s = 0; s < nslices; ++s) off[s + 1] = off[s] + (int64_t)(want2 ? (width[s] + 1) & ~1 : width[s]) * 32;
  M.nnz_padded = off[nslices];
  if (fmt == SPIS_FMT_AUTO) fmt = ((double)M.nnz_padded <= 1.25 * (double)(nnz > 0 ? nnz : 1) + 32.0 * 64.0) ? (want2 ? SPIS_FMT_SELL2 : SPIS_FMT_SELL) : SPIS_FMT_CSR;
  if (fmt == SPIS_FMT_SELLD) fmt = SPIS_FMT_SELL;      // built as SELL first, values coded afterwards
  M.fmt = fmt;
  if (fmt == SPIS_FMT_SELL || fmt == SPIS_FMT_SELL2) {
    TRY(dalloc(ctx, &M.slice_off, (size_t)nslices + 1, false));
    TRY(dalloc(ctx, &M.scols, (size_t)M.nnz_padded, false));
    TRY(dalloc(ctx, &M.svals, (size_t)M.nnz_padded, false));
    CU(cudaMemcpyAsync(M.slice_off, off.data(), ((size_t)nslices + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    if (fmt == SPIS_FMT_SELL2)
      sell2_fill_kernel<<<cgrid, 256, 0, s>>>(M.indptr, M.cols, M.vals, nrows, M.slice_off, M.scols, M.svals);
    else
      sell_fill_kernel<<<cgrid, 256, 0, s>>>(M.indptr, M.cols, M.vals, nrows, M.slice_off, M.rowperm, M.scols, M.svals);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));
    dfree(ctx, M.indptr); dfree(ctx, M.cols); dfree(ctx, M.vals);
    if (fmt == SPIS_FMT_SELL && (ctx->fmt_pref == SPIS_FMT_SELLD || (ctx->fmt_pref == SPIS_FMT_AUTO && ctx->auto_dict))) {
      int ok = 0;
      TRY(try_value_dictionary(ctx, M, s, off, &ok));
      REQUIRE(ok || ctx->fmt_pref != SPIS_FMT_SELLD, "matrix in slot %d has more than 256 distinct values: spmv_format=selld does not apply", slot);
    }
    if (fmt == SPIS_FMT_SELL && ctx->spmv_sellw && !M.rowperm && M.nnz_padded > 0) TRY(try_sellw(ctx, M, s));
  } else {
    const double avg = nrows ? (double)nnz / (double)nrows : 0.0;
    M.csr_lanes = avg <= 3 ? 2 : avg <= 6 ? 4 : avg <= 12 ? 8 : avg <= 24 ? 16 : 32;
    CU(cudaStreamSynchronize(s));
  }
  M.present = true;
  return SPIS_OK;
}

int spis_upload_vec(spis_ctx* ctx, int which, const double* host, int64_t n) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(host, "host pointer is null");
  REQUIRE(n == ctx->n, "vector length %lld != n %lld", (long long)n, (long long)ctx->n);
  CU(cudaSetDevice(ctx->device));
  double* dst = nullptr;
  switch (which) {
    case SPIS_VEC_B: dst = ctx->B; break;
    case SPIS_VEC_X0: dst = ctx->X0; break;
    case SPIS_VEC_PRE_DIAG:
      if (!ctx->pre_diag) TRY(dalloc(ctx, &ctx->pre_diag, (size_t)ctx->ld));
      dst = ctx->pre_diag; break;
    default: return fail(ctx, SPIS_E_INVALID, "vector id %d cannot be uploaded", which);
  }
  return h2d(ctx, dst, host, (size_t)n * sizeof(double));
}

int spis_upload_blocks(spis_ctx* ctx, int bs, int64_t nblk, int64_t sb, int64_t sf, const double* blocks) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(bs >= 1 && bs <= 8, "block size %d not supported (1..8)", bs);
  REQUIRE(blocks && nblk > 0, "null blocks");
  REQUIRE((nblk - 1) * sb + (bs - 1) * sf < ctx->n, "block index map exceeds n");
  CU(cudaSetDevice(ctx->device));
  dfree(ctx, ctx->pre_blocks);
  TRY(dalloc(ctx, &ctx->pre_blocks, (size_t)bs * bs * nblk, false));
  // host layout [i][r][c] -> device structure-of-arrays [(r*bs+c)][i]
  std::vector<double> soa((size_t)bs * bs * nblk);
  for (int64_t i = 0; i < nblk; ++i)
    for (int rc = 0; rc < bs * bs; ++rc) soa[(size_t)rc * nblk + i] = blocks[(size_t)i * bs * bs + rc];
  TRY(h2d(ctx, ctx->pre_blocks, soa.data(), soa.size() * sizeof(double)));
  ctx->pre_bs = bs; ctx->pre_nblk = nblk; ctx->pre_sb = sb; ctx->pre_sf = sf;
  return SPIS_OK;
}

int spis_set_precond(spis_ctx* ctx, int kind) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(kind >= SPIS_PRE_NONE && kind <= SPIS_PRE_HOST, "unknown preconditioner kind %d", kind);
  CU(cudaSetDevice(ctx->device));
  ctx->pre_kind = kind;
  ctx->began = false;
  return ensure_Z(ctx);
}

// ---- Krylov loop ----------------------------------------------------------------------
int spis_solve_begin(spis_ctx* ctx, double* beta_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(beta_out, "beta_out is null");
  CU(cudaSetDevice(ctx->device));
  TRY(ensure_Z(ctx));
  double* scal = ctx->d_small + 2 * ctx->K;
  TRY(do_halo(ctx, ctx->X0));
  TRY(launch_spmv(ctx, SPIS_SLOT_A, 1, ctx->X0, ctx->B, ctx->R0, scal + 1));
  TRY(prof_begin(ctx, SPIS_PROF_OTHER, 16.0 * (double)ctx->n));
  CU(cudaMemcpyAsync(ctx->V, ctx->R0, (size_t)ctx->hoff * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  TRY(prof_end(ctx));
  ctx->z_ready_index = -1;
  const bool fuse = ctx->pre_kind == SPIS_PRE_JACOBI && ctx->fuse_jacobi && ctx->pre_diag;
  TRY(launch_scale(ctx, ctx->V, scal + 1, fuse ? ctx->pre_diag : nullptr, fuse ? ctx->Z : nullptr));
  if (fuse) ctx->z_ready_index = 0;
  CU(cudaMemcpyAsync(ctx->h_small + 2 * ctx->K, scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *beta_out = std::sqrt(ctx->h_small[2 * ctx->K + 1]);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (auto& c : ctx->cons) { c.cols_done = 0; c.term0_done = false; }
  }
  ctx->began = true;
  ctx->arnoldi_inflight = -1;
  ctx->arnoldi_part1 = -1;
  ctx->resid_inflight = false;
  ctx->pipe_ready = false;
  return SPIS_OK;
}

// First half of Arnoldi step j: z_j, w = A z_j and (CGS2) h1 = V^T w, w -= V h1, h2 = V^T w.
// with_residual: the SpMV also measures ||A x - b|| of the iterate sitting in the X buffer (one pass over A for both).
static int arnoldi_begin_impl(spis_ctx* ctx, int j, bool with_residual) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->began, "spis_solve_begin has not been called");
  REQUIRE(j >= 0 && j < ctx->kmax, "Arnoldi index %d out of range [0,%d)", j, ctx->kmax);
  // (a FINISHED step whose column has not been collected yet may still be in flight: its copy into the pinned
  //  mirror is queued before anything this call launches; spis_arnoldi_finish is where the mirror is reused)
  REQUIRE(ctx->arnoldi_inflight < 0 || ctx->arnoldi_inflight == j - 1, "Arnoldi step %d is still in flight", ctx->arnoldi_inflight);
  REQUIRE(ctx->arnoldi_part1 < 0, "the first half of Arnoldi step %d is queued and was never finished", ctx->arnoldi_part1);
  CU(cudaSetDevice(ctx->device));
  const int m = j + 1;
  const size_t ld = (size_t)ctx->ld;
  double* qj = ctx->V + (size_t)j * ld;
  double* qn = ctx->V + (size_t)(j + 1) * ld;
  double* zj = zbase(ctx) + (size_t)j * ld;
  double* h1 = ctx->d_small;
  double* h2 = ctx->d_small + ctx->K;
  double* scal = ctx->d_small + 2 * ctx->K;
  // z[j] = P q[j]                                                  (solvers.py:190)
  if (ctx->pre_kind != SPIS_PRE_NONE && ctx->pre_kind != SPIS_PRE_HOST && ctx->z_ready_index != j)
    TRY(launch_precond(ctx, qj, zj));
  // w = A z[j]                                                     (solvers.py:191)
  TRY(do_halo(ctx, zj));
  if (with_residual) {
    REQUIRE(!ctx->resid_inflight, "an iterate/residual pair is still in flight");
    TRY(do_halo(ctx, ctx->X));
    TRY(launch_spmv_dual(ctx, zj, ctx->W, ctx->X, ctx->B, scal + 2));       // + ||A x - b||  (solvers.py:290)
    CU(cudaMemcpyAsync(ctx->h_resid, scal + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev_resid, ctx->stream));
    ctx->resid_inflight = true;
  } else {
    TRY(launch_spmv(ctx, SPIS_SLOT_A, 0, zj, nullptr, ctx->W, nullptr));
  }
  // orthogonalise w against q[0..j]                                (solvers.py:193-196)
  if (ctx->orth == SPIS_ORTH_MGS) {
    CU(cudaMemsetAsync(h2, 0, (size_t)ctx->K * sizeof(double), ctx->stream));
    for (int i = 0; i < m; ++i) {
      const double* qi = ctx->V + (size_t)i * ld;
      TRY(launch_mdot(ctx, qi, 1, nullptr, 0, ctx->W, h1 + i));
      const bool last = (i == m - 1);
      TRY(launch_lincomb(ctx, qi, 1, h1 + i, nullptr, -1.0, ctx->W, last ? qn : ctx->W, last ? 1 : 0, scal));
    }
  } else if (ctx->orth == SPIS_ORTH_CGS1) {
    CU(cudaMemsetAsync(h2, 0, (size_t)ctx->K * sizeof(double), ctx->stream));
    TRY(launch_mdot(ctx, ctx->V, m, nullptr, 0, ctx->W, h1));
    TRY(launch_lincomb(ctx, ctx->V, m, h1, nullptr, -1.0, ctx->W, qn, 1, scal));
  } else {
    TRY(launch_mdot(ctx, ctx->V, m, nullptr, 0, ctx->W, h1));
    int E_ = 0, st_ = 0, mb_ = 0; size_t sm_ = 0;
    if (ctx->orth_fused && orth_mid_plan(ctx, m, &E_, &st_, &mb_, &sm_)) {
      TRY(launch_orth_mid(ctx, ctx->V, m, h1, ctx->W, h2));           // one pass over V for both
    } else {
      TRY(launch_lincomb(ctx, ctx->V, m, h1, nullptr, -1.0, ctx->W, ctx->W, 0, nullptr));
      TRY(launch_mdot(ctx, ctx->V, m, nullptr, 0, ctx->W, h2));
    }
  }
  ctx->arnoldi_part1 = j;
  return SPIS_OK;
}

int spis_arnoldi_begin(spis_ctx* ctx, int j) { return arnoldi_begin_impl(ctx, j, false); }
int spis_arnoldi_begin_residual(spis_ctx* ctx, int j) { return arnoldi_begin_impl(ctx, j, true); }

// Second half: (CGS2) w -= V h2 -> q[j+1] with its norm, normalisation, Hessenberg column to the host.
// With m_it > 0 the same sweep over the basis also forms the iterate x = x0 + Z[:, :m_it] y of the PREVIOUS
// step (solvers.py:287) -- only without a preconditioner (Z is V) and m_it <= j + 1.
int spis_arnoldi_finish(spis_ctx* ctx, int j, int m_it, const double* y_it) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->arnoldi_part1 == j, "spis_arnoldi_begin(%d) has not been called (queued: %d)", j, ctx->arnoldi_part1);
  REQUIRE(ctx->arnoldi_inflight < 0, "the Hessenberg column of Arnoldi step %d has not been collected", ctx->arnoldi_inflight);
  CU(cudaSetDevice(ctx->device));
  const int m = j + 1;
  const size_t ld = (size_t)ctx->ld;
  double* qn = ctx->V + (size_t)(j + 1) * ld;
  double* h2 = ctx->d_small + ctx->K;
  double* scal = ctx->d_small + 2 * ctx->K;
  if (m_it > 0) {
    REQUIRE(y_it && m_it <= m, "bad iterate arguments (m_it=%d, step %d)", m_it, j);
    REQUIRE(ctx->orth == SPIS_ORTH_CGS2 && ctx->pre_kind == SPIS_PRE_NONE && ctx->fuse_iterate,
            "the fused iterate needs CGS2 and no preconditioner");
    memcpy(ctx->h_y, y_it, (size_t)m_it * sizeof(double));
    CU(cudaMemcpyAsync(ctx->d_y, ctx->h_y, (size_t)m_it * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRY(dl_fence(ctx));
    TRY(launch_lincomb2(ctx, ctx->V, m, h2, ctx->d_y, m_it, ctx->W, ctx->x0_is_zero ? nullptr : ctx->X0, qn, ctx->X, scal));
  } else if (ctx->orth == SPIS_ORTH_CGS2) {
    TRY(launch_lincomb(ctx, ctx->V, m, h2, nullptr, -1.0, ctx->W, qn, 1, scal));
  }
  // q[j+1] = w / ||w||                                             (solvers.py:196-198)
  const bool fuse = ctx->pre_kind == SPIS_PRE_JACOBI && ctx->fuse_jacobi && ctx->pre_diag && (j + 1 < ctx->kmax);
  TRY(launch_scale(ctx, qn, scal, fuse ? ctx->pre_diag : nullptr, fuse ? ctx->Z + (size_t)(j + 1) * ld : nullptr));
  ctx->z_ready_index = fuse ? j + 1 : -1;
  CU(cudaMemcpyAsync(ctx->h_small, ctx->d_small, ((size_t)2 * ctx->K + 8) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->xactive)   // peer-timeout word of the NVLink collectives rides along (slot scal[7] is unused)
    CU(cudaMemcpyAsync(ctx->h_small + 2 * ctx->K + 7, ctx->xbuf + ctx->xv.flags_off() + 4 * ctx->xv.world, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaEventRecord(ctx->ev_arnoldi, ctx->stream));
  ctx->arnoldi_part1 = -1;
  ctx->arnoldi_inflight = j;
  return SPIS_OK;
}

int spis_arnoldi_launch(spis_ctx* ctx, int j) {
  TRY(spis_arnoldi_begin(ctx, j));
  return spis_arnoldi_finish(ctx, j, 0, nullptr);
}

// ||A x - b|| of the iterate that is already in the X buffer (formed by spis_arnoldi_finish)
int spis_residual_launch(spis_ctx* ctx) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->began, "spis_solve_begin has not been called");
  REQUIRE(!ctx->resid_inflight, "an iterate/residual pair is still in flight");
  CU(cudaSetDevice(ctx->device));
  double* scal = ctx->d_small + 2 * ctx->K;
  TRY(do_halo(ctx, ctx->X));
  TRY(launch_spmv(ctx, SPIS_SLOT_A, 2, ctx->X, ctx->B, nullptr, scal + 2));
  CU(cudaMemcpyAsync(ctx->h_resid, scal + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaEventRecord(ctx->ev_resid, ctx->stream));
  ctx->resid_inflight = true;
  return SPIS_OK;
}

int spis_arnoldi_wait(spis_ctx* ctx, int j, double* hcol_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(hcol_out, "hcol_out is null");
  REQUIRE(ctx->arnoldi_inflight == j, "Arnoldi step %d was not launched (in flight: %d)", j, ctx->arnoldi_inflight);
  CU(cudaEventSynchronize(ctx->ev_arnoldi));
  if (ctx->xactive) {
    unsigned long long w; memcpy(&w, ctx->h_small + 2 * ctx->K + 7, sizeof(w));
    if (w >> 63) return fail(ctx, SPIS_E_CUDA, "NVLink collective %llu timed out waiting for a peer rank", w & ~(1ull << 63));
  }
  const int m = j + 1;
  for (int i = 0; i < m; ++i) hcol_out[i] = ctx->h_small[i] + ctx->h_small[ctx->K + i];
  hcol_out[m] = std::sqrt(ctx->h_small[2 * ctx->K]);
  ctx->arnoldi_inflight = -1;
  return SPIS_OK;
}

int spis_arnoldi_step(spis_ctx* ctx, int j, double* hcol_out) {
  TRY(spis_arnoldi_launch(ctx, j));
  return spis_arnoldi_wait(ctx, j, hcol_out);
}

static int form_iterate_impl(spis_ctx* ctx, int m, const double* y, double* dl_dst = nullptr, int chunks = 1) {
  REQUIRE(ctx->began, "spis_solve_begin has not been called");
  REQUIRE(m >= 0 && m <= ctx->kmax && (y || m == 0), "bad iterate arguments (m=%d)", m);
  CU(cudaSetDevice(ctx->device));
  TRY(dl_fence(ctx));
  if (m) {
    // h_y may still be the source of an earlier async copy only if the stream has not
    // drained; every path that uses it synchronises before returning, so it is free here.
    memcpy(ctx->h_y, y, (size_t)m * sizeof(double));
    CU(cudaMemcpyAsync(ctx->d_y, ctx->h_y, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  // x_j = x0 + Z y                                                 (solvers.py:287)
  if (!dl_dst) {
    TRY(launch_lincomb(ctx, zbase(ctx), m, ctx->d_y, nullptr, 1.0, ctx->x0_is_zero ? nullptr : ctx->X0, ctx->X, 0, nullptr));
    return SPIS_OK;
  }
  // the same sweep in row chunks (whole tiles, so every element sees the identical fma chain), each chunk handed to
  // the copy stream as soon as it is complete: the 8n-byte download of a final candidate overlaps its own formation
  // and the residual check instead of following them
  const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
  if (chunks < 1) chunks = 1;
  if ((int64_t)chunks > ntiles) chunks = (int)ntiles;
  const int64_t per = (ntiles + chunks - 1) / chunks * kTile;
  for (int64_t r0 = 0; r0 < ctx->n; r0 += per) {
    const int64_t nr = std::min<int64_t>(per, ctx->n - r0);
    const int64_t nt = (nr + kTile - 1) / kTile;
    const int grid = grid_for(ctx, nt, ctx->ctas_per_sm);
    const size_t smem = (size_t)(m + 2 + kWarps * 32) * sizeof(double);
    TRY(prof_begin(ctx, SPIS_PROF_LINCOMB, (double)(m + (ctx->x0_is_zero ? 0 : 1) + 1) * 8.0 * (double)nr));
    lincomb_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(zbase(ctx) + r0, ctx->ld, m, ctx->d_y, nullptr, 1.0,
                                                             ctx->x0_is_zero ? nullptr : ctx->X0 + r0, ctx->X + r0, nr, 0,
                                                             ctx->d_partial, ctx->d_counter, nullptr, XView(), 0ull);
    CU(cudaGetLastError());
    TRY(prof_end(ctx));
    CU(cudaEventRecord(ctx->ev_chunk, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->dstream, ctx->ev_chunk, 0));
    CU(cudaMemcpyAsync(dl_dst + r0, ctx->X + r0, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, ctx->dstream));
  }
  CU(cudaEventRecord(ctx->ev_dl, ctx->dstream));
  ctx->dl_inflight = true;
  return SPIS_OK;
}

int spis_form_iterate(spis_ctx* ctx, int m, const double* y) {
  if (!ctx) return SPIS_E_INVALID;
  TRY(form_iterate_impl(ctx, m, y));
  CU(cudaStreamSynchronize(ctx->stream));   // h_y staging is free again when this returns
  return SPIS_OK;
}

static int iterate_residual_launch_impl(spis_ctx* ctx, int m, const double* y, double* dl_dst, int chunks);
int spis_iterate_residual_launch(spis_ctx* ctx, int m, const double* y) {
  if (!ctx) return SPIS_E_INVALID;
  return iterate_residual_launch_impl(ctx, m, y, nullptr, 1);
}

// As spis_iterate_residual_launch, and x_j is also streamed to host_dst (page-locked, n doubles) while it is formed and
// checked: for an iterate the caller expects to be the last one (solvers.py:296-297 ends the loop on its residual).
// *started_out = 0: host_dst is not page-locked, nothing was copied (plain launch).  spis_download_join waits for the
// copy; until then host_dst must stay alive.  If the iterate is NOT the last one the caller joins and drops host_dst.
int spis_iterate_residual_launch_dl(spis_ctx* ctx, int m, const double* y, double* host_dst, int chunks, int* started_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(host_dst && started_out, "null argument");
  const bool ok = is_pinned_host(host_dst) && m > 0;
  *started_out = ok ? 1 : 0;
  return iterate_residual_launch_impl(ctx, m, y, ok ? host_dst : nullptr, chunks);
}

int spis_download_join(spis_ctx* ctx) {
  if (!ctx) return SPIS_E_INVALID;
  if (!ctx->dl_inflight) return SPIS_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventSynchronize(ctx->ev_dl));
  ctx->dl_inflight = false;
  return SPIS_OK;
}

static int iterate_residual_launch_impl(spis_ctx* ctx, int m, const double* y, double* dl_dst, int chunks) {
  REQUIRE(!ctx->resid_inflight, "an iterate/residual pair is still in flight");
  if (dl_dst && ctx->dl_inflight) TRY(spis_download_join(ctx));     // one early download at a time
  // chunks == 0: the copy is queued BEHIND the residual check (it starts when the check ends, without waiting for
  // the host to look at the result).  Row-sharded runs use this form: while a device-to-host copy is in flight the
  // peers' NVLink stores into this GPU are held up with it (measured: the 20-us residual check of a 4-GPU run took
  // 260 us, about the length of the copy, when the copy ran beside it), and the check ends in an exchange.
  const bool after = dl_dst && chunks == 0;
  TRY(form_iterate_impl(ctx, m, y, after ? nullptr : dl_dst, chunks));
  double* scal = ctx->d_small + 2 * ctx->K;
  // ||A x_j - b||                                                  (solvers.py:290)
  TRY(do_halo(ctx, ctx->X));
  TRY(launch_spmv(ctx, SPIS_SLOT_A, 2, ctx->X, ctx->B, nullptr, scal + 2));
  CU(cudaMemcpyAsync(ctx->h_resid, scal + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaEventRecord(ctx->ev_resid, ctx->stream));
  ctx->resid_inflight = true;
  if (after) {
    CU(cudaStreamWaitEvent(ctx->dstream, ctx->ev_resid, 0));
    CU(cudaMemcpyAsync(dl_dst, ctx->X, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->dstream));
    CU(cudaEventRecord(ctx->ev_dl, ctx->dstream));
    ctx->dl_inflight = true;
  }
  return SPIS_OK;
}

int spis_iterate_residual_wait(spis_ctx* ctx, double* resnorm_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(resnorm_out, "resnorm_out is null");
  REQUIRE(ctx->resid_inflight, "no iterate/residual pair in flight");
  CU(cudaEventSynchronize(ctx->ev_resid));      // later work queued on the stream (the next Arnoldi step) is not waited for
  ctx->resid_inflight = false;
  *resnorm_out = std::sqrt(*ctx->h_resid);
  return SPIS_OK;
}

int spis_iterate_residual(spis_ctx* ctx, int m, const double* y, double* resnorm_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(resnorm_out, "resnorm_out is null");
  TRY(spis_iterate_residual_launch(ctx, m, y));
  return spis_iterate_residual_wait(ctx, resnorm_out);
}

// ---- pipelined Krylov loop -------------------------------------------------------------
// The host queues whole Arnoldi steps ahead of time; everything an UNCONSTRAINED iteration needs (solvers.py:190-198,
// 231-235, 287, 290) is computed on the device by kernels that were queued before their inputs existed:
//   [halo]  one exchange for z_j and the iterate                                            (row-sharded runs)
//   SpMV    w = A z_j and, in the same pass, ||A x - b||^2 of the iterate in X               (dual kernels)
//   mdot    h1 = V^T w; its tail also finishes the residual norm, publishes it and flips the phase word
//   orth    w' = w - V h1, h2 = V^T w', ||w'||^2                                              (one staged pass)
//   hess    column j = h1 + h2, h[j+1,j]^2 = ||w'||^2 - |h2|^2, Givens update, y_j = argmin |beta e1 - H y|
//   sweep   q[j+1] = (w' - V h2) / h[j+1,j]  and  x_{j-1} = x0 + Z y_{j-1}                    (lincomb2n_kernel)
// i.e. four kernels and no host round trip per iteration (six kernels, a D2H, a host solve and an H2D before), and
// three cross-GPU exchanges instead of six.  The host follows through records in mapped page-locked memory.
namespace {
constexpr int kRecSlots = 8;

inline double* step_rec_host(spis_ctx* ctx, int j) { return ctx->h_rec + (size_t)(j % kRecSlots) * ctx->rec_stride; }
inline double* step_rec_dev(spis_ctx* ctx, int j) { return ctx->d_rec + (size_t)(j % kRecSlots) * ctx->rec_stride; }
inline double* res_rec_host(spis_ctx* ctx, long long t) { return ctx->h_rec + (size_t)kRecSlots * ctx->rec_stride + (size_t)(t % kRecSlots) * 8; }
inline double* res_rec_dev(spis_ctx* ctx, long long t) { return ctx->d_rec + (size_t)kRecSlots * ctx->rec_stride + (size_t)(t % kRecSlots) * 8; }

// spin until the device has published sequence word `want` into a mapped record
int wait_record(spis_ctx* ctx, const double* rec, unsigned long long want, const char* what) {
  const volatile unsigned long long* w = reinterpret_cast<const volatile unsigned long long*>(rec);
  const auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  while (*w != want) {
    if ((++spins & 0xfffu) == 0) {
      const cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady)
        return fail(ctx, SPIS_E_CUDA, "waiting for %s: %s", what, cudaGetErrorString(e));
      if (e == cudaSuccess && *w != want)
        return fail(ctx, SPIS_E_INVALID, "%s was never produced (the stream is idle)", what);
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0)
        return fail(ctx, SPIS_E_CUDA, "timed out waiting for %s", what);
    }
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return SPIS_OK;
}
}  // namespace

// Device side of the loop after spis_solve_begin: Givens state (g = beta e1), phase word (phase0 = 1: the very first
// iteration is already constrained, i.e. beta <= contol*tol), threshold thr on the residual NORM below which the
// device stops forming unconstrained iterates.
int spis_pipe_begin(spis_ctx* ctx, double thr, int phase0) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->began, "spis_solve_begin has not been called");
  REQUIRE(ctx->orth == SPIS_ORTH_CGS2, "the pipelined loop needs CGS2");
  REQUIRE(ctx->pre_kind != SPIS_PRE_HOST, "the pipelined loop cannot call a host preconditioner");
  REQUIRE(!ctx->allreduce && !ctx->halo, "the pipelined loop needs the peer-memory transport (or one GPU)");
  REQUIRE(ctx->arnoldi_inflight < 0 && ctx->arnoldi_part1 < 0 && !ctx->resid_inflight, "an Arnoldi step is in flight");
  CU(cudaSetDevice(ctx->device));
  const size_t km = (size_t)ctx->kmax;
  if (!ctx->hs_R) {
    TRY(dalloc(ctx, &ctx->hs_cs, km));
    TRY(dalloc(ctx, &ctx->hs_sn, km));
    TRY(dalloc(ctx, &ctx->hs_gv, km + 1));
    TRY(dalloc(ctx, &ctx->hs_R, km * km, false));
    TRY(dalloc(ctx, &ctx->hs_tracking, 4));
    TRY(dalloc(ctx, &ctx->d_phase, 4));
    TRY(dalloc(ctx, &ctx->d_ydev, (size_t)2 * ctx->K));
    TRY(dalloc(ctx, &ctx->d_rpartial, (size_t)ctx->max_grid));
  }
  if (!ctx->h_rec) {
    ctx->rec_stride = 2 * ctx->K + 8;
    const size_t bytes = ((size_t)kRecSlots * ctx->rec_stride + (size_t)kRecSlots * 8) * sizeof(double);
    if (spis_pinned_alloc(bytes, (void**)&ctx->h_rec) != SPIS_OK) return fail(ctx, SPIS_E_NOMEM, "pinned record buffer: %s", g_global_err);
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, ctx->h_rec, 0) != cudaSuccess) { cudaGetLastError(); dp = ctx->h_rec; }   // unified addressing: same pointer
    ctx->d_rec = static_cast<double*>(dp);
    memset(ctx->h_rec, 0, bytes);
  }
  {
    const size_t need = (size_t)(4 * ctx->K + 4 + 70 * 70) * sizeof(double);      // hess_kernel's dynamic shared memory at most
    REQUIRE(need <= 227 * 1024, "k_max = %d is too large for the pipelined loop", ctx->kmax);
    if (need > 48 * 1024) CU(cudaFuncSetAttribute(hess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
  }
  for (int i = 0; i < kRecSlots; ++i) { ctx->step_seq[i] = 0; ctx->res_seq[i] = 0; }
  ctx->res_tickets = 0;
  ctx->pipe_thr2 = thr * thr;
  HessState st{ctx->hs_cs, ctx->hs_sn, ctx->hs_gv, ctx->hs_R, ctx->hs_tracking, ctx->kmax};
  pipe_init_kernel<<<1, 32, 0, ctx->stream>>>(st, ctx->d_small + 2 * ctx->K + 1, ctx->d_phase, phase0 ? 1 : 0);
  CU(cudaGetLastError());
  ctx->prof_launch[SPIS_PROF_OTHER] += 1;
  ctx->pipe_ready = true;
  return SPIS_OK;
}

// Queue Arnoldi step j.  flags & 1: the SpMV also measures ||A x - b|| of the iterate in X (*ticket_out identifies the
// record, spis_resid_wait).  flags & 2: form an iterate with the device's least-squares coefficients if the phase word
// still says "unconstrained": x_{j-1} (from the same sweep, no preconditioner) or x_j (a sweep over Z, preconditioned).
int spis_step_enqueue(spis_ctx* ctx, int j, int flags, int64_t* ticket_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->pipe_ready, "spis_pipe_begin has not been called for this solve");
  REQUIRE(j >= 0 && j < ctx->kmax, "Arnoldi index %d out of range [0,%d)", j, ctx->kmax);
  REQUIRE(!ctx->resid_inflight || !(flags & 1), "an iterate/residual pair is still in flight");
  CU(cudaSetDevice(ctx->device));
  const int m = j + 1;
  const int K = ctx->K;
  const size_t ld = (size_t)ctx->ld;
  const bool nopre = ctx->pre_kind == SPIS_PRE_NONE;
  const bool want_res = (flags & 1) != 0, want_it = (flags & 2) != 0;
  double* qj = ctx->V + (size_t)j * ld;
  double* qn = ctx->V + (size_t)(j + 1) * ld;
  double* zj = zbase(ctx) + (size_t)j * ld;
  double* h1 = ctx->d_small;
  double* h2 = ctx->d_small + K;
  double* scal = ctx->d_small + 2 * K;
  if (ticket_out) *ticket_out = -1;
  // z[j] = P q[j]                                                  (solvers.py:190)
  if (!nopre && ctx->z_ready_index != j) TRY(launch_precond(ctx, qj, zj));
  TRY(do_halo2(ctx, zj, want_res ? ctx->X : nullptr));
  // w = A z[j]  (+ ||A x - b||^2)                                   (solvers.py:191, 290)
  TailExtra tx{};
  tx.reverse = ctx->sweep_reverse & 1;
  bool rides = false;
  if (want_res) {
    const long long t = ctx->res_tickets++;
    const unsigned long long sw = ctx->rec_counter++;
    ctx->res_seq[t % kRecSlots] = sw;
    *reinterpret_cast<volatile unsigned long long*>(res_rec_host(ctx, t)) = 0ull;
    tx.res_out = scal + 2; tx.host_rec = res_rec_dev(ctx, t); tx.rec_seq = sw; tx.phase = ctx->d_phase; tx.thr2 = ctx->pipe_thr2;
    int parts = 0;
    TRY(launch_spmv_dual(ctx, zj, ctx->W, ctx->X, ctx->B, scal + 2, ctx->d_rpartial, &parts));
    if (parts > 0) { tx.rpart = ctx->d_rpartial; tx.nrpart = parts; rides = true; }
    else {
      publish_res_kernel<<<1, 32, 0, ctx->stream>>>(scal + 2, tx);       // the norm is complete (and all-reduced) already
      CU(cudaGetLastError());
      ctx->prof_launch[SPIS_PROF_OTHER] += 1;
    }
    if (ticket_out) *ticket_out = (int64_t)t;
  } else {
    TRY(launch_spmv(ctx, SPIS_SLOT_A, 0, zj, nullptr, ctx->W, nullptr));
  }
  // CGS2: h1 = V^T w ; w' = w - V h1, h2 = V^T w', ||w'||^2        (solvers.py:193-196)
  TailExtra plain{};
  plain.reverse = tx.reverse;
  TRY(launch_mdot(ctx, ctx->V, m, nullptr, 0, ctx->W, h1, rides ? &tx : &plain));
  int E_ = 0, st_ = 0, mb_ = 0; size_t sm_ = 0;
  if (ctx->orth_fused && orth_mid_plan(ctx, m, &E_, &st_, &mb_, &sm_)) {
    TRY(launch_orth_mid(ctx, ctx->V, m, h1, ctx->W, h2, 1));
  } else {
    TRY(launch_lincomb(ctx, ctx->V, m, h1, nullptr, -1.0, ctx->W, ctx->W, 0, nullptr));
    TRY(launch_mdot(ctx, ctx->V, m, nullptr, 1, ctx->W, h2));              // h2[m] = w'.w'
  }
  // column j, h[j+1,j], Givens update, y_j                          (solvers.py:113 / 231-235)
  {
    const unsigned long long sw = ctx->rec_counter++;
    ctx->step_seq[j % kRecSlots] = sw;
    *reinterpret_cast<volatile unsigned long long*>(step_rec_host(ctx, j)) = 0ull;
    HessState st{ctx->hs_cs, ctx->hs_sn, ctx->hs_gv, ctx->hs_R, ctx->hs_tracking, ctx->kmax};
    // the leading block of R travels through shared memory while it fits 40 KB (m <= 64 and then some)
    const int rcache = m <= 70 ? (m + 1) / 2 * 2 : 0;
    const size_t smem = (size_t)(4 * K + 4 + rcache * rcache) * sizeof(double);
    const unsigned long long* errw = ctx->xactive ? reinterpret_cast<const unsigned long long*>(ctx->xbuf + ctx->xv.flags_off()) + 4 * ctx->xv.world : nullptr;
    // The sweep below needs nothing from this kernel (it forms h[j+1,j] itself, and takes y_{j-1} from the previous
    // step), so the ~10 us of sequential Givens arithmetic run beside it on a stream of their own; the main stream
    // joins again after the sweep, which orders everything queued later behind this step's state, y_j and phase word.
    cudaStream_t hs = ctx->hess_async ? ctx->hstream : ctx->stream;
    if (ctx->hess_async) { CU(cudaEventRecord(ctx->ev_orth, ctx->stream)); CU(cudaStreamWaitEvent(hs, ctx->ev_orth, 0)); }
    hess_kernel<<<1, 32, smem, hs>>>(j, st, h1, h2, scal, ctx->d_ydev + (size_t)(j & 1) * K, ctx->d_phase,
                                     step_rec_dev(ctx, j), sw, K, errw, rcache);
    CU(cudaGetLastError());
    if (ctx->hess_async) CU(cudaEventRecord(ctx->ev_hess, hs));
    ctx->prof_launch[SPIS_PROF_OTHER] += 1;
  }
  // q[j+1] = (w' - V h2) / h[j+1,j]  (+ x_{j-1} = x0 + Z y_{j-1})  (solvers.py:195-198, 287)
  if (want_it) TRY(dl_fence(ctx));
  {
    const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
    const int grid = grid_for(ctx, ntiles, ctx->lincomb2_ctas_per_sm);
    const size_t smem = (size_t)(2 * (m + 2)) * sizeof(double);
    const int mB = (nopre && want_it && j >= 1) ? j : 0;
    const bool fusej = ctx->pre_kind == SPIS_PRE_JACOBI && ctx->fuse_jacobi && ctx->pre_diag && (j + 1 < ctx->kmax);
    TRY(prof_begin(ctx, SPIS_PROF_LINCOMB, (double)(m + 2 + (mB ? 1 : 0) + ((mB && !ctx->x0_is_zero) ? 1 : 0) + (fusej ? 2 : 0)) * 8.0 * (double)ctx->n));
    lincomb2n_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(ctx->V, ctx->ld, m, h2, ctx->d_ydev + (size_t)((j + 1) & 1) * K, mB,
                                                               ctx->d_phase, ctx->W, ctx->x0_is_zero ? nullptr : ctx->X0, qn, ctx->X,
                                                               fusej ? ctx->pre_diag : nullptr, fusej ? ctx->Z + (size_t)(j + 1) * ld : nullptr, ctx->n,
                                                               (ctx->sweep_reverse >> 1) & 1);
    CU(cudaGetLastError());
    TRY(prof_end(ctx));
    ctx->z_ready_index = fusej ? j + 1 : -1;
  }
  if (ctx->hess_async) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_hess, 0));
  if (!nopre && want_it) {
    // preconditioned: Z is not V, the iterate x_j = x0 + Z y_j takes its own sweep (skipped once the phase word is set)
    const int64_t ntiles = (ctx->n + kTile - 1) / kTile;
    const int grid = grid_for(ctx, ntiles, ctx->ctas_per_sm);
    const size_t smem = (size_t)(m + 2 + kWarps * 32) * sizeof(double);
    TRY(prof_begin(ctx, SPIS_PROF_LINCOMB, (double)(m + (ctx->x0_is_zero ? 0 : 1) + 1) * 8.0 * (double)ctx->n));
    lincomb_kernel<4><<<grid, kThreads, smem, ctx->stream>>>(ctx->Z, ctx->ld, m, ctx->d_ydev + (size_t)(j & 1) * K, nullptr, 1.0,
                                                             ctx->x0_is_zero ? nullptr : ctx->X0, ctx->X, ctx->n, 0, ctx->d_partial,
                                                             ctx->d_counter, nullptr, XView(), 0ull, ctx->d_phase);
    CU(cudaGetLastError());
    TRY(prof_end(ctx));
  }
  return SPIS_OK;
}

// Blocks until step j's record has arrived.  col_out: h[0..j+1, j] (j + 2 doubles); y_out: argmin_y |beta e1 - H y|
// (j + 1 doubles, only meaningful when info_out[1] != 0); info_out[8]: [1] valid, [2] min_y |beta e1 - H y|,
// [3] h[j+1,j]^2, [4] ||w'||^2, [5] |h2|^2, [6] phase word when the step's last sweep ran.
int spis_step_wait(spis_ctx* ctx, int j, double* col_out, double* y_out, double* info_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->pipe_ready && j >= 0 && j < ctx->kmax && ctx->step_seq[j % kRecSlots] != 0, "step %d was not queued", j);
  REQUIRE(col_out && y_out && info_out, "null output");
  const double* rec = step_rec_host(ctx, j);
  TRY(wait_record(ctx, rec, ctx->step_seq[j % kRecSlots], "an Arnoldi step record"));
  const int K = ctx->K;
  for (int i = 0; i < 8; ++i) info_out[i] = rec[i];
  info_out[0] = (double)j;
  for (int i = 0; i <= j + 1; ++i) col_out[i] = rec[8 + i];
  for (int i = 0; i <= j; ++i) y_out[i] = rec[8 + K + i];
  if (rec[7] != 0.0) return fail(ctx, SPIS_E_CUDA, "an NVLink collective timed out waiting for a peer rank (before Arnoldi step %d)", j);
  return SPIS_OK;
}

// SQUARED residual norm measured by the step that returned `ticket` (the very number the device compared with thr^2);
// *go_out = 1 while the phase word still says "unconstrained"
int spis_resid_wait(spis_ctx* ctx, int64_t ticket, double* res_out, int* go_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->pipe_ready && ticket >= 0 && ticket < ctx->res_tickets && ticket + kRecSlots >= ctx->res_tickets,
          "residual ticket %lld is not pending", (long long)ticket);
  REQUIRE(res_out && go_out, "null output");
  const double* rec = res_rec_host(ctx, ticket);
  TRY(wait_record(ctx, rec, ctx->res_seq[ticket % kRecSlots], "a residual record"));
  *res_out = rec[1];
  *go_out = rec[2] == 0.0 ? 1 : 0;
  return SPIS_OK;
}

// ---- constraint stage -----------------------------------------------------------------
int spis_constraint_define(spis_ctx* ctx, int c, int mat_slot, const double* v, double cc) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(c >= 0 && c < SPIS_MAX_SLOTS, "constraint index %d out of range", c);
  REQUIRE(mat_slot < SPIS_MAX_SLOTS, "matrix slot %d out of range", mat_slot);
  REQUIRE(mat_slot < 0 || ctx->mats[mat_slot].present, "constraint matrix slot %d not uploaded", mat_slot);
  CU(cudaSetDevice(ctx->device));
  Constraint& C = ctx->cons[c];
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    C.defined = true; C.slot = mat_slot; C.cc = cc; C.cols_done = 0; C.term0_done = false; C.symmetric = -1;
  }
  if (v) {
    // an all-zero v is dropped (term1 then needs no v.Z pass).  Page-locked source: upload, then test the
    // device copy at HBM speed; pageable source: scan on the host first and skip the slow staged copy.
    // Row-sharded contexts never drop v on their own: whether v is zero is a GLOBAL question (a rank whose slice of v
    // is all zero still takes part in the all-reduced v.Z / v.x0 sums), so only the caller's NULL -- a decision it
    // takes collectively -- switches the v passes off there.
    const bool sharded = ctx->allreduce || ctx->halo || ctx->xbuf || ctx->n_halo > 0 || ctx->n_send > 0;
    int nz = 1;
    const bool pinned = is_pinned_host(v);
    if (!pinned && !sharded) TRY(spis_host_any_nonzero(v, (size_t)ctx->n, &nz));
    if (nz) {
      if (!C.v) TRY(dalloc(ctx, &C.v, (size_t)ctx->ld));
      TRY(h2d(ctx, C.v, v, (size_t)ctx->n * sizeof(double)));
      if (pinned && !sharded) TRY(device_any_nonzero(ctx, C.v, (size_t)ctx->n, &nz));
    }
    if (!nz && C.v) dfree(ctx, C.v);
  } else if (C.v) { dfree(ctx, C.v); }
  if (mat_slot < 0 && C.MZ) { dfree(ctx, C.MZ); }
  C.T1.assign((size_t)ctx->kmax, 0.0);
  C.T2.assign((size_t)ctx->kmax * ctx->kmax, 0.0);
  return SPIS_OK;
}

// New scalar c of a defined constraint (a time loop re-derives the invariants of every step's initial state,
// lkdv/Evolve.py:41 -> lkdv/lkdv.py:125-127, while M and v stay the same): nothing is uploaded again.
int spis_constraint_set_constant(spis_ctx* ctx, int c, double cc) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(c >= 0 && c < SPIS_MAX_SLOTS && ctx->cons[c].defined, "constraint %d is not defined", c);
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->cons[c].cc = cc;
  ctx->cons[c].term0_done = false;
  return SPIS_OK;
}

// New linear term v of a defined constraint (same matrix, same slot): lkdvRK's structured constraints depend on the
// step's initial state through v and c (wrappers/lkdvRK.py, conlist_structured), the matrix B'SB does not.
// v == NULL drops the linear term.
int spis_constraint_set_vector(spis_ctx* ctx, int c, const double* v) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(c >= 0 && c < SPIS_MAX_SLOTS && ctx->cons[c].defined, "constraint %d is not defined", c);
  CU(cudaSetDevice(ctx->device));
  Constraint& C = ctx->cons[c];
  if (v) {
    if (!C.v) TRY(dalloc(ctx, &C.v, (size_t)ctx->ld));
    TRY(h2d(ctx, C.v, v, (size_t)ctx->n * sizeof(double)));
  } else if (C.v) {
    dfree(ctx, C.v);
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  C.cols_done = 0; C.term0_done = false;
  return SPIS_OK;
}

// The whole staging of one class-form constraint -- is M identically zero (`0*A`, lkdv/LinearSolver.py:30)?
// upload + conversion of M, zero test and upload of v -- on a NATIVE helper thread and the context's auxiliary
// stream, so that the caller's thread can drive the Krylov loop meanwhile.  (Python helper threads did this
// before; every step of theirs had to win the interpreter lock from the thread running the loop, and the
// staging of a 10 ms job took 15-35 ms now and then, past the first constrained step.)
// The host arrays must stay valid until spis_constraint_setup_wait returns.
int spis_constraint_setup_async(spis_ctx* ctx, int c, int64_t nrows, int64_t ncols, int64_t nnz,
                                const int32_t* indptr, const int32_t* indices, const double* data,
                                const double* v, double cc) {
  return spis_constraint_setup_async2(ctx, c, nrows, ncols, nnz, indptr, indices, data, v, cc, -1);
}

// m_is_zero: -1 = find out (scan M's values), 0 / 1 = the caller knows -- a row-sharded session decides "is M zero?"
// for ALL ranks together (one collective for every such question of the set-up) and passes the answer down.
int spis_constraint_setup_async2(spis_ctx* ctx, int c, int64_t nrows, int64_t ncols, int64_t nnz,
                                 const int32_t* indptr, const int32_t* indices, const double* data,
                                 const double* v, double cc, int m_is_zero) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(c >= 0 && SPIS_SLOT_CON0 + c < SPIS_MAX_SLOTS, "constraint index %d out of range", c);
  REQUIRE(ctx->aux, "the context has no auxiliary stream");
  REQUIRE(indptr && (nnz == 0 || (indices && data)), "null CSR arrays");
  AsyncJob* job = new AsyncJob();
  ctx->jobs.push_back(job);
  job->th = std::thread([=]() {
    PhaseTrace pt;
    cudaSetDevice(ctx->device);
    tl_use_aux = true;
    int nz = 0;
    int rc = SPIS_OK;
    if (m_is_zero >= 0) nz = m_is_zero ? 0 : 1;
    else if (nnz > 0) rc = spis_host_any_nonzero(data, (size_t)nnz, &nz);
    pt.mark(nz ? "constraint: M is not zero" : "constraint: M is zero");
    int slot = -1;
    if (rc == SPIS_OK && nz) {
      slot = SPIS_SLOT_CON0 + c;
      rc = spis_upload_csr(ctx, slot, nrows, ncols, nnz, indptr, indices, data);
      pt.mark("constraint: upload M");
    }
    if (rc == SPIS_OK) rc = spis_constraint_define(ctx, c, slot, v, cc);
    if (cudaStreamSynchronize(ctx->aux) != cudaSuccess && rc == SPIS_OK) rc = SPIS_E_CUDA;
    pt.mark("constraint: v");
    job->rc = rc;
    if (rc != SPIS_OK) job->err = ctx->err;
    tl_use_aux = false;
  });
  return SPIS_OK;
}

int spis_constraint_setup_wait(spis_ctx* ctx) {
  if (!ctx) return SPIS_E_INVALID;
  int rc = SPIS_OK;
  for (AsyncJob* job : ctx->jobs) {
    if (job->th.joinable()) job->th.join();
    if (job->rc != SPIS_OK && rc == SPIS_OK) { rc = job->rc; snprintf(ctx->err, sizeof(ctx->err), "%s", job->err.c_str()); }
    delete job;
  }
  ctx->jobs.clear();
  return rc;
}

// Is the constraint matrix symmetric?  u^T (M w) == w^T (M u) for two pseudo-random vectors, to
// 1e-11 of |Mu||w| + |Mw||u| (a matrix whose asymmetry is >= 1e-6 relative fails this with
// overwhelming probability at any n; anything closer to symmetric is symmetric for our purposes:
// only the symmetric part of term2 enters y^T term2 y).  All mass/stiffness forms of the reference's
// experiments are symmetric; then the rows of Z^T (M Z) equal its columns and M Z need not be kept.
static int test_symmetry(spis_ctx* ctx, Constraint& C) {
  const size_t ld = (size_t)ctx->ld;
  double *u = ctx->G, *w = ctx->G + ld, *Mu = ctx->G + 2 * ld, *Mw = ctx->G + 3 * ld;
  CU(cudaMemsetAsync(ctx->G, 0, 4 * ld * sizeof(double), ctx->stream));
  fill_kernel<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(u, ctx->n, 0xA5A5ull);
  fill_kernel<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(w, ctx->n, 0x5A5Aull);
  CU(cudaGetLastError());
  TRY(do_halo(ctx, u));
  TRY(do_halo(ctx, w));
  TRY(launch_spmv(ctx, C.slot, 0, u, nullptr, Mu, nullptr));
  TRY(launch_spmv(ctx, C.slot, 0, w, nullptr, Mw, nullptr));
  double* o = ctx->d_cout;
  TRY(launch_mdot(ctx, Mu, 1, nullptr, 1, w, o));          // [Mu.w, w.w]
  TRY(launch_mdot(ctx, Mw, 1, nullptr, 1, u, o + 2));      // [Mw.u, u.u]
  TRY(launch_mdot(ctx, nullptr, 0, nullptr, 1, Mu, o + 4)); // [Mu.Mu]
  TRY(launch_mdot(ctx, nullptr, 0, nullptr, 1, Mw, o + 5)); // [Mw.Mw]
  TRY(d2h(ctx, ctx->h_cout, o, 6 * sizeof(double)));
  const double* h = ctx->h_cout;
  const double scale = std::sqrt(h[4] * h[1]) + std::sqrt(h[5] * h[3]);
  C.symmetric = std::fabs(h[0] - h[2]) <= 1e-11 * scale ? 1 : 0;
  return SPIS_OK;
}

static int constraint_terms_impl(spis_ctx* ctx, int c, int m, double* term0, double* term1, double* term2);

// A linear invariant (M == 0: term1 = v.Z only) asked for FIRST is served by the one-pass reduction of a quadratic
// constraint with the same columns pending: v rides as one more column of that pass (one more vector read instead of
// a pass of its own over Z), and the quadratic constraint's terms are kept for when they are asked for.
int spis_constraint_terms(spis_ctx* ctx, int c, int m, double* term0, double* term1, double* term2) {
  if (!ctx) return SPIS_E_INVALID;
  if (c >= 0 && c < SPIS_MAX_SLOTS && ctx->cons[c].defined && ctx->began && m >= 1 && m <= ctx->kmax && ctx->gram) {
    const Constraint& C = ctx->cons[c];
    if (C.slot < 0 && C.v && m - C.cols_done >= 8) {
      for (int o = 0; o < SPIS_MAX_SLOTS; ++o) {
        const Constraint& O = ctx->cons[o];
        if (o == c || !O.defined || O.slot < 0 || O.symmetric == 0 || O.cols_done != C.cols_done) continue;
        std::vector<double> t1((size_t)m), t2((size_t)m * m);
        double t0 = 0.0;
        ctx->piggy.assign(1, c);
        const int rc = constraint_terms_impl(ctx, o, m, &t0, t1.data(), t2.data());
        ctx->piggy.clear();
        if (rc != SPIS_OK) return rc;
        break;
      }
    }
  }
  return constraint_terms_impl(ctx, c, m, term0, term1, term2);
}

static int constraint_terms_impl(spis_ctx* ctx, int c, int m, double* term0, double* term1, double* term2) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(c >= 0 && c < SPIS_MAX_SLOTS && ctx->cons[c].defined, "constraint %d not defined", c);
  REQUIRE(ctx->began, "spis_solve_begin has not been called");
  REQUIRE(m >= 1 && m <= ctx->kmax, "m=%d out of range", m);
  REQUIRE(term0 && term1 && term2, "null output");
  CU(cudaSetDevice(ctx->device));
  Constraint& C = ctx->cons[c];
  const size_t ld = (size_t)ctx->ld;
  const int K = ctx->K;
  const int km = ctx->kmax;
  const bool hasM = C.slot >= 0;
  const bool x0nz = !ctx->x0_is_zero;
  double* Zb = zbase(ctx);
  if (hasM && !ctx->G) TRY(dalloc(ctx, &ctx->G, 4 * ld));
  if (hasM && C.symmetric < 0) {
    if (ctx->force_nonsymmetric) C.symmetric = 0; else TRY(test_symmetry(ctx, C));
  }
  const bool sym = hasM && C.symmetric == 1;
  if (hasM && !sym && !C.MZ) TRY(dalloc(ctx, &C.MZ, (size_t)ctx->kmax * ld));
  // term0 = 1/2 x0.M x0 + c + v.x0                                 (solvers.py:34)
  if (!C.term0_done) {
    double t = C.cc;
    if (x0nz && (hasM || C.v)) {
      double* o = ctx->d_cout;  // scratch: first 2 entries
      if (hasM) {
        TRY(do_halo(ctx, ctx->X0));
        TRY(launch_spmv(ctx, C.slot, 0, ctx->X0, nullptr, ctx->T, nullptr));
      }
      TRY(launch_mdot(ctx, ctx->T, hasM ? 1 : 0, C.v, 0, ctx->X0, o));
      TRY(d2h(ctx, ctx->h_cout, o, 2 * sizeof(double)));
      int idx = 0;
      if (hasM) t += 0.5 * ctx->h_cout[idx++];
      if (C.v) t += ctx->h_cout[idx++];
    }
    C.term0 = t; C.term0_done = true;
  }
  const int c0 = C.cols_done;
  if (m > c0) {
    // rows of the output block that are not written below must not carry stale data into the
    // batched all-reduce
    if (ctx->allreduce || ctx->xactive)
      CU(cudaMemsetAsync(ctx->d_cout + (size_t)c0 * 2 * K, 0, (size_t)(m - c0) * 2 * K * sizeof(double), ctx->stream));
    struct Group { int g0, g1, nr; };
    std::vector<Group> groups;
    ctx->defer_allreduce = true;
    int rc = SPIS_OK;
    // many new columns of a symmetric M (the switch to the constrained phase): M Z[c0..m) is kept, and ONE pass over
    // Z and M Z gives every entry (gram_kernel) instead of one pass over Z per four columns
    const double* gx = x0nz ? ctx->X0 : nullptr;
    bool use_gram = sym && (m - c0) >= 8 && gram_applicable(ctx, m, gx != nullptr);
    int gram_ra = 0;
    std::vector<const double*> gram_x; std::vector<int> gram_xc; bool gram_own_v_separate = false;
    if (use_gram) {
      const int want = (m - c0 + 7) / 8 * 8;
      if (ctx->gw_cols < want) {
        if (ctx->GW) dfree(ctx, ctx->GW);
        ctx->gw_cols = 0;
        const int rcw = dalloc(ctx, &ctx->GW, (size_t)want * ld, false);
        if (rcw == SPIS_OK) ctx->gw_cols = want;
        else if (rcw == SPIS_E_NOMEM) { ctx->err[0] = 0; use_gram = false; }     // no room for M Z: four columns at a time
        else { ctx->defer_allreduce = false; return rcw; }
      }
    }
    if (use_gram) {
      gram_ra = (m + (gx ? 1 : 0) + 7) / 8 * 8;
      for (int g0 = c0; g0 < m && rc == SPIS_OK;) {
        const int left = m - g0;
        const int nw = left >= 4 ? 4 : left >= 2 ? 2 : 1;
        double* dst = ctx->GW + (size_t)(g0 - c0) * ld;
        if (nw == 1) rc = launch_spmv(ctx, C.slot, 0, Zb + (size_t)g0 * ld, nullptr, dst, nullptr);
        else rc = launch_spmv_multi(ctx, C.slot, nw, Zb + (size_t)g0 * ld, (int64_t)ld, dst, (int64_t)ld);
        g0 += nw;
      }
      // v.Z of this constraint, and of the linear ones that asked to ride along, as extra columns of the same pass
      for (int pgy : ctx->piggy) {
        const Constraint& P = ctx->cons[pgy];
        if (gram_x.size() < 3 && P.defined && P.slot < 0 && P.v && P.cols_done == c0) { gram_x.push_back(P.v); gram_xc.push_back(pgy); }
      }
      if (C.v) { gram_x.push_back(C.v); gram_xc.push_back(c); }
      if ((size_t)(m - c0 + (int)gram_x.size()) * gram_ra > (size_t)(m - c0) * 2 * K) {     // no room in the output block: own v only, as a pass of its own
        gram_x.clear(); gram_xc.clear();
        if (rc == SPIS_OK && C.v)
          rc = launch_mdot(ctx, Zb + (size_t)c0 * ld, m - c0, nullptr, 0, C.v, ctx->d_cout + (size_t)(m - 1) * 2 * K + K);
        gram_own_v_separate = true;
      }
      if (rc == SPIS_OK) rc = launch_gram(ctx, Zb, m, gx, ctx->GW, m - c0, c0, 1, ctx->d_cout + (size_t)c0 * 2 * K, gram_ra,
                                          gram_x.data(), (int)gram_x.size());
    } else if (sym) {
      // symmetric M: groups of up to 4 new columns; M z_col lives only in the group buffer; one
      // pass over Z[0..g1) serves the whole group (each basis row is read once per group)
      for (int g0 = c0; g0 < m && rc == SPIS_OK;) {
        const int left = m - g0;
        const int nw = left >= 4 ? 4 : left >= 2 ? 2 : 1;
        const int g1 = g0 + nw;
        double* base = ctx->d_cout + (size_t)g0 * 2 * K;
        // M z_col for the whole group, one pass over M (ghosts filled by the Arnoldi step)          (:33)
        if (nw == 1) rc = launch_spmv(ctx, C.slot, 0, Zb + (size_t)g0 * ld, nullptr, ctx->G, nullptr);
        else rc = launch_spmv_multi(ctx, C.slot, nw, Zb + (size_t)g0 * ld, (int64_t)ld, ctx->G, (int64_t)ld);
        const double* extra = x0nz ? ctx->X0 : nullptr;
        const int nr = g1 + (extra ? 1 : 0);
        if (rc == SPIS_OK) {
          if (nw == 1) rc = launch_mdot(ctx, Zb, g1, extra, 0, ctx->G, base);        // (:35-36)
          else rc = launch_mdotm(ctx, nw, Zb, g1, extra, ctx->G, (int64_t)ld, base);
        }
        groups.push_back({g0, g1, nr});
        g0 = g1;
      }
      // v.z_col for ALL new columns in one pass: rows Z[c0..m) against v (the slot behind the last
      // column's block is not used by any group)
      if (rc == SPIS_OK && C.v)
        rc = launch_mdot(ctx, Zb + (size_t)c0 * ld, m - c0, nullptr, 0, C.v, ctx->d_cout + (size_t)(m - 1) * 2 * K + K);
    } else if (!hasM) {
      // M == 0 (mass-type invariants, lkdv/LinearSolver.py:28-32): term1 = v.Z only, one pass for all new columns
      if (C.v) rc = launch_mdot(ctx, Zb + (size_t)c0 * ld, m - c0, nullptr, 0, C.v, ctx->d_cout + (size_t)c0 * 2 * K);
    } else {
      for (int col = c0; col < m && rc == SPIS_OK; ++col) {
        double* zc = Zb + (size_t)col * ld;
        double* oA = ctx->d_cout + (size_t)col * 2 * K;
        double* oB = oA + K;
        if (hasM) {
          double* mz = C.MZ + (size_t)col * ld;
          rc = launch_spmv(ctx, C.slot, 0, zc, nullptr, mz, nullptr);                               // (:33)
          // column col of Z^T MZ, and x0.MZ_col                                                    (:35-36)
          if (rc == SPIS_OK) rc = launch_mdot(ctx, Zb, col + 1, x0nz ? ctx->X0 : nullptr, 0, mz, oA);
          // row col of Z^T MZ (MZ_i.z_col, i<col), and v.z_col
          if (rc == SPIS_OK && (col > 0 || C.v)) rc = launch_mdot(ctx, C.MZ, col, C.v, 0, zc, oB);
        }
      }
    }
    ctx->defer_allreduce = false;
    if (rc != SPIS_OK) return rc;
    // one all-reduce for every dot block of this call (row-sharded runs)
    TRY(do_allreduce(ctx, ctx->d_cout + (size_t)c0 * 2 * K, (int64_t)(m - c0) * 2 * K, true));
    TRY(d2h(ctx, ctx->h_cout + (size_t)c0 * 2 * K, ctx->d_cout + (size_t)c0 * 2 * K, (size_t)(m - c0) * 2 * K * sizeof(double)));
    if (use_gram) {
      const double* vz = ctx->h_cout + (size_t)(m - 1) * 2 * K + K;      // v.z_col, col = c0 .. m-1 (own pass)
      const double* base = ctx->h_cout + (size_t)c0 * 2 * K;
      const double* own = nullptr;                                       // v.z_row, row = 0 .. m-1 (rode along)
      for (size_t x = 0; x < gram_xc.size(); ++x) {
        const double* colx = base + (size_t)(m - c0 + (int)x) * gram_ra;
        if (gram_xc[x] == c) { own = colx; continue; }
        Constraint& P = ctx->cons[gram_xc[x]];
        for (int col = c0; col < m; ++col) P.T1[col] = colx[col];        // M == 0: term1 = v.Z (solvers.py:36)
        P.cols_done = m;
      }
      for (int col = c0; col < m; ++col) {
        const double* oA = base + (size_t)(col - c0) * gram_ra;
        for (int i = 0; i <= col; ++i) {
          C.T2[(size_t)i * km + col] = 0.5 * oA[i];
          C.T2[(size_t)col * km + i] = 0.5 * oA[i];
        }
        double t1 = 0.0;
        if (x0nz) t1 += oA[m];
        if (C.v) t1 += (gram_own_v_separate || !own) ? vz[col - c0] : own[col];
        C.T1[col] = t1;
      }
    } else if (sym) {
      const double* vz = ctx->h_cout + (size_t)(m - 1) * 2 * K + K;      // v.z_col, col = c0 .. m-1
      for (const Group& g : groups) {
        const int nw = g.g1 - g.g0;
        const double* base = ctx->h_cout + (size_t)g.g0 * 2 * K;
        for (int cc2 = 0; cc2 < nw; ++cc2) {
          const int col = g.g0 + cc2;
          const double* oA = base + (size_t)cc2 * g.nr;
          for (int i = 0; i <= col; ++i) {
            C.T2[(size_t)i * km + col] = 0.5 * oA[i];
            C.T2[(size_t)col * km + i] = 0.5 * oA[i];
          }
          double t1 = 0.0;
          if (x0nz) t1 += oA[g.g1];
          if (C.v) t1 += vz[col - c0];
          C.T1[col] = t1;
        }
      }
    } else if (!hasM) {
      const double* vz = ctx->h_cout + (size_t)c0 * 2 * K;
      for (int col = c0; col < m; ++col) C.T1[col] = C.v ? vz[col - c0] : 0.0;
    } else {
      for (int col = c0; col < m; ++col) {
        const double* oA = ctx->h_cout + (size_t)col * 2 * K;
        const double* oB = oA + K;
        double t1 = 0.0;
        for (int i = 0; i <= col; ++i) C.T2[(size_t)i * km + col] = 0.5 * oA[i];
        if (x0nz) t1 += oA[col + 1];
        for (int i = 0; i < col; ++i) C.T2[(size_t)col * km + i] = 0.5 * oB[i];
        if (C.v) t1 += oB[col];
        C.T1[col] = t1;
      }
    }
    C.cols_done = m;
  }
  *term0 = C.term0;
  for (int i = 0; i < m; ++i) term1[i] = C.T1[i];
  for (int i = 0; i < m; ++i)
    for (int k = 0; k < m; ++k) term2[(size_t)i * m + k] = C.T2[(size_t)i * ctx->kmax + k];
  return SPIS_OK;
}

// The reduced terms of SEVERAL constraints for the same m in one call.  The common case of a constrained iteration -- one
// new basis column, every constraint one column behind (solvers.py:242-247 rebuilds them all every iteration; cgmres_p
// does nothing else) -- is served in ONE pass over Z for all quadratic constraints together (mdotm_kernel: M_c z_col as
// up to four right-hand sides) with one cross-rank reduction, one copy to the host and one synchronisation, instead of
// a pass, a reduction and a synchronisation per constraint.  Anything else goes through spis_constraint_terms per
// constraint.  term1: nc x m, term2: nc x m x m (row-major, constraint by constraint).
int spis_constraint_terms_batch(spis_ctx* ctx, int nc, const int32_t* cs, int m, double* term0, double* term1, double* term2) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(nc >= 1 && nc <= SPIS_MAX_SLOTS && cs && term0 && term1 && term2, "bad batch arguments");
  REQUIRE(m >= 1 && m <= ctx->kmax, "m=%d out of range", m);
  for (int i = 0; i < nc; ++i)
    REQUIRE(cs[i] >= 0 && cs[i] < SPIS_MAX_SLOTS && ctx->cons[cs[i]].defined, "constraint %d not defined", cs[i]);
  REQUIRE(ctx->began, "spis_solve_begin has not been called");
  CU(cudaSetDevice(ctx->device));
  int quad[4]; int nq = 0;
  bool fast = ctx->batch_terms != 0;
  for (int i = 0; i < nc && fast; ++i) {
    const Constraint& C = ctx->cons[cs[i]];
    fast = C.term0_done && C.cols_done == m - 1;
    for (int k = 0; k < i && fast; ++k) fast = cs[k] != cs[i];
    if (fast && C.slot >= 0) {
      if (C.symmetric == 1 && nq < 4) quad[nq++] = i; else fast = false;
    }
  }
  const size_t ld = (size_t)ctx->ld;
  const bool x0nz = !ctx->x0_is_zero;
  const double* extra = x0nz ? ctx->X0 : nullptr;
  const int nrows = m + (extra ? 1 : 0);
  const int nw = nq <= 1 ? nq : nq == 2 ? 2 : 4;
  const int64_t count = (int64_t)nw * nrows + 2 * nc;
  if (fast && (nq == 0 || nw * nrows > ctx->pstride || count > (int64_t)ctx->kmax * 2 * ctx->K)) fast = false;
  if (!fast) {
    for (int i = 0; i < nc; ++i)
      TRY(spis_constraint_terms(ctx, cs[i], m, term0 + i, term1 + (size_t)i * m, term2 + (size_t)i * m * m));
    return SPIS_OK;
  }
  const int col = m - 1;
  const int km = ctx->kmax;
  double* Zb = zbase(ctx);
  double* zc = Zb + (size_t)col * ld;
  if (!ctx->G) TRY(dalloc(ctx, &ctx->G, 4 * ld));
  double* o_m = ctx->d_cout;                       // [nw][nrows]: row_i . (M_q z_col), then x0 . (M_q z_col)
  double* o_v = ctx->d_cout + (size_t)nw * nrows;  // [nc][2]: v_c . z_col
  if (ctx->allreduce || ctx->xactive) CU(cudaMemsetAsync(ctx->d_cout, 0, (size_t)count * sizeof(double), ctx->stream));
  ctx->defer_allreduce = true;
  int rc = SPIS_OK;
  for (int q = 0; q < nq && rc == SPIS_OK; ++q)      // M_q z_col (ghosts of z_col were filled by its Arnoldi step)   (:33)
    rc = launch_spmv(ctx, ctx->cons[cs[quad[q]]].slot, 0, zc, nullptr, ctx->G + (size_t)q * ld, nullptr);
  if (rc == SPIS_OK) {                               // column `col` of every Z^T M_q Z, and x0 . M_q z_col          (:35-36)
    if (nq == 1) rc = launch_mdot(ctx, Zb, m, extra, 0, ctx->G, o_m);
    else rc = launch_mdotm(ctx, nw, Zb, m, extra, ctx->G, (int64_t)ld, o_m);
  }
  for (int i = 0; i < nc && rc == SPIS_OK; ++i) {
    const Constraint& C = ctx->cons[cs[i]];
    if (C.v) rc = launch_mdot(ctx, zc, 1, nullptr, 0, C.v, o_v + 2 * i);
  }
  ctx->defer_allreduce = false;
  if (rc != SPIS_OK) return rc;
  TRY(do_allreduce(ctx, ctx->d_cout, count, true));
  TRY(d2h(ctx, ctx->h_cout, ctx->d_cout, (size_t)count * sizeof(double)));
  const double* hv = ctx->h_cout + (size_t)nw * nrows;
  for (int i = 0; i < nc; ++i) {
    Constraint& C = ctx->cons[cs[i]];
    double t1 = 0.0;
    int q = -1;
    for (int k = 0; k < nq; ++k) if (quad[k] == i) q = k;
    if (q >= 0) {
      const double* oA = ctx->h_cout + (size_t)q * nrows;
      for (int r = 0; r <= col; ++r) {
        C.T2[(size_t)r * km + col] = 0.5 * oA[r];
        C.T2[(size_t)col * km + r] = 0.5 * oA[r];
      }
      if (x0nz) t1 += oA[m];
    }
    if (C.v) t1 += hv[2 * i];
    C.T1[col] = t1;
    C.cols_done = m;
  }
  for (int i = 0; i < nc; ++i) {
    const Constraint& C = ctx->cons[cs[i]];
    term0[i] = C.term0;
    for (int r = 0; r < m; ++r) term1[(size_t)i * m + r] = C.T1[r];
    for (int r = 0; r < m; ++r)
      for (int k = 0; k < m; ++k) term2[((size_t)i * m + r) * m + k] = C.T2[(size_t)r * km + k];
  }
  return SPIS_OK;
}

// ---- downloads / host bridges ---------------------------------------------------------
int spis_download_vec(spis_ctx* ctx, int which, int j, double* host, int64_t n) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(host && n == ctx->n, "bad download arguments");
  CU(cudaSetDevice(ctx->device));
  const double* src = nullptr;
  switch (which) {
    case SPIS_VEC_B: src = ctx->B; break;
    case SPIS_VEC_X0: src = ctx->X0; break;
    case SPIS_VEC_R0: src = ctx->R0; break;
    case SPIS_VEC_X: src = ctx->X; break;
    case SPIS_VEC_W: src = ctx->W; break;
    case SPIS_VEC_Q: REQUIRE(j >= 0 && j <= ctx->kmax, "q index %d out of range", j); src = ctx->V + (size_t)j * ctx->ld; break;
    case SPIS_VEC_Z: REQUIRE(j >= 0 && j < ctx->kmax, "z index %d out of range", j); src = zbase(ctx) + (size_t)j * ctx->ld; break;
    default: return fail(ctx, SPIS_E_INVALID, "vector id %d cannot be downloaded", which);
  }
  return d2h(ctx, host, src, (size_t)n * sizeof(double));
}

int spis_download_Z(spis_ctx* ctx, int j0, int j1, double* host) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(host && j0 >= 0 && j0 <= j1 && j1 <= ctx->kmax, "bad Z range [%d,%d)", j0, j1);
  CU(cudaSetDevice(ctx->device));
  if (j1 == j0) return SPIS_OK;
  CU(cudaMemcpy2DAsync(host, (size_t)ctx->n * sizeof(double), zbase(ctx) + (size_t)j0 * ctx->ld, (size_t)ctx->ld * sizeof(double),
                       (size_t)ctx->n * sizeof(double), (size_t)(j1 - j0), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return SPIS_OK;
}

int spis_host_pre_get(spis_ctx* ctx, int j, double* q_host) {
  return spis_download_vec(ctx, SPIS_VEC_Q, j, q_host, ctx ? ctx->n : 0);
}

int spis_host_pre_put(spis_ctx* ctx, int j, const double* z_host) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->pre_kind == SPIS_PRE_HOST, "host preconditioner bridge is not selected");
  REQUIRE(z_host && j >= 0 && j < ctx->kmax, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  TRY(ensure_Z(ctx));
  return h2d(ctx, ctx->Z + (size_t)j * ctx->ld, z_host, (size_t)ctx->n * sizeof(double));
}

int spis_set_collectives(spis_ctx* ctx, spis_allreduce_fn allreduce, spis_halo_fn halo, void* user) {
  if (!ctx) return SPIS_E_INVALID;
  ctx->allreduce = allreduce; ctx->halo = halo; ctx->cuser = user;
  return SPIS_OK;
}

int spis_halo_set_plan(spis_ctx* ctx, const int32_t* send_idx, int64_t n_send) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(n_send >= 0 && (send_idx || n_send == 0), "bad halo plan");
  CU(cudaSetDevice(ctx->device));
  for (int64_t i = 0; i < n_send; ++i)
    REQUIRE(send_idx[i] >= 0 && send_idx[i] < ctx->n, "halo send index %lld out of range", (long long)send_idx[i]);
  dfree(ctx, ctx->d_send_idx); dfree(ctx, ctx->d_send);
  ctx->n_send = n_send;
  if (n_send) {
    TRY(dalloc(ctx, &ctx->d_send_idx, (size_t)n_send, false));
    TRY(dalloc(ctx, &ctx->d_send, (size_t)n_send));
    TRY(h2d(ctx, ctx->d_send_idx, send_idx, (size_t)n_send * sizeof(int32_t)));
  }
  return SPIS_OK;
}

int spis_xcomm_create(spis_ctx* ctx, int rank, int world, int64_t halo_cap, void* handle_out, int64_t handle_capacity) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank %d / world %d (max %d ranks)", rank, world, kMaxRanks);
  REQUIRE(handle_out && handle_capacity >= (int64_t)sizeof(cudaIpcMemHandle_t), "handle buffer too small (%zu bytes needed)", sizeof(cudaIpcMemHandle_t));
  REQUIRE(!ctx->xbuf, "peer-memory communicator already created");
  REQUIRE(halo_cap >= ctx->n_halo, "halo capacity %lld < n_halo %lld", (long long)halo_cap, (long long)ctx->n_halo);
  CU(cudaSetDevice(ctx->device));
  XView xv;
  xv.world = world; xv.rank = rank;
  xv.red_cap = ctx->K > 1024 ? ctx->K : 1024;
  xv.halo_cap = halo_cap > 0 ? 4 * roundup(halo_cap, 16) : 64;      // two vectors per exchange, 16 bytes per value (halo_xchg_kernel)
  const size_t bytes = xv.total_doubles() * sizeof(double);
  CU(cudaMalloc((void**)&ctx->xbuf, bytes));          // plain cudaMalloc: pool memory cannot be exported
  CU(cudaMemset(ctx->xbuf, 0, bytes));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ctx->xbuf));
  memcpy(handle_out, &h, sizeof(h));
  ctx->xv = xv;
  ctx->xv.base[rank] = ctx->xbuf;
  return SPIS_OK;
}

int spis_xcomm_connect(spis_ctx* ctx, const void* handles) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->xbuf && handles, "spis_xcomm_create must come first");
  CU(cudaSetDevice(ctx->device));
  const char* hb = static_cast<const char*>(handles);
  for (int r = 0; r < ctx->xv.world; ++r) {
    if (r == ctx->xv.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->xpeer[r] = p;
    ctx->xv.base[r] = static_cast<double*>(p);
  }
  ctx->xactive = ctx->xv.world > 1;
  ctx->xseq_own = 1; ctx->hseq_own = 1;
  return SPIS_OK;
}

// ---- persistent communicator -----------------------------------------------------------------------------
int spis_comm_create(int device, int rank, int world, int64_t red_cap, int64_t halo_cap, void* handle_out,
                     int64_t handle_capacity, spis_comm** comm_out) {
  spis_ctx* ctx = nullptr;
  if (!comm_out || !handle_out) return fail(ctx, SPIS_E_INVALID, "null argument");
  *comm_out = nullptr;
  REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank %d / world %d (max %d ranks)", rank, world, kMaxRanks);
  REQUIRE(handle_capacity >= (int64_t)sizeof(cudaIpcMemHandle_t), "handle buffer too small (%zu bytes needed)", sizeof(cudaIpcMemHandle_t));
  REQUIRE(red_cap >= 1 && red_cap < (1 << 24) && halo_cap >= 0, "bad capacities");
  CU(cudaSetDevice(device));
  spis_comm* c = new spis_comm();
  c->device = device;
  c->xv.world = world; c->xv.rank = rank;
  c->xv.red_cap = (int)(red_cap < 1024 ? 1024 : red_cap);
  c->xv.halo_cap = halo_cap > 0 ? 4 * roundup(halo_cap, 16) : 64;
  const size_t bytes = c->xv.total_doubles() * sizeof(double);
  cudaError_t e = cudaMalloc((void**)&c->xbuf, bytes);            // plain cudaMalloc: pool memory cannot be exported
  if (e == cudaSuccess) e = cudaMemset(c->xbuf, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->xbuf);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_tmp, (size_t)c->xv.red_cap * sizeof(double));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_tmp, (size_t)c->xv.red_cap * sizeof(double));
  if (e != cudaSuccess) {
    if (c->xbuf) cudaFree(c->xbuf);
    if (c->d_tmp) cudaFree(c->d_tmp);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return fail(ctx, e == cudaErrorMemoryAllocation ? SPIS_E_NOMEM : SPIS_E_CUDA, "communicator: %s", cudaGetErrorString(e));
  }
  memcpy(handle_out, &h, sizeof(h));
  c->xv.base[rank] = c->xbuf;
  *comm_out = c;
  return SPIS_OK;
}

int spis_comm_connect(spis_comm* comm, const void* handles) {
  spis_ctx* ctx = nullptr;
  if (!comm || !handles) return fail(ctx, SPIS_E_INVALID, "null argument");
  CU(cudaSetDevice(comm->device));
  const char* hb = static_cast<const char*>(handles);
  for (int r = 0; r < comm->xv.world; ++r) {
    if (r == comm->xv.rank || comm->xpeer[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    comm->xpeer[r] = p;
    comm->xv.base[r] = static_cast<double*>(p);
  }
  comm->connected = true;
  return SPIS_OK;
}

int spis_comm_destroy(spis_comm* comm) {
  if (!comm) return SPIS_OK;
  cudaSetDevice(comm->device);
  if (comm->stream) cudaStreamSynchronize(comm->stream);
  for (int r = 0; r < kMaxRanks; ++r)
    if (comm->xpeer[r]) cudaIpcCloseMemHandle(comm->xpeer[r]);
  if (comm->xbuf) cudaFree(comm->xbuf);
  if (comm->d_tmp) cudaFree(comm->d_tmp);
  if (comm->h_tmp) cudaFreeHost(comm->h_tmp);
  if (comm->stream) cudaStreamDestroy(comm->stream);
  delete comm;
  return SPIS_OK;
}

// capacities: doubles per reduction (>= k_max + 5 of every context that attaches) and ghost entries per vector
int spis_comm_capacity(const spis_comm* comm, int64_t* red_cap_out, int64_t* halo_cap_out) {
  if (!comm) return SPIS_E_INVALID;
  if (red_cap_out) *red_cap_out = comm->xv.red_cap;
  if (halo_cap_out) *halo_cap_out = comm->xv.halo_cap / 4;
  return SPIS_OK;
}

// in-place sum over all ranks of `count` host doubles (collective decisions of a session set-up: "is x0 zero on every
// rank?", ...): one tiny kernel on the communicator's own stream.  Every rank must call it, with the same count, at
// the same point of its sequence of collectives (the contexts' streams are idle between solves).
int spis_comm_allreduce(spis_comm* comm, double* vals, int count) {
  spis_ctx* ctx = nullptr;
  if (!comm || !vals) return fail(ctx, SPIS_E_INVALID, "null argument");
  REQUIRE(comm->connected || comm->xv.world == 1, "the communicator is not connected");
  REQUIRE(count >= 1 && count <= comm->xv.red_cap, "count %d exceeds the reduction capacity %d", count, comm->xv.red_cap);
  if (comm->xv.world == 1) return SPIS_OK;
  CU(cudaSetDevice(comm->device));
  memcpy(comm->h_tmp, vals, (size_t)count * sizeof(double));
  CU(cudaMemcpyAsync(comm->d_tmp, comm->h_tmp, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, comm->stream));
  xreduce_kernel<<<1, kThreads, 0, comm->stream>>>(comm->d_tmp, count, comm->xv, comm->xseq);
  CU(cudaGetLastError());
  comm->xseq += 1;
  CU(cudaMemcpyAsync(comm->h_tmp, comm->d_tmp, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, comm->stream));
  CU(cudaStreamSynchronize(comm->stream));
  memcpy(vals, comm->h_tmp, (size_t)count * sizeof(double));
  unsigned long long w = 0;
  CU(cudaMemcpy(&w, comm->xbuf + comm->xv.flags_off() + 4 * comm->xv.world, sizeof(w), cudaMemcpyDeviceToHost));
  if (w >> 63) return fail(ctx, SPIS_E_CUDA, "NVLink collective %llu timed out waiting for a peer rank", w & ~(1ull << 63));
  return SPIS_OK;
}

int spis_ctx_attach_comm(spis_ctx* ctx, spis_comm* comm) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(comm && (comm->connected || comm->xv.world == 1), "the communicator is not connected");
  REQUIRE(!ctx->xbuf && !ctx->comm, "the context already has a communicator");
  REQUIRE(comm->device == ctx->device, "communicator lives on device %d, context on %d", comm->device, ctx->device);
  REQUIRE(comm->xv.red_cap >= ctx->K + 1, "communicator reduces %d doubles at a time, k_max = %d needs %d", comm->xv.red_cap, ctx->kmax, ctx->K + 1);
  REQUIRE(comm->xv.halo_cap / 4 >= ctx->n_halo, "communicator holds %lld ghost entries per vector, the context has %lld", (long long)(comm->xv.halo_cap / 4), (long long)ctx->n_halo);
  ctx->comm = comm;
  ctx->xbuf = comm->xbuf;
  ctx->xv = comm->xv;
  ctx->xseq_p = &comm->xseq; ctx->hseq_p = &comm->hseq;
  ctx->xactive = comm->xv.world > 1;
  return SPIS_OK;
}

int spis_xcomm_set_halo(spis_ctx* ctx, const int32_t* dest_rank, const int32_t* dest_off,
                        const int32_t* send_to, const int32_t* recv_from) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ctx->xbuf && send_to && recv_from, "spis_xcomm_create must come first");
  REQUIRE(ctx->n_send == 0 || (dest_rank && dest_off), "null destination arrays");
  CU(cudaSetDevice(ctx->device));
  for (int64_t i = 0; i < ctx->n_send; ++i)
    REQUIRE(dest_rank[i] >= 0 && dest_rank[i] < ctx->xv.world && dest_off[i] >= 0, "bad halo destination at %lld", (long long)i);
  dfree(ctx, ctx->d_dest_rank); dfree(ctx, ctx->d_dest_off); dfree(ctx, ctx->d_send_to); dfree(ctx, ctx->d_recv_from);
  if (ctx->n_send) {
    TRY(dalloc(ctx, &ctx->d_dest_rank, (size_t)ctx->n_send, false));
    TRY(dalloc(ctx, &ctx->d_dest_off, (size_t)ctx->n_send, false));
    TRY(h2d(ctx, ctx->d_dest_rank, dest_rank, (size_t)ctx->n_send * sizeof(int32_t)));
    TRY(h2d(ctx, ctx->d_dest_off, dest_off, (size_t)ctx->n_send * sizeof(int32_t)));
  }
  TRY(dalloc(ctx, &ctx->d_send_to, (size_t)kMaxRanks));
  TRY(dalloc(ctx, &ctx->d_recv_from, (size_t)kMaxRanks));
  TRY(h2d(ctx, ctx->d_send_to, send_to, (size_t)ctx->xv.world * sizeof(int32_t)));
  TRY(h2d(ctx, ctx->d_recv_from, recv_from, (size_t)ctx->xv.world * sizeof(int32_t)));
  return SPIS_OK;
}

// out[0..3]: SM cycles this rank's reducing kernels spent waiting for their peers' contributions and the number of
// fused reductions, the same for halo exchanges -- since the last call (the counters are cleared).
int spis_xcomm_stats(spis_ctx* ctx, uint64_t* out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(out, "null output");
  for (int i = 0; i < 4; ++i) out[i] = 0;
  if (!ctx->xbuf) return SPIS_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  unsigned long long w[8] = {0};
  double* at = ctx->xbuf + ctx->xv.flags_off() + 4 * ctx->xv.world;
  CU(cudaMemcpy(w, at, sizeof(w), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) out[i] = w[1 + i];
  CU(cudaMemset(at + 1, 0, 6 * sizeof(double)));
  return SPIS_OK;
}

int spis_thread_use_aux_stream(spis_ctx* ctx, int on) {
  if (!ctx) return SPIS_E_INVALID;
  tl_use_aux = on != 0;
  if (!on) {                     // leaving: everything this thread queued on the auxiliary stream is complete
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->aux));
  }
  return SPIS_OK;
}

int spis_sync(spis_ctx* ctx) {
  if (!ctx) return SPIS_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return SPIS_OK;
}

// ---- measurement ----------------------------------------------------------------------
int spis_get_profile(spis_ctx* ctx, double* ms_out, double* bytes_out, int64_t* launches_out) {
  if (!ctx) return SPIS_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  TRY(prof_resolve(ctx));
  for (int i = 0; i < SPIS_PROF_CLASSES; ++i) {
    if (ms_out) ms_out[i] = ctx->prof_ms[i];
    if (bytes_out) bytes_out[i] = ctx->prof_bytes[i];
    if (launches_out) launches_out[i] = ctx->prof_launch[i];
  }
  return SPIS_OK;
}

int spis_reset_profile(spis_ctx* ctx) {
  if (!ctx) return SPIS_E_INVALID;
  TRY(prof_resolve(ctx));
  for (int i = 0; i < SPIS_PROF_CLASSES; ++i) { ctx->prof_ms[i] = 0; ctx->prof_bytes[i] = 0; ctx->prof_launch[i] = 0; ctx->prof_moved[i] = 0; ctx->prof_gap_ms[i] = 0; }
  return SPIS_OK;
}

// per class: bytes the launches moved with the storage format they ran on (16-bit stencil ids, 8-bit value codes, one
// pass over the matrix for two products ...); equals the algorithmic bytes of spis_get_profile except for SpMV
int spis_get_profile_moved(spis_ctx* ctx, double* moved_out) {
  if (!ctx || !moved_out) return SPIS_E_INVALID;
  for (int i = 0; i < SPIS_PROF_CLASSES; ++i) moved_out[i] = ctx->prof_moved[i];
  return SPIS_OK;
}

// profile mode: gaps_out[c] = milliseconds the device sat idle right before the launches of class c (end of the previous
// profiled kernel to the start of this one: waiting for the host, a copy, or another stream)
int spis_get_profile_gaps(spis_ctx* ctx, double* gaps_out) {
  if (!ctx || !gaps_out) return SPIS_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  TRY(prof_resolve(ctx));
  for (int i = 0; i < SPIS_PROF_CLASSES; ++i) gaps_out[i] = ctx->prof_gap_ms[i];
  return SPIS_OK;
}

// profile mode: the launches resolved by the last spis_get_profile / _gaps call, in launch order -- class, start (ms
// after the first of them) and duration.  A diagnostic: where on the timeline the device waits for the host.
int spis_get_profile_trace(spis_ctx* ctx, int32_t* cls_out, double* start_ms_out, double* dur_ms_out, int64_t cap, int64_t* n_out) {
  if (!ctx || !n_out) return SPIS_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->recs.empty()) TRY(prof_resolve(ctx));
  const int64_t n = (int64_t)ctx->trace.size();
  *n_out = n;
  for (int64_t i = 0; i < n && i < cap; ++i) {
    if (cls_out) cls_out[i] = ctx->trace[i].cls;
    if (start_ms_out) start_ms_out[i] = ctx->trace[i].start_ms;
    if (dur_ms_out) dur_ms_out[i] = ctx->trace[i].dur_ms;
  }
  return SPIS_OK;
}

int spis_timer_start(spis_ctx* ctx) {
  if (!ctx) return SPIS_E_INVALID;
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventRecord(ctx->ev_t0, ctx->stream));
  return SPIS_OK;
}

int spis_timer_stop(spis_ctx* ctx, double* ms_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ms_out, "ms_out is null");
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventRecord(ctx->ev_t1, ctx->stream));
  CU(cudaEventSynchronize(ctx->ev_t1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
  *ms_out = ms;
  return SPIS_OK;
}

// ---- single-kernel entry points ---------------------------------------------------------
int spis_op_spmv(spis_ctx* ctx, int slot, const double* x, double* y) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(x && y, "null argument");
  REQUIRE(slot >= 0 && slot < SPIS_MAX_SLOTS && ctx->mats[slot].present, "matrix slot %d not uploaded", slot);
  CU(cudaSetDevice(ctx->device));
  // x has ncols entries: owned part then ghost part
  const Matrix& M = ctx->mats[slot];
  CU(cudaMemsetAsync(ctx->T, 0, (size_t)ctx->ld * sizeof(double), ctx->stream));
  const int64_t nown = M.ncols < ctx->n ? M.ncols : ctx->n;
  CU(cudaMemcpyAsync(ctx->T, x, (size_t)nown * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (M.ncols > ctx->n)
    CU(cudaMemcpyAsync(ctx->T + ctx->hoff, x + ctx->n, (size_t)(M.ncols - ctx->n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  TRY(launch_spmv(ctx, slot, 0, ctx->T, nullptr, ctx->W, nullptr));
  return d2h(ctx, y, ctx->W, (size_t)ctx->n * sizeof(double));
}

int spis_op_mdot(spis_ctx* ctx, int m, const double* V, const double* w, double* out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(m >= 0 && m <= ctx->kmax && w && out && (V || m == 0), "bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (m)
    CU(cudaMemcpy2DAsync(ctx->V, (size_t)ctx->ld * sizeof(double), V, (size_t)ctx->n * sizeof(double), (size_t)ctx->n * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->W, w, (size_t)ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  TRY(launch_mdot(ctx, ctx->V, m, nullptr, 1, ctx->W, ctx->d_small));
  return d2h(ctx, out, ctx->d_small, (size_t)(m + 1) * sizeof(double));
}

int spis_op_lincomb(spis_ctx* ctx, int m, const double* V, const double* base, const double* coef,
                    double sign, double* out, double* sumsq_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(m >= 0 && m <= ctx->kmax && out && (m == 0 || (V && coef)), "bad arguments");
  CU(cudaSetDevice(ctx->device));
  if (m) {
    CU(cudaMemcpy2DAsync(ctx->V, (size_t)ctx->ld * sizeof(double), V, (size_t)ctx->n * sizeof(double), (size_t)ctx->n * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_y, coef, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (base) CU(cudaMemcpyAsync(ctx->W, base, (size_t)ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  double* scal = ctx->d_small + 2 * ctx->K;
  TRY(launch_lincomb(ctx, ctx->V, m, ctx->d_y, nullptr, sign, base ? ctx->W : nullptr, ctx->T, sumsq_out ? 1 : 0, scal + 3));
  TRY(d2h(ctx, out, ctx->T, (size_t)ctx->n * sizeof(double)));
  if (sumsq_out) TRY(d2h(ctx, sumsq_out, scal + 3, sizeof(double)));
  return SPIS_OK;
}

int spis_op_orth_mid(spis_ctx* ctx, int m, const double* V, const double* w, const double* coef,
                     double* w_out, double* dots_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(m >= 1 && m <= ctx->kmax && V && w && coef && w_out && dots_out, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpy2DAsync(ctx->V, (size_t)ctx->ld * sizeof(double), V, (size_t)ctx->n * sizeof(double), (size_t)ctx->n * sizeof(double), (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_y, coef, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->W, w, (size_t)ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  TRY(launch_orth_mid(ctx, ctx->V, m, ctx->d_y, ctx->W, ctx->d_small));
  TRY(d2h(ctx, w_out, ctx->W, (size_t)ctx->n * sizeof(double)));
  return d2h(ctx, dots_out, ctx->d_small, (size_t)m * sizeof(double));
}

int spis_op_precond(spis_ctx* ctx, const double* q, double* z) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(q && z, "null argument");
  REQUIRE(ctx->pre_kind == SPIS_PRE_JACOBI || ctx->pre_kind == SPIS_PRE_CSR || ctx->pre_kind == SPIS_PRE_BLOCK, "no device preconditioner selected");
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemsetAsync(ctx->T, 0, (size_t)ctx->ld * sizeof(double), ctx->stream));
  CU(cudaMemsetAsync(ctx->W, 0, (size_t)ctx->ld * sizeof(double), ctx->stream));
  CU(cudaMemcpyAsync(ctx->T, q, (size_t)ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  TRY(launch_precond(ctx, ctx->T, ctx->W));
  return d2h(ctx, z, ctx->W, (size_t)ctx->n * sizeof(double));
}

int spis_bench_kernel(spis_ctx* ctx, int cls, int m, int reps, double* ms_out, double* bytes_out) {
  if (!ctx) return SPIS_E_INVALID;
  REQUIRE(ms_out && reps >= 1 && m >= 0 && m <= ctx->kmax, "bad arguments");
  CU(cudaSetDevice(ctx->device));
  const size_t ld = (size_t)ctx->ld;
  // resident pseudo-random operands; pads stay zero because only [0,n) of each row is filled
  for (int i = 0; i <= (m < ctx->kmax ? m : ctx->kmax); ++i)
    fill_kernel<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(ctx->V + (size_t)i * ld, ctx->n, 0x1234ull + (uint64_t)i);
  fill_kernel<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(ctx->W, ctx->n, 0x9999ull);
  fill_kernel<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(ctx->d_y, ctx->K, 0x77ull);
  CU(cudaGetLastError());
  double* scal = ctx->d_small + 2 * ctx->K;
  const double one = 1.0;
  CU(cudaMemcpyAsync(scal + 4, &one, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  const int saved_profile = ctx->profile;
  spis_allreduce_fn saved_ar = ctx->allreduce; spis_halo_fn saved_halo = ctx->halo;
  const bool saved_x = ctx->xactive;
  ctx->profile = 0; ctx->allreduce = nullptr; ctx->halo = nullptr; ctx->xactive = false;
  double bytes0[SPIS_PROF_CLASSES]; for (int i = 0; i < SPIS_PROF_CLASSES; ++i) bytes0[i] = ctx->prof_bytes[i];
  int rc = SPIS_OK;
  for (int rep = -2; rep < reps && rc == SPIS_OK; ++rep) {
    if (rep == 0) { for (int i = 0; i < SPIS_PROF_CLASSES; ++i) bytes0[i] = ctx->prof_bytes[i]; cudaEventRecord(e0, ctx->stream); }
    switch (cls) {
      case SPIS_PROF_SPMV: rc = launch_spmv(ctx, SPIS_SLOT_A, m % 3, ctx->V, ctx->W, ctx->T, scal + 3); break;   // m selects the mode
      case SPIS_PROF_MDOT:
        if (ctx->bench_mdotm_nw) {       // right-hand sides = the last nw basis rows (resident, distinct from rows 0..m)
          if (!ctx->G) { rc = dalloc(ctx, &ctx->G, 4 * ld); if (rc != SPIS_OK) break; for (int q = 0; q < 4; ++q) fill_kernel<<<ctx->nsm * 8, 256, 0, ctx->stream>>>(ctx->G + (size_t)q * ld, ctx->n, 0x4242ull + q); }
          rc = launch_mdotm(ctx, ctx->bench_mdotm_nw, ctx->V, m, nullptr, ctx->G, (int64_t)ld, ctx->d_cout);
        } else rc = launch_mdot(ctx, ctx->V, m, nullptr, 0, ctx->W, ctx->d_small);
        break;
      case SPIS_PROF_LINCOMB: rc = launch_lincomb(ctx, ctx->V, m, ctx->d_y, nullptr, -1.0, ctx->W, ctx->T, 1, scal + 3); break;
      case SPIS_PROF_SCALE: rc = launch_scale(ctx, ctx->V, scal + 4, nullptr, nullptr); break;
      case SPIS_PROF_ORTHMID: rc = launch_orth_mid(ctx, ctx->V, m, ctx->d_y, ctx->W, ctx->d_small); break;
      case SPIS_PROF_PRECOND: rc = launch_precond(ctx, ctx->V, ctx->T); break;
      default: rc = fail(ctx, SPIS_E_INVALID, "class %d cannot be benchmarked", cls);
    }
  }
  cudaEventRecord(e1, ctx->stream);
  ctx->profile = saved_profile; ctx->allreduce = saved_ar; ctx->halo = saved_halo; ctx->xactive = saved_x;
  if (rc != SPIS_OK) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
  CU(cudaEventSynchronize(e1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_out = (double)ms / reps;
  if (bytes_out) *bytes_out = (ctx->prof_bytes[cls] - bytes0[cls]) / reps;
  return SPIS_OK;
}

}  // extern "C"
