// l2s_b200 C ABI (include/l2s_b200.h): handle management and kernel launches.  No torch, no Python: a plain
// CUDA shared library (libl2s_b200.so) that the Python drop-in binds with ctypes.
#include "../../include/l2s_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "l2s_kernels.cuh"

using namespace l2s;

namespace {

constexpr size_t kSmemPerSm = 233472;      // 228 KB
constexpr size_t kSmemPerCtaMax = 232448;  // 227 KB opt-in
constexpr size_t kSmemCtaReserve = 1024;

inline int ru(long long x, int a) { return (int)((x + a - 1) / a * a); }

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "l2s_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return L2S_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

Layout make_layout(int n_j, int n_m, bool stage_durw = false, int pmax = 0, bool decouple = false) {
  Layout L;
  memset(&L, 0, sizeof(L));
  L.n_j = n_j;
  L.n_m = n_m;
  L.n = n_j * n_m;
  L.n1 = L.n + 2;
  L.n1p = ru(L.n1, 4);
  L.walk_stride = ru(L.n, 4);
  L.node_stride = ru(L.n1, 8);
  L.mch_stride = ru(L.n, 16);
  L.e_mc = n_m * (n_j - 1);
  int off = 0;
  L.off_tq = off;   off += 2 * L.n1p * 4;
  L.off_walk = off; off += L.walk_stride * 4;
  L.off_crit = off; off += L.node_stride * 2;
  L.off_misc = off; off += 32;   // parked scalars (8 B) + two mbarriers
  // small instances (warp-per-instance step kernel) keep the static node words in shared memory too (cheaper
  // feature loop); large ones (CTA-per-instance kernel) read them from HBM so that one more instance fits an SM
  L.off_durw = -1;
  if (stage_durw) {
    L.off_durw = off;
    off += L.node_stride * 2;
  }
  // host record staging (4 + pmax words): the critical-predecessor array is dead by then; tiny instances get
  // their own region
  // ... and so do the n5_moves scratch bits (one per path position), behind the record staging
  const int rec_bytes = ru(4 + pmax, 16) * 4, bits_bytes = ru(4 * (L.n / 32 + 2), 16);
  L.off_rec = L.off_crit;
  L.off_bits = L.off_crit + rec_bytes;
  if (rec_bytes + bits_bytes > L.node_stride * 2) {
    L.off_rec = off;
    off += rec_bytes;
    L.off_bits = L.off_crit;
    if (bits_bytes > L.node_stride * 2) {   // (only degenerate shapes)
      L.off_bits = off;
      off += bits_bytes;
    }
  }
  // latency regime of the CTA kernel: shared memory is plentiful, so the arrays that otherwise overlay the dead halves
  // of tq get regions of their own and the outputs can be written while warp 0 traces the critical path
  L.off_msu = L.off_rp = L.off_j2 = L.off_j4 = -1;
  if (decouple) {
    off = ru(off, 16);
    L.off_msu = off; off += L.node_stride * 2;
    L.off_j2 = off;  off += L.node_stride * 2;
    L.off_j4 = off;  off += L.node_stride * 2;
    L.off_rp = off;  off += ru(L.n * 2, 16);
  }
  L.magic_nm = (unsigned)((1ull << 32) / (unsigned)n_m) + 1u;
  L.bytes = ru(off, 16);
  return L;
}

// warps (= instances) per CTA that maximise resident warps per SM; 0 if one instance does not fit
int pick_warps(size_t bytes_per_inst, int max_w = 16) {
  int best = 0, best_w = 0;
  const int cand[] = {1, 2, 4, 8, 16};
  for (int w : cand) {
    if (w > max_w) break;
    const size_t need = (size_t)w * bytes_per_inst;
    if (need > kSmemPerCtaMax) continue;
    long long ctas = (long long)(kSmemPerSm / (need + kSmemCtaReserve));
    if (ctas > 32) ctas = 32;
    long long warps = ctas * w;
    if (warps > 64) warps = 64;
    if (warps > best) { best = (int)warps; best_w = w; }
  }
  return best_w;
}

// Smallest number of FMA residual corrections (1 or 2) after which x*r reproduces the IEEE quotient x/c for
// EVERY integer numerator 0 <= x <= max_num (the only values the kernels ever divide), or 0 if neither does
// (then the kernels fall back to __fdiv_rn).  Same operations as div_rn() on the device: fmaf is correctly
// rounded on both sides.
int verify_divisor(float c, float* r_out, int max_num) {
  if (!(c > 0.f) || !(c < 1e30f)) { *r_out = 0.f; return 0; }
  const float r = (float)(1.0 / (double)c);   // correctly rounded: double rounding cannot occur for a float c
  *r_out = r;
  for (int mode = 1; mode <= 2; ++mode) {
    bool ok = true;
    for (int n = 0; n <= max_num && ok; ++n) {
      const float x = (float)n;
      float q = x * r;
      q = fmaf(fmaf(-c, q, x), r, q);
      if (mode == 2) q = fmaf(fmaf(-c, q, x), r, q);
      ok = (q == x / c);
    }
    if (ok) return mode;
  }
  return 0;
}

}  // namespace

struct l2s_handle_s {
  int device;
  int B, pmax;
  Layout L;
  State st;
  long long inst_off;
  long long itr;
  int warps_step, warps_run, warps_init;
  int cta_warps;   // > 0: large instances -> step_kernel_cta with this many warps per instance
  int fast_sweep;  // CTA kernel: latency variant of the sweep (at most 4 instances resident per SM)
  int incremental; // CTA kernel: keep fin / q of the last evaluation in HBM and resume the sweeps from them
  size_t run_stride, run_stride_greedy;
  int run_off_where, run_off_plist, run_off_walk2;
  int warps_run_greedy;
  InitLayout IL;
  // host-driven step: pinned action staging and the per-instance result records
  int32_t* d_actions;
  int32_t* h_actions;          // pinned + mapped
  int32_t* h_actions_dev;      // device-side address of h_actions
  uint32_t* h_rec;             // pinned + mapped: [B][4 + pmax] words, written by the step kernel
  uint32_t* h_rec_dev;         // device-side address of h_rec
  uint32_t* d_rec;             // device staging, only for rec_direct == 0
  size_t rec_bytes;
  int rec_stride;              // words per record: 4 + pmax rounded up to a multiple of 16 (64-byte lines)
  int rec_lines;               // 1: the kernel writes whole 64-byte lines (padding words are stale)
  int act_prefetch;            // 1: a small kernel copies the host actions to the device with 16-byte coalesced reads
  // divisors already verified for the FMA division (x features)
  float div_c[2], div_r[2];
  int div_mode[2];
  // zero-copy host path (defaults 1/1; L2S_HOST_ZEROCOPY=0 stages the actions with an H2D copy,
  // L2S_HOST_RECORDS=0 writes the records to device memory and copies them back with one bulk D2H)
  int zero_copy;
  int rec_direct;
  // completion of a host-driven step: 0 = cudaStreamSynchronize; 1 = a stream-ordered 32-bit write into mapped pinned
  // host memory (cuStreamWriteValue32, fenced behind the kernel's own host writes) that the calling thread polls
  int wait_flag;
  int host_debug;              // experiments (L2S_HOST_DEBUG): 1 = the kernel writes no host records, 2 = actions are
                               // taken from device memory without any transfer (both make the results meaningless)
  uint32_t* h_flag;            // pinned + mapped
  uint32_t* h_flag_dev;
  uint32_t flag_epoch;
  // l2s_step_host_begin .. l2s_step_host_wait
  cudaEvent_t ev_done;
  int pending;                 // a begun step has not been waited for yet
  cudaStream_t pending_stream;
};

// cuStreamWriteValue32 through the runtime's driver entry-point lookup (no link-time dependency on libcuda)
typedef int (*l2s_write32_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static l2s_write32_fn resolve_write32() {
  static l2s_write32_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = (l2s_write32_fn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

extern "C" {

const char* l2s_status_string(int s) {
  switch (s) {
    case L2S_OK: return "ok";
    case L2S_ERR_INVALID_ARG: return "invalid argument";
    case L2S_ERR_CUDA: return "CUDA error";
    case L2S_ERR_UNSUPPORTED_SHAPE: return "unsupported shape (n_m > 32, n+2 > 65535 or state exceeds shared memory)";
    case L2S_ERR_ILLEGAL_ACTION: return "illegal action: not an adjacent pair on one machine";
    case L2S_ERR_BAD_GRAPH: return "cyclic or malformed disjunctive graph";
    case L2S_ERR_DURATION_RANGE: return "duration out of range [1,4095] or total duration >= 2^24";
    case L2S_ERR_MACHINE_ROWS: return "every job must visit machines 1..n_m exactly once";
    case L2S_ERR_PAIR_OVERFLOW: return "more N5 pairs than pmax";
    case L2S_ERR_SHORT_PATH: return "critical path of two operations on one machine (reference raises IndexError)";
    case L2S_ERR_BAD_SOLUTION: return "machine order rows are not the operations of that machine";
    default: return "unknown status";
  }
}

int l2s_version(void) { return 100; }

#ifdef L2S_TIMELINE
int l2s_debug_timeline(unsigned long long* dev_buf) {
  return cudaMemcpyToSymbol(g_timeline, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -2;
}
int l2s_debug_levels(unsigned long long* out2, int reset) {   // {sum of levels, sweeps} since the last reset
  if (out2 && cudaMemcpyFromSymbol(out2, g_dbg_levels, 16) != cudaSuccess) return -2;
  if (reset) { unsigned long long z[2] = {0, 0}; if (cudaMemcpyToSymbol(g_dbg_levels, z, 16) != cudaSuccess) return -2; }
  return 0;
}
#endif

static int create_into(l2s_handle h, int n_j, int n_m, int batch, int pmax, int device) {
  CU(cudaSetDevice(device));
  h->device = device;
  h->B = batch;
  h->pmax = pmax;
  // kernel + layout choice: the warp-per-instance step kernel (with staged node words) if at least 16 such
  // instances fit an SM, else one CTA per instance
  h->L = make_layout(n_j, n_m, true, pmax);
  // With more than 16 machines the warp kernel runs its forward and backward sweeps one after the other, the CTA
  // kernel side by side on two warps: for small batches (latency regime) the CTA kernel is the faster one there
  // (30x20, 512 instances: 15.4 vs 18.5 us per step; equal at 4096; with n_m <= 16 the warp kernel wins at any batch)
  const bool fits16 = (long long)(kSmemPerSm / ((size_t)h->L.bytes + kSmemCtaReserve)) >= 16;
  const bool small = fits16 && (n_m <= 16 || ((long long)batch + 147) / 148 > 16);
  if (!small) h->L = make_layout(n_j, n_m, false, pmax);
  const Layout& L = h->L;
  h->run_off_where = L.bytes;
  h->run_off_plist = L.bytes + L.node_stride * 2;
  h->run_stride = ru((long long)h->run_off_plist + (long long)pmax * 4, 16);
  // GREEDY with n_m <= 16 evaluates two neighbours per sweep and keeps a second copy of the walk words for that
  h->run_off_walk2 = n_m <= 16 ? (int)h->run_stride : -1;
  h->run_stride_greedy = n_m <= 16 ? h->run_stride + (size_t)L.walk_stride * 4 + 16 : h->run_stride;   // +16: a sweep prefetches one word past its last row
  // initial-solution builder layout
  {
    InitLayout& I = h->IL;
    int off = 0;
    I.off_starts = off; off += ru((long long)L.n * 4, 16);
    I.off_ops = off;    off += ru((long long)L.n * 2, 16);
    I.off_cnt = off;    off += ru((long long)n_m * 4, 16);
    I.off_nxt = off;    off += ru((long long)n_j * 4, 16);
    I.off_rdy = off;    off += ru((long long)n_j * 4, 16);
    I.off_work = off;   off += ru((long long)n_j * 4, 16);
    I.off_tot = off;    off += ru((long long)n_j * 4, 16);
    I.off_pri = off;    off += ru((long long)n_j * 8, 16);
    I.bytes = off;
  }
  h->warps_step = pick_warps(L.bytes, 4);   // step_kernel is compiled with __launch_bounds__(128, 8)
  if (const char* e = getenv("L2S_WARPS_STEP")) { int v = atoi(e); if (v > 0 && v <= 4 && (size_t)v * L.bytes <= kSmemPerCtaMax) h->warps_step = v; }
  h->warps_run = pick_warps(h->run_stride);
  h->warps_run_greedy = pick_warps(h->run_stride_greedy);
  if (!h->warps_run_greedy) { h->warps_run_greedy = h->warps_run; h->run_stride_greedy = h->run_stride; h->run_off_walk2 = -1; }
  h->warps_init = pick_warps(h->IL.bytes);
  if (!h->warps_step || !h->warps_run || !h->warps_init) return L2S_ERR_UNSUPPORTED_SHAPE;
  // few instances per SM (large graphs): give every instance a whole CTA so that the SM is not left idle
  {
    const long long per_sm = (long long)(kSmemPerSm / ((size_t)L.bytes + kSmemCtaReserve));
    // warps per instance so that the co-resident instances of an SM add up to about 32 warps (measured, after the
    // sweeps were fixed to run converged: 50x20 B=2048 best at 2, 100x20 B=1024 at 4, 150x25 B=1024 at 8;
    // 30x20 and smaller: warp per instance)
    long long resident = (batch + 147) / 148;
    if (resident > per_sm) resident = per_sm;
    if (resident < 1) resident = 1;
    h->cta_warps = small ? 0 : (resident >= 12 ? 2 : resident >= 5 ? 4 : 8);
    if (const char* e = getenv("L2S_CTA_WARPS")) {      // experiments: force the CTA kernel (0 keeps the choice above)
      const int v = atoi(e);
      if (v > 0) { h->cta_warps = v; }
    }
    if (h->cta_warps > 8) h->cta_warps = 8;
    // few resident instances: the sweep warps are alone on their schedulers and a level costs its dependent chain ->
    // latency variant; many: the schedulers are saturated by sweep instructions -> lean variant (l2s_device.cuh)
    // (2 = latency variant with two-operation look-ahead: fewer, longer levels; pays for long machine rows when an
    // instance has an SM to itself: 150x25 at 128 instances 21.6 -> 20.6 us, 100x20 16.5 -> 16.4, 30x20 12.3 -> 12.8)
    h->fast_sweep = h->cta_warps > 0 && resident <= 4 ? (resident <= 2 && n_j >= 100 ? 2 : 1) : 0;
    if (const char* e = getenv("L2S_FAST_SWEEP")) h->fast_sweep = atoi(e);      // 0 lean, 1 latency, 2 latency + look-ahead
    // ... and shared memory is plentiful: decoupled regions (make_layout) if the resident instances still fit
    bool decouple = h->cta_warps > 1 && resident <= 4;
    if (const char* e = getenv("L2S_DECOUPLE")) decouple = h->cta_warps > 1 && atoi(e) != 0;
    if (decouple) {
      const Layout L2 = make_layout(n_j, n_m, false, pmax, true);
      if ((long long)(kSmemPerSm / ((size_t)L2.bytes + kSmemCtaReserve)) >= resident && (size_t)L2.bytes + 1024 <= kSmemPerCtaMax)
        h->L = L2;
    }
  }
  CU(cudaFuncSetAttribute(step_kernel_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPerCtaMax - 1024));   // it also has static shared memory
  CU(cudaFuncSetAttribute(step_kernel_cta, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPerCtaMax));
  CU(cudaFuncSetAttribute(run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPerCtaMax));
  CU(cudaFuncSetAttribute(build_initial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPerCtaMax));
  // these kernels live in shared memory, not in L1: ask for the largest shared-memory carve-out
  CU(cudaFuncSetAttribute(step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(run_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(build_initial_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const size_t B = batch;
  State& st = h->st;
  CU(cudaMalloc(&st.durw, B * L.node_stride * 2));
  CU(cudaMalloc(&st.mch, B * L.mch_stride));
  CU(cudaMalloc(&st.walk, B * L.walk_stride * 4));
  CU(cudaMalloc(&st.where, B * L.node_stride * 2));
  CU(cudaMalloc(&st.tabu, B * 8));
  CU(cudaMalloc(&st.cur, B * 4));
  CU(cudaMalloc(&st.inc, B * 4));
  CU(cudaMalloc(&st.err, 4));
  // incremental re-evaluation behind the step API: only the CTA kernel (large instances, where sweep levels are a
  // third to a half of a step) keeps the last evaluation; L2S_INCREMENTAL=0 turns it off (experiments, tests)
  // Measured: the sweep depth of a resumed evaluation halves, but the forward and the backward sweep
  // run side by side, so the wall-clock depth is max(u, 1-u) ~ 0.75 of a full one and the saving (0.6 us of 7.8 at
  // 100x20) is eaten by the extra state traffic and passes: OFF by default, l2s_set_option("incremental", 1) or
  // L2S_INCREMENTAL=1 turns it on.
  if (h->cta_warps > 0) {
    CU(cudaMalloc(&st.fq, B * 2 * (size_t)L.n1p * 4));
    CU(cudaMalloc(&st.evalid, B * 4));
    CU(cudaMemset(st.evalid, 0, B * 4));
  }
  h->incremental = 0;
  if (const char* e = getenv("L2S_INCREMENTAL")) h->incremental = st.fq != nullptr && atoi(e) != 0;
  {
    int32_t* page = nullptr;
    CU(cudaMalloc(&page, (size_t)2 * L.n1p * 4));
    CU(cudaMemset(page, 0xff, (size_t)2 * L.n1p * 4));
    st.pending_page = page;
  }
  CU(cudaMemset(st.durw, 0, B * L.node_stride * 2));
  CU(cudaMemset(st.mch, 0, B * L.mch_stride));
  CU(cudaMemset(st.walk, 0, B * L.walk_stride * 4));
  CU(cudaMemset(st.where, 0, B * L.node_stride * 2));
  CU(cudaMemset(st.tabu, 0, B * 8));
  CU(cudaMemset(st.cur, 0, B * 4));
  CU(cudaMemset(st.inc, 0, B * 4));
  CU(cudaMemset(st.err, 0, 4));
  // host-step staging.  Measured on B200 (20x15, 8192 instances per step): records written by the kernel straight
  // into mapped host memory (one coalesced store of 16 + 4*npairs bytes per instance, overlapping the rest of the
  // launch) 57 us per step; packed SoA block + one bulk D2H 84 us; scattered 4-byte zero-copy stores: slower still.
  h->rec_stride = ru(L2S_REC_HEADER_WORDS + pmax, 16);     // whole 64-byte lines per record
  h->rec_bytes = B * (size_t)h->rec_stride * 4;
  CU(cudaHostAlloc((void**)&h->h_rec, h->rec_bytes, cudaHostAllocMapped));
  memset(h->h_rec, 0, h->rec_bytes);
  CU(cudaHostGetDevicePointer((void**)&h->h_rec_dev, h->h_rec, 0));
  CU(cudaMalloc(&h->d_rec, h->rec_bytes));
  CU(cudaMemset(h->d_rec, 0, h->rec_bytes));
  CU(cudaMalloc(&h->d_actions, B * 8 + 16));
  CU(cudaHostAlloc((void**)&h->h_actions, B * 8, cudaHostAllocMapped));
  memset(h->h_actions, 0, B * 8);
  CU(cudaHostGetDevicePointer((void**)&h->h_actions_dev, h->h_actions, 0));
  h->zero_copy = 1;
  h->rec_direct = 1;
  if (const char* e = getenv("L2S_HOST_ZEROCOPY")) h->zero_copy = atoi(e) != 0;
  if (const char* e = getenv("L2S_HOST_RECORDS")) h->rec_direct = atoi(e) != 0;
  CU(cudaHostAlloc((void**)&h->h_flag, 64, cudaHostAllocMapped));
  memset(h->h_flag, 0, 64);
  CU(cudaHostGetDevicePointer((void**)&h->h_flag_dev, h->h_flag, 0));
  h->flag_epoch = 0;
  CU(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  h->pending = 0;
  h->wait_flag = 0;
  h->host_debug = 0;
  h->rec_lines = 0;
  h->act_prefetch = 0;
  if (const char* e = getenv("L2S_HOST_DEBUG")) h->host_debug = atoi(e);
  if (const char* e = getenv("L2S_HOST_LINES")) h->rec_lines = atoi(e) != 0;
  if (const char* e = getenv("L2S_HOST_PREFETCH")) h->act_prefetch = atoi(e) != 0;
  if (const char* e = getenv("L2S_HOST_WAIT")) h->wait_flag = (strcmp(e, "flag") == 0) && resolve_write32() != nullptr;
  return L2S_OK;
}


int l2s_destroy(l2s_handle h);

int l2s_create(int n_j, int n_m, int batch, int pmax, int device, l2s_handle* out) {
  if (!out || n_j < 2 || n_m < 1 || batch < 1 || pmax < 0) return L2S_ERR_INVALID_ARG;
  if (pmax == 0) pmax = 32;
  if (n_m > 32 || (long long)n_j * n_m + 2 > 65535) return L2S_ERR_UNSUPPORTED_SHAPE;
  if ((long long)batch * ((long long)n_j * n_m + 2) >= (1ll << 31)) return L2S_ERR_INVALID_ARG;   // 32-bit node ids per slice
  l2s_handle h = (l2s_handle)calloc(1, sizeof(l2s_handle_s));
  if (!h) return L2S_ERR_INVALID_ARG;
  h->device = device;
  const int r = create_into(h, n_j, n_m, batch, pmax, device);
  if (r != L2S_OK) {   // release whatever was allocated before the failure
    l2s_destroy(h);
    return r;
  }
  *out = h;
  return L2S_OK;
}

int l2s_destroy(l2s_handle h) {
  if (!h) return L2S_OK;
  cudaSetDevice(h->device);
  cudaFree(h->st.durw); cudaFree(h->st.mch); cudaFree(h->st.walk); cudaFree(h->st.where);
  cudaFree(h->st.tabu); cudaFree(h->st.cur); cudaFree(h->st.inc); cudaFree(h->st.err);
  cudaFree((void*)h->st.pending_page);
  cudaFree(h->st.fq); cudaFree(h->st.evalid);
  cudaFree(h->d_rec); cudaFreeHost(h->h_rec); cudaFree(h->d_actions); cudaFreeHost(h->h_actions);
  cudaFreeHost(h->h_flag);
  cudaEventDestroy(h->ev_done);
  free(h);
  return L2S_OK;
}

int l2s_pmax(l2s_handle h) { return h ? h->pmax : L2S_ERR_INVALID_ARG; }
int64_t l2s_edges_mc_per_instance(l2s_handle h) { return h ? h->L.e_mc : L2S_ERR_INVALID_ARG; }
int64_t l2s_itr(l2s_handle h) { return h ? h->itr : L2S_ERR_INVALID_ARG; }
int l2s_set_itr(l2s_handle h, int64_t itr) {
  if (!h || itr < 0) return L2S_ERR_INVALID_ARG;
  h->itr = itr;
  return L2S_OK;
}

int l2s_launch_info(l2s_handle h, int* warps_per_cta_step, int* smem_per_instance, int* warps_per_cta_run) {
  if (!h) return L2S_ERR_INVALID_ARG;
  if (warps_per_cta_step) *warps_per_cta_step = h->warps_step;
  if (smem_per_instance) *smem_per_instance = h->L.bytes;
  if (warps_per_cta_run) *warps_per_cta_run = h->warps_run;
  return L2S_OK;
}

static inline dim3 warp_grid(int B, int threads) { return dim3((unsigned)(((long long)B * 32 + threads - 1) / threads)); }

int l2s_load_instances(l2s_handle h, const int32_t* dur, const int32_t* mch, int64_t instance_offset, void* stream) {
  if (!h || !dur || !mch) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  h->inst_off = instance_offset;
  load_kernel<<<warp_grid(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->L, h->st, h->B, dur, mch);
  CU(cudaGetLastError());
  return L2S_OK;
}

int l2s_set_solution(l2s_handle h, const int32_t* mseq, void* stream) {
  if (!h || !mseq) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  h->itr = 0;
  set_solution_kernel<<<warp_grid(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->L, h->st, h->B, mseq);
  CU(cudaGetLastError());
  return L2S_OK;
}

int l2s_get_solution(l2s_handle h, int32_t* mseq, void* stream) {
  if (!h || !mseq) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  get_solution_kernel<<<warp_grid(h->B, 128), 128, 0, (cudaStream_t)stream>>>(h->L, h->st, h->B, mseq);
  CU(cudaGetLastError());
  return L2S_OK;
}

int l2s_build_initial(l2s_handle h, int rule, int high, void* stream) {
  if (!h || (rule != L2S_RULE_SPT && rule != L2S_RULE_FDD_DIVIDE_MWKR)) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  h->itr = 0;
  const int W = h->warps_init;
  build_initial_kernel<<<(h->B + W - 1) / W, W * 32, (size_t)W * h->IL.bytes, (cudaStream_t)stream>>>(
      h->L, h->IL, h->st, h->B, rule, high);
  CU(cudaGetLastError());
  return L2S_OK;
}

static int launch_step(l2s_handle h, const StepArgs& a, void* stream) {
  if (h->cta_warps > 0) {
    step_kernel_cta<<<h->B, h->cta_warps * 32, (size_t)h->L.bytes, (cudaStream_t)stream>>>(a);
    CU(cudaGetLastError());
    return L2S_OK;
  }
  const int W = h->warps_step;
  static int pad = -1;   // experiments: extra dynamic shared memory per CTA (bytes) to cap the CTAs resident per SM
  if (pad < 0) { const char* e = getenv("L2S_STEP_SMEM_PAD"); pad = e ? atoi(e) : 0; }
  step_kernel<<<(h->B + W - 1) / W, W * 32, (size_t)W * h->L.bytes + (size_t)pad, (cudaStream_t)stream>>>(a);
  CU(cudaGetLastError());
  return L2S_OK;
}

static StepArgs step_args(l2s_handle h, const l2s_step_io* io) {
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.L = h->L;
  a.st = h->st;
  a.B = h->B;
  a.pmax = h->pmax;
  a.incremental = h->incremental;
  a.fast_sweep = h->fast_sweep;
  if (io) {
    a.actions = io->actions;
    a.high = io->high;
    a.fea = io->fea_norm_const;
    // feature divisions: dur/high has numerators <= kDurMax, est|lst/fea numerators < 2^24
    if (h->div_c[0] != io->high) { h->div_c[0] = io->high; h->div_mode[0] = verify_divisor(io->high, &h->div_r[0], kDurMax); }
    if (h->div_c[1] != io->fea_norm_const) {
      h->div_c[1] = io->fea_norm_const;
      h->div_mode[1] = verify_divisor(io->fea_norm_const, &h->div_r[1], (1 << 24) - 1);
    }
    a.r_high = h->div_r[0]; a.div_high = h->div_mode[0];
    a.r_fea = h->div_r[1];  a.div_fea = h->div_mode[1];
    a.reward_type = io->reward_type;
    a.is_reset = io->is_reset;
    a.x = io->x;
    a.emc = io->edge_index_mc;
    a.makespan = io->makespan;
    a.incumbent = io->incumbent;
    a.reward = io->reward;
    a.done = io->done;
    a.pairs = io->pairs;
    a.npairs = io->npairs;
    a.status = io->status;
  }
  return a;
}

int l2s_step(l2s_handle h, const l2s_step_io* io, void* stream) {
  if (!h || !io) return L2S_ERR_INVALID_ARG;
  if (io->reward_type != L2S_REWARD_YAOXIN && io->reward_type != L2S_REWARD_CONSECUTIVE) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  StepArgs a = step_args(h, io);
  int r = launch_step(h, a, stream);
  if (r == L2S_OK && !io->is_reset && a.actions) h->itr += 1;
  return r;
}

int l2s_set_option(l2s_handle h, const char* name, int value) {
  if (!h || !name) return L2S_ERR_INVALID_ARG;
  if (!strcmp(name, "host_wait_flag")) h->wait_flag = value != 0 && resolve_write32() != nullptr;
  else if (!strcmp(name, "host_zero_copy")) h->zero_copy = value != 0;
  else if (!strcmp(name, "host_records_direct")) h->rec_direct = value != 0;
  else if (!strcmp(name, "host_record_lines")) h->rec_lines = value != 0;
  else if (!strcmp(name, "host_action_prefetch")) h->act_prefetch = value != 0;
  else if (!strcmp(name, "host_debug")) h->host_debug = value;
  else if (!strcmp(name, "incremental")) h->incremental = value != 0 && h->st.fq != nullptr;
  else if (!strcmp(name, "fast_sweep")) h->fast_sweep = value < 0 ? 0 : value > 2 ? 2 : value;
  else return L2S_ERR_INVALID_ARG;
  return L2S_OK;
}

int l2s_host_records(l2s_handle h, const uint32_t** records, int* stride_words, int32_t** actions) {
  if (!h) return L2S_ERR_INVALID_ARG;
  if (records) *records = h->h_rec;
  if (stride_words) *stride_words = h->rec_stride;
  if (actions) *actions = h->h_actions;
  return L2S_OK;
}

int l2s_step_host(l2s_handle h, const l2s_step_io* io, const int32_t* actions_host, int32_t* pairs_host,
                  int32_t* npairs_host, uint8_t* done_host, float* reward_host, float* makespan_host,
                  int32_t* status_host, void* stream) {
  const int r = l2s_step_host_begin(h, io, actions_host, stream);
  if (r != L2S_OK) return r;
  return l2s_step_host_wait(h, pairs_host, npairs_host, done_host, reward_host, makespan_host, status_host);
}

int l2s_step_host_begin(l2s_handle h, const l2s_step_io* io, const int32_t* actions_host, void* stream) {
  if (!h || !io) return L2S_ERR_INVALID_ARG;
  if (h->pending) return L2S_ERR_INVALID_ARG;   // one step in flight per handle: wait for it first
  if (io->reward_type != L2S_REWARD_YAOXIN && io->reward_type != L2S_REWARD_CONSECUTIVE) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t B = h->B;
  StepArgs a = step_args(h, io);
  a.actions = nullptr;
  // Host path: the kernel reads the 8-byte actions straight from mapped pinned host memory and writes one
  // coalesced result record per instance straight back into mapped pinned host memory: no copy calls at all,
  // the PCIe traffic overlaps the launch.  (L2S_HOST_ZEROCOPY=0 / L2S_HOST_RECORDS=0 stage through device
  // memory with explicit copies instead.)
  if (actions_host) {
    const int32_t* in_place = nullptr;   // device-visible address of the caller's buffer, if it is pinned host memory
    if (actions_host == h->h_actions) {
      in_place = h->h_actions_dev;
    } else if (h->zero_copy) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, actions_host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
        in_place = (const int32_t*)at.devicePointer;
      else
        cudaGetLastError();
    }
    if (!in_place || !h->zero_copy) memcpy(h->h_actions, actions_host, B * 8);   // pageable caller memory: stage it
    const int32_t* src_dev = in_place ? in_place : h->h_actions_dev;
    if (h->zero_copy && h->act_prefetch && !((uintptr_t)src_dev & 15u) && !((B * 8) & 15u)) {
      // one small kernel pulls the whole action block over PCIe with 16-byte coalesced reads (512 requests of 128
      // bytes instead of one 8-byte request per instance from the step kernel's warps)
      const int n16 = (int)((B * 8 + 15) / 16);
      copy16_kernel<<<(n16 + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4*>(in_place ? in_place : h->h_actions_dev),
                                                      reinterpret_cast<uint4*>(h->d_actions), n16);
      CU(cudaGetLastError());
      a.actions = h->d_actions;
    } else if (h->zero_copy) {
      a.actions = in_place ? in_place : h->h_actions_dev;
    } else {
      CU(cudaMemcpyAsync(h->d_actions, h->h_actions, B * 8, cudaMemcpyHostToDevice, s));
      a.actions = h->d_actions;
    }
  }
  a.rec = h->rec_direct ? h->h_rec_dev : h->d_rec;
  a.rec_stride = h->rec_stride;
  a.rec_lines = h->rec_lines;
  if (h->host_debug & 1) a.rec = h->d_rec;
  if ((h->host_debug & 2) && a.actions) a.actions = h->d_actions;
  int r = launch_step(h, a, stream);
  if (r != L2S_OK) return r;
  if (!io->is_reset && a.actions) h->itr += 1;
  if (!h->rec_direct) CU(cudaMemcpyAsync(h->h_rec, h->d_rec, h->rec_bytes, cudaMemcpyDeviceToHost, s));
  if (h->wait_flag) {
    if (resolve_write32()(s, (unsigned long long)(uintptr_t)h->h_flag_dev, ++h->flag_epoch, 0u) != 0) return L2S_ERR_CUDA;
  } else {
    CU(cudaEventRecord(h->ev_done, s));
  }
  h->pending = 1;
  h->pending_stream = s;
  return L2S_OK;
}

int l2s_step_host_wait(l2s_handle h, int32_t* pairs_host, int32_t* npairs_host, uint8_t* done_host, float* reward_host,
                       float* makespan_host, int32_t* status_host) {
  if (!h || !h->pending) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  const size_t B = h->B;
  cudaStream_t s = h->pending_stream;
  h->pending = 0;
  if (h->wait_flag) {
    const uint32_t epoch = h->flag_epoch;
    volatile uint32_t* flag = h->h_flag;
    for (unsigned long long spins = 1; *flag != epoch; ++spins) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((spins & 0xfffffull) == 0) {   // every ~million polls: a faulted launch would never write the flag
        const cudaError_t q = cudaStreamQuery(s);
        if (q != cudaSuccess && q != cudaErrorNotReady) {
          fprintf(stderr, "l2s_b200: CUDA error %s while waiting for a step\n", cudaGetErrorString(q));
          return L2S_ERR_CUDA;
        }
      }
    }
  } else {
    CU(cudaEventSynchronize(h->ev_done));
  }
  if (pairs_host || npairs_host || status_host || reward_host || makespan_host || done_host) {
    const size_t stride = (size_t)h->rec_stride;
    for (size_t b = 0; b < B; ++b) {   // unpack the records; only the valid prefix of every pair row is defined
      const uint32_t* rec = h->h_rec + b * stride;
      const int np = (int)L2S_REC_NPAIRS(rec[0]);
      if (pairs_host) {
        int32_t* row = pairs_host + b * (size_t)h->pmax * 2;
        for (int i = 0; i < np; ++i) {
          row[2 * i] = (int32_t)(rec[L2S_REC_HEADER_WORDS + i] & 0xffffu);
          row[2 * i + 1] = (int32_t)(rec[L2S_REC_HEADER_WORDS + i] >> 16);
        }
      }
      if (npairs_host) npairs_host[b] = np;
      if (status_host) status_host[b] = L2S_REC_STATUS(rec[0]);
      if (done_host) done_host[b] = (uint8_t)L2S_REC_DONE(rec[0]);
      if (makespan_host) memcpy(makespan_host + b, rec + 1, 4);
      if (reward_host) memcpy(reward_host + b, rec + 2, 4);
    }
  }
  return L2S_OK;
}

int l2s_run(l2s_handle h, int policy, int K, const int32_t* tape, uint64_t seed, int32_t* taken,
            const int32_t* log_steps_host, int n_log, float* incumbent_log, void* stream) {
  if (!h || K < 0 || policy < 0 || policy > 2 || n_log < 0 || n_log > kMaxLog) return L2S_ERR_INVALID_ARG;
  if (policy == L2S_POLICY_REPLAY && !tape) return L2S_ERR_INVALID_ARG;
  if (n_log > 0 && (!log_steps_host || !incumbent_log)) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  RunArgs a;
  memset(&a, 0, sizeof(a));
  a.L = h->L;
  a.st = h->st;
  a.B = h->B;
  a.policy = policy;
  a.K = K;
  a.tape = tape;
  a.seed = seed;
  a.inst_off = h->inst_off;
  a.itr0 = h->itr;
  a.taken = taken;
  a.n_log = n_log;
  for (int i = 0; i < n_log; ++i) {
    if (i && log_steps_host[i] <= log_steps_host[i - 1]) return L2S_ERR_INVALID_ARG;
    a.log_steps[i] = log_steps_host[i];
  }
  a.inc_log = incumbent_log;
  a.pmax = h->pmax;
  a.off_where = h->run_off_where;
  a.off_plist = h->run_off_plist;
  const bool two = policy == L2S_POLICY_GREEDY && h->run_off_walk2 >= 0;
  const size_t stride = two ? h->run_stride_greedy : h->run_stride;
  a.off_walk2 = two ? h->run_off_walk2 : -1;
  a.stride = (int)stride;
  const int W = two ? h->warps_run_greedy : h->warps_run;
  run_kernel<<<(h->B + W - 1) / W, W * 32, (size_t)W * stride, (cudaStream_t)stream>>>(a);
  CU(cudaGetLastError());
  h->itr += K;
  return L2S_OK;
}

int l2s_export_state(l2s_handle h, int32_t* est, int32_t* lst, int32_t* path, int32_t* path_len, int32_t* tabu,
                     uint8_t* jobfirst, float* makespan, float* incumbent, void* stream) {
  if (!h) return L2S_ERR_INVALID_ARG;
  if ((est == nullptr) != (lst == nullptr) || (path == nullptr) != (path_len == nullptr)) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (est || path) {
    StepArgs a = step_args(h, nullptr);
    a.export_only = 1;
    a.high = 1.f;
    a.fea = 1.f;
    a.ex_est = est;
    a.ex_lst = lst;
    a.ex_path = path;
    a.ex_path_len = path_len;
    int r = launch_step(h, a, stream);
    if (r != L2S_OK) return r;
  }
  if (tabu) CU(cudaMemcpyAsync(tabu, h->st.tabu, (size_t)h->B * 8, cudaMemcpyDeviceToDevice, s));
  if (jobfirst || makespan || incumbent) {
    export_misc_kernel<<<warp_grid(h->B, 128), 128, 0, s>>>(h->L, h->st, h->B, jobfirst, makespan, incumbent);
    CU(cudaGetLastError());
  }
  return L2S_OK;
}

int l2s_poll_error(l2s_handle h, int* instance_out, void* stream) {
  if (!h) return L2S_ERR_INVALID_ARG;
  CU(cudaSetDevice(h->device));
  int word = 0;
  CU(cudaMemcpyAsync(&word, h->st.err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  if (!word) return L2S_OK;
  CU(cudaMemsetAsync(h->st.err, 0, 4, (cudaStream_t)stream));
  if (instance_out) *instance_out = word & 0xffffff;
  return -(int)(((unsigned)word) >> 24);
}

// ---- stateless COO evaluator -------------------------------------------------------------------------

int64_t l2s_eval_coo_workspace_bytes(int n_j, int n_m, int64_t batch) {
  if (n_j < 1 || n_m < 1 || batch < 1) return L2S_ERR_INVALID_ARG;
  const long long node_stride = ru((long long)n_j * n_m + 2, 8);
  return batch * node_stride * 2 + 16;
}

int l2s_eval_coo(const int64_t* edge_index, int64_t n_edges, const void* duration, int duration_is_i64, int n_j,
                 int n_m, int64_t batch, float* est, float* lst, float* makespan, void* workspace,
                 int64_t workspace_bytes, int workspace_is_clean, int32_t* status_flag, int device, void* stream) {
  if (!edge_index || !duration || !est || !lst || !makespan || !workspace || !status_flag || batch < 1 ||
      n_j < 2 || n_m < 1 || n_edges < 0)
    return L2S_ERR_INVALID_ARG;
  const long long n = (long long)n_j * n_m;
  if (n + 2 > 65535 || n_m > 32) return L2S_ERR_UNSUPPORTED_SHAPE;
  if (batch * (n + 2) >= (1ll << 31)) return L2S_ERR_INVALID_ARG;   // 32-bit node ids per call
  if (workspace_bytes < l2s_eval_coo_workspace_bytes(n_j, n_m, batch) || ((unsigned long long)workspace & 15ull))
    return L2S_ERR_INVALID_ARG;
  // a batch of complete selections has exactly this many edges (duplicates are caught on the device)
  if (n_edges != batch * ((long long)n_j * (n_m + 1) + (long long)n_m * (n_j - 1))) return L2S_ERR_BAD_GRAPH;
  CU(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  CooEvalArgs a;
  memset(&a, 0, sizeof(a));
  a.L = make_layout(n_j, n_m);
  const Layout& L = a.L;
  a.B = batch;
  a.duration = duration;
  a.dur_i64 = duration_is_i64;
  a.dur_tma = duration_is_i64 && ((n + 2) % 2 == 0) && (((unsigned long long)duration & 15ull) == 0);
  if (const char* e = getenv("L2S_EVAL_DUR_TMA")) a.dur_tma = a.dur_tma && atoi(e) != 0;   // experiments
  const size_t links = (size_t)batch * L.node_stride * 2;
  a.ws.msucc = (uint16_t*)workspace;
  a.ws.conj_count = (unsigned long long*)((char*)workspace + links);
  a.ws.flag = status_flag;
  a.est = est; a.lst = lst; a.ms = makespan;
  a.expect_conj = batch * ((long long)n_j * (n_m + 1));
  // shared memory: tq | walk (offsets of the env layout), then node words, successors and the head list
  int off = L.off_crit;
  a.off_durw = off;  off += L.node_stride * 2;
  a.off_succ = off;  off += L.node_stride * 2;
  a.off_heads = off; off += 64;
  a.off_bar = off;   off += 16;
  a.bytes = ru(off, 16);
  // warp per graph while at least 16 graphs fit an SM, else one CTA per graph (coo_eval_kernel_cta) -- except for the
  // band where 4-7 graphs fit and the batch keeps them all busy (100x20 at batch 1024: warp kernel 52 us, CTA kernel
  // 56-69 us; measured table in profiles/README_r2.md)
  const long long fits_warp = (long long)(kSmemPerSm / ((size_t)a.bytes + kSmemCtaReserve));
  const long long busy_warp = fits_warp < (batch + 147) / 148 ? fits_warp : (batch + 147) / 148;
  bool cta = fits_warp < 16 && !(fits_warp >= 4 && fits_warp < 8 && busy_warp > 2);
  if (const char* e = getenv("L2S_EVAL_CTA")) cta = atoi(e) != 0;     // experiments / tests: 0 = warp kernel, 2|4|8 = CTA kernel
  int W = 0, cta_warps = 0;
  if (cta) {
    a.off_anchor = a.bytes;
    a.bytes = ru((long long)a.bytes + 2ll * n_m * ((n_j + 3) / 4), 16);
    if ((size_t)a.bytes + 1024 > kSmemPerCtaMax) return L2S_ERR_UNSUPPORTED_SHAPE;
    const long long per_sm = (long long)(kSmemPerSm / ((size_t)a.bytes + kSmemCtaReserve));
    long long resident = (batch + 147) / 148;
    if (resident > per_sm) resident = per_sm;
    if (resident < 1) resident = 1;
    cta_warps = resident >= 12 ? 2 : resident >= 5 ? 4 : 8;
    if (const char* e = getenv("L2S_EVAL_CTA")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) cta_warps = v; }
    a.fast_sweep = resident <= 4 ? (resident <= 2 && n_j >= 100 ? 2 : 1) : 0;
    if (const char* e = getenv("L2S_FAST_SWEEP")) a.fast_sweep = atoi(e);
  } else {
    W = pick_warps(a.bytes, 4);
    if (!W) return L2S_ERR_UNSUPPORTED_SHAPE;
  }
  static bool attr_set[64] = {false};
  if (device >= 0 && device < 64 && !attr_set[device]) {
    CU(cudaFuncSetAttribute(coo_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPerCtaMax));
    CU(cudaFuncSetAttribute(coo_eval_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(coo_eval_kernel_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPerCtaMax - 1024));
    CU(cudaFuncSetAttribute(coo_eval_kernel_cta, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr_set[device] = true;
  }
  if (!workspace_is_clean) CU(cudaMemsetAsync(workspace, 0, links + 16, s));
  CooIngestArgs g;
  memset(&g, 0, sizeof(g));
  g.ei = (const long long*)edge_index;
  g.E = n_edges; g.B = batch; g.n_j = n_j; g.n_m = n_m; g.n = (int)n; g.n1 = (int)n + 2; g.node_stride = L.node_stride;
  g.magic_n1 = (unsigned)((1ull << 32) / (unsigned)(n + 2)) + 1u;
  g.magic_nm = (unsigned)((1ull << 32) / (unsigned)n_m) + 1u;
  g.ws = a.ws;
  g.lim = (unsigned)(batch * (n + 2));
  g.vec_ok = (n_edges % 2 == 0) && (((unsigned long long)edge_index & 15ull) == 0);
  const long long items = g.vec_ok ? n_edges / 2 : n_edges;           // loads per row
  long long blocks = (items + 256 * kIngestTrips - 1) / (256 * kIngestTrips);
  if (blocks < 1) blocks = 1;
  coo_ingest_kernel<<<(unsigned)blocks, 256, 0, s>>>(g);
  CU(cudaGetLastError());
  {
    // programmatic dependent launch: the per-graph kernel's prologue overlaps the tail of the ingest kernel
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static int pdl = -1;
    if (pdl < 0) { const char* e = getenv("L2S_EVAL_PDL"); pdl = e ? atoi(e) != 0 : 1; }
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (cta) {
      cfg.gridDim = dim3((unsigned)batch);
      cfg.blockDim = dim3((unsigned)cta_warps * 32);
      cfg.dynamicSmemBytes = (size_t)a.bytes;
      CU(cudaLaunchKernelEx(&cfg, coo_eval_kernel_cta, a));
    } else {
      cfg.gridDim = dim3((unsigned)((batch + W - 1) / W));
      cfg.blockDim = dim3((unsigned)W * 32);
      cfg.dynamicSmemBytes = (size_t)W * a.bytes;
      CU(cudaLaunchKernelEx(&cfg, coo_eval_kernel, a));
    }
  }
  CU(cudaGetLastError());
  return L2S_OK;
}

}  // extern "C"
