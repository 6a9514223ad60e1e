// l2s_b200 kernels (sm_100a): environment step, fused K-step run, setup, initial-solution builder and the
// stateless COO evaluator.  See l2s_device.cuh for the execution model.
#pragma once
#include "l2s_device.cuh"

namespace l2s {

// ------------------------------------------------------------------------------------------------------
// setup kernels (one warp per instance; not on the hot path)
// ------------------------------------------------------------------------------------------------------

// int32 durations / machine ids -> node words + byte machine ids, with input validation.
__global__ void load_kernel(Layout L, State st, int B, const int32_t* __restrict__ dur,
                            const int32_t* __restrict__ mch) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  int bad = 0;
  long long sum = 0;
  for (int v = lane; v < L.node_stride; v += 32) {
    int word = 0;
    if (v >= 1 && v <= L.n) {
      const int d = dur[(size_t)b * L.n + v - 1];
      if (d < 1 || d > kDurMax) bad |= 1;
      const int pos = (v - 1) % L.n_m;
      word = (d & kDurMask) | (pos == 0 ? kFirstFlag : 0) | (pos == L.n_m - 1 ? kLastFlag : 0);
      sum += d;
    }
    st.durw[(size_t)b * L.node_stride + v] = (uint16_t)word;
  }
  for (int i = lane; i < L.mch_stride; i += 32) {
    int mc = 0;
    if (i < L.n) {
      mc = mch[(size_t)b * L.n + i];
      if (mc < 1 || mc > L.n_m) { bad |= 2; mc = 1; }
    }
    st.mch[(size_t)b * L.mch_stride + i] = (uint8_t)mc;
  }
  const unsigned full = L.n_m == 32 ? 0xffffffffu : ((1u << L.n_m) - 1u);
  for (int j = lane; j < L.n_j; j += 32) {
    unsigned mask = 0;
    for (int k = 0; k < L.n_m; ++k) mask |= 1u << ((mch[(size_t)b * L.n + j * L.n_m + k] - 1) & 31);
    if (mask != full) bad |= 2;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
  if (sum >= (1ll << 24)) bad |= 1;
  bad = __reduce_or_sync(kFull, bad);
  if (bad && lane == 0) record_error(st.err, (bad & 2) ? -7 : -6, b);
}

__device__ __forceinline__ void reset_dynamic(const State& st, int b, int lane) {
  if (lane == 0) {
    st.tabu[2 * b] = 0;
    st.tabu[2 * b + 1] = 0;
    st.cur[b] = 0;
    st.inc[b] = 0;
    if (st.evalid) st.evalid[b] = 0;   // a new solution: the cached evaluation is void
  }
}

// mseq [B, n_m, n_j] int32 node ids -> walk words + where index; validates rows.
__global__ void set_solution_kernel(Layout L, State st, int B, const int32_t* __restrict__ mseq) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  uint16_t* where = st.where + (size_t)b * L.node_stride;
  const uint16_t* durw = st.durw + (size_t)b * L.node_stride;
  uint32_t* walk = st.walk + (size_t)b * L.walk_stride;
  for (int i = lane; i < L.node_stride; i += 32) where[i] = 0xffff;
  __syncwarp();
  int bad = 0;
  for (int m = 0; m < L.n_m; ++m) {
    for (int k = lane; k < L.n_j; k += 32) {
      const int idx = m * L.n_j + k;
      int v = mseq[(size_t)b * L.n + idx];
      if (v < 1 || v > L.n || (int)st.mch[(size_t)b * L.mch_stride + v - 1] - 1 != m) {
        bad = 1;
        v = 1;
      } else {
        const unsigned short old = atomicCAS(reinterpret_cast<unsigned short*>(&where[v]), (unsigned short)0xffff,
                                             (unsigned short)idx);
        if (old != 0xffff) bad = 1;
      }
      walk[idx] = make_walk(v, durw[v], false, k == L.n_j - 1);
    }
  }
  for (int i = L.n + lane; i < L.walk_stride; i += 32) walk[i] = 0;
  reset_dynamic(st, b, lane);
  bad = __reduce_or_sync(kFull, bad);
  if (bad && lane == 0) record_error(st.err, -10, b);
}

__global__ void get_solution_kernel(Layout L, State st, int B, int32_t* __restrict__ mseq) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  for (int i = lane; i < L.n; i += 32) mseq[(size_t)b * L.n + i] = (int)(st.walk[(size_t)b * L.walk_stride + i] & kWId);
}

__global__ void export_misc_kernel(Layout L, State st, int B, uint8_t* jobfirst, float* makespan, float* incumbent) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  if (jobfirst) {
    if (lane == 0) { jobfirst[(size_t)b * L.n1] = 0; jobfirst[(size_t)b * L.n1 + L.n + 1] = 0; }
    for (int i = lane; i < L.n; i += 32) {
      const unsigned w = st.walk[(size_t)b * L.walk_stride + i];
      jobfirst[(size_t)b * L.n1 + (w & kWId)] = (w & kWJf) ? 1 : 0;
    }
  }
  if (lane == 0) {
    if (makespan) makespan[b] = (float)st.cur[b];
    if (incumbent) incumbent[b] = (float)st.inc[b];
  }
}

// ------------------------------------------------------------------------------------------------------
// Initial solution: priority dispatching with permissible left shift.
// JsspN5._rules_solver (env/environment.py:203-255) + permissibleLeftShift (env/permissible_LS.py:5-78).
// One warp per instance; the machine gantt rows live in shared memory.  Priorities change only for the
// job just dispatched, so they are cached (one IEEE double division per dispatched operation).
// Shared memory per instance: starts int32[n], ops uint16[n], cnt int32[n_m], nxt/jobrdy/done int32[n_j]x3,
// pri double[n_j].
// ------------------------------------------------------------------------------------------------------
struct InitLayout {
  int off_starts, off_ops, off_cnt, off_nxt, off_rdy, off_work, off_tot, off_pri, bytes;
};

__global__ void build_initial_kernel(Layout L, InitLayout IL, State st, int B, int rule, int high) {
  extern __shared__ __align__(16) unsigned char smem_init[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + w;
  if (b >= B) return;
  unsigned char* base = smem_init + (size_t)w * IL.bytes;
  int32_t* starts = reinterpret_cast<int32_t*>(base + IL.off_starts);   // [n_m][n_j]
  uint16_t* ops = reinterpret_cast<uint16_t*>(base + IL.off_ops);       // [n_m][n_j] node ids (1-based)
  int32_t* cnt = reinterpret_cast<int32_t*>(base + IL.off_cnt);         // ops placed per machine
  int32_t* nxt = reinterpret_cast<int32_t*>(base + IL.off_nxt);         // next position per job
  int32_t* rdy = reinterpret_cast<int32_t*>(base + IL.off_rdy);         // completion of the job's last placed op
  int32_t* work = reinterpret_cast<int32_t*>(base + IL.off_work);       // remaining work per job
  int32_t* total = reinterpret_cast<int32_t*>(base + IL.off_tot);       // total work per job
  double* pri = reinterpret_cast<double*>(base + IL.off_pri);
  const uint16_t* durw = st.durw + (size_t)b * L.node_stride;
  const uint8_t* mch = st.mch + (size_t)b * L.mch_stride;
  const int n_j = L.n_j, n_m = L.n_m;
  const double kInf = __longlong_as_double(0x7ff0000000000000ll);

  for (int m = lane; m < n_m; m += 32) cnt[m] = 0;
  for (int j = lane; j < n_j; j += 32) {
    int tot = 0;
    for (int k = 0; k < n_m; ++k) tot += durw[j * n_m + k + 1] & kDurMask;
    nxt[j] = 0;
    rdy[j] = 0;
    work[j] = tot;
    total[j] = tot;
    const int d0 = durw[j * n_m + 1] & kDurMask;
    // fdd/mwkr: cumulative duration up to and including the candidate / remaining work (:234-236)
    pri[j] = rule == 0 ? (double)d0 : __ddiv_rn((double)d0, (double)tot);
  }
  __syncwarp();
  for (int it = 0; it < L.n; ++it) {
    // argmin over jobs, first minimum in job order (reference: np.random.choice among ties, :237)
    double best = kInf;
    int bj = 0x7fffffff;
    for (int j = lane; j < n_j; j += 32) {
      const double p = pri[j];
      if (p < best) { best = p; bj = j; }
    }
    for (int o = 16; o; o >>= 1) {
      const double ob = __shfl_xor_sync(kFull, best, o);
      const int oj = __shfl_xor_sync(kFull, bj, o);
      if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    // malformed input that load_kernel has flagged (NaN priorities from an all-zero job, a job that repeats a
    // machine) must not turn into out-of-bounds accesses here, whatever order the host polls errors in
    if (bj >= n_j) break;
    const int j = bj;
    const int kpos = nxt[j];
    const int v = j * n_m + kpos + 1;           // node id
    const int d = durw[v] & kDurMask;
    const int m = (int)mch[v - 1] - 1;
    int32_t* rs = starts + m * n_j;
    uint16_t* ro = ops + m * n_j;
    const int c = min(cnt[m], n_j - 1);
    const int job_rdy = rdy[j];
    const int mch_rdy = c ? rs[c - 1] + (durw[ro[c - 1]] & kDurMask) : 0;
    // first placed op that starts after job_rdy (rows are sorted by start time)
    int p0 = c;
    for (int k0 = 0; k0 < c; k0 += 32) {
      const int k = k0 + lane;
      const unsigned hit = __ballot_sync(kFull, k < c && rs[k] > job_rdy);
      if (hit) { p0 = k0 + __ffs(hit) - 1; break; }
    }
    int ins = -1, start_a = 0;
    if (p0 < c) {
      // end of the op before the first candidate slot; for p0 == 0 the reference reads index -1, i.e. the
      // unfilled row tail: -high + dur[flat index n - n_j] (env/permissible_LS.py:41)
      const int before = p0 > 0 ? rs[p0 - 1] + (durw[ro[p0 - 1]] & kDurMask)
                                : -high + (durw[L.n - n_j + 1] & kDurMask);
      for (int k0 = p0; k0 < c; k0 += 32) {
        const int k = k0 + lane;
        bool fits = false;
        int open = 0;
        if (k < c) {
          open = (k == p0) ? max(job_rdy, before) : rs[k - 1] + (durw[ro[k - 1]] & kDurMask);
          fits = d <= rs[k] - open;
        }
        const unsigned hit = __ballot_sync(kFull, fits);
        if (hit) {
          const int src = __ffs(hit) - 1;
          ins = k0 + src;
          start_a = __shfl_sync(kFull, open, src);
          break;
        }
      }
    }
    __syncwarp();
    if (ins >= 0) {
      // shift [ins, c) one slot to the right, back to front in chunks so reads precede overwrites
      for (int hi = c; hi > ins; hi -= 32) {
        const int k = hi - 1 - lane;           // source index
        int sv = 0, ov = 0;
        const bool act = k >= ins;
        if (act) { sv = rs[k]; ov = ro[k]; }
        __syncwarp();
        if (act) { rs[k + 1] = sv; ro[k + 1] = (uint16_t)ov; }
        __syncwarp();
      }
      if (lane == 0) { rs[ins] = start_a; ro[ins] = (uint16_t)v; }
    } else {
      start_a = max(job_rdy, mch_rdy);
      if (lane == 0) { rs[c] = start_a; ro[c] = (uint16_t)v; }
    }
    if (lane == 0) {
      cnt[m] = c + 1;
      rdy[j] = start_a + d;
      const int wk = work[j] - d;
      work[j] = wk;
      nxt[j] = kpos + 1;
      if (kpos + 1 < n_m) {
        const int dn = durw[v + 1] & kDurMask;
        // cum(candidate) = work done so far + dur(candidate)
        pri[j] = rule == 0 ? (double)dn : __ddiv_rn((double)(total[j] - wk + dn), (double)wk);
      } else {
        pri[j] = kInf;
      }
    }
    __syncwarp();
  }
  uint32_t* walk = st.walk + (size_t)b * L.walk_stride;
  for (int m = 0; m < n_m; ++m)
    for (int k = lane; k < n_j; k += 32) {
      const int idx = m * n_j + k;
      const int v = min(max((int)ops[idx], 1), L.n);   // (garbage only for instances load_kernel rejected)
      walk[idx] = make_walk(v, durw[v], false, k == n_j - 1);
    }
  for (int i = L.n + lane; i < L.walk_stride; i += 32) walk[i] = 0;
  reset_dynamic(st, b, lane);
}

// 16-byte coalesced copy (host action block -> device staging, see l2s_step_host)
__global__ void copy16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n16) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------------
// The environment step: JsspN5.step (env/environment.py:356-387), one launch for the whole slice.
// ------------------------------------------------------------------------------------------------------
// IEEE round-to-nearest float32 quotient x/c for a launch-constant divisor.  mode 1/2: x*r with one/two
// FMA residual corrections (r = correctly rounded 1/c); the host verified exhaustively over the numerator
// range that this equals x/c bit for bit for THIS c (l2s_capi.cu: verify_divisor); mode 0: __fdiv_rn.
__device__ __forceinline__ float div_rn(float x, float c, float r, int mode) {
  if (mode == 0) return __fdiv_rn(x, c);
  float q = __fmul_rn(x, r);
  q = __fmaf_rn(__fmaf_rn(-c, q, x), r, q);
  if (mode == 2) q = __fmaf_rn(__fmaf_rn(-c, q, x), r, q);
  return q;
}

#ifdef L2S_TIMELINE
__device__ unsigned long long* g_timeline = nullptr;   // [B][12]: smid, globaltimer start, clock64 stamps
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TL(k) do { if (lane == 0 && g_timeline) g_timeline[(size_t)b * 12 + (k)] = (k) <= 1 ? gtimer() : (unsigned long long)clock64(); } while (0)
#else
#define TL(k) do {} while (0)
#endif
#define TLC(k) do { if (tid == 0) { const int lane = 0; TL(k); (void)lane; } } while (0)

struct StepArgs {
  Layout L;
  State st;
  int B;
  const int32_t* actions;
  float high, fea;
  float r_high, r_fea;     // correctly rounded reciprocals
  int div_high, div_fea;   // 1/2: FMA-corrected reciprocal division verified bit-exact on the host, 0: __fdiv_rn
  int reward_type, is_reset;
  float* x;
  int64_t* emc;
  float* makespan;
  float* incumbent;
  float* reward;
  uint8_t* done;
  int32_t* pairs;
  int32_t* npairs;
  int32_t* status;
  int pmax;
  // host path (l2s_step_host): ONE record per instance, staged in shared memory and written by the whole warp as
  // one coalesced store -- normally straight into mapped pinned host memory (see include/l2s_b200.h):
  //   word 0: npairs | done << 16 | (status & 0xff) << 24;  words 1-3: makespan, reward, incumbent (float32);
  //   words 4..4+npairs: the feasible pairs, a0 | a1 << 16
  uint32_t* rec;
  int rec_stride;          // words per record: 4 + pmax rounded up to a multiple of 16
  int rec_lines;           // write whole 64-byte lines (the words behind 4 + npairs are padding with stale contents)
  int fast_sweep;          // CTA kernel: 1 = latency variant of the wavefront (few resident instances per SM), 2 = + look-ahead
  int incremental;         // CTA kernel: resume the sweeps from the cached evaluation st.fq (0: always sweep everything)
  // test introspection (l2s_export_state): evaluate only, no mutation
  int export_only;
  int32_t* ex_est;
  int32_t* ex_lst;
  int32_t* ex_path;
  int32_t* ex_path_len;
};

// Warp-per-instance step for instances of which >= 16 fit an SM; the static node words are always staged in
// shared memory here (l2s_create picks the layout together with the kernel).
// <= 4 warps per CTA, >= 8 CTAs per SM: 64 registers (measured: the kernel is issue-bound, not occupancy-bound)
__global__ void __launch_bounds__(128, 8) step_kernel(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_step[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + w;
  if (b >= a.B) return;
  const Layout& L = a.L;
  const Inst s = inst_view(L, smem_addr(smem_step) + (uint32_t)w * (uint32_t)L.bytes);
  const int n = L.n, n1 = L.n1;

  TL(0); TL(2);
  int a0 = 0, a1 = 0;
  if (a.actions && !a.export_only) {
    if ((reinterpret_cast<uintptr_t>(a.actions) & 7u) == 0) {   // [B,2] int32 rows: one 8-byte load
      const int2 act = reinterpret_cast<const int2*>(a.actions)[b];
      a0 = act.x;
      a1 = act.y;
    } else {
      a0 = a.actions[2 * b];
      a1 = a.actions[2 * b + 1];
    }
  }
  const int2 tabu_in = reinterpret_cast<const int2*>(a.st.tabu)[b];
  int tabu0 = tabu_in.x, tabu1 = tabu_in.y;
  const int inc_in = a.st.inc[b], cur_in = a.st.cur[b];
  stage_instance(L, s, a.st, b, lane, true);
  TL(11);   // staged
  __syncwarp();
  // position of a0 in the machine orders: scan the staged walk words (no index array to keep up to date in HBM;
  // looking up a0's machine row first costs a dependent global load and measured 5 % slower)
  int i0 = -1;
  if (a0 >= 1 && a0 <= n) {
    for (int i = 4 * lane; i < L.walk_stride; i += 128) {   // four walk words per lane and trip (the padding is 0)
      const uint4 wq = lds128(s.walk + 4u * i);
      if ((int)(wq.x & kWId) == a0) i0 = i;
      if ((int)(wq.y & kWId) == a0) i0 = i + 1;
      if ((int)(wq.z & kWId) == a0) i0 = i + 2;
      if ((int)(wq.w & kWId) == a0) i0 = i + 3;
    }
    i0 = __reduce_max_sync(kFull, i0);
  }

  TL(3);   // staged + searched
  int status = 0;
  // 1. N5 swap (change_nxgraph_topology :318-354) + tabu update (:370-380)
  {
    bool has_t = false;
    const int r = apply_swap(L, s, a0, a1, i0, lane, has_t);
    if (r == 0) {
      tabu0 = a1;
      tabu1 = a0;
      if (lane == 0) {
        uint32_t* gw = a.st.walk + (size_t)b * L.walk_stride;
        gw[i0] = (uint32_t)lds32(s.walk + 4u * i0);
        gw[i0 + 1] = (uint32_t)lds32(s.walk + 4u * (i0 + 1));
        if (has_t) gw[i0 + 2] = (uint32_t)lds32(s.walk + 4u * (i0 + 2));
        a.st.tabu[2 * b] = tabu0;
        a.st.tabu[2 * b + 1] = tabu1;
      }
      __syncwarp();
    } else if (r < 0) {
      status = r;
    }
  }

  TL(4);   // swapped
  // park the scalars that are only needed at the very end in shared memory (registers are the occupancy limit);
  // not earlier, so that their global loads do not gate the action lookup and the swap
  if (lane == 0) { sts32(s.misc, inc_in); sts32(s.misc + 4u, cur_in); }
  stage_wait_rest(s);
  // 2. evaluate (Evaluator.forward, env/message_passing_evl.py:161-199): forward + backward sweep
  const int ms = evaluate_instance(L, s, lane, true);
  if (ms < 0) {
    status = -5;
    if (lane == 0) {
      record_error(a.st.err, -5, b);
      if (a.status) a.status[b] = status;
      if (a.npairs) a.npairs[b] = 0;
      if (a.done) a.done[b] = 1;
      if (a.rec) a.rec[(size_t)b * a.rec_stride] = 0x10000u | (0xfbu << 24);   // done, no pairs, status -5
    }
    return;
  }
  TL(5);   // evaluated
  const uint32_t fin = s.tq, q = s.tq + 4u * L.n1p;
  // latest start of S = min over job-first ops (:195 on the S row)
  int qs = 0;
  for (int j = lane; j < L.n_j; j += 32) qs = max(qs, lds32(q + 4u * (j * L.n_m + 1)));
  qs = __reduce_max_sync(kFull, qs);
  // S and T rows: est[S]=0, lst[S]=ms-qs, est[T]=lst[T]=ms (their node words are 0)
  if (lane == 0) {
    sts32(fin, 0);
    sts32(q, qs);
    sts32(fin + 4u * (n + 1), ms);
    sts32(q + 4u * (n + 1), 0);
  }
  __syncwarp();

  // 3. node features x = [dur/high, est/c, lst/c] (env/environment.py:312): float32 IEEE-RN quotients
  if (a.x) {
    float* xo = a.x + (size_t)b * n1 * 3 + 3 * lane;
    const float fms = (float)ms;
    if (a.div_high == 1 && a.div_fea == 1 && !(n1 & 1)) {
      // common case (one FMA residual correction reproduces the IEEE quotient, verified on the host; even node
      // count => 8-byte aligned rows): two adjacent nodes per lane, 64-bit shared loads, 64-bit global stores
      const float ch = a.high, rh = a.r_high, cf = a.fea, rf = a.r_fea;
      float2* const xb = reinterpret_cast<float2*>(a.x + (size_t)b * n1 * 3);
      L2S_LOOP
      for (int p = lane; p < (n1 >> 1); p += 32) {
        float2* const xo2 = xb + 3 * p;
        const unsigned dw = (unsigned)lds32(s.durw + 4u * p);
        const int2 f = lds64(fin + 8u * p), t = lds64(q + 8u * p);
        const int d0 = (int)(dw & kDurMask), d1 = (int)((dw >> 16) & kDurMask);
        const float fd0 = (float)d0, fe0 = (float)(f.x - d0), fl0 = fms - (float)t.x;
        const float fd1 = (float)d1, fe1 = (float)(f.y - d1), fl1 = fms - (float)t.y;
        const float q00 = __fmul_rn(fd0, rh), q01 = __fmul_rn(fe0, rf), q02 = __fmul_rn(fl0, rf);
        const float q10 = __fmul_rn(fd1, rh), q11 = __fmul_rn(fe1, rf), q12 = __fmul_rn(fl1, rf);
        xo2[0] = make_float2(__fmaf_rn(__fmaf_rn(-ch, q00, fd0), rh, q00), __fmaf_rn(__fmaf_rn(-cf, q01, fe0), rf, q01));
        xo2[1] = make_float2(__fmaf_rn(__fmaf_rn(-cf, q02, fl0), rf, q02), __fmaf_rn(__fmaf_rn(-ch, q10, fd1), rh, q10));
        xo2[2] = make_float2(__fmaf_rn(__fmaf_rn(-cf, q11, fe1), rf, q11), __fmaf_rn(__fmaf_rn(-cf, q12, fl1), rf, q12));
      }
    } else if (a.div_high == 1 && a.div_fea == 1) {
      const float ch = a.high, rh = a.r_high, cf = a.fea, rf = a.r_fea;
      for (int i = lane; i < n1; i += 32, xo += 96) {
        const int d = lds16(s.durw + 2u * i) & kDurMask;
        const float fd = (float)d, fe = (float)(lds32(fin + 4u * i) - d), fl = fms - (float)lds32(q + 4u * i);
        float q0 = __fmul_rn(fd, rh), q1 = __fmul_rn(fe, rf), q2 = __fmul_rn(fl, rf);
        xo[0] = __fmaf_rn(__fmaf_rn(-ch, q0, fd), rh, q0);
        xo[1] = __fmaf_rn(__fmaf_rn(-cf, q1, fe), rf, q1);
        xo[2] = __fmaf_rn(__fmaf_rn(-cf, q2, fl), rf, q2);
      }
    } else {
      for (int i = lane; i < n1; i += 32, xo += 96) {
        const int d = lds16(s.durw + 2u * i) & kDurMask;
        const int f = lds32(fin + 4u * i), t = lds32(q + 4u * i);
        xo[0] = div_rn((float)d, a.high, a.r_high, a.div_high);
        xo[1] = div_rn((float)(f - d), a.fea, a.r_fea, a.div_fea);
        xo[2] = div_rn(fms - (float)t, a.fea, a.r_fea, a.div_fea);
      }
    }
  }
  if (a.ex_est) {
    for (int i = lane; i < n1; i += 32) {
      const int d = lds16(s.durw + 2u * i) & kDurMask;
      a.ex_est[(size_t)b * n1 + i] = lds32(fin + 4u * i) - d;
      a.ex_lst[(size_t)b * n1 + i] = ms - lds32(q + 4u * i);
    }
  }

  TL(6);   // x written
  // 4. machine edges ascending by source (dag2pyg :300-305).  q is dead now: its storage holds the node-indexed
  //    machine successors (msu) and, after them, the reversed critical path.
  //    Node ids of a slice fit 32 bits (checked in l2s_create): 32-bit arithmetic, 64-bit stores.
  __syncwarp();
  const uint32_t msu = q, rp = q + 2u * (uint32_t)n1;
  if (a.emc) link_pass<true>(L, s, msu, lane); else link_pass<false>(L, s, msu, lane);
  __syncwarp();
  if (a.emc) {
    const unsigned off = (unsigned)b * (unsigned)n1;
    uint2* src = reinterpret_cast<uint2*>(a.emc) + (size_t)b * L.e_mc;
    uint2* dst = src + (size_t)a.B * L.e_mc;
    const unsigned below = (1u << lane) - 1u;
    L2S_LOOP
    for (int v0 = 1; v0 <= n; v0 += 32) {
      const int v = v0 + lane;
      const unsigned sc = v <= n ? (unsigned)lds16(msu + 2u * v) : 0u;
      const unsigned bal = __ballot_sync(kFull, sc != 0u);
      if (sc) {
        const int pos = __popc(bal & below);
        src[pos] = make_uint2(off + (unsigned)v, 0u);
        dst[pos] = make_uint2(off + sc, 0u);
      }
      const int adv = __popc(bal);
      src += adv;
      dst += adv;
    }
  }
  __syncwarp();

  TL(7);   // link + emc
  // 5. critical path (nx.dag_longest_path, :77) and N5 move set (_get_pairs :84-105)
  const int len = trace_path(L, s, ms, lane, rp);
  if (a.ex_path) {
    for (int i = lane; i < len; i += 32) a.ex_path[(size_t)b * n + i] = lds16(rp + 2u * (len - 1 - i));
    if (lane == 0) a.ex_path_len[b] = len;
  }
  TL(8);   // traced
  int32_t* po = a.pairs ? a.pairs + (size_t)b * a.pmax * 2 : nullptr;
  const int pmax = a.pmax;
  const uint32_t stage = s.rec;
  uint32_t* const rec = a.rec;
  int np = n5_moves(L, s, rp, len, tabu0, tabu1, lane, [&](int idx, int u, int v) {
    if (idx < pmax) {
      if (po) {
        po[2 * idx] = u;
        po[2 * idx + 1] = v;
      }
      if (rec) sts32(stage + 16u + 4u * idx, (int)((uint32_t)u | ((uint32_t)v << 16)));
    }
  });
  if (np < 0) { status = status ? status : np; np = 0; }
  if (np > pmax) { status = status ? status : -8; np = pmax; }

  TL(9);   // moves
  // 6. reward / incumbent / current (:359-367)
  if (lane == 0) {
    const int inc_old = lds32(s.misc), cur_old = lds32(s.misc + 4u);
    int reward, inc;
    if (a.is_reset) {
      reward = 0;
      inc = ms;
    } else {
      reward = a.reward_type == 1 ? cur_old - ms : max(inc_old - ms, 0);
      inc = min(inc_old, ms);
    }
    if (!a.export_only) {
      a.st.cur[b] = ms;
      a.st.inc[b] = inc;
    }
    if (a.makespan) a.makespan[b] = (float)ms;
    if (a.incumbent) a.incumbent[b] = (float)inc;
    if (a.reward) a.reward[b] = (float)reward;
    if (a.done) a.done[b] = np == 0;
    if (a.npairs) a.npairs[b] = np;
    if (a.status) a.status[b] = status;
    if (status) record_error(a.st.err, status, b);
    if (rec) {
      sts32(stage, (int)((uint32_t)np | (np == 0 ? 0x10000u : 0u) | ((uint32_t)(status & 0xff) << 24)));
      sts32(stage + 4u, __float_as_int((float)ms));
      sts32(stage + 8u, __float_as_int((float)reward));
      sts32(stage + 12u, __float_as_int((float)inc));
    }
  }
  if (rec) {
    __syncwarp();
    uint32_t* r = rec + (size_t)b * a.rec_stride;
    const int nw_rec = a.rec_lines ? ((4 + np + 15) & ~15) : 4 + np;
    for (int i = lane; i < nw_rec; i += 32) r[i] = (uint32_t)lds32(stage + 4u * i);
  }
  TL(10); TL(1);
}

// ------------------------------------------------------------------------------------------------------
// The same step for LARGE instances (few instances fit an SM, so one warp per instance would leave the SM
// almost empty): one CTA of kW warps owns one instance.  Warp 0 runs the forward sweep while warp 1 runs the
// backward sweep; staging, the link pass and the x / edge outputs are spread over all threads; the critical
// path trace and the move enumeration stay on warp 0.  Synchronisation is __syncthreads (one instance per CTA).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) step_kernel_cta(const __grid_constant__ StepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_step[];
  __shared__ int sh[8];   // 0: i0, 1: swap result, 2: has_t, 3: ms, 4: fwd ok, 5: bwd ok, 6: qs, 7: scratch
  __shared__ int wcount[8];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
  const int b = blockIdx.x;
  const Layout& L = a.L;
  const Inst s = inst_view(L, smem_addr(smem_step));
  const int n = L.n, n1 = L.n1;

  TLC(0); TLC(2);
  int a0 = 0, a1 = 0;
  if (a.actions && !a.export_only) {
    a0 = a.actions[2 * b];
    a1 = a.actions[2 * b + 1];
  }
  int tabu0 = a.st.tabu[2 * b], tabu1 = a.st.tabu[2 * b + 1];
  if (tid < 8) sh[tid] = tid == 0 ? -1 : 0;
  // the static node words (durations by node id) are not staged for large instances: fetch this thread's share
  // for the feature loop NOW (clamped index: no select that would wait for the data), so that the loads are in
  // flight during staging and the sweeps instead of being a chain of dependent global-load latencies inside that
  // loop (measured at 100x20: 9-11 us -> 1-3 us)
  constexpr int kPairRegs = 9;
  const uint16_t* gdurw = a.st.durw + (size_t)b * L.node_stride;
  // fast feature path: both divisors pass the one-correction test, even node count (8-byte aligned rows), and a
  // thread's share of node-word PAIRS fits the registers
  // decoupled regions (latency regime, see make_layout): warp 0 goes straight from the link pass to the path trace and
  // the move set while warps 1.. write the features and the edge list; otherwise all threads write the features first
  const bool decoupled = L.off_msu >= 0 && nw > 1;
  const int xt = decoupled ? tid - 32 : tid, xn = decoupled ? nthr - 32 : nthr;     // feature-writer index / count
  const bool x_fast = a.x && a.div_high == 1 && a.div_fea == 1 && !(n1 & 1) && ((n1 >> 1) + xn - 1) / xn <= kPairRegs;
  unsigned dpair[kPairRegs];
  if (x_fast && xt >= 0) {
#pragma unroll
    for (int k = 0; k < kPairRegs; ++k)
      dpair[k] = __ldg(reinterpret_cast<const unsigned*>(gdurw) + min(xt + k * xn, (n1 >> 1) - 1));
  }
  // incremental re-evaluation: the completion times / tails of the previous evaluation come back from HBM
  // (one more bulk copy on the second barrier) instead of "pending" sentinels
  const bool cached = a.incremental && a.st.evalid && a.st.evalid[b] != 0;
  stage_instance(L, s, a.st, b, tid, true, nthr, cached);
  __syncthreads();
  if (a0 >= 1 && a0 <= n) {
    for (int i = tid; i < n; i += nthr)
      if ((lds32(s.walk + 4u * i) & (int)kWId) == a0) sh[0] = i;     // a node id occurs exactly once
  }
  __syncthreads();
  TLC(3);
  const int i0 = sh[0];
  int status = 0;
  if (w == 0) {
    bool has_t = false;
    const int r = apply_swap(L, s, a0, a1, i0, lane, has_t);
    if (lane == 0) { sh[1] = r; sh[2] = has_t; }
    if (r == 0 && lane == 0) {
      uint32_t* gw = a.st.walk + (size_t)b * L.walk_stride;
      gw[i0] = (uint32_t)lds32(s.walk + 4u * i0);
      gw[i0 + 1] = (uint32_t)lds32(s.walk + 4u * (i0 + 1));
      if (has_t) gw[i0 + 2] = (uint32_t)lds32(s.walk + 4u * (i0 + 2));
      a.st.tabu[2 * b] = a1;
      a.st.tabu[2 * b + 1] = a0;
      if (a.st.evalid) a.st.evalid[b] = 0;   // until this launch has stored the new evaluation
    }
  }
  __syncthreads();
  const int swap_r = sh[1];
  if (swap_r == 0) { tabu0 = a1; tabu1 = a0; } else if (swap_r < 0) status = swap_r;
  TLC(4);
  stage_wait_rest(s);
  TLC(11);   // sentinels / cached evaluation landed
  // evaluation: forward sweep on warp 0, backward sweep on warp 1 (or both on warp 0 if the CTA has one warp).
  // With a cached evaluation nothing is swept when nothing moved, and after a swap only the part of the schedule
  // that can change (see resume_point): the rows resume behind their kept prefix (forward) / suffix (backward).
  int startv = 0, prevv = 0, start2 = 0, prev2 = 0;     // resume points of this lane's row (warp 0: fwd[, bwd]; warp 1: bwd)
  const bool sweep = !cached || swap_r == 0;
  if (cached && swap_r == 0) {
    const unsigned wa1 = (unsigned)lds32(s.walk + 4u * i0), wa0 = (unsigned)lds32(s.walk + 4u * (i0 + 1));
    const int thr_f = lds32(s.tq + 4u * (wa0 & kWId)) - (int)(wa0 >> kWDurShift);                 // old est(a0)
    const int thr_b = lds32(s.tq + 4u * (L.n1p + (wa1 & kWId))) - (int)(wa1 >> kWDurShift);       // old tail after a1
    if (w == 0) {
      resume_point(L, s, lane, 0, lane < L.n_m, thr_f, startv, prevv);
      if (nw == 1) resume_point(L, s, lane, 1, lane < L.n_m, thr_b, start2, prev2);
    } else if (w == 1) {
      resume_point(L, s, lane, 1, lane < L.n_m, thr_b, startv, prevv);
    }
    __syncthreads();                       // every resume point is read from the OLD values
    invalidate_from(L, s, thr_f, thr_b, tid, nthr);
    __syncthreads();
  }
  // __syncthreads does not reconverge a warp whose lanes arrived separately (lane 0 took the write-back branch
  // above): without the explicit __syncwarp the sweep below would run with the warp split in two, every level
  // paying the divergent path of its own __syncwarp (measured: 255 ns instead of 64 ns per level)
  __syncwarp();
  if (sweep) {
    if (w == 0) {
      int last;
      const bool ok = a.fast_sweep == 2 ? wavefront_fast2(L, s, lane, 0, lane < L.n_m, last, -1, 0u, startv, prevv)
                      : a.fast_sweep ? wavefront_fast(L, s, lane, 0, lane < L.n_m, last, -1, 0u, startv, prevv)
                                     : wavefront(L, s, lane, 0, lane < L.n_m, last, -1, 0u, startv, prevv);
      const int ms0 = __reduce_max_sync(kFull, last);
      if (lane == 0) { sh[3] = ms0; sh[4] = ok; }
      if (nw == 1) {
        int l2;
        const bool ok2 = a.fast_sweep == 2 ? wavefront_fast2(L, s, lane, 1, lane < L.n_m, l2, -1, 0u, start2, prev2)
                         : a.fast_sweep ? wavefront_fast(L, s, lane, 1, lane < L.n_m, l2, -1, 0u, start2, prev2)
                                        : wavefront(L, s, lane, 1, lane < L.n_m, l2, -1, 0u, start2, prev2);
        if (lane == 0) sh[5] = ok2;
      }
    } else if (w == 1) {
      int l2;
      const bool ok2 = a.fast_sweep == 2 ? wavefront_fast2(L, s, lane, 1, lane < L.n_m, l2, -1, 0u, startv, prevv)
                       : a.fast_sweep ? wavefront_fast(L, s, lane, 1, lane < L.n_m, l2, -1, 0u, startv, prevv)
                                      : wavefront(L, s, lane, 1, lane < L.n_m, l2, -1, 0u, startv, prevv);
      if (lane == 0) sh[5] = ok2;
    }
  } else if (tid == 0) {                   // nothing moved: the cached evaluation stands (fin[T] = makespan)
    sh[3] = lds32(s.tq + 4u * (n + 1));
    sh[4] = 1;
    sh[5] = 1;
  }
  __syncthreads();
  TLC(5);
  const int ms = sh[3];
  if (!sh[4] || !sh[5]) {
    if (tid == 0) {
      record_error(a.st.err, -5, b);
      if (a.status) a.status[b] = -5;
      if (a.npairs) a.npairs[b] = 0;
      if (a.done) a.done[b] = 1;
      if (a.rec) a.rec[(size_t)b * a.rec_stride] = 0x10000u | (0xfbu << 24);   // done, no pairs, status -5
    }
    return;
  }
  const uint32_t fin = s.tq, q = s.tq + 4u * L.n1p;
  {
    int qs = 0;
    for (int j = tid; j < L.n_j; j += nthr) qs = max(qs, lds32(q + 4u * (j * L.n_m + 1)));
    qs = __reduce_max_sync(kFull, qs);
    if (lane == 0 && qs > 0) atomicMax(&sh[6], qs);
  }
  __syncthreads();
  if (tid == 0) {
    sts32(fin, 0);
    sts32(q, sh[6]);
    sts32(fin + 4u * (n + 1), ms);
    sts32(q + 4u * (n + 1), 0);
  }
  __syncthreads();
  if (a.incremental && a.st.fq && sweep && !a.export_only) {
    // keep this evaluation for the next launch: 16-byte coalesced copies of fin | q, then the flag (the next launch
    // on the stream sees both; a kernel that changes the machine orders clears the flag)
    uint4* g = reinterpret_cast<uint4*>(a.st.fq + (size_t)b * 2 * L.n1p);
    for (int i = tid; i < L.n1p / 2; i += nthr) g[i] = lds128(s.tq + 16u * i);
    if (tid == 0) a.st.evalid[b] = 1;
  }
  auto write_features = [&]() {
    if (a.x) {
      float* xo = a.x + (size_t)b * n1 * 3;
      const float fms = (float)ms;
      if (x_fast) {
        // two adjacent nodes per thread: 64-bit shared loads, 64-bit global stores, one FMA residual correction
        const float ch = a.high, rh = a.r_high, cf = a.fea, rf = a.r_fea;
        float2* const xb = reinterpret_cast<float2*>(xo);
  #pragma unroll
        for (int k = 0; k < kPairRegs; ++k) {
          const int p = xt + k * xn;
          if (p < (n1 >> 1)) {
            const unsigned dw = dpair[k];
            const int2 f = lds64(fin + 8u * p), t = lds64(q + 8u * p);
            const int d0 = (int)(dw & kDurMask), d1 = (int)((dw >> 16) & kDurMask);
            const float fd0 = (float)d0, fe0 = (float)(f.x - d0), fl0 = fms - (float)t.x;
            const float fd1 = (float)d1, fe1 = (float)(f.y - d1), fl1 = fms - (float)t.y;
            const float q00 = __fmul_rn(fd0, rh), q01 = __fmul_rn(fe0, rf), q02 = __fmul_rn(fl0, rf);
            const float q10 = __fmul_rn(fd1, rh), q11 = __fmul_rn(fe1, rf), q12 = __fmul_rn(fl1, rf);
            float2* const xo2 = xb + 3 * p;
            xo2[0] = make_float2(__fmaf_rn(__fmaf_rn(-ch, q00, fd0), rh, q00), __fmaf_rn(__fmaf_rn(-cf, q01, fe0), rf, q01));
            xo2[1] = make_float2(__fmaf_rn(__fmaf_rn(-cf, q02, fl0), rf, q02), __fmaf_rn(__fmaf_rn(-ch, q10, fd1), rh, q10));
            xo2[2] = make_float2(__fmaf_rn(__fmaf_rn(-cf, q11, fe1), rf, q11), __fmaf_rn(__fmaf_rn(-cf, q12, fl1), rf, q12));
          }
        }
      } else if (a.div_high == 1 && a.div_fea == 1) {   // odd node count or a very large instance: node by node
        const float ch = a.high, rh = a.r_high, cf = a.fea, rf = a.r_fea;
  #pragma unroll 4
        for (int i = xt; i < n1; i += xn) {
          const int d = __ldg(gdurw + i) & kDurMask;
          const float fd = (float)d, fe = (float)(lds32(fin + 4u * i) - d), fl = fms - (float)lds32(q + 4u * i);
          const float q0 = __fmul_rn(fd, rh), q1 = __fmul_rn(fe, rf), q2 = __fmul_rn(fl, rf);
          xo[3 * i] = __fmaf_rn(__fmaf_rn(-ch, q0, fd), rh, q0);
          xo[3 * i + 1] = __fmaf_rn(__fmaf_rn(-cf, q1, fe), rf, q1);
          xo[3 * i + 2] = __fmaf_rn(__fmaf_rn(-cf, q2, fl), rf, q2);
        }
      } else {
  #pragma unroll 4
        for (int i = xt; i < n1; i += xn) {
          const int d = __ldg(gdurw + i) & kDurMask;
          const int f = lds32(fin + 4u * i), t = lds32(q + 4u * i);
          xo[3 * i] = div_rn((float)d, a.high, a.r_high, a.div_high);
          xo[3 * i + 1] = div_rn((float)(f - d), a.fea, a.r_fea, a.div_fea);
          xo[3 * i + 2] = div_rn(fms - (float)t, a.fea, a.r_fea, a.div_fea);
        }
      }
    }
    if (a.ex_est) {
      for (int i = xt; i < n1; i += xn) {
        const int d = __ldg(gdurw + i) & kDurMask;
        a.ex_est[(size_t)b * n1 + i] = lds32(fin + 4u * i) - d;
        a.ex_lst[(size_t)b * n1 + i] = ms - lds32(q + 4u * i);
      }
    }
  };
  if (!decoupled) write_features();
  TLC(6);
  __syncthreads();   // (overlay mode) q is dead from here on: its storage holds msu and then the reversed critical path
  const uint32_t base0 = smem_addr(smem_step);
  const uint32_t msu = decoupled ? base0 + L.off_msu : q, rp = decoupled ? base0 + L.off_rp : q + 2u * (uint32_t)n1;
  if (a.emc) link_pass<true>(L, s, msu, tid, nthr); else link_pass<false>(L, s, msu, tid, nthr);
  if (tid == 0) { sts16(s.crit, 0); sts16(s.crit + 2u * (n + 1), 0); }                 // S and T map to S
  if (w == nw - 1) { const int st0 = path_start(L, s, ms, lane); if (lane == 0) sh[7] = st0; }   // reads fin: before it is reused
  __syncthreads();
  // jump pointers for the path trace, in the (now dead) fin half of tq unless they have regions of their own
  const uint32_t jump2 = decoupled ? base0 + L.off_j2 : s.tq;
  const uint32_t jump4 = decoupled ? base0 + L.off_j4 : s.tq + 2u * (uint32_t)L.n1p;
  square_map(jump2, s.crit, n1, tid, nthr);
  __syncthreads();
  square_map(jump4, jump2, n1, tid, nthr);
  __syncthreads();
  TLC(7);
  if (decoupled && w > 0) write_features();
  // the edge list is written by warps 1.. while warp 0 already traces the critical path (a single-warp CTA does
  // both in turn); msu and rp are disjoint parts of the dead q half
  const int ew = nw > 1 ? nw - 1 : 1, myw = nw > 1 ? w - 1 : 0;
  if (a.emc && (nw == 1 || w > 0)) {
    // ordered compaction over contiguous node chunks, one per writer warp: count, exchange, write
    const int chunk = ((n + ew - 1) / ew + 31) & ~31;
    const int vlo = 1 + myw * chunk, vhi = min(n, vlo + chunk - 1);
    int cnt = 0;
    for (int v0 = vlo; v0 <= vhi; v0 += 32) {
      const int v = v0 + lane;
      cnt += __popc(__ballot_sync(kFull, v <= vhi && lds16(msu + 2u * v) != 0));
    }
    if (lane == 0) wcount[myw] = cnt;
    if (nw > 1) asm volatile("bar.sync 1, %0;" ::"r"(ew * 32));   // writer warps only (named barrier 1)
    else __syncwarp();
    int base = 0;
    for (int k = 0; k < myw; ++k) base += wcount[k];
    const unsigned off = (unsigned)b * (unsigned)n1;
    uint2* src = reinterpret_cast<uint2*>(a.emc) + (size_t)b * L.e_mc + base;
    uint2* dst = src + (size_t)a.B * L.e_mc;
    const unsigned below = (1u << lane) - 1u;
    for (int v0 = vlo; v0 <= vhi; v0 += 32) {
      const int v = v0 + lane;
      const unsigned sc = v <= vhi ? (unsigned)lds16(msu + 2u * v) : 0u;
      const unsigned bal = __ballot_sync(kFull, sc != 0u);
      if (sc) {
        const int pos = __popc(bal & below);
        src[pos] = make_uint2(off + (unsigned)v, 0u);
        dst[pos] = make_uint2(off + sc, 0u);
      }
      const int adv = __popc(bal);
      src += adv;
      dst += adv;
    }
  }
  if (w != 0) return;
  __syncwarp();
  const int len = trace_path_jump(L, s, sh[7], lane, rp, jump2, jump4);
  TLC(8);
  if (a.ex_path) {
    for (int i = lane; i < len; i += 32) a.ex_path[(size_t)b * n + i] = lds16(rp + 2u * (len - 1 - i));
    if (lane == 0) a.ex_path_len[b] = len;
  }
  int32_t* po = a.pairs ? a.pairs + (size_t)b * a.pmax * 2 : nullptr;
  const int pmax = a.pmax;
  const uint32_t stage = s.rec;
  uint32_t* const rec = a.rec;
  int np = n5_moves(L, s, rp, len, tabu0, tabu1, lane, [&](int idx, int u, int v) {
    if (idx < pmax) {
      if (po) {
        po[2 * idx] = u;
        po[2 * idx + 1] = v;
      }
      if (rec) sts32(stage + 16u + 4u * idx, (int)((uint32_t)u | ((uint32_t)v << 16)));
    }
  });
  TLC(9);
  if (np < 0) { status = status ? status : np; np = 0; }
  if (np > pmax) { status = status ? status : -8; np = pmax; }
  if (lane == 0) {
    const int inc_old = a.st.inc[b], cur_old = a.st.cur[b];
    int reward, inc;
    if (a.is_reset) {
      reward = 0;
      inc = ms;
    } else {
      reward = a.reward_type == 1 ? cur_old - ms : max(inc_old - ms, 0);
      inc = min(inc_old, ms);
    }
    if (!a.export_only) {
      a.st.cur[b] = ms;
      a.st.inc[b] = inc;
    }
    if (a.makespan) a.makespan[b] = (float)ms;
    if (a.incumbent) a.incumbent[b] = (float)inc;
    if (a.reward) a.reward[b] = (float)reward;
    if (a.done) a.done[b] = np == 0;
    if (a.npairs) a.npairs[b] = np;
    if (a.status) a.status[b] = status;
    if (status) record_error(a.st.err, status, b);
    if (rec) {
      sts32(stage, (int)((uint32_t)np | (np == 0 ? 0x10000u : 0u) | ((uint32_t)(status & 0xff) << 24)));
      sts32(stage + 4u, __float_as_int((float)ms));
      sts32(stage + 8u, __float_as_int((float)reward));
      sts32(stage + 12u, __float_as_int((float)inc));
    }
  }
  if (rec) {
    __syncwarp();
    uint32_t* r = rec + (size_t)b * a.rec_stride;
    const int nw_rec = a.rec_lines ? ((4 + np + 15) & ~15) : 4 + np;
    for (int i = lane; i < nw_rec; i += 32) r[i] = (uint32_t)lds32(stage + 4u * i);
  }
  TLC(10); TLC(1);
}

// ------------------------------------------------------------------------------------------------------
// K fused steps with an on-device policy; state stays in shared memory for the whole launch.
// ------------------------------------------------------------------------------------------------------
constexpr int kMaxLog = 16;
struct RunArgs {
  Layout L;
  State st;
  int B;
  int policy, K;
  const int32_t* tape;
  unsigned long long seed;
  long long inst_off, itr0;
  int32_t* taken;
  int n_log;
  int log_steps[kMaxLog];
  float* inc_log;
  int pmax;
  int off_where;   // byte offsets (within the instance's smem) of the where index (uint16 [node_stride]) ...
  int off_plist;   // ... and of the uint16 pair list [pmax][2]
  int off_walk2;   // >= 0: second copy of the walk words (GREEDY with n_m <= 16: two neighbours per sweep)
  int stride;      // shared-memory bytes per instance (state + where + pair list)
};

__global__ void run_kernel(const __grid_constant__ RunArgs a) {
  extern __shared__ __align__(16) unsigned char smem_run[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * (blockDim.x >> 5) + w;
  if (b >= a.B) return;
  const Layout& L = a.L;
  const uint32_t base = smem_addr(smem_run) + (uint32_t)w * (uint32_t)a.stride;
  const Inst s = inst_view(L, base);
  const uint32_t swhere = base + a.off_where, plist = base + a.off_plist;
  const uint32_t rp = s.tq + 4u * L.n1p;   // the backward half of tq is unused here
  const int pmax = a.pmax;

  int tabu0 = a.st.tabu[2 * b], tabu1 = a.st.tabu[2 * b + 1];
  int inc = a.st.inc[b], cur = a.st.cur[b];
  stage_instance(L, s, a.st, b, lane, false);
  __syncwarp();
  for (int i = lane; i < L.n; i += 32) {
    const int wd = lds32(s.walk + 4u * i);
    sts16(swhere + 2u * (wd & (int)kWId), i);
    if (a.off_walk2 >= 0) sts32(base + (uint32_t)a.off_walk2 + 4u * i, wd);
  }
  stage_wait_rest(s);
  __syncwarp();
  int status = 0;
  int ms = evaluate_instance(L, s, lane, false);
  if (ms < 0) status = -5;
  int np = 0;
  auto enumerate = [&]() {
    link_pass<false>(L, s, 0u, lane);
    __syncwarp();
    const int len = trace_path(L, s, ms, lane, rp);
    int c = n5_moves(L, s, rp, len, tabu0, tabu1, lane, [&](int idx, int u, int v) {
      if (idx < pmax) { sts16(plist + 4u * idx, u); sts16(plist + 4u * idx + 2u, v); }
    });
    if (c < 0) { status = status ? status : c; c = 0; }
    if (c > pmax) { status = status ? status : -8; c = pmax; }
    __syncwarp();
    return c;
  };
  if (!status) np = enumerate();
  int li = 0;
  for (int step = 0; step < a.K; ++step) {
    int a0 = 0, a1 = 0;
    if (!status) {
      if (a.policy == 0) {
        a0 = a.tape[((size_t)b * a.K + step) * 2];
        a1 = a.tape[((size_t)b * a.K + step) * 2 + 1];
      } else if (np > 0) {
        int pick = 0;
        if (a.policy == 1) {
          pick = random_pick(a.seed, (uint64_t)(a.inst_off + b), (uint64_t)(a.itr0 + step), np);
        } else {
          // GREEDY: evaluate every neighbour, first minimum wins (conventional_baselines...:186-197)
          int best = 0x7fffffff;
          auto trial = [&](int i) -> int {   // makespan of neighbour i by a forward sweep over a trial swap
            const int p0 = lds16(swhere + 2u * lds16(plist + 4u * i));
            const uint32_t aw = s.walk + 4u * p0;
            const unsigned w0 = (unsigned)lds32(aw), w1 = (unsigned)lds32(aw + 4u);
            __syncwarp();
            if (lane == 0) {   // trial swap (row-end bit stays with the position, history bits untouched)
              sts32(aw, (int)(w1 & ~kWRowEnd));
              sts32(aw + 4u, (int)((w0 & ~kWRowEnd) | (w1 & kWRowEnd)));
            }
            fill_pending(L, s, lane, false);
            __syncwarp();
            const int nms = evaluate_instance(L, s, lane, false);
            if (lane == 0) { sts32(aw, (int)w0); sts32(aw + 4u, (int)w1); }
            __syncwarp();
            return nms;
          };
          if (a.off_walk2 >= 0) {
            // two neighbours per sweep: half-warp h evaluates neighbour i+h forward in tq half h over its own
            // copy of the walk words (a forward sweep needs only n_m <= 16 lanes)
            const uint32_t walk2 = base + (uint32_t)a.off_walk2;
            const int h = lane >> 4, m = lane & 15;
            for (int i = 0; i < np; i += 2) {
              const bool have = i + h < np;
              uint32_t aw = 0u;
              unsigned w0 = 0u, w1 = 0u;
              if (have) {
                const int p0 = lds16(swhere + 2u * lds16(plist + 4u * (i + h)));
                aw = (h ? walk2 : s.walk) + 4u * p0;
                w0 = (unsigned)lds32(aw);
                w1 = (unsigned)lds32(aw + 4u);
              }
              __syncwarp();
              if (have && m == 0) {
                sts32(aw, (int)(w1 & ~kWRowEnd));
                sts32(aw + 4u, (int)((w0 & ~kWRowEnd) | (w1 & kWRowEnd)));
              }
              fill_pending(L, s, lane, true);
              __syncwarp();
              int last = 0;
              const bool ok = wavefront(L, s, m, 0, have && m < L.n_m, last, h, h ? walk2 : s.walk);
              const int msA = __reduce_max_sync(kFull, h == 0 ? last : 0);
              const int msB = __reduce_max_sync(kFull, h == 1 ? last : 0);
              if (have && m == 0) { sts32(aw, (int)w0); sts32(aw + 4u, (int)w1); }
              __syncwarp();
              if (ok) {
                if (msA < best) { best = msA; pick = i; }
                if (i + 1 < np && msB < best) { best = msB; pick = i + 1; }
              } else {   // a cyclic trial (cannot happen for N5 moves): judge the two neighbours one by one
                for (int k = i; k < min(i + 2, np); ++k) {
                  const int nms = trial(k);
                  if (nms >= 0 && nms < best) { best = nms; pick = k; }
                }
              }
            }
          } else {
            for (int i = 0; i < np; ++i) {
              const int nms = trial(i);
              if (nms >= 0 && nms < best) { best = nms; pick = i; }
            }
          }
        }
        a0 = lds16(plist + 4u * pick);
        a1 = lds16(plist + 4u * pick + 2u);
      }
      const int i0 = (a0 >= 1 && a0 <= L.n) ? lds16(swhere + 2u * a0) : -1;
      bool has_t = false;
      const int r = apply_swap(L, s, a0, a1, i0, lane, has_t);
      if (r == 0) {
        tabu0 = a1;
        tabu1 = a0;
        if (lane == 0) {
          sts16(swhere + 2u * a0, i0 + 1);
          sts16(swhere + 2u * a1, i0);
          if (a.off_walk2 >= 0) {   // mirror the three touched words into the second copy
            const uint32_t walk2 = base + (uint32_t)a.off_walk2;
            sts32(walk2 + 4u * i0, lds32(s.walk + 4u * i0));
            sts32(walk2 + 4u * (i0 + 1), lds32(s.walk + 4u * (i0 + 1)));
            if (has_t) sts32(walk2 + 4u * (i0 + 2), lds32(s.walk + 4u * (i0 + 2)));
          }
        }
      } else if (r < 0) {
        status = r;
      }
      if (!status && r == 0) {   // a dummy move leaves the schedule, its makespan and its move set unchanged
        if (a.policy != 2 && L.n_m <= 32) {
          // REPLAY / RANDOM: tq's forward half still holds the completion times of the previous evaluation:
          // only the part of the schedule behind the swapped pair is swept again
          ms = evaluate_forward_from(L, s, lane, i0);
        } else {   // GREEDY: the trial evaluations overwrote them
          fill_pending(L, s, lane, false);
          __syncwarp();
          ms = evaluate_instance(L, s, lane, false);
        }
        if (ms < 0) status = -5;
      }
      if (!status && r == 0) {
        cur = ms;
        inc = min(inc, ms);
        np = enumerate();
      }
    }
    if (a.taken && lane == 0) {
      a.taken[((size_t)b * a.K + step) * 2] = a0;
      a.taken[((size_t)b * a.K + step) * 2 + 1] = a1;
    }
    if (li < a.n_log && step + 1 == a.log_steps[li]) {
      if (lane == 0) a.inc_log[(size_t)li * a.B + b] = (float)inc;
      ++li;
    }
  }
  // write the search state back
  __syncwarp();
  uint32_t* gw = a.st.walk + (size_t)b * L.walk_stride;
  for (int i = lane; i < L.walk_stride; i += 32) gw[i] = (uint32_t)lds32(s.walk + 4u * i);
  if (lane == 0) {
    a.st.tabu[2 * b] = tabu0;
    a.st.tabu[2 * b + 1] = tabu1;
    a.st.cur[b] = cur;
    a.st.inc[b] = inc;
    if (a.st.evalid) a.st.evalid[b] = 0;   // the machine orders moved on: the step kernel's cached evaluation is void
    if (status) record_error(a.st.err, status, b);
  }
}

// ------------------------------------------------------------------------------------------------------
// Stateless evaluator for COO input: drop-in for Evaluator.forward (env/message_passing_evl.py:161-199).
//   pass 1 (coo_ingest_kernel): edges are read once, coalesced (int64 pairs); conjunctive arcs are implicit in
//           the node numbering and only counted, machine arcs are scattered into a uint16 successor array.
//   pass 2 (coo_eval_kernel): one warp per graph: stage successors + durations, follow the n_m machine chains
//           from their heads into walk rows (the machine of an operation is not part of this API), then run
//           the SAME wavefront as the environment and write est / lst / makespan as float32.
// ------------------------------------------------------------------------------------------------------
struct CooWs {
  uint16_t* msucc;      // [B][node_stride]
  unsigned long long* conj_count;
  int32_t* flag;        // device status word
};

struct CooIngestArgs {
  const long long* ei;
  long long E, B;
  int n_j, n_m, n, n1, node_stride;
  unsigned magic_n1, magic_nm;   // floor(2^32 / d) + 1, see fast_div
  unsigned lim;                  // batch * (n+2)
  int vec_ok;                    // rows 16-byte aligned and E even: 2 edges per load
  CooWs ws;
};

// x / d for 0 <= x < 2^31 via one 32x32->64 multiply (magic = floor(2^32/d)+1, exact while x*d' error < 1;
// the host only enables the 32-bit path when N = B*n1 < 2^31 and falls back to 64-bit division otherwise).
__device__ __forceinline__ unsigned fast_div(unsigned x, unsigned magic, unsigned d) {
  unsigned qd = __umulhi(x, magic);
  // magic is floor(2^32/d)+1: q can overshoot by one for large x; correct it
  return qd - ((qd * d > x) ? 1u : 0u);
}

// classify one edge (u -> v): returns 1 for a conjunctive arc, 0 after scattering a machine arc, -1 if malformed.
// Branch-free on purpose (the ingest kernel is instruction-bound and a warp sees a mix of arc kinds): 32-bit
// arithmetic (the host guarantees batch*(n+2) < 2^31), one multiply-shift division to find the graph and one exact
// multiply-shift remainder (operand < 2^16) for "first / last operation of a job".
//   machine arc  : op -> op, unless the ids are consecutive inside one job
//   conjunctive  : S -> first op of a job | last op of a job -> T | op -> next op of the same job
__device__ __forceinline__ int coo_edge(const CooIngestArgs& a, long long u64, long long v64) {
  const unsigned n = (unsigned)a.n, n1 = (unsigned)a.n1, n_m = (unsigned)a.n_m;
  const unsigned u = (unsigned)u64, v = (unsigned)v64;
  const bool small = (((unsigned long long)(u64 | v64)) >> 31) == 0ull;     // neither negative nor >= 2^31
  const unsigned bu = fast_div(u, a.magic_n1, n1);
  const unsigned base = bu * n1;
  const unsigned ul = u - base, vl = v - base;                               // vl wraps around if v is in an earlier graph
  const bool inrange = small & (u < a.lim) & (vl < n1);                      // same graph (implies v < lim)
  const bool uop = (ul - 1u) < n, vop = (vl - 1u) < n;
  const bool consec = vl == ul + 1u;
  const unsigned x = ul == 0u ? vl - 1u : ul;                                // S -> v: is v first?   u -> .: is u last?
  const bool divis = (x == __umulhi(x, a.magic_nm) * n_m) | (n_m == 1u);     // exact for x < 2^16; no 32-bit magic for n_m = 1
  const bool machine = uop & vop & !(consec & !divis);
  const bool conj = ((ul == 0u) & vop & divis) | (uop & (vl == n + 1u) & divis) | (uop & vop & consec & !divis);
  if (inrange & machine) a.ws.msucc[bu * (unsigned)a.node_stride + ul] = (uint16_t)vl;
  return inrange & (machine | conj) ? (machine ? 0 : 1) : -1;
}

// Each thread takes kIngestTrips x 2 edges as 16-byte loads from both rows of the [2,E] list, all issued before the
// first use: the kernel is a pure HBM stream (16 B read per edge, 2 B scattered write per machine arc).  The rows
// are 16-byte aligned when E is even; otherwise (or for a misaligned view) every thread takes single edges.
constexpr int kIngestTrips = 4;
__global__ void __launch_bounds__(256) coo_ingest_kernel(const __grid_constant__ CooIngestArgs a) {
  // programmatic dependent launch: the per-graph kernel may be scheduled as soon as every CTA of this grid has started
  // (its prologue -- barrier set-up, the bulk copy of the durations -- does not depend on this kernel; it waits with
  // griddepcontrol.wait before it touches the successor array)
  asm volatile("griddepcontrol.launch_dependents;");
  int conj = 0, bad = 0;
  const unsigned E = (unsigned)a.E;
  const unsigned tid0 = blockIdx.x * (256u * kIngestTrips) + threadIdx.x;
  if (a.vec_ok) {
    const longlong2* us = reinterpret_cast<const longlong2*>(a.ei);
    const longlong2* vs = reinterpret_cast<const longlong2*>(a.ei + a.E);
    const unsigned E2 = E >> 1;
    longlong2 u[kIngestTrips], v[kIngestTrips];
#pragma unroll
    for (int k = 0; k < kIngestTrips; ++k) {
      const unsigned e = tid0 + 256u * k;
      if (e < E2) { u[k] = us[e]; v[k] = vs[e]; }
    }
#pragma unroll
    for (int k = 0; k < kIngestTrips; ++k) {
      const unsigned e = tid0 + 256u * k;
      if (e < E2) {
        const int r0 = coo_edge(a, u[k].x, v[k].x), r1 = coo_edge(a, u[k].y, v[k].y);
        conj += (r0 > 0) + (r1 > 0);
        bad |= (r0 < 0) | (r1 < 0);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < kIngestTrips; ++k) {
      const unsigned e = tid0 + 256u * k;
      if (e < E) {
        const int r = coo_edge(a, a.ei[e], a.ei[a.E + e]);
        conj += r > 0;
        bad |= r < 0;
      }
    }
  }
  if (bad) atomicCAS(a.ws.flag, 0, -5);
  // one atomic per CTA
  __shared__ int warp_sums[8];
  for (int o = 16; o; o >>= 1) conj += __shfl_xor_sync(kFull, conj, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = conj;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = threadIdx.x < 8 ? warp_sums[threadIdx.x] : 0;
    for (int o = 4; o; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    if (threadIdx.x == 0 && t) atomicAdd(a.ws.conj_count, (unsigned long long)t);
  }
}

struct CooEvalArgs {
  Layout L;            // tq + walk offsets are reused
  long long B;
  const void* duration;
  int dur_i64;
  int dur_tma;         // int64 rows 16-byte aligned: the raw durations are bulk-copied into tq and converted from there
  CooWs ws;
  float* est;
  float* lst;
  float* ms;
  int bytes;           // smem per graph
  int off_durw, off_succ, off_heads, off_bar;
  int off_anchor;      // CTA kernel: uint16 [n_m][ceil(n_j/4)] chain anchors
  int fast_sweep;      // CTA kernel: latency variant of the wavefront
  long long expect_conj;
};

__global__ void __launch_bounds__(128, 8) coo_eval_kernel(const __grid_constant__ CooEvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem_coo[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + w;
  if (b >= a.B) return;
  const Layout& L = a.L;
  const uint32_t base = smem_addr(smem_coo) + (uint32_t)w * (uint32_t)a.bytes;
  const Inst s = inst_view(L, base);
  const uint32_t succ = base + a.off_succ, heads = base + a.off_heads;   // uint16 [node_stride], uint16 [32]
  const uint32_t durw = base + a.off_durw;                               // uint16 [node_stride] node words
  const int n = L.n, n1 = L.n1, n_m = L.n_m, n_j = L.n_j;
  TL(0); TL(2);

  // stage: machine successors (and, for the plain path, the "pending" sentinels) by TMA bulk copies issued by one
  // lane.  Durations: when the int64 rows are 16-byte aligned they are bulk-copied RAW into tq (8*n1 <= 8*n1p bytes)
  // on a second barrier and converted after the chain heads have been found, so that the largest input of the call
  // is in flight while the warp works; otherwise they are loaded and converted here.
  // Everything up to griddepcontrol.wait is independent of the ingest kernel and may overlap its tail.
  const uint32_t bar = base + a.off_bar, bar2 = bar + 8u;
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    if (a.dur_tma) {
      mbar_expect_tx(bar2, (uint32_t)n1 * 8u);
      bulk_g2s(s.tq, reinterpret_cast<const long long*>(a.duration) + b * n1, (uint32_t)n1 * 8u, bar2);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (b == 0 && lane == 0) {
    if (*a.ws.conj_count != (unsigned long long)a.expect_conj) atomicCAS(a.ws.flag, 0, -5);
    *a.ws.conj_count = 0ull;   // the workspace is handed back zero-filled (see l2s_eval_coo)
  }
  if (lane == 0) {
    const uint32_t sbytes = (uint32_t)L.node_stride * 2u;
    mbar_expect_tx(bar, sbytes);
    bulk_g2s(succ, a.ws.msucc + b * L.node_stride, sbytes, bar);
  }
  if (!a.dur_tma) fill_pending(L, s, lane, true);   // tq is free: "pending" sentinels right away
  int bad = 0;
  // node word of node i from its duration d; pos = (i-1) mod n_m
  auto node_word = [&](int i, long long d, int pos) -> int {
    if (d < 0 || d > kDurMax) { bad = 1; d = 0; }
    int word = (int)d;
    if (i >= 1 && i <= n) word |= (pos == 0 ? kFirstFlag : 0) | (pos == n_m - 1 ? kLastFlag : 0);
    else if (d != 0) bad = 1;
    return word;
  };
  const int pos0 = lane == 0 ? n_m - 1 : (lane - 1) % n_m;   // (i-1) mod n_m for i = lane, advanced by 32 per trip
  const int step = 32 % n_m;
  if (!a.dur_tma) {
    int pos = pos0;
    // all loads of a group of kDurTrips trips are issued before the first shared-memory store (which is an asm
    // with a memory clobber and would otherwise serialise the trips into a chain of global-load latencies)
    constexpr int kDurTrips = 6;
    for (int i0 = lane; i0 < n1; i0 += 32 * kDurTrips) {
      long long dv[kDurTrips];
#pragma unroll
      for (int k = 0; k < kDurTrips; ++k) {
        const int i = min(i0 + 32 * k, n1 - 1);
        dv[k] = a.dur_i64 ? reinterpret_cast<const long long*>(a.duration)[b * n1 + i]
                          : (long long)reinterpret_cast<const int32_t*>(a.duration)[b * n1 + i];
      }
#pragma unroll
      for (int k = 0; k < kDurTrips; ++k) {
        const int i = i0 + 32 * k;
        if (i < n1) {
          sts16(durw + 2u * i, node_word(i, dv[k], pos));
          pos += step;
          if (pos >= n_m) pos -= n_m;
        }
      }
    }
  }
  __syncwarp();
  TL(3);   // durations converted (plain path)
  mbar_wait(bar, 0);
  TL(4);   // staged
  {   // the successors are in shared memory: hand the row of the workspace back zero-filled
    uint4* grow = reinterpret_cast<uint4*>(a.ws.msucc + b * L.node_stride);
    for (int i = lane; i < L.node_stride / 8; i += 32) grow[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // heads = operations nobody points to.  One flag BYTE per node in the (still unused) walk words: plain stores, no
  // atomics (a bitmap needed shared-memory atomics on a handful of words: ~10-way serialised)
  const uint32_t flags = s.walk;                      // n1 bytes <= 4*walk_stride bytes
  const int nflag16 = (n1 + 15) >> 4;
  for (int i = lane; i < nflag16; i += 32) sts128(flags + 16u * i, make_uint4(0u, 0u, 0u, 0u));
  __syncwarp();
  int narc = 0;
  for (int v = 1 + lane; v <= n; v += 32) {
    const int sc = lds16(succ + 2u * v);
    if (sc) { sts8(flags + (uint32_t)sc, 1); ++narc; }
  }
  narc = __reduce_add_sync(kFull, narc);
  __syncwarp();
  int nh = 0;
  for (int v0 = 1; v0 <= n; v0 += 32) {
    const int v = v0 + lane;
    const bool ish = v <= n && lds8(flags + (uint32_t)v) == 0;
    const unsigned bh = __ballot_sync(kFull, ish);
    const int ph = nh + __popc(bh & ((1u << lane) - 1u));
    if (ish && ph < 32) sts16(heads + 2u * ph, v);
    nh += __popc(bh);
  }
  if (a.dur_tma) {   // the raw durations have landed in tq by now: convert, then turn tq into "pending" sentinels
    mbar_wait(bar2, 0);
    int pos = pos0;
    for (int i = lane; i < n1; i += 32) {
      const int2 dd = lds64(s.tq + 8u * i);
      const long long d = (long long)(((unsigned long long)(unsigned)dd.y << 32) | (unsigned)dd.x);
      sts16(durw + 2u * i, node_word(i, d, pos));
      pos += step;
      if (pos >= n_m) pos -= n_m;
    }
    __syncwarp();
    fill_pending(L, s, lane, true);
  }
  __syncwarp();
  bad = __reduce_or_sync(kFull, bad);
  // a complete selection has exactly n_m machine chains of n_j operations: n_m heads, n - n_m machine arcs
  if (bad || nh != n_m || narc != n - n_m) {
    if (lane == 0) atomicCAS(a.ws.flag, 0, bad ? -6 : -5);
    return;
  }
  TL(5);   // heads found
  // follow chain `lane` from its head into walk row `lane` (pointer chase, n_j steps)
  int chain_ok = 1;
  if (lane < n_m) {
    int v = lds16(heads + 2u * lane);
    for (int k = 0; k < n_j; ++k) {
      if (v == 0) { chain_ok = 0; break; }
      const int nx = lds16(succ + 2u * v);
      sts32(s.walk + 4u * (lane * n_j + k), (int)make_walk(v, lds16(durw + 2u * v), false, k == n_j - 1));
      v = nx;
    }
    if (v != 0) chain_ok = 0;      // longer than n_j (or a cycle)
  }
  chain_ok = __reduce_and_sync(kFull, chain_ok);
  __syncwarp();
  TL(6);   // walk rows built
  const int ms = chain_ok ? evaluate_instance(L, s, lane, true) : -1;
  TL(7);   // swept
  if (ms < 0) {
    if (lane == 0) atomicCAS(a.ws.flag, 0, -5);
    return;
  }
  const uint32_t fin = s.tq, q = s.tq + 4u * L.n1p;
  int qs = 0;
  for (int j = lane; j < n_j; j += 32) qs = max(qs, lds32(q + 4u * (j * n_m + 1)));
  qs = __reduce_max_sync(kFull, qs);
  if (lane == 0) {
    sts32(fin, 0);
    sts32(q, qs);
    sts32(fin + 4u * (n + 1), ms);
    sts32(q + 4u * (n + 1), 0);
  }
  __syncwarp();
  float* eo = a.est + b * n1;
  float* lo = a.lst + b * n1;
  for (int i = lane; i < n1; i += 32) {
    const int d = lds16(durw + 2u * i) & kDurMask;
    eo[i] = (float)(lds32(fin + 4u * i) - d);
    lo[i] = (float)(ms - lds32(q + 4u * i));
  }
  if (lane == 0) a.ms[b] = (float)ms;
  TL(8); TL(1);
}


// ------------------------------------------------------------------------------------------------------
// The same evaluator for LARGE graphs (fewer than 16 fit an SM, so a warp per graph would leave the SM almost empty:
// 3 warps at 150x25): one CTA of 4-8 warps per graph.  Staging, head finding, the duration conversion and the outputs
// are spread over all threads; the machine chains are turned into walk rows with the jump-pointer trick of the step
// kernel's path trace (successor map squared twice, one lane per chain drops an anchor every fourth position, all
// threads fill the three operations behind every anchor); warp 0 sweeps forward while warp 1 sweeps backward.
// Shared memory: the warp kernel's regions plus the anchors (uint16 [n_m][ceil(n_j/4)]).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) coo_eval_kernel_cta(const __grid_constant__ CooEvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem_coo[];
  __shared__ int sh[8];   // 0: heads, 1: machine arcs, 2: bad, 3: makespan, 4: fwd ok, 5: bwd ok, 6: qs
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
  const long long b = blockIdx.x;
  const Layout& L = a.L;
  const uint32_t base = smem_addr(smem_coo);
  const Inst s = inst_view(L, base);
  const uint32_t succ = base + a.off_succ, heads = base + a.off_heads, durw = base + a.off_durw;
  const uint32_t anchors = base + a.off_anchor;
  const int n = L.n, n1 = L.n1, n_m = L.n_m, n_j = L.n_j;
  const int n4 = (n_j + 3) >> 2;                         // anchors per chain
  if (tid < 8) sh[tid] = 0;
  const uint32_t bar = base + a.off_bar, bar2 = bar + 8u;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    if (a.dur_tma) {
      mbar_expect_tx(bar2, (uint32_t)n1 * 8u);
      bulk_g2s(s.tq, reinterpret_cast<const long long*>(a.duration) + b * n1, (uint32_t)n1 * 8u, bar2);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");      // everything above may overlap the tail of the ingest kernel
  if (tid == 0) {
    if (b == 0) {
      if (*a.ws.conj_count != (unsigned long long)a.expect_conj) atomicCAS(a.ws.flag, 0, -5);
      *a.ws.conj_count = 0ull;   // the workspace is handed back zero-filled (see l2s_eval_coo)
    }
    const uint32_t sbytes = (uint32_t)L.node_stride * 2u;
    mbar_expect_tx(bar, sbytes);
    bulk_g2s(succ, a.ws.msucc + b * L.node_stride, sbytes, bar);
  }
  int bad = 0;
  auto node_word = [&](int i, long long d) -> int {     // node word of node i from its duration d
    if (d < 0 || d > kDurMax) { bad = 1; d = 0; }
    int word = (int)d;
    if (i >= 1 && i <= n) {
      const int pos = (i - 1) % n_m;
      word |= (pos == 0 ? kFirstFlag : 0) | (pos == n_m - 1 ? kLastFlag : 0);
    } else if (d != 0) {
      bad = 1;
    }
    return word;
  };
  if (!a.dur_tma) {
    for (int i = tid; i < n1; i += nthr) {
      const long long d = a.dur_i64 ? reinterpret_cast<const long long*>(a.duration)[b * n1 + i]
                                    : (long long)reinterpret_cast<const int32_t*>(a.duration)[b * n1 + i];
      sts16(durw + 2u * i, node_word(i, d));
    }
  }
  // flag bytes (in the still unused walk words): 1 = somebody points to this node
  const uint32_t flags = s.walk;
  for (int i = tid; i < (n1 + 15) >> 4; i += nthr) sts128(flags + 16u * i, make_uint4(0u, 0u, 0u, 0u));
  __syncthreads();                                         // barriers initialised, flags cleared
  mbar_wait(bar, 0);
  {   // the successors are in shared memory: hand the row of the workspace back zero-filled
    uint4* grow = reinterpret_cast<uint4*>(a.ws.msucc + b * L.node_stride);
    for (int i = tid; i < L.node_stride / 8; i += nthr) grow[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  int narc = 0;
  for (int v = 1 + tid; v <= n; v += nthr) {
    const int sc = lds16(succ + 2u * v);
    if (sc) { sts8(flags + (uint32_t)sc, 1); ++narc; }
  }
  narc = __reduce_add_sync(kFull, narc);
  if (lane == 0 && narc) atomicAdd(&sh[1], narc);
  if (a.dur_tma) {   // the raw durations have landed in tq by now: convert them
    mbar_wait(bar2, 0);
    for (int i = tid; i < n1; i += nthr) {
      const int2 dd = lds64(s.tq + 8u * i);
      sts16(durw + 2u * i, node_word(i, (long long)(((unsigned long long)(unsigned)dd.y << 32) | (unsigned)dd.x)));
    }
  }
  __syncthreads();                                         // flags complete, tq free
  // heads = operations nobody points to (any order: a chain may take any walk row); successor map squared twice
  const uint32_t s2 = s.tq, s4 = s.tq + 2u * (uint32_t)L.n1p;
  for (int v = 1 + tid; v <= n; v += nthr) {
    if (lds8(flags + (uint32_t)v) == 0) {
      const int slot = atomicAdd(&sh[0], 1);
      if (slot < 32) sts16(heads + 2u * slot, v);
    }
  }
  square_map(s2, succ, n1, tid, nthr);
  if (__syncthreads_or(bad)) {
    if (tid == 0) atomicCAS(a.ws.flag, 0, -6);
    return;
  }
  square_map(s4, s2, n1, tid, nthr);
  // a complete selection has exactly n_m machine chains of n_j operations: n_m heads, n - n_m machine arcs
  if (sh[0] != n_m || sh[1] != n - n_m) {
    if (tid == 0) atomicCAS(a.ws.flag, 0, -5);
    return;
  }
  __syncthreads();
  // one lane per chain follows the fourth-successor map: an anchor every fourth operation
  if (w == 0 && lane < n_m) {
    int v = lds16(heads + 2u * lane), k = 0;
    while (v != 0 && k < n4) {
      sts16(anchors + 2u * (uint32_t)(lane * n4 + k), v);
      v = lds16(s4 + 2u * v);
      ++k;
    }
    if (k != n4 || v != 0) sh[2] = 1;                       // shorter or longer than n_j (or a cycle)
  }
  __syncthreads();
  if (sh[2]) {
    if (tid == 0) atomicCAS(a.ws.flag, 0, -5);
    return;
  }
  // all threads: the walk words of every anchor and of the (up to) three operations behind it
  int chain_bad = 0;
  for (int idx = tid; idx < n_m * n4; idx += nthr) {
    const int k = idx / n4, i = idx - k * n4;
    const int v0 = lds16(anchors + 2u * idx);
    const int v1 = lds16(succ + 2u * v0), v2 = lds16(s2 + 2u * v0);
    const int v3 = lds16(succ + 2u * v2);
    const int vs[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pos = 4 * i + j;
      if (pos < n_j) {
        if (vs[j] == 0) chain_bad = 1;
        else sts32(s.walk + 4u * (uint32_t)(k * n_j + pos), (int)make_walk(vs[j], lds16(durw + 2u * vs[j]), false, pos == n_j - 1));
      } else if (vs[j] != 0) {
        chain_bad = 1;
      }
    }
  }
  if (__syncthreads_or(chain_bad)) {                        // (also: walk rows complete, the jump maps are dead)
    if (tid == 0) atomicCAS(a.ws.flag, 0, -5);
    return;
  }
  fill_pending(L, s, tid, true, nthr);
  __syncthreads();
  __syncwarp();
  if (w == 0) {
    int last;
    const bool ok = a.fast_sweep == 2 ? wavefront_fast2(L, s, lane, 0, lane < n_m, last)
                    : a.fast_sweep ? wavefront_fast(L, s, lane, 0, lane < n_m, last) : wavefront(L, s, lane, 0, lane < n_m, last);
    const int ms0 = __reduce_max_sync(kFull, last);
    if (lane == 0) { sh[3] = ms0; sh[4] = ok; }
  } else if (w == 1) {
    int l2;
    const bool ok2 = a.fast_sweep == 2 ? wavefront_fast2(L, s, lane, 1, lane < n_m, l2)
                     : a.fast_sweep ? wavefront_fast(L, s, lane, 1, lane < n_m, l2) : wavefront(L, s, lane, 1, lane < n_m, l2);
    if (lane == 0) sh[5] = ok2;
  }
  __syncthreads();
  const int ms = sh[3];
  if (!sh[4] || !sh[5]) {
    if (tid == 0) atomicCAS(a.ws.flag, 0, -5);
    return;
  }
  const uint32_t fin = s.tq, q = s.tq + 4u * L.n1p;
  {
    int qs = 0;
    for (int j = tid; j < n_j; j += nthr) qs = max(qs, lds32(q + 4u * (j * n_m + 1)));
    qs = __reduce_max_sync(kFull, qs);
    if (lane == 0 && qs > 0) atomicMax(&sh[6], qs);
  }
  __syncthreads();
  if (tid == 0) {
    sts32(fin, 0);
    sts32(q, sh[6]);
    sts32(fin + 4u * (n + 1), ms);
    sts32(q + 4u * (n + 1), 0);
    a.ms[b] = (float)ms;
  }
  __syncthreads();
  float* eo = a.est + b * n1;
  float* lo = a.lst + b * n1;
  for (int i = tid; i < n1; i += nthr) {
    const int d = lds16(durw + 2u * i) & kDurMask;
    eo[i] = (float)(lds32(fin + 4u * i) - d);
    lo[i] = (float)(ms - lds32(q + 4u * i));
  }
  (void)nw;
}

}  // namespace l2s
