// l2s_b200 device-side building blocks (sm_100a).  Integer graph work: no tensor cores by design.
//
// Execution model: ONE WARP OWNS ONE INSTANCE.  The instance's whole search state lives in that warp's
// slice of shared memory, so no block-level barrier is ever needed; every synchronisation is a
// __syncwarp / vote.  A CTA is just a bundle of independent warps (so that > 32 instances fit one SM).
//
// Longest paths are computed WORK-EFFICIENTLY (O(n) node updates instead of the reference's O(n * depth)
// Jacobi sweeps, env/message_passing_evl.py:176-197): every lane "walks" one machine's processing
// sequence.  Its machine predecessor's value is in a register; only the job neighbour is read from
// shared memory (sentinel -1 = not final yet).  One operation per lane per level => the number of
// iterations equals the hop depth of the graph (27..210 at the reference's sizes) and each iteration is a
// couple of dozen warp instructions regardless of n.  With n_m <= 16 the forward sweep (lanes 0-15) and the
// backward sweep (tails q, lanes 16-31) run in the SAME iterations.
//
// State format ("walk words"): the processing order of machine m is the row walk[m*n_j .. m*n_j+n_j) of
// 32-bit words, one per operation:
//     bits  0..15  node id (1..n)
//     bit  16      first operation of its job   (job predecessor = S)
//     bit  17      last operation of its job    (job successor = T)
//     bit  18      "jobfirst": the job predecessor precedes the machine predecessor in networkx's
//                  insertion order (set when a swap re-adds the machine in-edge)
//     bit  19      last position of its machine row
//     bits 20..31  duration (<= 4095; the top bits, so that a sweep extracts it with one shift)
// so that one shared-memory load gives a walker everything about its next operation.  where[v] is the index
// of node v's word.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// The per-node loops of the step kernel are kept rolled: unrolled by the compiler they grow the kernel by 600
// instructions, and with 40 warps at 40 different places of a long straight-line kernel the instruction cache is a
// measured cost (5 % of the stall samples are instruction fetches; 10x10 at 8192 instances: 22.5 -> 20.6 us rolled).
#define L2S_LOOP _Pragma("unroll 1")

namespace l2s {

constexpr int kDurMax = 4095;
// node words (uint16, indexed by node id, static): duration | flags
constexpr int kFirstFlag = 0x8000;
constexpr int kLastFlag = 0x4000;
constexpr int kDurMask = 0x0FFF;
// walk words
constexpr unsigned kWId = 0xffffu;
constexpr int kWDurShift = 20;
constexpr unsigned kWFirst = 1u << 16;
constexpr unsigned kWLast = 1u << 17;
constexpr unsigned kWJf = 1u << 18;
constexpr unsigned kWRowEnd = 1u << 19;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ __forceinline__ unsigned make_walk(int v, int durw, bool jf, bool rowend) {
  return (unsigned)v | ((unsigned)(durw & kDurMask) << kWDurShift) | ((durw & kFirstFlag) ? kWFirst : 0u) |
         ((durw & kLastFlag) ? kWLast : 0u) | (jf ? kWJf : 0u) | (rowend ? kWRowEnd : 0u);
}

// Shapes and shared-memory layout of one instance (all offsets in bytes, 16-byte aligned).
struct Layout {
  int n_j, n_m, n, n1;   // n = n_j*n_m operations, n1 = n+2 nodes
  int n1p;               // n1 rounded up to 4 (int32 elements per tq half)
  int walk_stride;       // uint32 words per instance (multiple of 4, >= n)
  int node_stride;       // uint16 elements per instance of node-indexed arrays (multiple of 8, >= n1)
  int mch_stride;        // bytes per instance of the machine-id array (multiple of 16)
  int e_mc;              // machine edges per instance n_m*(n_j-1)
  int off_tq, off_walk, off_crit, off_misc;
  int off_durw;          // >= 0: static node words staged in shared memory (small instances); -1: read from HBM
  int off_rec;           // staging of the host record (4 + pmax words): overlays crit when that is large enough
  int off_msu;           // >= 0 (CTA kernel, few resident instances): own regions for the machine successors, the path
  int off_rp, off_j2, off_j4;   //   and the jump maps, so that feature / edge output overlaps the path trace; -1: overlays
  int off_bits;          // n5_moves scratch (one bit per path position): in crit, behind the record staging if that overlays it
  unsigned magic_nm;     // floor(2^32 / n_m) + 1: (x mod n_m) by multiply-shift for x < 2^16
  int bytes;             // per instance (step kernel)
};

// Persistent per-slice search state in HBM (SoA, one row per instance).
struct State {
  uint32_t* walk;   // [B][walk_stride]  machine orders as walk words
  uint16_t* where;  // [B][node_stride]  scratch (duplicate detection in l2s_set_solution); kernels rebuild positions
  uint16_t* durw;   // [B][node_stride]  node words (duration | flags), indexed by node id; 0 for S,T
  uint8_t* mch;     // [B][mch_stride]   1-based machine of op v at [v-1] (setup/validation only)
  int32_t* tabu;    // [B][2]
  int32_t* cur;     // [B] current makespan
  int32_t* inc;     // [B] incumbent makespan
  int32_t* err;     // [1] first device-side error: (code<<24)|instance  (0 = none)
  const int32_t* pending_page;   // [2*n1p] words of -1: source of the TMA fill of tq
  // Incremental re-evaluation behind the step API (CTA-per-instance kernel only; nullptr otherwise): completion
  // times and tails of the last evaluation, kept between launches, and a per-instance "matches walk" flag
  int32_t* fq;      // [B][2*n1p]  fin | q exactly as they stand in shared memory after an evaluation
  int32_t* evalid;  // [B]         1: fq[b] belongs to the current machine orders
};

// 32-bit shared-memory addressing (keeps the hot loops free of 64-bit generic pointer arithmetic)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds32(uint32_t a) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int2 lds64(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds16(uint32_t a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return (int)v;
}
__device__ __forceinline__ void sts16(uint32_t a, int v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}

__device__ __forceinline__ int lds8(uint32_t a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return (int)v;
}
__device__ __forceinline__ void sts8(uint32_t a, int v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

struct Inst {            // shared-memory view of one instance (32-bit shared addresses)
  uint32_t tq;           // int32 [2][n1p]: [v] = est[v]+dur[v] ("fin"), [n1p+v] = tail q[v]; -1 = pending
  uint32_t walk;         // uint32 [walk_stride]
  uint32_t crit;         // uint16 [node_stride] chosen critical predecessor (networkx tie-break); 0 = S
  uint32_t durw;         // uint16 [node_stride] node words, valid iff Layout::off_durw >= 0
  uint32_t misc;         // int32 [4] parked scalars
  uint32_t rec;          // uint32 [4 + pmax] host record staging (valid after the critical path has been traced)
  uint32_t bits;         // uint32 [n/32 + 2] n5_moves scratch (valid after the critical path has been traced)
};

__device__ __forceinline__ Inst inst_view(const Layout& L, uint32_t base) {
  Inst s;
  s.tq = base + L.off_tq;
  s.walk = base + L.off_walk;
  s.crit = base + L.off_crit;
  s.durw = base + (L.off_durw >= 0 ? L.off_durw : 0);
  s.misc = base + L.off_misc;
  s.rec = base + L.off_rec;
  s.bits = base + L.off_bits;
  return s;
}

__device__ __forceinline__ void record_error(int32_t* err, int code, int b) {
  atomicCAS(err, 0, (int)(((unsigned)(-code) << 24) | ((unsigned)b & 0xffffffu)));   // keep the first one
}

__device__ __forceinline__ uint4 ld_nc16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void fill_pending(const Layout& L, const Inst& s, int lane, bool fill_q, int nthr = 32) {
  const uint4 pending = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  const int cnt = (fill_q ? 2 : 1) * (L.n1p / 4);
  for (int i = lane; i < cnt; i += nthr) sts128(s.tq + 16 * i, pending);
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: one lane moves a whole array into shared memory
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

// Staging of one instance's state into shared memory: ONE elected lane issues TMA bulk copies (walk words and,
// for small instances, the static node words; 16-byte aligned, sizes multiples of 16) that complete on an
// mbarrier in the instance's scratch words; the "pending" sentinels of tq are bulk-copied from a page of -1 words
// (L2 resident).  Every lane then waits on the barrier (phase 0; each instance slot is used once per kernel).
// preload (CTA kernel): tq is bulk-copied from the cached evaluation st.fq instead of being filled with sentinels
__device__ __forceinline__ void stage_instance(const Layout& L, const Inst& s, const State& st, int b, int lane,
                                               bool fill_q, int nthr = 32, bool preload = false) {
  const uint32_t bar = s.misc + 8u, bar2 = s.misc + 16u;
  if (lane == 0) {
    const uint32_t wbytes = (uint32_t)L.walk_stride * 4u, dbytes = L.off_durw >= 0 ? (uint32_t)L.node_stride * 2u : 0u;
    // warp-per-instance: the "pending" sentinels come from a -1 page in L2 too; a whole CTA fills them itself
    const uint32_t tbytes = nthr > 32 ? 0u : (fill_q ? 2u : 1u) * (uint32_t)L.n1p * 4u;
    // two barriers: the walk words gate the start (action lookup, swap); the sentinels and the node words are only
    // needed from the evaluation on (stage_wait_rest) and land while the swap is being applied
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    mbar_expect_tx(bar, wbytes);
    bulk_g2s(s.walk, st.walk + (size_t)b * L.walk_stride, wbytes, bar);
    const uint32_t pbytes = preload ? 2u * (uint32_t)L.n1p * 4u : 0u;
    mbar_expect_tx(bar2, dbytes + tbytes + pbytes);
    if (tbytes) bulk_g2s(s.tq, st.pending_page, tbytes, bar2);
    if (pbytes) bulk_g2s(s.tq, st.fq + (size_t)b * 2 * L.n1p, pbytes, bar2);
    if (dbytes) bulk_g2s(s.durw, st.durw + (size_t)b * L.node_stride, dbytes, bar2);
  }
  if (nthr > 32 && !preload) fill_pending(L, s, lane, fill_q, nthr);
  if (nthr > 32) __syncthreads(); else __syncwarp();   // the barrier words are initialised before anyone polls them
  mbar_wait(bar, 0);
}
// second half of the staging: sentinels + node words have landed
__device__ __forceinline__ void stage_wait_rest(const Inst& s) { mbar_wait(s.misc + 16u, 0); }

// ------------------------------------------------------------------------------------------------------
// Level-synchronous wavefront.  Lane role: machine row `m`, direction dir (0: forward est/fin from the head
// of the row, 1: backward tail q from the end of the row); inactive lanes just vote.
// Returns false if the selection is cyclic (no progress possible).  `last` = final value of the lane's
// chain (forward: completion time of the machine's last operation).
// The loop body is kept minimal (one shared load, one shared store, a handful of ALU ops per level): it is
// the serial spine of the whole step.  Critical predecessors / machine successors are derived afterwards in
// one data-parallel pass (link_pass).
// ------------------------------------------------------------------------------------------------------
#ifdef L2S_TIMELINE
__device__ unsigned long long g_dbg_levels[2] = {0ull, 0ull};   // diagnostic build: sum of level trips, number of sweeps
#endif
#ifndef L2S_LEVELS_PER_TRIP
#define L2S_LEVELS_PER_TRIP 4
#endif
// half / walk: which half of tq the lane's values live in and which copy of the walk words it reads (defaults:
// half = dir, the instance's own walk words); the greedy policy evaluates two neighbours at once with both
// half-warps sweeping FORWARD, each in its own tq half over its own copy of the walk words.
template <bool kOnePtr = false>
__device__ __forceinline__ bool wavefront(const Layout& L, const Inst& s, int m, int dir, bool active, int& last,
                                          int half = -1, uint32_t walk = 0u, int start = 0, int prev0 = 0) {
  const int n_j = L.n_j;
  if (half < 0) half = dir;
  if (walk == 0u) walk = s.walk;
  const uint32_t a_self = s.tq + (half ? 4u * L.n1p : 0u);           // &tq[base + 0]
  const uint32_t a_nb = dir ? a_self + 4u : a_self - 4u;             // &tq[base +- 1]
  const unsigned bflag = dir ? kWLast : kWFirst;
  const int stepb = dir ? -4 : 4;                                    // walk step in bytes; also a_self - a_nb
  // start / prev0: resume the lane's row behind its first `start` operations in sweep order (forward: from the head
  // of the row, backward: from its end) with the value prev0 of the last operation that is kept (incremental
  // re-evaluation: the operations before `start` keep their values)
  uint32_t a_w = walk + 4u * (uint32_t)(m * n_j + (dir ? n_j - 1 - start : start));
  int left = active ? n_j - start : 0;                               // operations still to place
  int prevval = prev0;
  unsigned w = 0, wn = 0;
  if (left > 0) w = (unsigned)lds32(a_w);
  if (left > 1) wn = (unsigned)lds32(a_w + stepb);
  // kOnePtr (per-lane direction, step in a register): a_w becomes the address of the word that turns into `wn` at
  // the next placement, which saves an add per level; with a compile-time direction the +-4 folds into the load
  // and the two-pointer form schedules better (measured: 30x20 1.9 % faster)
  if (kOnePtr) a_w += 2 * stepb;
  int budget = (L.n + 3) >> 1;  // a DAG needs at most n levels; two levels per trip
  // One level: place the lane's next operation if its job neighbour is final.  Straight-line (predicated)
  // code: the warp stays converged, so shared-memory writes of one level are visible to the next.
  auto level = [&]() {
    // Hand-scheduled: 19-20 instructions, fully predicated.  Equivalent C:
    //   a = a_nb + 4*(w & 0xffff); nb = left > 0 ? tq[a] : -1; if (w & bflag) nb = 0;
    //   adv = left > 0 && nb >= 0; val = max(prevval, nb) + dur(w);
    //   if (adv) { tq[a + dself] = val; prevval = val; w = wn; a_w += stepb; --left; wn = walk[a_w + stepb]; }
    //   (the prefetch past the end of a row reads a neighbouring word of the instance's own shared memory: never used)
    if (kOnePtr) {
      asm volatile(
          "{\n\t"
          ".reg .pred pa, pb, pv;\n\t"
          ".reg .b32 a, nb, t, d, val;\n\t"
          "setp.gt.s32 pa, %2, 0;\n\t"
          "and.b32 a, %0, 0xffff;\n\t"
          "mad.lo.u32 a, a, 4, %5;\n\t"
          "mov.b32 nb, -1;\n\t"
          "@pa ld.shared.b32 nb, [a];\n\t"
          "and.b32 t, %0, %6;\n\t"
          "setp.ne.u32 pb, t, 0;\n\t"
          "selp.b32 nb, 0, nb, pb;\n\t"
          "setp.ge.and.s32 pv, nb, 0, pa;\n\t"
          "shr.u32 d, %0, 20;\n\t"
          "max.s32 val, %3, nb;\n\t"
          "add.s32 val, val, d;\n\t"
          "add.s32 a, a, %7;\n\t"
          "@pv st.shared.b32 [a], val;\n\t"
          "@pv mov.b32 %3, val;\n\t"
          "@pv mov.b32 %0, %1;\n\t"
          "@pv add.s32 %2, %2, -1;\n\t"
          "@pv ld.shared.b32 %1, [%4];\n\t"
          "@pv add.s32 %4, %4, %7;\n\t"
          "}\n"
          : "+r"(w), "+r"(wn), "+r"(left), "+r"(prevval), "+r"(a_w)
          : "r"(a_nb), "r"(bflag), "r"(stepb)
          : "memory");
    } else {
      asm volatile(
          "{\n\t"
          ".reg .pred pa, pb, pv;\n\t"
          ".reg .b32 a, nb, t, d, val;\n\t"
          "setp.gt.s32 pa, %2, 0;\n\t"
          "and.b32 a, %0, 0xffff;\n\t"
          "mad.lo.u32 a, a, 4, %5;\n\t"
          "mov.b32 nb, -1;\n\t"
          "@pa ld.shared.b32 nb, [a];\n\t"
          "and.b32 t, %0, %6;\n\t"
          "setp.ne.u32 pb, t, 0;\n\t"
          "selp.b32 nb, 0, nb, pb;\n\t"
          "setp.ge.and.s32 pv, nb, 0, pa;\n\t"
          "shr.u32 d, %0, 20;\n\t"
          "max.s32 val, %3, nb;\n\t"
          "add.s32 val, val, d;\n\t"
          "add.s32 a, a, %7;\n\t"
          "@pv st.shared.b32 [a], val;\n\t"
          "@pv mov.b32 %3, val;\n\t"
          "@pv mov.b32 %0, %1;\n\t"
          "@pv add.s32 %4, %4, %7;\n\t"
          "@pv add.s32 %2, %2, -1;\n\t"
          "add.s32 t, %4, %7;\n\t"
          "@pv ld.shared.b32 %1, [t];\n\t"
          "}\n"
          : "+r"(w), "+r"(wn), "+r"(left), "+r"(prevval), "+r"(a_w)
          : "r"(a_nb), "r"(bflag), "r"(stepb)
          : "memory");
    }
  };
  // L2S_LEVELS_PER_TRIP levels between two votes.  The level is straight-line predicated code, so a warp that enters
  // converged stays converged and executes the levels in lock step (shared-memory stores of one level are visible to
  // the loads of the next: same warp, program order); the warp is re-converged explicitly once per trip.
  budget = (L.n + 2 * L2S_LEVELS_PER_TRIP - 1) / L2S_LEVELS_PER_TRIP;
#ifdef L2S_TIMELINE
  const int budget0 = budget;
#endif
  while (true) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < L2S_LEVELS_PER_TRIP; ++k) level();
    if (!__any_sync(kFull, left > 0)) break;
    if (--budget < 0) { last = prevval; return false; }
  }
#ifdef L2S_TIMELINE
  if ((threadIdx.x & 31) == 0) { atomicAdd(&g_dbg_levels[0], (unsigned long long)L2S_LEVELS_PER_TRIP * (unsigned long long)(budget0 - budget + 1)); atomicAdd(&g_dbg_levels[1], 1ull); }
#endif
  last = prevval;
  return true;
}

// LATENCY variant of the wavefront (same arguments, same results).  When an SM holds only a few large instances
// (batch of the order of the SM count), the sweep warps are alone on their schedulers and a level costs its dependent
// chain, not its instruction count (measured at 100x20, 128 instances: 58 -> 41 ns per level, evaluation 7.8 -> 5.6 us).
// With many resident instances the schedulers are saturated by sweep instructions instead and the leaner level of
// `wavefront` (21 instead of 23 instructions) wins (100x20, 1024 instances: 37.4 vs 39.2 us), so the CTA kernel picks
// per launch (StepArgs::fast_sweep).  Here the level keeps its dependent chain minimal:
//   load neighbour value -> compare / add / max -> store own value, advance.
// Everything else is decoded ONE OPERATION AHEAD, off the chain, from the prefetched next walk word: the address of
// the value to wait for (A), of the own slot (As) and the duration (D).  Two reserved slots per direction replace the
// flag tests: a job-first (forward) / job-last (backward) operation waits on the ZERO slot (S resp. T, value 0 while
// the sweep runs) and a lane that has placed its whole row waits on the DEAD slot (T resp. S, value -1 = pending for
// ever), so "ready" is always just nb >= 0.  The slots are written on entry (and S/T get their real values after the
// sweeps, as before).
__device__ __forceinline__ bool wavefront_fast(const Layout& L, const Inst& s, int m, int dir, bool active, int& last,
                                          int half = -1, uint32_t walk = 0u, int start = 0, int prev0 = 0) {
  const int n_j = L.n_j;
  if (half < 0) half = dir;
  if (walk == 0u) walk = s.walk;
  const uint32_t base = s.tq + (half ? 4u * L.n1p : 0u);               // &val[0] of the lane's half
  const uint32_t zero_a = base + (dir ? 4u * (uint32_t)(L.n + 1) : 0u);
  const uint32_t dead_a = base + (dir ? 0u : 4u * (uint32_t)(L.n + 1));
  const int nboff = dir ? 4 : -4;                                      // job neighbour: v+1 backward, v-1 forward
  const unsigned bflag = dir ? kWLast : kWFirst;
  const int stepb = dir ? -4 : 4;                                      // walk step in bytes
  if (m == 0) { sts32(zero_a, 0); sts32(dead_a, -1); }
  // start / prev0: resume the lane's row behind its first `start` operations in sweep order (forward: from the head
  // of the row, backward: from its end) with the value prev0 of the last operation that is kept (incremental
  // re-evaluation: the operations before `start` keep their values)
  uint32_t a_w = walk + 4u * (uint32_t)(m * n_j + (dir ? n_j - 1 - start : start));
  int left = active ? n_j - start : 0;                                 // operations still to place
  unsigned w = 0, wn = 0;
  if (left > 0) w = (unsigned)lds32(a_w);
  if (left > 1) wn = (unsigned)lds32(a_w + stepb);
  a_w += 2 * stepb;                                                    // the word that follows wn
  uint32_t As = base + 4u * (w & kWId);
  uint32_t A = (w & bflag) ? zero_a : As + nboff;
  int D = (int)(w >> kWDurShift);
  if (left <= 0) { A = dead_a; D = 0; }
  int prevD = prev0 + D;                                               // value of the current operation if its machine predecessor decides
  __syncwarp();
  int budget = (L.n + 2 * L2S_LEVELS_PER_TRIP - 1) / L2S_LEVELS_PER_TRIP;   // a DAG needs at most n levels
  // One level: place the lane's current operation if the value it waits for is final.  Straight-line predicated code
  // (23 instructions): a warp that enters converged stays converged, so the shared-memory stores of one level are
  // visible to the loads of the next (same warp, program order).  Equivalent C:
  //   nb = *A;  [decode wn -> An, Asn, Dn; An = dead, Dn = 0 if wn is not an operation of this row]
  //   if (nb >= 0) { val = max(prevD, nb + D); *As = val; prevD = val + Dn; A = An; As = Asn; D = Dn; --left;
  //                  if (operations remain behind wn) wn = *a_w; a_w += stepb; }
  auto level = [&]() {
    asm volatile(
        "{\n\t"
        ".reg .pred pv, pf, pl, pq;\n\t"
        ".reg .b32 nb, t, val, idn, asn, an, dn, tf;\n\t"
        "ld.shared.b32 nb, [%0];\n\t"
        "and.b32 idn, %5, 0xffff;\n\t"
        "mad.lo.u32 asn, idn, 4, %7;\n\t"
        "and.b32 tf, %5, %11;\n\t"
        "setp.ne.u32 pf, tf, 0;\n\t"
        "add.s32 an, asn, %10;\n\t"
        "selp.b32 an, %8, an, pf;\n\t"
        "setp.le.s32 pl, %4, 1;\n\t"
        "selp.b32 an, %9, an, pl;\n\t"
        "shr.u32 dn, %5, 20;\n\t"
        "selp.b32 dn, 0, dn, pl;\n\t"
        "setp.ge.s32 pv, nb, 0;\n\t"
        "add.s32 t, nb, %2;\n\t"
        "max.s32 val, %3, t;\n\t"
        "@pv st.shared.b32 [%1], val;\n\t"
        "setp.gt.and.s32 pq, %4, 2, pv;\n\t"
        "@pv add.s32 %3, val, dn;\n\t"
        "@pv mov.b32 %0, an;\n\t"
        "@pv mov.b32 %1, asn;\n\t"
        "@pv mov.b32 %2, dn;\n\t"
        "@pv add.s32 %4, %4, -1;\n\t"
        "@pq ld.shared.b32 %5, [%6];\n\t"
        "@pv add.s32 %6, %6, %12;\n\t"
        "}\n"
        : "+r"(A), "+r"(As), "+r"(D), "+r"(prevD), "+r"(left), "+r"(wn), "+r"(a_w)
        : "r"(base), "r"(zero_a), "r"(dead_a), "r"(nboff), "r"(bflag), "r"(stepb)
        : "memory");
  };
#ifdef L2S_TIMELINE
  const int budget0 = budget;
#endif
  while (true) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < L2S_LEVELS_PER_TRIP; ++k) level();
    if (!__any_sync(kFull, left > 0)) break;
    if (--budget < 0) { last = prevD; return false; }
  }
#ifdef L2S_TIMELINE
  if ((threadIdx.x & 31) == 0) { atomicAdd(&g_dbg_levels[0], (unsigned long long)L2S_LEVELS_PER_TRIP * (unsigned long long)(budget0 - budget + 1)); atomicAdd(&g_dbg_levels[1], 1ull); }
#endif
  last = prevD;     // a finished lane carries the value of its last operation (Dn = 0 behind the row)
  return true;
}

// LOOK-AHEAD variant of the latency sweep: every lane may place TWO consecutive operations of its row per level.  The
// value the second one waits for is loaded at the START of the level together with the first one's (conservative: it
// is placed only if that value was already final then), so the number of levels falls by 40-45 % (simulated and
// measured: 100x20 132 -> 74-77, 150x25 197 -> 111, 20x15 37 -> 26) while the dependent chain of a level only grows by
// one add / max: both values are computed from the two loads, and which operations become "current" and "next"
// afterwards is a pair of selects on the two predicates.
// State per lane: the current operation (A, As, D), the next one (An, Asn, Dn), both decoded, and the raw walk words of
// the two operations after them (w3, w4), decoded at the start of every level in the shadow of the loads; zero / dead
// slots as in wavefront_fast.
__device__ __forceinline__ bool wavefront_fast2(const Layout& L, const Inst& s, int m, int dir, bool active, int& last,
                                                int half = -1, uint32_t walk = 0u, int start = 0, int prev0 = 0) {
  const int n_j = L.n_j;
  if (half < 0) half = dir;
  if (walk == 0u) walk = s.walk;
  const uint32_t base = s.tq + (half ? 4u * L.n1p : 0u);
  const uint32_t zero_a = base + (dir ? 4u * (uint32_t)(L.n + 1) : 0u);
  const uint32_t dead_a = base + (dir ? 0u : 4u * (uint32_t)(L.n + 1));
  const int nboff = dir ? 4 : -4;
  const unsigned bflag = dir ? kWLast : kWFirst;
  const int stepb = dir ? -4 : 4;
  if (m == 0) { sts32(zero_a, 0); sts32(dead_a, -1); }
  uint32_t a_w = walk + 4u * (uint32_t)(m * n_j + (dir ? n_j - 1 - start : start));
  int left = active ? n_j - start : 0;
  unsigned w1 = 0, w2 = 0, w3 = 0, w4 = 0;
  if (left > 0) w1 = (unsigned)lds32(a_w);
  if (left > 1) w2 = (unsigned)lds32(a_w + stepb);
  if (left > 2) w3 = (unsigned)lds32(a_w + 2 * stepb);
  if (left > 3) w4 = (unsigned)lds32(a_w + 3 * stepb);
  a_w += 4 * stepb;                                                    // the word that follows w4
  if (left <= 0) a_w = walk;                                           // idle lanes: the unconditional prefetches of a level stay in bounds
  uint32_t As = base + 4u * (w1 & kWId), Asn = base + 4u * (w2 & kWId);
  uint32_t A = (w1 & bflag) ? zero_a : As + nboff, An = (w2 & bflag) ? zero_a : Asn + nboff;
  int D = (int)(w1 >> kWDurShift), Dn = (int)(w2 >> kWDurShift);
  if (left <= 0) { A = dead_a; D = 0; }
  if (left <= 1) { An = dead_a; Dn = 0; }
  int prevD = prev0 + D;
  __syncwarp();
  int budget = (L.n + L2S_LEVELS_PER_TRIP) / (L2S_LEVELS_PER_TRIP / 2);      // a DAG needs at most n levels; n/2 + 2 trips of 2
  // (an active lane prefetches at most 5 words past its row: the words of the next row, or of the arrays that follow /
  // precede the walk words in the instance's own shared memory -- never used)
  // Equivalent C of one level (all predicated, no branch):
  //   nb1 = *A; nb2 = *An; l0 = *a_w; l1 = *(a_w + stepb);
  //   op3 = decode(w3) (dead, duration 0 unless left > 2); op4 = decode(w4) (... unless left > 3)
  //   p1 = nb1 >= 0; v1 = max(prevD, nb1 + D);        if (p1) *As = v1;
  //   p2 = p1 && nb2 >= 0; v2 = max(v1, nb2) + Dn;     if (p2) *Asn = v2;
  //   p2: cur = op3, next = op4, prevD = v2 + D3, left -= 2, (w3, w4) = (l0, l1), a_w += 2 steps
  //   p1: cur = next, next = op3, prevD = v1 + Dn, left -= 1, (w3, w4) = (w4, l0), a_w += 1 step
  auto level = [&]() {
    asm volatile(
        "{\n\t"
        ".reg .pred p1, p2, pf, pl;\n\t"
        ".reg .b32 nb1, nb2, l0, l1, t, v1, v2, id, as3, a3, d3, as4, a4, d4, tf, x;\n\t"
        "ld.shared.b32 nb1, [%0];\n\t"
        "ld.shared.b32 nb2, [%3];\n\t"
        "ld.shared.b32 l0, [%10];\n\t"
        "add.s32 x, %10, %16;\n\t"
        "ld.shared.b32 l1, [x];\n\t"
        // decode w3 -> (a3, as3, d3), valid iff left > 2
        "and.b32 id, %8, 0xffff;\n\t"
        "mad.lo.u32 as3, id, 4, %11;\n\t"
        "and.b32 tf, %8, %15;\n\t"
        "setp.ne.u32 pf, tf, 0;\n\t"
        "add.s32 a3, as3, %14;\n\t"
        "selp.b32 a3, %12, a3, pf;\n\t"
        "setp.le.s32 pl, %7, 2;\n\t"
        "selp.b32 a3, %13, a3, pl;\n\t"
        "shr.u32 d3, %8, 20;\n\t"
        "selp.b32 d3, 0, d3, pl;\n\t"
        // decode w4 -> (a4, as4, d4), valid iff left > 3
        "and.b32 id, %9, 0xffff;\n\t"
        "mad.lo.u32 as4, id, 4, %11;\n\t"
        "and.b32 tf, %9, %15;\n\t"
        "setp.ne.u32 pf, tf, 0;\n\t"
        "add.s32 a4, as4, %14;\n\t"
        "selp.b32 a4, %12, a4, pf;\n\t"
        "setp.le.s32 pl, %7, 3;\n\t"
        "selp.b32 a4, %13, a4, pl;\n\t"
        "shr.u32 d4, %9, 20;\n\t"
        "selp.b32 d4, 0, d4, pl;\n\t"
        // the two placements
        "setp.ge.s32 p1, nb1, 0;\n\t"
        "add.s32 t, nb1, %2;\n\t"
        "max.s32 v1, %6, t;\n\t"
        "@p1 st.shared.b32 [%1], v1;\n\t"
        "setp.ge.and.s32 p2, nb2, 0, p1;\n\t"
        "max.s32 v2, v1, nb2;\n\t"
        "add.s32 v2, v2, %5;\n\t"
        "@p2 st.shared.b32 [%4], v2;\n\t"
        // advance by one (p1) or two (p2) operations
        "@p1 add.s32 %6, v1, %5;\n\t"
        "@p2 add.s32 %6, v2, d3;\n\t"
        "@p1 mov.b32 %0, %3;\n\t"
        "@p1 mov.b32 %1, %4;\n\t"
        "@p1 mov.b32 %2, %5;\n\t"
        "@p1 mov.b32 %3, a3;\n\t"
        "@p1 mov.b32 %4, as3;\n\t"
        "@p1 mov.b32 %5, d3;\n\t"
        "@p2 mov.b32 %0, a3;\n\t"
        "@p2 mov.b32 %1, as3;\n\t"
        "@p2 mov.b32 %2, d3;\n\t"
        "@p2 mov.b32 %3, a4;\n\t"
        "@p2 mov.b32 %4, as4;\n\t"
        "@p2 mov.b32 %5, d4;\n\t"
        "@p1 add.s32 %7, %7, -1;\n\t"
        "@p2 add.s32 %7, %7, -1;\n\t"
        "@p1 mov.b32 %8, %9;\n\t"
        "@p1 mov.b32 %9, l0;\n\t"
        "@p2 mov.b32 %8, l0;\n\t"
        "@p2 mov.b32 %9, l1;\n\t"
        "@p1 add.s32 %10, %10, %16;\n\t"
        "@p2 add.s32 %10, %10, %16;\n\t"
        "}\n"
        : "+r"(A), "+r"(As), "+r"(D), "+r"(An), "+r"(Asn), "+r"(Dn), "+r"(prevD), "+r"(left), "+r"(w3), "+r"(w4), "+r"(a_w)
        : "r"(base), "r"(zero_a), "r"(dead_a), "r"(nboff), "r"(bflag), "r"(stepb)
        : "memory");
  };
#ifdef L2S_TIMELINE
  const int budget0 = budget;
#endif
  while (true) {
    __syncwarp();
#pragma unroll
    for (int k = 0; k < L2S_LEVELS_PER_TRIP / 2; ++k) level();
    if (!__any_sync(kFull, left > 0)) break;
    if (--budget < 0) { last = prevD; return false; }
  }
#ifdef L2S_TIMELINE
  if ((threadIdx.x & 31) == 0) { atomicAdd(&g_dbg_levels[0], (unsigned long long)(L2S_LEVELS_PER_TRIP / 2) * (unsigned long long)(budget0 - budget + 1)); atomicAdd(&g_dbg_levels[1], 1ull); }
#endif
  last = prevD;
  return true;
}

// Forward (+ backward) evaluation by one warp.  Returns makespan, or -1 for a cyclic selection.
// need_q: also compute the tails (latest start times).
__device__ __forceinline__ int evaluate_instance(const Layout& L, const Inst& s, int lane, bool need_q) {
  int last = 0;
  bool ok;
  int fwd_last = 0;
  if (L.n_m <= 16 && need_q) {
    const int m = lane & 15, dir = lane >> 4;
    ok = wavefront<true>(L, s, m, dir, m < L.n_m, last);
    fwd_last = dir ? 0 : last;
  } else {
    ok = wavefront(L, s, lane, 0, lane < L.n_m, last);
    fwd_last = last;
    if (need_q && ok) {
      int l2;
      ok = wavefront(L, s, lane, 1, lane < L.n_m, l2);
    }
  }
  const int ms = __reduce_max_sync(kFull, fwd_last);
  return ok ? ms : -1;
}

// Incremental forward re-evaluation after the adjacent pair at walk positions i0, i0+1 was swapped, for callers that
// kept the completion times of the previous evaluation in tq's forward half (the fused K-step kernel).
// Every operation whose value can change is a descendant of the pair, so its earliest start -- before and after the
// move -- is >= t0 = completion of the pair's machine predecessor (0 at a row start).  Operations with an old
// earliest start < t0 therefore keep their values (the set is closed under predecessors); along a machine row
// earliest starts increase, so these are a prefix of every row: each lane finds its prefix by binary search, the
// other completion times go back to "pending", and the sweep resumes every row behind its prefix.
__device__ __forceinline__ int evaluate_forward_from(const Layout& L, const Inst& s, int lane, int i0) {
  const int n_j = L.n_j;
  const uint32_t fin = s.tq;
  int t0 = 0;
  {
    const unsigned wp = i0 > 0 ? (unsigned)lds32(s.walk + 4u * (i0 - 1)) : kWRowEnd;
    if (!(wp & kWRowEnd)) t0 = lds32(fin + 4u * (wp & kWId));
  }
  const bool active = lane < L.n_m;
  const uint32_t row = s.walk + 4u * (uint32_t)((active ? lane : 0) * n_j);
  int lo = 0, hi = n_j;
  for (int span = n_j; span > 0; span >>= 1) {   // uniform trip count >= ceil(log2(n_j + 1))
    if (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const unsigned w = (unsigned)lds32(row + 4u * mid);
      const int est = lds32(fin + 4u * (w & kWId)) - (int)(w >> kWDurShift);
      if (est < t0) lo = mid + 1; else hi = mid;
    }
  }
  const int start = lo;
  int prev0 = 0;
  if (start > 0) prev0 = lds32(fin + 4u * ((unsigned)lds32(row + 4u * (start - 1)) & kWId));
  __syncwarp();
  for (int i = lane; i < L.n; i += 32) {
    const unsigned w = (unsigned)lds32(s.walk + 4u * i);
    const uint32_t af = fin + 4u * (w & kWId);
    if (lds32(af) - (int)(w >> kWDurShift) >= t0) sts32(af, -1);
  }
  __syncwarp();
  int last = 0;
  const bool ok = wavefront(L, s, lane, 0, active, last, -1, 0u, start, prev0);
  const int ms = __reduce_max_sync(kFull, last);
  return ok ? ms : -1;
}

// ---- incremental re-evaluation behind the step API (CTA kernel) --------------------------------------------------
// tq holds the completion times fin and tails q of the evaluation BEFORE the swap of the adjacent pair a0 -> a1 (now
// a1 at walk position i0, a0 at i0+1).  Only descendants of the pair can change their earliest start, only ancestors
// their tail:
//   forward : every descendant has an old earliest start >= thr_f = old est(a0); operations below that threshold keep
//             fin (the set is closed under predecessors) and form a PREFIX of every machine row (est grows along a row);
//   backward: every ancestor has an old tail-after-completion q - dur >= thr_b = old q(a1) - dur(a1); operations below
//             keep q and form a SUFFIX of every row.
// resume_point: binary search of the lane's row (uniform trip count) for the number of kept operations in sweep order
// and the value of the last kept one.  Must run BEFORE invalidate_from resets the others to "pending".
__device__ __forceinline__ void resume_point(const Layout& L, const Inst& s, int m, int dir, bool active, int thr,
                                             int& start, int& prev0) {
  const int n_j = L.n_j;
  const uint32_t val = s.tq + (dir ? 4u * L.n1p : 0u);
  const uint32_t row = s.walk + 4u * (uint32_t)((active ? m : 0) * n_j);
  int lo = 0, hi = active ? n_j : 0;
  for (int span = n_j; span > 0; span >>= 1) {
    if (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const unsigned w = (unsigned)lds32(row + 4u * (dir ? n_j - 1 - mid : mid));
      const int head = lds32(val + 4u * (w & kWId)) - (int)(w >> kWDurShift);
      if (head < thr) lo = mid + 1; else hi = mid;
    }
  }
  start = lo;
  prev0 = 0;
  if (lo > 0) prev0 = lds32(val + 4u * ((unsigned)lds32(row + 4u * (dir ? n_j - lo : lo - 1)) & kWId));
}
// every value at or above its threshold goes back to "pending" (all threads of the CTA)
__device__ __forceinline__ void invalidate_from(const Layout& L, const Inst& s, int thr_f, int thr_b, int tid, int nthr) {
  const uint32_t fin = s.tq, q = s.tq + 4u * L.n1p;
  for (int i = tid; i < L.n; i += nthr) {
    const unsigned w = (unsigned)lds32(s.walk + 4u * i);
    const uint32_t o = 4u * (w & kWId);
    const int d = (int)(w >> kWDurShift);
    if (lds32(fin + o) - d >= thr_f) sts32(fin + o, -1);
    if (lds32(q + o) - d >= thr_b) sts32(q + o, -1);
  }
}

// Data-parallel pass over the walk words after an evaluation: for every operation record
//   crit[v] = its critical predecessor = the FIRST predecessor in networkx's G.pred[v] insertion order that
//             attains est[v] (job predecessor listed first iff the jobfirst bit is set or jp < mp; S, = 0, only
//             if there is no other predecessor) -- this is what nx.dag_longest_path keeps per node.
// kMsu: also scatter every node's machine successor into the node-indexed uint16 array at `msu` (which lives in
// the backward half of tq, so the pass must run after the latest start times have been consumed).
template <bool kMsu>
__device__ __forceinline__ void link_pass(const Layout& L, const Inst& s, uint32_t msu, int lane, int nthr = 32) {
  L2S_LOOP
  for (int i = lane; i < L.n; i += nthr) {
    const unsigned w = (unsigned)lds32(s.walk + 4u * i);
    const unsigned wp = i > 0 ? (unsigned)lds32(s.walk + 4u * (i - 1)) : kWRowEnd;
    const unsigned v = w & kWId;
    const unsigned mp = (wp & kWRowEnd) ? 0u : (wp & kWId);       // row start <=> previous word ends a row
    const unsigned jp = v - 1u;
    const int fj = (w & kWFirst) ? -1 : lds32(s.tq + 4u * jp);     // completion of the job predecessor
    const int fm = mp ? lds32(s.tq + 4u * mp) : -1;
    // job predecessor wins iff it exists and (fj, tie) > (fm, 0) lexicographically, tie = listed first in G.pred[v]
    // (fm = -1 when there is no machine predecessor); one integer comparison of 2*f + tie (f < 2^24)
    const int tie = (((w & kWJf) != 0u) | (jp < mp)) ? 1 : 0;
    const bool take_jp = (fj >= 0) & (2 * fj + tie > 2 * fm);
    sts16(s.crit + 2u * v, (int)(take_jp ? jp : mp));
    if (kMsu) {   // msu[v] = machine successor (0 = none), node-indexed, for the ascending-source edge list
      const unsigned ms = (w & kWRowEnd) ? 0u : ((unsigned)lds32(s.walk + 4u * (i + 1)) & kWId);
      sts16(msu + 2u * v, (int)ms);
    }
  }
}

// first operation of its job <=> (v-1) mod n_m == 0  (multiply-shift, exact for v < 2^16)
__device__ __forceinline__ bool is_job_first(const Layout& L, int v) {
  const unsigned x = (unsigned)(v - 1);
  unsigned qd = __umulhi(x, L.magic_nm);
  qd -= (qd * (unsigned)L.n_m > x) ? 1u : 0u;
  return (x == qd * (unsigned)L.n_m) | (L.n_m == 1);   // (the 32-bit magic number does not exist for n_m = 1)
}

// Swap the adjacent machine pair a0 -> a1 (JsspN5.change_nxgraph_topology, env/environment.py:318-354) in the
// shared-memory walk array and set the tie-break history bits of a0, a1 and a1's old machine successor.
// i0 = where[a0] (caller supplies it).  Returns 0, 1 (dummy) or -4 (not an adjacent pair on one machine); on
// error nothing is modified.  Uniform across the warp; lane 0 writes.
__device__ __forceinline__ int apply_swap(const Layout& L, const Inst& s, int a0, int a1, int i0, int lane,
                                          bool& has_t) {
  if (a0 == 0 && a1 == 0) return 1;
  if (a0 < 1 || a0 > L.n || a1 < 1 || a1 > L.n || a0 == a1 || i0 < 0 || i0 >= L.n) return -4;
  const unsigned w0 = (unsigned)lds32(s.walk + 4u * i0);
  if ((int)(w0 & kWId) != a0 || (w0 & kWRowEnd)) return -4;
  const unsigned w1 = (unsigned)lds32(s.walk + 4u * (i0 + 1));
  if ((int)(w1 & kWId) != a1) return -4;
  has_t = !(w1 & kWRowEnd);
  __syncwarp();
  if (lane == 0) {
    sts32(s.walk + 4u * i0, (int)((w1 & ~kWRowEnd) | kWJf));
    sts32(s.walk + 4u * (i0 + 1), (int)((w0 & ~kWRowEnd) | kWJf | (w1 & kWRowEnd)));
    if (has_t) sts32(s.walk + 4u * (i0 + 2), (int)((unsigned)lds32(s.walk + 4u * (i0 + 2)) | kWJf));
  }
  __syncwarp();
  return 0;
}

// Critical path == nx.dag_longest_path(G)[1:-1] (env/environment.py:77) by chasing the recorded critical
// predecessors from T.  Written REVERSED (rp[0] = op next to T) as uint16 at shared address `rp`.
__device__ __forceinline__ int trace_path(const Layout& L, const Inst& s, int ms, int lane, uint32_t rp) {
  // T lists the job-last operations by ascending job: the first one that attains the makespan wins
  int start = 0;
  for (int j0 = 0; j0 < L.n_j; j0 += 32) {
    const int j = j0 + lane;
    const unsigned hit = __ballot_sync(kFull, j < L.n_j && lds32(s.tq + 4u * (j * L.n_m + L.n_m)) == ms);
    if (hit) { start = (j0 + __ffs(hit) - 1) * L.n_m + L.n_m; break; }
  }
  int len = 0;
  int v = start;
  while (v != 0 && len < L.n) {   // uniform across the warp (broadcast loads)
    if (lane == 0) sts16(rp + 2u * len, v);
    ++len;
    v = lds16(s.crit + 2u * v);
  }
  __syncwarp();
  return len;
}

// The same critical path for the CTA kernel, where the chase from T is the longest serial chain of a large instance
// (113 / 161 hops at 100x20 / 150x25): all threads first square the critical-predecessor map twice (jump2 = crit o crit,
// jump4 = jump2 o jump2; crit[S] = crit[T] = 0, so S is a fixed point), then ONE warp follows jump4 -- a quarter of the
// dependent loads -- dropping an anchor every fourth path position, and fills the three operations behind every anchor
// in parallel (crit[a], jump2[a], crit[jump2[a]]).
__device__ __forceinline__ int path_start(const Layout& L, const Inst& s, int ms, int lane) {
  // T lists the job-last operations by ascending job: the first one that attains the makespan wins
  for (int j0 = 0; j0 < L.n_j; j0 += 32) {
    const int j = j0 + lane;
    const unsigned hit = __ballot_sync(kFull, j < L.n_j && lds32(s.tq + 4u * (j * L.n_m + L.n_m)) == ms);
    if (hit) return (j0 + __ffs(hit) - 1) * L.n_m + L.n_m;
  }
  return 0;
}
__device__ __forceinline__ void square_map(uint32_t dst, uint32_t src, int n1, int tid, int nthr) {
  for (int v = tid; v < n1; v += nthr) sts16(dst + 2u * v, lds16(src + 2u * lds16(src + 2u * v)));
}
__device__ __forceinline__ int trace_path_jump(const Layout& L, const Inst& s, int start, int lane, uint32_t rp,
                                               uint32_t jump2, uint32_t jump4) {
  int v = start, na = 0;
  const int cap = (L.n + 3) / 4 + 1;
  while (v != 0 && na < cap) {     // uniform across the warp (broadcast loads)
    if (lane == 0) sts16(rp + 8u * na, v);
    ++na;
    v = lds16(jump4 + 2u * v);
  }
  __syncwarp();
  int len = 0;
  for (int a0 = 0; a0 < na; a0 += 32) {
    const int a = a0 + lane;
    if (a < na) {
      const int A = lds16(rp + 8u * a);
      const int c1 = lds16(s.crit + 2u * A), c2 = lds16(jump2 + 2u * A);
      const int c3 = lds16(s.crit + 2u * c2);
      if (c1) sts16(rp + 8u * a + 2u, c1);
      if (c2) sts16(rp + 8u * a + 4u, c2);
      if (c3) sts16(rp + 8u * a + 6u, c3);
      if (a == na - 1) len = 4 * a + 1 + (c1 != 0) + (c2 != 0) + (c3 != 0);
    }
  }
  len = __reduce_max_sync(kFull, len);
  __syncwarp();
  return len;
}

// N5 block-boundary pairs in path order with the tabu filter: JsspN5._get_pairs (env/environment.py:84-105).
// Two consecutive path operations are on the same machine iff the hop between them is not a job arc.  With
// s_i = "path operations i and i+1 share a machine" (0 <= i < rg = len-1) and the sentinels s_(-1) = s_rg = 1 the
// reference's case analysis collapses to   take_i = s_i and not (s_(i-1) and s_(i+1))   (first pair of every block
// except a >= 3-long first block, last pair of every block except the last block, each pair once).  Pass 1 writes the
// s bits as ballot words to the scratch `bits` (in the dead critical-predecessor array), pass 2 is branch-free bit
// logic plus the tabu filter.  emit(idx, a0, a1) is called by the owning lane for pair number idx; returns the pair
// count, or a negative status (-9: the reference's IndexError corner, a two-operation path on one machine).
template <typename Emit>
__device__ __forceinline__ int n5_moves(const Layout& L, const Inst& s, uint32_t rp, int len, int tabu0, int tabu1,
                                        int lane, Emit emit) {
  const int rg = len - 1;
  if (rg <= 0) return 0;
  const uint32_t bits = s.bits;
  const uint32_t last = rp + 2u * (uint32_t)(len - 1);          // path position i lives at last - 2*i (rp is reversed)
  const int kmax = rg >> 5;                                      // the word that holds the sentinel bit rg
  for (int k = 0; k <= kmax; ++k) {
    const int i = 32 * k + lane;
    bool same = i == rg;
    if (i < rg) {
      const int u = lds16(last - 2u * i), w = lds16(last - 2u * i - 2u);
      same = !(u == w - 1 && !is_job_first(L, w));
    }
    const unsigned bal = __ballot_sync(kFull, same);
    if (lane == 0) sts32(bits + 4u * k, (int)bal);
  }
  __syncwarp();
  const unsigned below = (1u << lane) - 1u;
  int count = 0;
  unsigned prevw = 0xffffffffu, curw = (unsigned)lds32(bits);
  const int err = (len == 2 && (curw & 1u)) ? -9 : 0;
  for (int k = 0; k <= kmax; ++k) {
    const unsigned nextw = k < kmax ? (unsigned)lds32(bits + 4u * (k + 1)) : 0u;
    const unsigned prev = (curw << 1) | (prevw >> 31), next = (curw >> 1) | (nextw << 31);
    const int i = 32 * k + lane;
    bool take = i < rg && (((curw & ~(prev & next)) >> lane) & 1u);
    int u = 0, w = 0;
    if (take) {
      u = lds16(last - 2u * i);
      w = lds16(last - 2u * i - 2u);
      take = !(u == tabu0 && w == tabu1);
    }
    const unsigned bal = __ballot_sync(kFull, take);
    if (take) emit(count + __popc(bal & below), u, w);
    count += __popc(bal);
    prevw = curw;
    curw = nextw;
  }
  return err < 0 ? err : count;
}

// Counter-based RNG of the RANDOM policy (mirrored in oracle/l2s_oracle.py:mix64).
__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t inst, uint64_t step) {
  uint64_t z = seed ^ (inst * 0x9E3779B97F4A7C15ull) ^ (step * 0xD1B54A32D192ED03ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ int random_pick(uint64_t seed, uint64_t inst, uint64_t step, int count) {
  return (int)(((mix64(seed, inst, step) >> 32) * (uint64_t)count) >> 32);
}

}  // namespace l2s
