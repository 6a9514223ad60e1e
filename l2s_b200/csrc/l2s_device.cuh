// l2s_b200 device-side building blocks (sm_100a).  Integer graph work: no tensor cores by design.
//
// Execution model: ONE WARP OWNS ONE INSTANCE.  The instance's whole search state lives in that warp's
// slice of shared memory, so no block-level barrier is ever needed; every synchronisation is a
// __syncwarp / ballot.  A CTA is just a bundle of independent warps (so that > 32 instances fit one SM).
//
// Longest paths are computed WORK-EFFICIENTLY (O(n) node updates instead of the reference's O(n * depth)
// Jacobi sweeps, env/message_passing_evl.py:176-197): every lane "walks" one machine's processing
// sequence.  Its machine predecessor's value is in a register; only the job neighbour is read from
// shared memory (sentinel -1 = not final yet).  One operation per lane per level => the number of
// iterations equals the hop depth of the graph (27..210 at the reference's sizes) and each iteration
// costs ~20 warp instructions regardless of n.  With n_m <= 16 the forward sweep (heads, lanes 0-15)
// and the backward sweep (tails q, lanes 16-31) run in the SAME iterations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2s {

constexpr int kFirstFlag = 0x8000;  // node word: operation is the first of its job (job predecessor = S)
constexpr int kLastFlag = 0x4000;   //            operation is the last of its job (job successor = T)
constexpr int kDurMask = 0x3FFF;    //            duration (<= 16383)
constexpr unsigned kFull = 0xffffffffu;

// Shapes and shared-memory layout of one instance (all offsets in bytes, 16-byte aligned).
struct Layout {
  int n_j, n_m, n, n1;   // n = n_j*n_m operations, n1 = n+2 nodes
  int n1p;               // n1 rounded up to 4 (int32 elements per tq half)
  int seq_stride;        // uint16 elements per instance in HBM/smem machine-order array (multiple of 8)
  int node_stride;       // uint16 elements per instance of node-indexed arrays (multiple of 8, >= n1)
  int mch_stride;        // bytes per instance of the machine-id array (multiple of 16)
  int jf_words;          // uint32 words of tie-break history bits (multiple of 4)
  int e_mc;              // machine edges per instance n_m*(n_j-1)
  int off_tq, off_seq, off_durw, off_crit, off_mch, off_jf, off_misc;
  int bytes;             // per instance
};

constexpr int kMiscInts = 8;

// Persistent per-slice search state in HBM (SoA, one row per instance).
struct State {
  uint16_t* durw;   // [B][node_stride] node words (duration | flags), indexed by node id; 0 for S,T
  uint8_t* mch;     // [B][mch_stride]  1-based machine of op v at [v-1]
  uint16_t* seq;    // [B][seq_stride]  machine orders, row m at [m*n_j .. m*n_j+n_j)
  uint32_t* jf;     // [B][jf_words]    bit v: job predecessor listed first (networkx insertion order)
  int32_t* tabu;    // [B][2]
  int32_t* cur;     // [B] current makespan
  int32_t* inc;     // [B] incumbent makespan
  int32_t* err;     // [1] first device-side error: (code<<24)|instance  (0 = none)
  int32_t* est;     // [B][n1] last evaluation (test introspection)   -- may be null
  int32_t* lst;     // [B][n1]
  int32_t* path;    // [B][n] critical path (forward order)            -- may be null
  int32_t* path_len;// [B]
};

struct Inst {            // shared-memory view of one instance
  int32_t* tq;           // [2][n1p]: tq[v] = est[v]+dur[v] ("fin"), tq[n1p+v] = tail q[v]; -1 = pending
  uint16_t* seq;
  uint16_t* durw;
  uint16_t* crit;        // chosen critical predecessor (networkx tie-break applied); 0 = S
  uint8_t* mch;
  uint32_t* jf;
  int32_t* misc;
};

__device__ __forceinline__ Inst inst_view(const Layout& L, unsigned char* base) {
  Inst s;
  s.tq = reinterpret_cast<int32_t*>(base + L.off_tq);
  s.seq = reinterpret_cast<uint16_t*>(base + L.off_seq);
  s.durw = reinterpret_cast<uint16_t*>(base + L.off_durw);
  s.crit = reinterpret_cast<uint16_t*>(base + L.off_crit);
  s.mch = reinterpret_cast<uint8_t*>(base + L.off_mch);
  s.jf = reinterpret_cast<uint32_t*>(base + L.off_jf);
  s.misc = reinterpret_cast<int32_t*>(base + L.off_misc);
  return s;
}

__device__ __forceinline__ void record_error(int32_t* err, int code, int b) {
  // code is negative; keep the first one
  atomicCAS(err, 0, (int)(((unsigned)(-code) << 24) | ((unsigned)b & 0xffffffu)));
}

__device__ __forceinline__ uint4 ld_nc16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Coalesced 16-byte staging of one instance's state into shared memory + sentinel fill of tq.
__device__ __forceinline__ void stage_instance(const Layout& L, const Inst& s, const State& st, int b, int lane,
                                               bool fill_q) {
  const uint4* gd = reinterpret_cast<const uint4*>(st.durw + (size_t)b * L.node_stride);
  const uint4* gs = reinterpret_cast<const uint4*>(st.seq + (size_t)b * L.seq_stride);
  const uint4* gm = reinterpret_cast<const uint4*>(st.mch + (size_t)b * L.mch_stride);
  const uint4* gj = reinterpret_cast<const uint4*>(st.jf + (size_t)b * L.jf_words);
  uint4* sd = reinterpret_cast<uint4*>(s.durw);
  uint4* ss = reinterpret_cast<uint4*>(s.seq);
  uint4* sm = reinterpret_cast<uint4*>(s.mch);
  uint4* sj = reinterpret_cast<uint4*>(s.jf);
  for (int i = lane; i < L.node_stride / 8; i += 32) sd[i] = ld_nc16(gd + i);   // static for the whole search
  for (int i = lane; i < L.mch_stride / 16; i += 32) sm[i] = ld_nc16(gm + i);
  for (int i = lane; i < L.seq_stride / 8; i += 32) ss[i] = gs[i];
  for (int i = lane; i < L.jf_words / 4; i += 32) sj[i] = gj[i];
  const int4 pending = make_int4(-1, -1, -1, -1);
  int4* t4 = reinterpret_cast<int4*>(s.tq);
  const int cnt = (fill_q ? 2 : 1) * (L.n1p / 4);
  for (int i = lane; i < cnt; i += 32) t4[i] = pending;
}

__device__ __forceinline__ void fill_pending(const Layout& L, const Inst& s, int lane, bool fill_q) {
  const int4 pending = make_int4(-1, -1, -1, -1);
  int4* t4 = reinterpret_cast<int4*>(s.tq);
  const int cnt = (fill_q ? 2 : 1) * (L.n1p / 4);
  for (int i = lane; i < cnt; i += 32) t4[i] = pending;
}

// ------------------------------------------------------------------------------------------------------
// Level-synchronous wavefront.  Lane role: machine `m` (row of seq), direction dir (0: forward est/fin
// from the head of the row, 1: backward tail q from the end of the row), `active` lanes only.
// Returns false if the selection is cyclic (no lane can advance).  `last` = final value of this lane's
// chain (forward: completion time of the machine's last operation).
// kCrit: also record the critical predecessor of every operation (forward lanes).
// ------------------------------------------------------------------------------------------------------
template <bool kCrit>
__device__ __forceinline__ bool wavefront(const Layout& L, const Inst& s, int m, int dir, bool active, int& last) {
  volatile int32_t* tq = s.tq;
  const int n_j = L.n_j;
  const int base = dir ? L.n1p : 0;
  const int nbr = dir ? 1 : -1;
  const int bflag = dir ? kLastFlag : kFirstFlag;
  const int row = m * n_j;
  int k = 0, prevval = 0, prevop = 0, cur = 0, dw = 0;
  bool more = active;
  if (more) {
    cur = s.seq[row + (dir ? n_j - 1 : 0)];
    dw = s.durw[cur];
  }
  int budget = L.n + 2;  // a DAG needs at most n levels
  while (true) {
    if (more) {
      int nb = tq[base + cur + nbr];
      nb = (dw & bflag) ? 0 : nb;
      if (nb >= 0) {
        const int val = max(prevval, nb) + (dw & kDurMask);
        if (kCrit && !dir) {
          const int jp = cur - 1;
          int c;
          if (dw & kFirstFlag) c = prevop;                        // machine predecessor, else S (=0)
          else if (prevop == 0 || nb > prevval) c = jp;
          else if (prevval > nb) c = prevop;
          else c = (((s.jf[cur >> 5] >> (cur & 31)) & 1u) || jp < prevop) ? jp : prevop;  // tie: insertion order
          s.crit[cur] = (uint16_t)c;
        }
        tq[base + cur] = val;
        prevval = val;
        prevop = cur;
        ++k;
        more = k < n_j;
        if (more) {
          cur = s.seq[row + (dir ? n_j - 1 - k : k)];
          dw = s.durw[cur];
        }
      }
    }
    __syncwarp();
    if (!__any_sync(kFull, more)) break;
    if (--budget < 0) { last = prevval; return false; }
  }
  last = prevval;
  return true;
}

// Forward+backward evaluation by one warp.  Returns makespan, or -1 for a cyclic selection.
// need_q: also compute the tails (latest start times).
template <bool kCrit>
__device__ __forceinline__ int evaluate_instance(const Layout& L, const Inst& s, int lane, bool need_q) {
  int last = 0;
  bool ok;
  int fwd_last = 0;
  if (L.n_m <= 16 && need_q) {
    const int m = lane & 15, dir = lane >> 4;
    ok = wavefront<kCrit>(L, s, m, dir, m < L.n_m, last);
    fwd_last = dir ? 0 : last;
  } else {
    ok = wavefront<kCrit>(L, s, lane, 0, lane < L.n_m, last);
    fwd_last = last;
    if (need_q && ok) {
      int l2;
      ok = wavefront<false>(L, s, lane, 1, lane < L.n_m, l2);
    }
  }
  const int ms = __reduce_max_sync(kFull, fwd_last);
  return ok ? ms : -1;
}

// Swap the adjacent machine pair a0 -> a1 in shared memory (JsspN5.change_nxgraph_topology,
// env/environment.py:318-354) and record the tie-break history bits.  Returns 0, 1 (dummy) or a negative
// status; on error nothing is modified.  `pos_out` = position of a0 before the swap, `m_out` its machine.
__device__ __forceinline__ int apply_swap(const Layout& L, const Inst& s, int a0, int a1, int lane, int& m_out,
                                          int& pos_out, int& t_out) {
  if (a0 == 0 && a1 == 0) return 1;
  if (a0 < 1 || a0 > L.n || a1 < 1 || a1 > L.n || a0 == a1) return -4;
  const int m = (int)s.mch[a0 - 1] - 1;
  if ((int)s.mch[a1 - 1] - 1 != m || m < 0 || m >= L.n_m) return -4;
  uint16_t* row = s.seq + m * L.n_j;
  int p = -1;
  for (int k0 = 0; k0 < L.n_j; k0 += 32) {
    const int k = k0 + lane;
    const unsigned hit = __ballot_sync(kFull, k < L.n_j && row[k] == a0);
    if (hit) { p = k0 + __ffs(hit) - 1; break; }
  }
  if (p < 0 || p + 1 >= L.n_j || row[p + 1] != a1) return -4;
  const int t = (p + 2 < L.n_j) ? row[p + 2] : 0;
  __syncwarp();
  if (lane == 0) {
    row[p] = (uint16_t)a1;
    row[p + 1] = (uint16_t)a0;
    s.jf[a0 >> 5] |= 1u << (a0 & 31);
    s.jf[a1 >> 5] |= 1u << (a1 & 31);
    if (t) s.jf[t >> 5] |= 1u << (t & 31);
  }
  __syncwarp();
  m_out = m; pos_out = p; t_out = t;
  return 0;
}

// Critical path == nx.dag_longest_path(G)[1:-1] (env/environment.py:77) by chasing the recorded critical
// predecessors from T.  Written REVERSED (rp[0] = op next to T) as uint16 into `rp`.  Returns its length.
__device__ __forceinline__ int trace_path(const Layout& L, const Inst& s, int ms, int lane, uint16_t* rp) {
  // T lists the job-last operations by ascending job: the first one that attains the makespan wins
  int start = 0;
  for (int j0 = 0; j0 < L.n_j; j0 += 32) {
    const int j = j0 + lane;
    const unsigned hit = __ballot_sync(kFull, j < L.n_j && s.tq[j * L.n_m + L.n_m] == ms);
    if (hit) { start = (j0 + __ffs(hit) - 1) * L.n_m + L.n_m; break; }
  }
  int len = 0;
  int v = start;
  while (v != 0 && len < L.n) {   // uniform across the warp (broadcast loads)
    if (lane == 0) rp[len] = (uint16_t)v;
    ++len;
    v = s.crit[v];
  }
  __syncwarp();
  return len;
}

// N5 block-boundary pairs in path order with the tabu filter: JsspN5._get_pairs (env/environment.py:84-105).
// Two consecutive path operations are on the same machine iff the hop between them is not a job arc.
// emit(idx, a0, a1) is called by the owning lane for pair number idx; returns the pair count, or a negative
// status (-9: the reference's IndexError corner).
template <typename Emit>
__device__ __forceinline__ int n5_moves(const Layout& L, const Inst& s, const uint16_t* rp, int len, int tabu0,
                                        int tabu1, int lane, Emit emit) {
  const int rg = len - 1;
  int count = 0;
  int err = 0;
  auto pf = [&](int i) -> int { return rp[len - 1 - i]; };
  auto same = [&](int i) -> bool {   // ops i and i+1 of the path share a machine (0 <= i < rg)
    const int u = pf(i), w = pf(i + 1);
    return !(u == w - 1 && !(s.durw[w] & kFirstFlag));
  };
  for (int i0 = 0; i0 < rg; i0 += 32) {
    const int i = i0 + lane;
    bool take = false;
    int u = 0, w = 0;
    if (i < rg && same(i)) {
      u = pf(i); w = pf(i + 1);
      if (i == 0) {
        if (len < 3) err = -9; else take = !same(1);
      } else if (!same(i - 1)) {
        take = true;
      } else if (i + 1 == rg) {
        take = false;
      } else {
        take = !same(i + 1);
      }
      if (take && u == tabu0 && w == tabu1) take = false;
    }
    const unsigned bal = __ballot_sync(kFull, take);
    if (take) emit(count + __popc(bal & ((1u << lane) - 1u)), u, w);
    count += __popc(bal);
  }
  err = __reduce_min_sync(kFull, err);
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
