/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement ("oracle") of the L2S evaluator + N5 local-search step.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker / timed CPU baseline.  The product (l2s_b200) never links it.
 *
 * Parity status: pinned through oracle/l2s_oracle.py, which is itself pinned against outputs of the
 * unmodified reference (tests/golden/, see oracle/make_golden.py); tests/test_cpu_oracle.py asserts that this
 * C restatement and the numpy restatement agree on every golden trajectory.
 *
 * Each function cites the reference lines it follows (paths relative to the reference root).
 * Node ids: S=0, operations 1..n row-major by (job, position), T=n+1.  mseq[m*n_j+k] = node id of the k-th
 * operation on machine m.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n_j, n_m, n;
  const int32_t* dur;  /* [n]   durations of ops 1..n at [v-1]            */
  const int32_t* mch;  /* [n]   1-based machine of op v at [v-1]          */
  int32_t* mseq;       /* [n_m*n_j] machine orders                         */
  int32_t* mpred;      /* [n+2] machine predecessor / 0                    */
  int32_t* msucc;      /* [n+2]                                            */
  uint8_t* jobfirst;   /* [n+2] tie-break history        */
  int32_t* est;        /* [n+2]                                            */
  int32_t* lst;        /* [n+2]                                            */
  int32_t* order;      /* [n]   scratch: topological order                 */
  int32_t* indeg;      /* [n+2] scratch                                    */
  int32_t* path;       /* [n]   scratch: critical path                     */
  int32_t tabu[2];
  int32_t makespan;
} orc_inst;

static int dur_of(const orc_inst* s, int v) { return (v >= 1 && v <= s->n) ? s->dur[v - 1] : 0; }

/* machine arcs ops[k] -> ops[k+1]: env/environment.py:254-259 */
static void orc_links(orc_inst* s) {
  memset(s->mpred, 0, sizeof(int32_t) * (s->n + 2));
  memset(s->msucc, 0, sizeof(int32_t) * (s->n + 2));
  for (int m = 0; m < s->n_m; ++m)
    for (int k = 0; k + 1 < s->n_j; ++k) {
      int a = s->mseq[m * s->n_j + k], b = s->mseq[m * s->n_j + k + 1];
      s->msucc[a] = b;
      s->mpred[b] = a;
    }
}

/* Evaluator.forward (env/message_passing_evl.py:161-199): the fixed point of its max-aggregation sweeps is
 * the longest-path recurrence, evaluated here once in topological order.  Returns makespan or -1 (cycle). */
static int orc_evaluate(orc_inst* s) {
  const int n = s->n, n_m = s->n_m;
  int head = 0, tail = 0;
  for (int v = 1; v <= n; ++v) {
    s->est[v] = 0;
    s->indeg[v] = ((v - 1) % n_m != 0) + (s->mpred[v] != 0);
    if (s->indeg[v] == 0) s->order[tail++] = v;
  }
  while (head < tail) {
    int v = s->order[head++];
    int fin = s->est[v] + dur_of(s, v);
    int succ[2] = {(v % n_m != 0) ? v + 1 : 0, s->msucc[v]};
    for (int i = 0; i < 2; ++i) {
      int t = succ[i];
      if (!t) continue;
      if (fin > s->est[t]) s->est[t] = fin;
      if (--s->indeg[t] == 0) s->order[tail++] = t;
    }
  }
  if (tail != n) return -1;
  int ms = 0;
  for (int j = 0; j < s->n_j; ++j) {
    int last = j * n_m + n_m, f = s->est[last] + dur_of(s, last);
    if (f > ms) ms = f;
  }
  s->est[0] = 0;
  s->est[n + 1] = ms;
  s->lst[n + 1] = ms;
  for (int i = n - 1; i >= 0; --i) {
    int v = s->order[i];
    int best = (v % n_m != 0) ? s->lst[v + 1] : ms;
    if (s->msucc[v] && s->lst[s->msucc[v]] < best) best = s->lst[s->msucc[v]];
    s->lst[v] = best - dur_of(s, v);
  }
  int l0 = s->lst[1];
  for (int j = 1; j < s->n_j; ++j)
    if (s->lst[j * n_m + 1] < l0) l0 = s->lst[j * n_m + 1];
  s->lst[0] = l0;
  s->makespan = ms;
  return ms;
}

/* nx.dag_longest_path(G)[1:-1] (env/environment.py:77) with networkx's first-max-in-insertion-order rule. */
static int orc_critical_path(orc_inst* s) {
  const int n_m = s->n_m;
  int v = 0, len = 0;
  for (int j = 0; j < s->n_j; ++j) {
    int last = j * n_m + n_m;
    if (s->est[last] + dur_of(s, last) == s->makespan) { v = last; break; }
  }
  while (v) {
    s->path[len++] = v;
    int jp = ((v - 1) % n_m != 0) ? v - 1 : 0, mp = s->mpred[v];
    int c0 = jp, c1 = mp;
    if (jp && mp && !(s->jobfirst[v] || jp < mp)) { c0 = mp; c1 = jp; }
    if (!c0) { c0 = c1; c1 = 0; }
    int best_len = jp ? -1 : 0, best = 0;   /* job-first ops also carry the weight-0 S edge, listed last */
    if (c0 && s->est[c0] + dur_of(s, c0) > best_len) { best_len = s->est[c0] + dur_of(s, c0); best = c0; }
    if (c1 && s->est[c1] + dur_of(s, c1) > best_len) { best_len = s->est[c1] + dur_of(s, c1); best = c1; }
    v = best;
  }
  for (int i = 0; i < len / 2; ++i) { int t = s->path[i]; s->path[i] = s->path[len - 1 - i]; s->path[len - 1 - i] = t; }
  return len;
}

/* JsspN5._get_pairs (env/environment.py:84-105); returns count, -9 for the reference's IndexError corner */
static int orc_pairs(const orc_inst* s, int len, int32_t* out, int pmax) {
  int rg = len - 1, cnt = 0;
  for (int i = 0; i < rg; ++i) {
    int ci = s->mch[s->path[i] - 1], cn = s->mch[s->path[i + 1] - 1];
    if (ci != cn) continue;
    int take;
    if (i == 0) {
      if (len < 3) return -9;
      take = s->mch[s->path[1] - 1] != s->mch[s->path[2] - 1];
    } else if (ci != s->mch[s->path[i - 1] - 1]) take = 1;
    else if (i + 1 == rg) take = 0;
    else take = cn != s->mch[s->path[i + 2] - 1];
    if (take && !(s->path[i] == s->tabu[0] && s->path[i + 1] == s->tabu[1])) {
      if (cnt < pmax) { out[2 * cnt] = s->path[i]; out[2 * cnt + 1] = s->path[i + 1]; }
      ++cnt;
    }
  }
  return cnt;
}

/* change_nxgraph_topology (env/environment.py:318-354) + tabu (:370-380); 0 ok, 1 dummy, -4 illegal */
static int orc_apply(orc_inst* s, int a0, int a1, int set_history) {
  if (a0 == 0 && a1 == 0) return 1;
  if (a0 < 1 || a0 > s->n || a1 < 1 || a1 > s->n || s->msucc[a0] != a1) return -4;
  int p = s->mpred[a0], t = s->msucc[a1];
  if (p) s->msucc[p] = a1;
  s->mpred[a1] = p;
  s->msucc[a1] = a0;
  s->mpred[a0] = a1;
  s->msucc[a0] = t;
  if (t) s->mpred[t] = a0;
  int m = s->mch[a0 - 1] - 1;
  int32_t* row = s->mseq + m * s->n_j;
  for (int k = 0; k + 1 < s->n_j; ++k)
    if (row[k] == a0 && row[k + 1] == a1) { row[k] = a1; row[k + 1] = a0; break; }
  if (set_history) {
    s->jobfirst[a0] = s->jobfirst[a1] = 1;
    if (t) s->jobfirst[t] = 1;
    s->tabu[0] = a1;
    s->tabu[1] = a0;
  }
  return 0;
}

static uint64_t mix64(uint64_t seed, uint64_t inst, uint64_t step) {
  uint64_t z = seed ^ (inst * 0x9E3779B97F4A7C15ull) ^ (step * 0xD1B54A32D192ED03ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static orc_inst* orc_new(int n_j, int n_m, const int32_t* dur, const int32_t* mch, int32_t* mseq) {
  orc_inst* s = (orc_inst*)calloc(1, sizeof(orc_inst));
  int n = n_j * n_m;
  s->n_j = n_j; s->n_m = n_m; s->n = n; s->dur = dur; s->mch = mch; s->mseq = mseq;
  s->mpred = (int32_t*)calloc(n + 2, 4); s->msucc = (int32_t*)calloc(n + 2, 4);
  s->jobfirst = (uint8_t*)calloc(n + 2, 1);
  s->est = (int32_t*)calloc(n + 2, 4); s->lst = (int32_t*)calloc(n + 2, 4);
  s->order = (int32_t*)calloc(n, 4); s->indeg = (int32_t*)calloc(n + 2, 4); s->path = (int32_t*)calloc(n, 4);
  orc_links(s);
  return s;
}
static void orc_free(orc_inst* s) {
  free(s->mpred); free(s->msucc); free(s->jobfirst); free(s->est); free(s->lst); free(s->order); free(s->indeg);
  free(s->path); free(s);
}

/* ---------------------------------------------------------------------------------------------------
 * Batch driver: K steps of every instance with policy 0 REPLAY (tape [B,K,2]) / 1 RANDOM (counter RNG) /
 * 2 GREEDY (conventional_baselines_with_long-term-mem.py:186-197).  mseq [B,n_m,n_j] is updated in place.
 * With write_state != 0 every step also produces what JsspN5.step returns (features x, machine edges, pair
 * list: env/environment.py:291-316,411-423) into the *_out buffers of the LAST step -- this is the work the
 * CPU baseline timing has to include.  Returns 0 or the first negative status.
 * OpenMP parallel over instances (threads = omp default or `threads` if > 0).
 * --------------------------------------------------------------------------------------------------- */
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_run_batch(int B, int n_j, int n_m, const int32_t* dur, const int32_t* mch, int32_t* mseq, int policy, int K,
                  const int32_t* tape, uint64_t seed, int64_t inst_off, int64_t itr0, int pmax, int32_t* taken,
                  const int32_t* log_steps, int n_log, float* inc_log, float* makespan_out, float* incumbent_out,
                  int write_state, float high, float fea, float* x_out, int64_t* emc_out, int32_t* pairs_out,
                  int32_t* npairs_out, int32_t* est_out, int32_t* lst_out, uint8_t* jobfirst_io, int32_t* tabu_io,
                  int threads) {
  const int n = n_j * n_m, n1 = n + 2, e_mc = n_m * (n_j - 1);
  int status = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
  for (int b = 0; b < B; ++b) {
    orc_inst* s = orc_new(n_j, n_m, dur + (size_t)b * n, mch + (size_t)b * n, mseq + (size_t)b * n);
    int32_t* pl = (int32_t*)malloc(sizeof(int32_t) * 2 * (pmax > 0 ? pmax : 1));
    if (jobfirst_io) memcpy(s->jobfirst, jobfirst_io + (size_t)b * n1, n1);
    if (tabu_io) { s->tabu[0] = tabu_io[2 * b]; s->tabu[1] = tabu_io[2 * b + 1]; }
    int st = 0, np = 0, li = 0;
    int ms = orc_evaluate(s);
    if (ms < 0) st = -5;
    int inc = incumbent_out && incumbent_out[b] > 0 ? (int)incumbent_out[b] : ms;
    if (!st) { np = orc_pairs(s, orc_critical_path(s), pl, pmax); if (np < 0) { st = np; np = 0; } if (np > pmax) { st = -8; np = pmax; } }
    for (int k = 0; k < K; ++k) {
      int a0 = 0, a1 = 0;
      if (!st) {
        if (policy == 0) { a0 = tape[((size_t)b * K + k) * 2]; a1 = tape[((size_t)b * K + k) * 2 + 1]; }
        else if (np > 0) {
          int pick = 0;
          if (policy == 1) pick = (int)(((mix64(seed, (uint64_t)(inst_off + b), (uint64_t)(itr0 + k)) >> 32) * (uint64_t)np) >> 32);
          else {
            int best = 0x7fffffff;
            for (int i = 0; i < np; ++i) {
              orc_apply(s, pl[2 * i], pl[2 * i + 1], 0);
              int nms = orc_evaluate(s);
              orc_apply(s, pl[2 * i + 1], pl[2 * i], 0);
              if (nms >= 0 && nms < best) { best = nms; pick = i; }
            }
          }
          a0 = pl[2 * pick]; a1 = pl[2 * pick + 1];
        }
        int r = orc_apply(s, a0, a1, 1);
        if (r < 0) st = r;
        if (!st) {
          ms = orc_evaluate(s);
          if (ms < 0) st = -5;
        }
        if (!st) {
          if (ms < inc) inc = ms;
          int len = orc_critical_path(s);
          np = orc_pairs(s, len, pl, pmax);
          if (np < 0) { st = np; np = 0; }
          if (np > pmax) { st = -8; np = pmax; }
          if (write_state && x_out) {
            /* x = [dur/high, est/c, lst/c] float32: env/environment.py:312 */
            float* xo = x_out + (size_t)b * n1 * 3;
            for (int v = 0; v < n1; ++v) {
              xo[3 * v] = (float)dur_of(s, v) / high;
              xo[3 * v + 1] = (float)s->est[v] / fea;
              xo[3 * v + 2] = (float)s->lst[v] / fea;
            }
          }
          if (write_state && emc_out) {
            int64_t* src = emc_out + (size_t)b * e_mc; int64_t* dst = emc_out + (size_t)B * e_mc + (size_t)b * e_mc;
            int c = 0;
            for (int v = 1; v <= n; ++v) if (s->msucc[v]) { src[c] = (int64_t)b * n1 + v; dst[c] = (int64_t)b * n1 + s->msucc[v]; ++c; }
          }
        }
      }
      if (taken) { taken[((size_t)b * K + k) * 2] = a0; taken[((size_t)b * K + k) * 2 + 1] = a1; }
      if (li < n_log && k + 1 == log_steps[li]) { inc_log[(size_t)li * B + b] = (float)inc; ++li; }
    }
    if (makespan_out) makespan_out[b] = (float)ms;
    if (incumbent_out) incumbent_out[b] = (float)inc;
    if (pairs_out) { memset(pairs_out + (size_t)b * pmax * 2, 0, sizeof(int32_t) * 2 * pmax); memcpy(pairs_out + (size_t)b * pmax * 2, pl, sizeof(int32_t) * 2 * np); }
    if (npairs_out) npairs_out[b] = np;
    if (est_out) memcpy(est_out + (size_t)b * n1, s->est, sizeof(int32_t) * n1);
    if (lst_out) memcpy(lst_out + (size_t)b * n1, s->lst, sizeof(int32_t) * n1);
    if (jobfirst_io) memcpy(jobfirst_io + (size_t)b * n1, s->jobfirst, n1);
    if (tabu_io) { tabu_io[2 * b] = s->tabu[0]; tabu_io[2 * b + 1] = s->tabu[1]; }
    if (st) {
#ifdef _OPENMP
#pragma omp critical
#endif
      if (!status) status = st;
    }
    free(pl);
    orc_free(s);
  }
  return status;
}

/* ---------------------------------------------------------------------------------------------------
 * Initial solution: JsspN5._rules_solver (env/environment.py:203-255) + permissibleLeftShift
 * (env/permissible_LS.py:5-78).  rule 0 = spt, 1 = fdd-divide-mwkr.  Priority ties: first tied job in job
 * order (the reference draws np.random.choice); *ties_out counts tie events.  mseq_out [n_m*n_j] node ids.
 * --------------------------------------------------------------------------------------------------- */
int orc_rules_solver(int n_j, int n_m, const int32_t* dur, const int32_t* mch, int rule, int high, int32_t* mseq_out,
                     int32_t* ties_out) {
  const int n = n_j * n_m;
  int32_t* starts = (int32_t*)malloc(4 * n); int32_t* ops = (int32_t*)malloc(4 * n);
  int32_t* cnt = (int32_t*)calloc(n_m, 4); int32_t* nxt = (int32_t*)calloc(n_j, 4); int32_t* rdy = (int32_t*)calloc(n_j, 4);
  int64_t* total = (int64_t*)calloc(n_j, 8); int64_t* donew = (int64_t*)calloc(n_j, 8);
  for (int i = 0; i < n; ++i) { starts[i] = -high; ops[i] = -n_j; }
  for (int j = 0; j < n_j; ++j) for (int k = 0; k < n_m; ++k) total[j] += dur[j * n_m + k];
  int ties = 0;
  for (int it = 0; it < n; ++it) {
    double best = 0; int bj = -1, tied = 0;
    for (int j = 0; j < n_j; ++j) {
      if (nxt[j] >= n_m) continue;
      int d = dur[j * n_m + nxt[j]];
      double p = rule == 0 ? (double)d : (double)(donew[j] + d) / (double)(total[j] - donew[j]);
      if (bj < 0 || p < best) { best = p; bj = j; tied = 0; }
      else if (p == best) tied = 1;
    }
    ties += tied;
    int j = bj, a = j * n_m + nxt[j], d = dur[a], m = mch[a] - 1;
    int32_t* rs = starts + m * n_j; int32_t* ro = ops + m * n_j;
    int c = cnt[m], job_rdy = rdy[j];
    int mch_rdy = c ? rs[c - 1] + dur[ro[c - 1]] : 0;
    int p0 = c;
    for (int k = 0; k < c; ++k) if (rs[k] > job_rdy) { p0 = k; break; }
    int ins = -1, start_a = 0;
    if (p0 < c) {
      /* index -1 wraps to the unfilled row tail exactly as in the reference (env/permissible_LS.py:41) */
      int before = p0 > 0 ? rs[p0 - 1] + dur[ro[p0 - 1]] : -high + dur[n - n_j];
      for (int k = p0; k < c; ++k) {
        int open = (k == p0) ? (job_rdy > before ? job_rdy : before) : rs[k - 1] + dur[ro[k - 1]];
        if (d <= rs[k] - open) { ins = k; start_a = open; break; }
      }
    }
    if (ins >= 0) {
      for (int k = c; k > ins; --k) { rs[k] = rs[k - 1]; ro[k] = ro[k - 1]; }
      rs[ins] = start_a; ro[ins] = a;
    } else {
      start_a = job_rdy > mch_rdy ? job_rdy : mch_rdy;
      rs[c] = start_a; ro[c] = a;
    }
    cnt[m] = c + 1; rdy[j] = start_a + d; donew[j] += d; nxt[j] += 1;
  }
  for (int i = 0; i < n; ++i) mseq_out[i] = ops[i] + 1;
  if (ties_out) *ties_out = ties;
  free(starts); free(ops); free(cnt); free(nxt); free(rdy); free(total); free(donew);
  return 0;
}
