/* Memory-safety driver: a plain-C host on the C ABI that runs every kernel family of libl2s_b200.so on small
 * batches -- load / build_initial / set_solution / get_solution, the warp and the CTA step kernel (device and host
 * path, incremental re-evaluation on and off, dummy moves), the fused run kernel with all three policies, the state
 * export and the stateless COO evaluator -- for a list of shapes.  No Python, so it starts in milliseconds under
 * `compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck` (see profiles/sanitize.sh) where that tool is
 * available.  On its own it checks what it can without the tool: every device buffer the library writes is
 * allocated between two 512-byte GUARD ZONES filled with a pattern and verified after every phase (out-of-bounds
 * global writes), the whole sequence runs TWICE from the same seed and every output must be bit-identical (races
 * show up as nondeterminism), and the evaluator's makespans must equal the step kernel's.
 *
 *   gcc -O1 -I include -I /usr/local/cuda/include profiles/sanitize_host.c -o /tmp/l2s_sanitize \
 *       -L l2s_b200/lib -ll2s_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/l2s_b200/lib
 *   usage: l2s_sanitize n_j n_m batch steps
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "l2s_b200.h"

#define CHECK(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "%s failed: %d (%s) at line %d\n", #x, r_, l2s_status_string(r_), __LINE__); return 1; } } while (0)
#define CUCHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

/* guarded device allocations */
#define GUARD 512
#define MAXG 64
static struct { unsigned char* base; size_t bytes; const char* name; } g_allocs[MAXG];
static int g_nalloc = 0;
static int gmalloc(void** p, size_t bytes, const char* name) {
  unsigned char* base;
  const size_t padded = (bytes + 15) & ~(size_t)15;
  if (cudaMalloc((void**)&base, padded + 2 * GUARD) != cudaSuccess) return 1;
  cudaMemset(base, 0xA5, padded + 2 * GUARD);
  cudaMemset(base + GUARD, 0, padded);
  g_allocs[g_nalloc].base = base; g_allocs[g_nalloc].bytes = padded; g_allocs[g_nalloc].name = name; ++g_nalloc;
  *p = base + GUARD;
  return 0;
}
static int check_guards(const char* where) {
  unsigned char buf[GUARD];
  int bad = 0;
  cudaDeviceSynchronize();
  for (int i = 0; i < g_nalloc; ++i)
    for (int side = 0; side < 2; ++side) {
      cudaMemcpy(buf, g_allocs[i].base + (side ? GUARD + g_allocs[i].bytes : 0), GUARD, cudaMemcpyDeviceToHost);
      for (int k = 0; k < GUARD; ++k)
        if (buf[k] != 0xA5) { fprintf(stderr, "GUARD ZONE of %s (%s side, byte %d) overwritten after %s\n", g_allocs[i].name, side ? "high" : "low", k, where); bad = 1; break; }
    }
  return bad;
}
static void gfree_all(void) { for (int i = 0; i < g_nalloc; ++i) cudaFree(g_allocs[i].base); g_nalloc = 0; }
#define GMALLOC(p, bytes) do { if (gmalloc((void**)&(p), (bytes), #p)) { fprintf(stderr, "cudaMalloc %s failed\n", #p); return 1; } } while (0)
#define GUARDS(where) do { if (check_guards(where)) return 1; } while (0)
static uint64_t fnv(uint64_t h, const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } return h; }
static uint64_t dev_hash(uint64_t h, const void* d, size_t n) { void* t = malloc(n); cudaMemcpy(t, d, n, cudaMemcpyDeviceToHost); h = fnv(h, t, n); free(t); return h; }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 32); }

static int run_all(int argc, char** argv, uint64_t* digest);

int main(int argc, char** argv) {
  uint64_t d1 = 0, d2 = 0;
  rng_state = 0x9E3779B97F4A7C15ull;
  if (run_all(argc, argv, &d1)) return 1;
  rng_state = 0x9E3779B97F4A7C15ull;
  if (run_all(argc, argv, &d2)) return 1;
  if (d1 != d2) { fprintf(stderr, "NONDETERMINISTIC: output digests of two identical runs differ (%llx vs %llx)\n", (unsigned long long)d1, (unsigned long long)d2); return 1; }
  printf("two identical runs: output digest %016llx both times; all guard zones intact\n", (unsigned long long)d1);
  return 0;
}

static int run_all(int argc, char** argv, uint64_t* digest) {
  uint64_t hsh = 1469598103934665603ull;
  const int NJ = argc > 1 ? atoi(argv[1]) : 10, NM = argc > 2 ? atoi(argv[2]) : 5, B = argc > 3 ? atoi(argv[3]) : 8;
  const int STEPS = argc > 4 ? atoi(argv[4]) : 6;
  const int N = NJ * NM, N1 = N + 2, EMC = NM * (NJ - 1), EPC = NJ * (NM + 1);
  int32_t* dur = (int32_t*)malloc(sizeof(int32_t) * B * N);
  int32_t* mch = (int32_t*)malloc(sizeof(int32_t) * B * N);
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < NJ; ++j) {
      for (int k = 0; k < NM; ++k) { dur[b * N + j * NM + k] = 1 + rnd() % 98; mch[b * N + j * NM + k] = k + 1; }
      for (int k = NM - 1; k > 0; --k) {   /* random permutation of the machines */
        const int r = rnd() % (k + 1), t = mch[b * N + j * NM + k];
        mch[b * N + j * NM + k] = mch[b * N + j * NM + r];
        mch[b * N + j * NM + r] = t;
      }
    }
  int32_t *d_dur, *d_mch, *d_sol, *d_pairs, *d_npairs, *d_status, *d_actions, *d_taken, *d_flag;
  float *d_x, *d_ms, *d_inc, *d_rew, *d_log, *d_est, *d_lst, *d_evms;
  uint8_t* d_done;
  int64_t *d_emc, *d_ei, *d_dur64;
  void* d_ws;
  CUCHECK(cudaMalloc((void**)&d_dur, sizeof(int32_t) * B * N));
  CUCHECK(cudaMalloc((void**)&d_mch, sizeof(int32_t) * B * N));
  GMALLOC(d_sol, sizeof(int32_t) * B * N);
  CUCHECK(cudaMemcpy(d_dur, dur, sizeof(int32_t) * B * N, cudaMemcpyHostToDevice));
  CUCHECK(cudaMemcpy(d_mch, mch, sizeof(int32_t) * B * N, cudaMemcpyHostToDevice));
  GMALLOC(d_x, sizeof(float) * B * N1 * 3);
  GMALLOC(d_emc, sizeof(int64_t) * 2 * B * EMC);
  GMALLOC(d_ms, sizeof(float) * B);
  GMALLOC(d_inc, sizeof(float) * B);
  GMALLOC(d_rew, sizeof(float) * B);
  GMALLOC(d_done, B);
  GMALLOC(d_npairs, sizeof(int32_t) * B);
  GMALLOC(d_status, sizeof(int32_t) * B);
  CUCHECK(cudaMalloc((void**)&d_actions, sizeof(int32_t) * B * 2));
  GMALLOC(d_taken, sizeof(int32_t) * B * STEPS * 2);
  GMALLOC(d_log, sizeof(float) * B);
  GMALLOC(d_flag, sizeof(int32_t));

  for (int variant = 0; variant < 2; ++variant) {   /* 0: defaults, 1: flag wait + record lines + action prefetch + look-ahead sweep */
    l2s_handle h;
    CHECK(l2s_create(NJ, NM, B, NJ > 30 ? 128 : 0, 0, &h));
    if (variant == 1) {
      l2s_set_option(h, "incremental", 1);
      l2s_set_option(h, "host_wait_flag", 1);
      l2s_set_option(h, "host_record_lines", 1);
      l2s_set_option(h, "host_action_prefetch", 1);
      l2s_set_option(h, "fast_sweep", 2);
    }
    const int pmax = l2s_pmax(h);
    GMALLOC(d_pairs, sizeof(int32_t) * B * pmax * 2);
    int32_t* pairs = (int32_t*)malloc(sizeof(int32_t) * B * pmax * 2);
    int32_t* npairs = (int32_t*)malloc(sizeof(int32_t) * B);
    int32_t* status = (int32_t*)malloc(sizeof(int32_t) * B);
    int32_t* actions = (int32_t*)malloc(sizeof(int32_t) * B * 2);
    float* ms = (float*)malloc(sizeof(float) * B);
    float* reward = (float*)malloc(sizeof(float) * B);
    uint8_t* done = (uint8_t*)malloc(B);
    CHECK(l2s_load_instances(h, d_dur, d_mch, 0, NULL));
    CHECK(l2s_build_initial(h, variant ? L2S_RULE_SPT : L2S_RULE_FDD_DIVIDE_MWKR, 99, NULL));
    CHECK(l2s_poll_error(h, NULL, NULL));
    CHECK(l2s_get_solution(h, d_sol, NULL));
    CHECK(l2s_set_solution(h, d_sol, NULL));
    CHECK(l2s_poll_error(h, NULL, NULL));
    l2s_step_io io;
    memset(&io, 0, sizeof(io));
    io.high = 99.f; io.fea_norm_const = 1000.f; io.reward_type = L2S_REWARD_YAOXIN;
    io.x = d_x; io.edge_index_mc = d_emc; io.makespan = d_ms; io.incumbent = d_inc; io.reward = d_rew; io.done = d_done;
    io.is_reset = 1;
    CHECK(l2s_step_host(h, &io, NULL, pairs, npairs, done, reward, ms, status, NULL));
    io.is_reset = 0;
    for (int k = 0; k < STEPS; ++k) {
      for (int b = 0; b < B; ++b) {       /* a random feasible pair; every 5th instance-step a dummy move */
        const int pick = npairs[b] ? (int)(rnd() % npairs[b]) : 0;
        const int dummy = !npairs[b] || (rnd() % 5 == 0);
        actions[2 * b] = dummy ? 0 : pairs[(b * pmax + pick) * 2];
        actions[2 * b + 1] = dummy ? 0 : pairs[(b * pmax + pick) * 2 + 1];
      }
      if (k & 1) {                       /* device path: actions and every output in device memory */
        CUCHECK(cudaMemcpy(d_actions, actions, sizeof(int32_t) * B * 2, cudaMemcpyHostToDevice));
        l2s_step_io io2 = io;
        io2.actions = d_actions; io2.pairs = d_pairs; io2.npairs = d_npairs; io2.status = d_status;
        CHECK(l2s_step(h, &io2, NULL));
        CUCHECK(cudaMemcpy(pairs, d_pairs, sizeof(int32_t) * B * pmax * 2, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(npairs, d_npairs, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
        CUCHECK(cudaMemcpy(status, d_status, sizeof(int32_t) * B, cudaMemcpyDeviceToHost));
      } else {
        CHECK(l2s_step_host(h, &io, actions, pairs, npairs, done, reward, ms, status, NULL));
      }
      for (int b = 0; b < B; ++b)
        if (status[b]) { fprintf(stderr, "device status %d for instance %d at step %d\n", status[b], b, k); return 1; }
      GUARDS("a step");
      hsh = dev_hash(hsh, d_x, sizeof(float) * B * N1 * 3);
      hsh = dev_hash(hsh, d_emc, sizeof(int64_t) * 2 * B * EMC);
      hsh = fnv(hsh, npairs, sizeof(int32_t) * B);
      for (int b = 0; b < B; ++b) hsh = fnv(hsh, pairs + (size_t)b * pmax * 2, sizeof(int32_t) * 2 * npairs[b]);
    }
    /* state export (evaluate-only launch), fused runs with the three policies, one more step afterwards */
    GMALLOC(d_est, sizeof(int32_t) * B * N1);
    GMALLOC(d_lst, sizeof(int32_t) * B * N1);
    CHECK(l2s_export_state(h, (int32_t*)d_est, (int32_t*)d_lst, NULL, NULL, NULL, NULL, d_ms, d_inc, NULL));
    int32_t log_step = STEPS;
    CHECK(l2s_run(h, L2S_POLICY_RANDOM, STEPS, NULL, 7, d_taken, &log_step, 1, d_log, NULL));
    CHECK(l2s_run(h, L2S_POLICY_REPLAY, STEPS, d_taken, 0, NULL, NULL, 0, NULL, NULL));   /* (replays stale moves: */
    cudaDeviceSynchronize();                                                              /*  illegal ones only flag) */
    l2s_poll_error(h, NULL, NULL);
    CHECK(l2s_run(h, L2S_POLICY_GREEDY, 3, NULL, 0, d_taken, NULL, 0, NULL, NULL));
    CHECK(l2s_poll_error(h, NULL, NULL));
    GUARDS("the fused runs");
    hsh = dev_hash(hsh, d_taken, sizeof(int32_t) * B * STEPS * 2);
    hsh = dev_hash(hsh, d_log, sizeof(float) * B);
    CHECK(l2s_step_host(h, &io, NULL, pairs, npairs, done, reward, ms, status, NULL));
    /* stateless evaluator on the current graphs: conjunctive arcs + the machine arcs the step just wrote */
    {
      const int64_t E = (int64_t)B * (EPC + EMC);
      int64_t* ei = (int64_t*)malloc(sizeof(int64_t) * 2 * E);
      int64_t* emc = (int64_t*)malloc(sizeof(int64_t) * 2 * B * EMC);
      int64_t* d64 = (int64_t*)calloc((size_t)B * N1, sizeof(int64_t));
      CUCHECK(cudaMemcpy(emc, d_emc, sizeof(int64_t) * 2 * B * EMC, cudaMemcpyDeviceToHost));
      int64_t e = 0;
      for (int b = 0; b < B; ++b) {
        const int64_t off = (int64_t)b * N1;
        for (int j = 0; j < NJ; ++j) { ei[e] = off; ei[E + e] = off + j * NM + 1; ++e; }
        for (int v = 1; v <= N; ++v) { ei[e] = off + v; ei[E + e] = (v % NM) ? off + v + 1 : off + N + 1; ++e; }
        for (int v = 1; v <= N; ++v) d64[off + v] = dur[b * N + v - 1];
      }
      for (int64_t k = 0; k < (int64_t)B * EMC; ++k) { ei[e] = emc[k]; ei[E + e] = emc[(int64_t)B * EMC + k]; ++e; }
      for (int64_t k = E - 1; k > 0; --k) {   /* any edge order */
        const int64_t r = rnd() % (k + 1), t0 = ei[k], t1 = ei[E + k];
        ei[k] = ei[r]; ei[E + k] = ei[E + r]; ei[r] = t0; ei[E + r] = t1;
      }
      const int64_t wsb = l2s_eval_coo_workspace_bytes(NJ, NM, B);
      CUCHECK(cudaMalloc((void**)&d_ei, sizeof(int64_t) * 2 * E));
      CUCHECK(cudaMalloc((void**)&d_dur64, sizeof(int64_t) * B * N1));
      GMALLOC(d_ws, (size_t)wsb);
      GMALLOC(d_evms, sizeof(float) * B);
      CUCHECK(cudaMemcpy(d_ei, ei, sizeof(int64_t) * 2 * E, cudaMemcpyHostToDevice));
      CUCHECK(cudaMemcpy(d_dur64, d64, sizeof(int64_t) * B * N1, cudaMemcpyHostToDevice));
      CUCHECK(cudaMemset(d_flag, 0, 4));
      CHECK(l2s_eval_coo(d_ei, E, d_dur64, 1, NJ, NM, B, d_est, d_lst, d_evms, d_ws, wsb, 0, d_flag, 0, NULL));
      CHECK(l2s_eval_coo(d_ei, E, d_dur64, 1, NJ, NM, B, d_est, d_lst, d_evms, d_ws, wsb, 1, d_flag, 0, NULL));
      int32_t flag = 0;
      float* evms = (float*)malloc(sizeof(float) * B);
      CUCHECK(cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost));
      CUCHECK(cudaMemcpy(evms, d_evms, sizeof(float) * B, cudaMemcpyDeviceToHost));
      if (flag) { fprintf(stderr, "evaluator status %d\n", flag); return 1; }
      GUARDS("the evaluator");
      hsh = dev_hash(hsh, d_est, sizeof(float) * B * N1);
      hsh = dev_hash(hsh, d_lst, sizeof(float) * B * N1);
      for (int b = 0; b < B; ++b)
        if (evms[b] != ms[b]) { fprintf(stderr, "evaluator makespan %f != step makespan %f (instance %d)\n", evms[b], ms[b], b); return 1; }
      cudaFree(d_ei); cudaFree(d_dur64);
      free(ei); free(emc); free(d64); free(evms);
    }
    printf("%dx%d B=%d variant %d: ok, makespan[0] = %.0f\n", NJ, NM, B, variant, ms[0]);
    CHECK(l2s_destroy(h));
    free(pairs); free(npairs); free(status); free(actions); free(ms); free(reward); free(done);
  }
  gfree_all();
  cudaFree(d_dur); cudaFree(d_mch); cudaFree(d_actions);
  free(dur); free(mch);
  *digest = hsh;
  return 0;
}
