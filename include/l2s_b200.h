/*
 * l2s_b200 -- C ABI of the B200-native (sm_100a) L2S hot path: batched disjunctive-graph evaluator and
 * N5 local-search step for job-shop scheduling.
 *
 * This is the drop-in boundary.  The reference (zcaicaros/L2S) is pure Python; the calls below are what a
 * Python binding (ctypes) needs in order to replace, with identical results:
 *
 *   l2s_eval_coo            <- Evaluator.forward                 env/message_passing_evl.py:161-199
 *   l2s_load_instances      <- JsspN5.reset (instance ingest)     env/environment.py:389-390
 *   l2s_build_initial       <- JsspN5._rules_solver               env/environment.py:203-255
 *                              + permissibleLeftShift             env/permissible_LS.py:5-78
 *   l2s_set_solution        <- (inject a machine order; what _rules_solver produces in opIDsOnMchs,
 *                               env/environment.py:221,254-255)
 *   l2s_step / l2s_step_host<- JsspN5.step                        env/environment.py:356-387
 *                              = change_nxgraph_topology :318-354, dag2pyg :291-316, reward/incumbent
 *                                :359-367, tabu :370-380, feasible_actions :411-423 (_gen_moves :72-81,
 *                                nx.dag_longest_path, _get_pairs :84-105)
 *   l2s_run                 <- the rollout loops of test_learned.py:185-189 / Greedy_baselines
 *                              (conventional_baselines_with_long-term-mem.py:154-221) with an on-device
 *                              policy, K steps per launch and no host round trip
 *
 * Conventions: plain C, no torch types.  Unless a parameter name ends in `_host`, every pointer is a DEVICE
 * pointer owned by the caller; the library owns only the opaque handle.  Calls are asynchronous and ordered
 * on the `cudaStream_t` passed as `void* stream` (NULL = default stream) unless stated otherwise.
 * Every function returns L2S_OK (0) or a negative l2s_status.  Per-instance problems found on the device
 * (illegal action, cyclic selection, ...) are reported through the `status` output array / l2s_poll_error.
 *
 * Node ids follow the reference: S = 0, operations 1..n row-major by (job, position), T = n+1, n = n_j*n_m;
 * in batched tensors instance b is offset by b*(n+2).
 */
#ifndef L2S_B200_H_
#define L2S_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct l2s_handle_s* l2s_handle;

typedef enum {
  L2S_OK = 0,
  L2S_ERR_INVALID_ARG = -1,
  L2S_ERR_CUDA = -2,
  L2S_ERR_UNSUPPORTED_SHAPE = -3, /* n_m > 32, n+2 > 65535, or state does not fit shared memory        */
  L2S_ERR_ILLEGAL_ACTION = -4,    /* (a0,a1) is not an adjacent pair on one machine; reference: networkx
                                     NetworkXError from remove_edge, env/environment.py:348             */
  L2S_ERR_BAD_GRAPH = -5,         /* cyclic selection / malformed COO (reference: evaluator runs N sweeps
                                     and nx.dag_longest_path raises)                                    */
  L2S_ERR_DURATION_RANGE = -6,    /* durations must be in [1, 4095], sum per instance < 2^24              */
  L2S_ERR_MACHINE_ROWS = -7,      /* every job must visit machines 1..n_m exactly once                    */
  L2S_ERR_PAIR_OVERFLOW = -8,     /* more N5 pairs than `pmax`; recreate the handle with a larger pmax    */
  L2S_ERR_SHORT_PATH = -9,        /* critical path of exactly two ops on one machine; the reference raises
                                     IndexError in _get_pairs, env/environment.py:89-90                   */
  L2S_ERR_BAD_SOLUTION = -10      /* l2s_set_solution: rows are not the operations of that machine        */
} l2s_status;

enum { L2S_REWARD_YAOXIN = 0, L2S_REWARD_CONSECUTIVE = 1 };        /* env/environment.py:359-364 */
enum { L2S_RULE_SPT = 0, L2S_RULE_FDD_DIVIDE_MWKR = 1 };           /* env/environment.py:227-238 */
enum { L2S_POLICY_REPLAY = 0, L2S_POLICY_RANDOM = 1, L2S_POLICY_GREEDY = 2 };

const char* l2s_status_string(int status);
int l2s_version(void);

/* ---- environment handle -------------------------------------------------------------------------- */

/* One handle per (rank, batch slice).  `pmax` = capacity of the per-instance N5 pair list (>= 1; 0 picks
 * the default 32).  Allocates all device state once; no allocation happens in any later call. */
int l2s_create(int n_j, int n_m, int batch, int pmax, int device, l2s_handle* out);
int l2s_destroy(l2s_handle h);

/* Queries (host only). */
int l2s_pmax(l2s_handle h);
int64_t l2s_edges_mc_per_instance(l2s_handle h); /* n_m*(n_j-1) */
int64_t l2s_itr(l2s_handle h);                   /* steps applied since the last reset (JsspN5.itr)      */
/* Set the step counter (JsspN5.itr): callers that advance the search outside l2s_step / l2s_run -- replaying a
 * captured CUDA graph -- keep the handle's counter (which keys the RANDOM policy's RNG) in step with theirs. */
int l2s_set_itr(l2s_handle h, int64_t itr);
int l2s_launch_info(l2s_handle h, int* warps_per_cta_step, int* smem_per_instance, int* warps_per_cta_run);

/* Durations and 1-based machine ids, both [batch, n_j*n_m] int32 (the reference's instances[:,0], [:,1]).
 * `instance_offset` is the global index of this slice's first instance (keys the RANDOM policy's counter
 * RNG so that sharding a batch over ranks does not change trajectories). */
int l2s_load_instances(l2s_handle h, const int32_t* dur, const int32_t* mch, int64_t instance_offset, void* stream);

/* Install a solution: mseq[b, m, k] = node id (1..n) of the k-th operation on machine m (0-based m).
 * Clears tabu lists, tie-break history and the step counter; does not evaluate (call l2s_step with
 * actions == NULL and is_reset = 1). */
int l2s_set_solution(l2s_handle h, const int32_t* mseq, void* stream);

/* Build the initial solution on the device by priority dispatching with permissible left shift.
 * Priority ties take the first tied job in job order (the reference draws np.random.choice there). */
int l2s_build_initial(l2s_handle h, int rule, int high, void* stream);

/* Copy the current machine orders out: mseq [batch, n_m, n_j] int32. */
int l2s_get_solution(l2s_handle h, int32_t* mseq, void* stream);

typedef struct {
  /* inputs */
  const int32_t* actions; /* [batch,2] (a0,a1); (0,0) = dummy; NULL = no move (re-evaluate only)          */
  float high;             /* x[:,0] = dur / high            env/environment.py:312                         */
  float fea_norm_const;   /* x[:,1:] = est|lst / const                                                     */
  int32_t reward_type;    /* L2S_REWARD_*                                                                  */
  int32_t is_reset;       /* 1: incumbent = current = makespan, reward = 0 (JsspN5.reset :403-405)         */
  /* outputs (any may be NULL) */
  float* x;               /* [batch*(n+2), 3]                                                              */
  int64_t* edge_index_mc; /* [2, batch*n_m*(n_j-1)], ascending source per instance                         */
  float* makespan;        /* [batch] current_objs                                                          */
  float* incumbent;       /* [batch] incumbent_objs                                                        */
  float* reward;          /* [batch]                                                                       */
  uint8_t* done;          /* [batch] 1 = no feasible move ([[0,0]] in the reference)                       */
  int32_t* pairs;         /* [batch, pmax, 2] feasible N5 pairs in critical-path order                     */
  int32_t* npairs;        /* [batch]                                                                       */
  int32_t* status;        /* [batch] 0 or a negative l2s_status for that instance                          */
} l2s_step_io;

/* One environment step for the whole slice, one kernel launch: swap, re-evaluate, features, machine edges,
 * reward/incumbent/tabu update, critical path, N5 move set. */
int l2s_step(l2s_handle h, const l2s_step_io* io, void* stream);

/* Same step driven from HOST memory (the call the Python drop-in's JsspN5.step makes, env/environment.py:356-387).
 * `actions_host` ([batch,2] int32, may be NULL = evaluate only) is read by the kernel in place when it is pinned host
 * memory (cudaHostAlloc / cudaHostRegister, or the handle's own action buffer); pageable memory is first copied
 * into the handle's pinned buffer; the device
 * outputs in `io` are written as in l2s_step (io->actions is ignored); the per-instance host results -- move set,
 * counts, reward, makespan, incumbent, done, status -- are written BY THE KERNEL as one coalesced record per
 * instance straight into mapped pinned host memory (l2s_host_records), so that a step needs no copy call; the
 * stream is synchronised before returning.  The optional host output arrays are unpacked from the records on the
 * host; pass NULL and read the records in place to skip that.  pairs_host is [batch,pmax,2] int32; only the
 * first npairs_host[b] pairs of row b are written. */
int l2s_step_host(l2s_handle h, const l2s_step_io* io, const int32_t* actions_host, int32_t* pairs_host,
                  int32_t* npairs_host, uint8_t* done_host, float* reward_host, float* makespan_host,
                  int32_t* status_host, void* stream);

/* The same call in two halves, for a host that drives SEVERAL independent environments (handles) from one thread and
 * wants the launch + completion latency of one hidden behind the kernels of the others: `begin` enqueues the step
 * (action transfer, kernel, record write-back) on `stream` and returns at once; `wait` blocks until that step's
 * records are in host memory (an event recorded behind the kernel, so later work on the same stream is not waited
 * for) and unpacks them like l2s_step_host.  One step in flight per handle (a second `begin` before `wait` is
 * L2S_ERR_INVALID_ARG); `actions_host` must stay untouched until `wait` returns when the kernel reads it in place.
 * l2s_step_host(...) == begin(...) followed by wait(...). */
int l2s_step_host_begin(l2s_handle h, const l2s_step_io* io, const int32_t* actions_host, void* stream);
int l2s_step_host_wait(l2s_handle h, int32_t* pairs_host, int32_t* npairs_host, uint8_t* done_host, float* reward_host,
                       float* makespan_host, int32_t* status_host);

/* Host result records of l2s_step_host: `records` is [batch][stride_words] uint32 in pinned host memory owned by
 * the handle, stride_words = L2S_REC_HEADER_WORDS + pmax rounded up to a multiple of 16 (every record starts on a
 * 64-byte line).  Record of instance b:
 *   word 0      npairs (bits 0-15) | done (bit 16) | per-instance l2s_status as int8 (bits 24-31)
 *   word 1..3   makespan, reward, incumbent as float32 bit patterns
 *   word 4+i    feasible pair i (i < npairs) in critical-path order: a0 | a1 << 16 (node ids fit 16 bits);
 *               words beyond npairs are stale
 * `actions` is the pinned action staging buffer [batch,2] int32: write actions there and pass the same pointer
 * as actions_host to skip a host copy.  Any out-pointer may be NULL. */
#define L2S_REC_HEADER_WORDS 4
#define L2S_REC_NPAIRS(w0) ((w0) & 0xffffu)
#define L2S_REC_DONE(w0) (((w0) >> 16) & 1u)
#define L2S_REC_STATUS(w0) ((int)(int8_t)((w0) >> 24))
int l2s_host_records(l2s_handle h, const uint32_t** records, int* stride_words, int32_t** actions);

/* Tuning knobs of a handle (all have measured defaults; also settable through the environment at l2s_create):
 *   "host_wait_flag"       l2s_step_host waits on a polled stream-ordered flag instead of cudaStreamSynchronize
 *   "host_zero_copy"       the kernel reads host actions in place (1) / they are copied H2D first (0)
 *   "host_action_prefetch" with zero copy: a small kernel pulls the action block with 16-byte coalesced reads
 *   "host_records_direct"  the kernel writes the records straight into host memory (1) / device staging + D2H (0)
 *   "host_record_lines"    records are written as whole 64-byte lines
 *   "incremental"          large instances: resume the sweeps from the evaluation kept in HBM
 * Returns L2S_ERR_INVALID_ARG for an unknown name. */
int l2s_set_option(l2s_handle h, const char* name, int value);

/* K fused steps in ONE launch with an on-device policy; the search state never leaves shared memory.
 *   REPLAY: tape [batch, K, 2] int32 actions.
 *   RANDOM: uniformly among the feasible pairs, index = mix64(seed, global instance, step) (see oracle).
 *   GREEDY: best neighbour by makespan, first minimum in move-list order, always moves.
 * taken (nullable) [batch, K, 2] receives the actions applied; incumbent_log (nullable)
 * [n_log, batch] float receives incumbents after log_steps_host[i] steps (ascending, 1-based counts). */
int l2s_run(l2s_handle h, int policy, int K, const int32_t* tape, uint64_t seed, int32_t* taken,
            const int32_t* log_steps_host, int n_log, float* incumbent_log, void* stream);

/* Test introspection (re-evaluates the current solution, mutates nothing): any pointer may be NULL.
 * est/lst [batch, n+2] int32 (both or neither), path [batch, n] int32 + path_len [batch] (both or neither),
 * tabu [batch,2], jobfirst [batch, n+2] uint8, makespan/incumbent [batch] float. */
int l2s_export_state(l2s_handle h, int32_t* est, int32_t* lst, int32_t* path, int32_t* path_len,
                     int32_t* tabu, uint8_t* jobfirst, float* makespan, float* incumbent, void* stream);

/* Synchronises `stream` and returns the first device-side error recorded since the last poll
 * (0 if none); *instance_out (nullable) receives the slice-local instance index. */
int l2s_poll_error(l2s_handle h, int* instance_out, void* stream);

/* ---- stateless evaluator (drop-in for Evaluator.forward) ------------------------------------------- */

/* Bytes of scratch the caller must provide for a batch of `batch` graphs (16-byte aligned device memory). */
int64_t l2s_eval_coo_workspace_bytes(int n_j, int n_m, int64_t batch);

/* edge_index [2,E] int64 COO in any order (conjunctive + machine arcs of `batch` complete selections),
 * duration [batch*(n+2)] int32 or int64 (`duration_is_i64`), outputs est/lst [batch*(n+2)] and
 * makespan [batch] as float32 (integer valued, like the reference).
 * workspace: every call hands it back ZERO-FILLED (for any shape).  workspace_is_clean != 0 states that it already
 * is zero-filled on entry (a cudaMemset once after allocation, or any earlier call on it) and saves the call its
 * own memset; pass 0 when in doubt.
 * status_flag: one int32 on the device that the CALLER zeroes; it is set (and then left) to a negative l2s_status if
 * an input is not a batch of acyclic complete JSSP selections -- check it after every call or once after many. */
int l2s_eval_coo(const int64_t* edge_index, int64_t n_edges, const void* duration, int duration_is_i64,
                 int n_j, int n_m, int64_t batch, float* est, float* lst, float* makespan,
                 void* workspace, int64_t workspace_bytes, int workspace_is_clean, int32_t* status_flag, int device,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* L2S_B200_H_ */
