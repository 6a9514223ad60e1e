/* Plain-C host using the l2s_b200 C ABI directly (no Python, no torch): two 4x3 instances, fdd-divide-mwkr
 * initial solutions built on the device, then N5 steps that always take the first feasible pair.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/capi_demo.c -o capi_demo \
 *       -L l2s_b200/lib -ll2s_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/l2s_b200/lib
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "l2s_b200.h"

#define CHECK(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "%s failed: %d (%s)\n", #x, r_, l2s_status_string(r_)); return 1; } } while (0)

int main(void) {
  enum { B = 2, NJ = 4, NM = 3, N = NJ * NM, N1 = N + 2, EMC = NM * (NJ - 1), STEPS = 6 };
  /* durations and 1-based machine ids, [B][n_j*n_m] row-major by (job, position) */
  const int32_t dur[B][N] = {{5, 3, 7, 2, 8, 4, 6, 1, 9, 3, 3, 5}, {4, 4, 2, 9, 1, 6, 3, 7, 5, 2, 8, 6}};
  const int32_t mch[B][N] = {{1, 2, 3, 2, 1, 3, 3, 1, 2, 1, 3, 2}, {2, 3, 1, 1, 2, 3, 3, 2, 1, 2, 1, 3}};
  int32_t *d_dur, *d_mch;
  float *d_x, *d_ms;
  int64_t* d_emc;
  cudaMalloc((void**)&d_dur, sizeof(dur)); cudaMalloc((void**)&d_mch, sizeof(mch));
  cudaMalloc((void**)&d_x, sizeof(float) * B * N1 * 3); cudaMalloc((void**)&d_emc, sizeof(int64_t) * 2 * B * EMC);
  cudaMalloc((void**)&d_ms, sizeof(float) * B);
  cudaMemcpy(d_dur, dur, sizeof(dur), cudaMemcpyHostToDevice);
  cudaMemcpy(d_mch, mch, sizeof(mch), cudaMemcpyHostToDevice);

  l2s_handle h;
  CHECK(l2s_create(NJ, NM, B, 0, 0, &h));
  CHECK(l2s_load_instances(h, d_dur, d_mch, 0, NULL));
  CHECK(l2s_build_initial(h, L2S_RULE_FDD_DIVIDE_MWKR, 99, NULL));
  CHECK(l2s_poll_error(h, NULL, NULL));

  const int pmax = l2s_pmax(h);
  int32_t* pairs = (int32_t*)malloc(sizeof(int32_t) * B * pmax * 2);
  int32_t npairs[B], status[B], actions[B][2];
  float ms[B], reward[B];
  uint8_t done[B];
  l2s_step_io io;
  memset(&io, 0, sizeof(io));
  io.high = 99.f; io.fea_norm_const = 1000.f; io.reward_type = L2S_REWARD_YAOXIN;
  io.x = d_x; io.edge_index_mc = d_emc; io.makespan = d_ms;
  io.is_reset = 1;
  CHECK(l2s_step_host(h, &io, NULL, pairs, npairs, done, reward, ms, status, NULL));
  printf("reset: makespans %.0f %.0f, moves %d %d\n", ms[0], ms[1], npairs[0], npairs[1]);
  io.is_reset = 0;
  for (int k = 0; k < STEPS; ++k) {
    for (int b = 0; b < B; ++b) {       /* first feasible pair, or the dummy action */
      actions[b][0] = npairs[b] ? pairs[(b * pmax) * 2] : 0;
      actions[b][1] = npairs[b] ? pairs[(b * pmax) * 2 + 1] : 0;
    }
    CHECK(l2s_step_host(h, &io, &actions[0][0], pairs, npairs, done, reward, ms, status, NULL));
    if (status[0] || status[1]) { fprintf(stderr, "device status %d %d\n", status[0], status[1]); return 1; }
    printf("step %d: actions (%d,%d) (%d,%d) -> makespans %.0f %.0f, rewards %.0f %.0f, moves %d %d\n", k + 1,
           actions[0][0], actions[0][1], actions[1][0], actions[1][1], ms[0], ms[1], reward[0], reward[1], npairs[0],
           npairs[1]);
  }
  float x[3];
  cudaMemcpy(x, d_x + 3, sizeof(x), cudaMemcpyDeviceToHost);      /* features of operation 1 of instance 0 */
  printf("x[op 1] = %.6f %.6f %.6f\n", x[0], x[1], x[2]);
  CHECK(l2s_destroy(h));
  free(pairs);
  printf("ok\n");
  return 0;
}
