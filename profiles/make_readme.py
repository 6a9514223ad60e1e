"""Regenerates the measured tables of profiles/README_r1.md from the artefacts brought back by gpurun:
profiles/bench_r1_1gpu.json, profiles/bench_r1_8gpu.json, profiles/launches_r1_bench_steps40.csv and (scratch)
gpurun_out/prof_step_r1_v2.ncu-rep + gpurun_out/cs_v2.csv.  Run from the repo root."""
import collections
import csv
import json
import subprocess

rows = [r for r in csv.reader(open('profiles/launches_r1_bench_steps40.csv')) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[iv].replace(',', ''))
    except ValueError:
        continue
    agg[r[ik]][0] += 1
    agg[r[ik]][1] += v
tot = sum(v[1] for v in agg.values())
raw = subprocess.run("ncu -i gpurun_out/prof_step_r1_v2.ncu-rep --page raw --csv", shell=True, capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]


def m(name):
    i = h.index(name)
    return [x[i] for x in rr[2:]]


N = float(m('smsp__inst_executed.sum')[0])
phase = subprocess.run("python profiles/by_phase.py gpurun_out/cs_v2.csv %d 8192" % int(N), shell=True, capture_output=True, text=True).stdout
traffic = (float(m('dram__bytes_read.sum')[0]) + float(m('dram__bytes_write.sum')[0])) * 1e6
json.dump({"20x15_b8192": traffic,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of l2s::step_kernel from the final ncu --set full "
                   "capture; below the algorithmic bytes because the x/edge outputs of the single replayed launch stay in "
                   "the 126 MB L2 (write-back)"}, open('profiles/step_kernel_traffic.json', 'w'))
b = json.loads(open('profiles/bench_r1_1gpu.json').read().strip().split('\n')[-1])
b8 = json.loads(open('profiles/bench_r1_8gpu.json').read().strip().split('\n')[-1])
old = open('profiles/README_r1.md').read()
keep_sweeps = old[old.index("Other shapes, same kernel family"):old.index("## Launch list")]
tail = old[old.index("History of this kernel during the round"):]
with open('profiles/README_r1.md', 'w') as f:
    f.write(f"""# Round-1 profiles (B200, sm_100a)

All numbers come from runs on the GPU box via `gpurun`; bench values are never taken under ncu.
Files: `bench_r1_1gpu.json`, `bench_r1_8gpu.json` (bench lines), `launches_r1_bench_steps40.csv` (ncu launch list),
`step_kernel_traffic.json` (DRAM bytes per launch from the `--set full` capture), `step_kernel_ncu_details_r1.txt` / `step_kernel_ncu_raw_r1.csv` (details and raw pages of that capture, exported with `ncu -i`), `timeline.py` / `timeline_eval.py` (+ `timeline_r1.txt`, `timeline_eval_r1.txt`), `shapes_r1.txt`, `by_phase.py` / `by_line.py` /
`sass_of.py` / `raw_metrics.py` / `make_readme.py` (the scripts that produced the tables below from the `.ncu-rep` exports).

## Bench line (`python bench.py`, defaults: 2000 steps, 8 warm-up)

* value (device-resident): **{b['value']:.3e} inst-steps/s**, {b['ms_per_step']*1e3:.1f} us per `l2s::step_kernel` launch over 8192 instances of 20x15
* roofline: achieved {b['roofline']['achieved']:.0f} GB/s algorithmic = **{b['roofline']['frac']:.3f} of the measured HBM peak** ({b['roofline']['peak']} GB/s)
* e2e (C-ABI `l2s_step_host`: the kernel reads host actions from and writes per-instance result records to pinned host memory, sync per step): {b['e2e']['value']:.3e} inst-steps/s
* e2e through the Python drop-in (`JsspN5.step`, Python lists in/out): {b['e2e_python_api']['value']:.3e} inst-steps/s
* evaluator (`Evaluator.forward` on int64 COO): {b['evaluator']['graphs_per_s']:.3e} graphs/s = {b['evaluator']['achieved_gbs']:.0f} GB/s ({b['evaluator']['frac_of_peak']:.3f} of peak)
* fused K-step kernel (`l2s_run`): {b['fused_run']['random_policy_inst_steps_per_s']:.3e} inst-steps/s (random policy), {b['fused_run']['greedy_policy_inst_steps_per_s']:.3e} (greedy best-neighbour)
* CPU port of the reference path (numpy/Python oracle, 1 core): {b['cpu_baseline']['value']:.0f} inst-steps/s
* clocks during the run: median SM {b['clocks']['sm_mhz']} MHz of {b['clocks']['sm_max_mhz']}, throttle reasons {b['clocks']['reasons']} ({b['clocks']['samples']} samples)
* 8 GPUs (weak scaling, 8192 instances per GPU, `bench_r1_8gpu.json`): {b8['value']:.3e} inst-steps/s; e2e {b8['e2e']['value']:.3e}; evaluator {b8['evaluator']['graphs_per_s']:.3e} graphs/s

{keep_sweeps}## Launch list

`ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv python bench.py --steps 40 --warmup 4 --no-cpu-baseline --e2e-steps 2`
(cold-cache, serialised: compare shares, not absolutes)

| kernel | launches | total us | avg us | share |
|---|---|---|---|---|
""")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:9]:
        f.write(f"| `{k[:64]}` | {v[0]} | {v[1]/1e3:.1f} | {v[1]/v[0]/1e3:.1f} | {100*v[1]/tot:.1f}% |\n")
    f.write(f"""
A bench step is exactly one `l2s::step_kernel` launch (share of the timed step: 100 %); the other launches are
set-up (`build_initial_kernel`, `load_kernel`, `set_solution_kernel`), action-tape recording and the fused phase
(`run_kernel`), the e2e phases (same `step_kernel`) and the evaluator phase (`coo_ingest_kernel` + `coo_eval_kernel`).

## `ncu --set full` of `l2s::step_kernel` (20x15, 8192 instances)

| metric | value |
|---|---|
| duration (under ncu, cold) | {m('gpu__time_duration.sum')[0]} us |
| warp instructions / launch | {N:.3e} ({N/8192:.0f} per instance-step) |
| IPC per SM (active) | {m('sm__inst_executed.avg.per_cycle_active')[0]} |
| issue slots busy | {m('smsp__issue_active.avg.pct_of_peak_sustained_active')[0]} % |
| registers / thread | {m('launch__registers_per_thread')[0]} |
| block / grid | {m('launch__block_size')[0]} / {m('launch__grid_size')[0]} |
| CTAs per SM (limit regs / smem) | {m('launch__occupancy_limit_registers')[0]} / {m('launch__occupancy_limit_shared_mem')[0]} |
| achieved warps active | {m('sm__warps_active.avg.pct_of_peak_sustained_active')[0]} % |
| LSU data-pipe wavefronts | {m('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')[0]} % of peak |
| dram read + write / launch | {m('dram__bytes_read.sum')[0]} MB + {m('dram__bytes_write.sum')[0]} MB |

Reading: the kernel is **instruction-issue bound** (issue slots ~70 % busy, LSU pipe ~55 %, DRAM ~6 %), not HBM
bound: the dependent integer chains of the longest-path sweep plus output formatting cost ~3.3 k warp
instructions per instance-step. DRAM traffic per launch is BELOW the algorithmic 110 MB because ncu replays one
launch with the 67 MB of x / edge_index_mc output still resident in the 126 MB L2 (write-back); there are no
wasted re-reads (state is staged once into shared memory by TMA bulk copies). Register caps of 64, 56 and 48 per
thread (32-40 resident warps) measure the same within 1 %: the kernel is issue-bound, not occupancy-bound; a
40-register cap spills and is slower.

Per-phase shares (`by_phase.py` on the source-correlated export of the same capture):

```
{phase}```

{tail}""")
print("ok")
