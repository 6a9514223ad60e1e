#!/bin/bash
# Memory-safety pass over every kernel family (run on the GPU box):
#     bash profiles/sanitize.sh [outdir] [tools...]
# tools: "plain" (default: the driver's own guard zones + run-twice determinism + evaluator cross-check) and/or
#        compute-sanitizer tools: memcheck racecheck synccheck initcheck.
# Shapes: warp kernel (10x5, 20x15), CTA kernel by default (50x20, 100x20) and forced with 1/2/4/8 warps per instance
# (L2S_CTA_WARPS) on shapes that would otherwise take the warp kernel; n_m = 32 and tiny 3x2 as edge cases.
OUT=${1:-gpurun_out/sanitizer_r2}
shift
TOOLS=${@:-plain}
mkdir -p "$OUT"
cd "$(dirname "$0")/.."
gcc -O1 -I include -I /usr/local/cuda/include profiles/sanitize_host.c -o /tmp/l2s_sanitize \
    -L l2s_b200/lib -ll2s_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/l2s_b200/lib || exit 1
run() {  # tool  tag  [env...] -- args
  local tool=$1 tag=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  if [ "$tool" = plain ]; then
    env "${envs[@]}" timeout 300 /tmp/l2s_sanitize "$@" > "$OUT/${tool}_${tag}.log" 2>&1
  else
    env "${envs[@]}" timeout 900 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 5 /tmp/l2s_sanitize "$@" \
        > "$OUT/${tool}_${tag}.log" 2>&1
  fi
  echo "$tool $tag ${envs[*]} args=$* rc=$? : $(tail -1 "$OUT/${tool}_${tag}.log")"
}
for tool in $TOOLS; do
  run $tool warp_10x5 -- 10 5 12 6
  run $tool warp_20x15 -- 20 15 9 4
  run $tool warp_20x15_big -- 20 15 700 6
  run $tool warp_3x2 -- 3 2 5 4
  run $tool warp_30x20 -- 30 20 300 4
  run $tool cta_50x20 -- 50 20 5 4
  run $tool cta_100x20 -- 100 20 3 3
  run $tool cta_150x25 -- 150 25 160 3
  run $tool nm32_8x32 -- 8 32 4 3
  for w in 1 2 4 8; do run $tool ctaw${w}_12x7 L2S_CTA_WARPS=$w -- 12 7 6 5; done
  run $tool ctaw4_20x15 L2S_CTA_WARPS=4 -- 20 15 5 4
  run $tool ctaw2_20x15_big L2S_CTA_WARPS=2 -- 20 15 600 5
done
