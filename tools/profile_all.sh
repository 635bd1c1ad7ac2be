#!/bin/bash
# One-call measurement sweep for the GPU box (run from the repo root under gpurun); everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/profile_all.sh r2a'
# Cost: ~6 min of box time.  Pass "quick" as 2nd argument to skip the ncu launch lists (~2.5 min).
set -u
TAG=${1:-run}
MODE=${2:-full}
OUT=gpurun_out
mkdir -p $OUT
{
  echo "== decode step (CUDA-graph replay, large-v3, 128 slots)"
  timeout 300 python tools/step_probe.py --active 128,64,16,1 2>&1 | tail -4
  echo "== decode-time GEMMs (graph-chained), modes 0 / 4 (no main loop) / 5 (launch only)"
  timeout 300 python tools/gemm_probe.py skinny 2>&1 | grep "mode=[045]"
  echo "== encoder GEMMs"
  timeout 300 python tools/gemm_probe.py big 2>&1 | grep "mode=0"
  echo "== encoder pass, 120 windows"
  timeout 300 python tools/encoder_probe.py --windows 120 --reps 2 2>&1 | tail -1
  echo "== alignment micro-workload"
  timeout 300 python bench.py --workload align --no-cpu-baseline 2>&1 | tail -1
  echo "== 1-h bench (2 steps)"
  WTS_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
} > $OUT/sweep_$TAG.txt 2>&1
if [ "$MODE" != "quick" ]; then
  for A in 128 1; do
    timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv \
      --log-file $OUT/step_${TAG}_a$A.csv python tools/step_probe.py --eager --steps 2 --active $A > /dev/null 2>&1
  done
  timeout 400 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv \
    --log-file $OUT/enc_$TAG.csv python tools/encoder_probe.py --windows 120 --reps 1 > /dev/null 2>&1
  for f in step_${TAG}_a128 step_${TAG}_a1 enc_$TAG; do
    python tools/summarize_launches.py $OUT/$f.csv 2 > $OUT/$f.md 2>/dev/null
  done
fi
tail -n 40 $OUT/sweep_$TAG.txt
