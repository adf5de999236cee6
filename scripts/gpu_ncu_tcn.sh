#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:tcn_stage -s 60 -c 2 -o gpurun_out/prof_tcn_stage -f $CMD > gpurun_out/ncu_tcn.log 2>&1
echo "tcn full rc=$?"
