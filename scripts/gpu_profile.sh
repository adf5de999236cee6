#!/bin/bash
# ncu evidence for profiles/: (1) launch list with per-launch device time for one bench step,
# (2) one --set full capture of the dominant kernel, (3) full captures of the STFT / enhance / stem / head kernels.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.err; exit 1; }
echo "plain ok"
# 46 launches per step (stft, 2 stem, 41 stages, head, enhance): skip the 3 warm-up steps, list the timed step
ncu --metrics gpu__time_duration.sum --clock-control none -s 138 -c 46 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tcn_stage -s 60 -c 2 -o gpurun_out/prof_tcn_stage -f $CMD > gpurun_out/ncu_tcn.log 2>&1
echo "tcn full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"stft_kernel|istft_kernel|stem_umma|head_umma" -s 4 -c 5 -o gpurun_out/prof_signal -f $CMD > gpurun_out/ncu_signal.log 2>&1
echo "signal full rc=$?"
ls -la gpurun_out | head -40
