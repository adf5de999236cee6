#!/bin/bash
# Round-2 ncu evidence for profiles/: (1) launch list with per-launch device time of one bench step (3 launches: stft, the network,
# map + gain + istft), (2) --set full captures of the network kernel and of the two signal kernels.  Plain run first, as the recipe asks.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sustained --no-extra-configs"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.err; exit 1; }
echo "plain ok"
ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 3 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tcn_chain -s 3 -c 1 -o gpurun_out/r02_prof_tcn_chain -f $CMD > gpurun_out/ncu_tcn.log 2>&1
echo "tcn full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"stft_kernel|istft_kernel" -s 6 -c 2 -o gpurun_out/r02_prof_signal -f $CMD > gpurun_out/ncu_signal.log 2>&1
echo "signal full rc=$?"
# MHANetV3 (C3 shape): one block of the network = QKV projection (with the fused K / V images), attention, projection, feed-forward in / out
MCMD="python scripts/mhanet_time.py 64 1875 f16x3"
$MCMD > gpurun_out/mha_plain.log 2>&1 || { echo "mha plain run failed"; tail -20 gpurun_out/mha_plain.log; exit 1; }
tail -3 gpurun_out/mha_plain.log
ncu --set full --clock-control none --import-source on -k regex:"lin_umma|attn_umma" -s 28 -c 5 -o gpurun_out/r02_prof_mhanet -f $MCMD > gpurun_out/ncu_mha.log 2>&1
echo "mha full rc=$?"
ls -la gpurun_out | head -30
