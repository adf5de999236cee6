#!/bin/bash
# ncu --set full of the MHANetV3 tensor-core kernels (config C3 shape)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python scripts/mhanet_time.py 64 1875 f16x3"
$CMD > gpurun_out/mha_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/mha_plain.log; exit 1; }
tail -3 gpurun_out/mha_plain.log
ncu --set full --clock-control none --import-source on -k regex:"lin_umma|attn_umma|attn_pack" -s 33 -c 7 -o gpurun_out/prof_mha -f $CMD > gpurun_out/ncu_mha.log 2>&1
echo "mha full rc=$?"
ls -la gpurun_out/prof_mha.ncu-rep
