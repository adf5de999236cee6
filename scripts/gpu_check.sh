#!/bin/bash
# One GPU-box pass: tcgen05 self test first (short timeout: a wrong barrier protocol must not hang the box),
# then the parity suite, smoke, a short bench and the ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== selftest" ; timeout -k 5 300 python -m pytest tests/test_gpu_network.py -q -k umma_selftest --maxfail=50 > gpurun_out/selftest.log 2>&1 ; echo "rc=$?" | tee -a gpurun_out/selftest.log ; tail -15 gpurun_out/selftest.log
echo "== signal" ; timeout -k 5 600 python -m pytest tests/test_gpu_signal.py tests/test_train_tgt.py -m gpu -q --maxfail=50 > gpurun_out/signal.log 2>&1 ; echo "rc=$?" | tee -a gpurun_out/signal.log ; tail -30 gpurun_out/signal.log
echo "== network" ; timeout -k 5 900 python -m pytest tests/test_gpu_network.py -q -k "not umma_selftest" --maxfail=50 > gpurun_out/network.log 2>&1 ; echo "rc=$?" | tee -a gpurun_out/network.log ; tail -40 gpurun_out/network.log
echo "== smoke" ; timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "rc=$?" | tee -a gpurun_out/smoke.log ; tail -5 gpurun_out/smoke.log
for prec in f16x3 f16 f32; do
  echo "== bench $prec" ; timeout -k 5 600 python bench.py --steps 5 --warmup 3 --precision $prec $( [ $prec != f16x3 ] && echo --no-cpu-baseline ) > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err ; echo "rc=$?" ; tail -c 3000 gpurun_out/bench_$prec.json ; tail -5 gpurun_out/bench_$prec.err
done
echo "== reference arm" ; timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ; tail -c 1500 gpurun_out/bench_reference.json
