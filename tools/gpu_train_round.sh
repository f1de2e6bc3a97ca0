#!/bin/bash
# One GPU trip: every -m gpu test, the eval bench line, the training step benchmark + per-kernel launch list + ncu captures
# of the EdgeConv forward/backward kernels, and the end-to-end training benchmark.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python tools/train_bench.py --steps 50 --warmup 5 --torch-baseline > gpurun_out/train_bench.jsonl 2> gpurun_out/train_bench.err
timeout 300 python tools/train_bench.py --steps 50 --warmup 5 --two-sided >> gpurun_out/train_bench.jsonl 2>> gpurun_out/train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py --steps 2 --warmup 2 > gpurun_out/ncu_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgpr_train_edge_fwd -s 1 -c 1 -f -o gpurun_out/prof_train_fwd python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgpr_train_edge_bwd -s 1 -c 1 -f -o gpurun_out/prof_train_bwd python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_bwd.log 2>&1
timeout 600 python tools/train_e2e_bench.py --graphs 1000 --pairs 2560 --cpu-steps 1 2> gpurun_out/train_e2e.err | grep -E "^\{" > gpurun_out/train_e2e.jsonl
if [ "$1" == "full" ]; then
  timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgpr_embed -s 5 -c 2 -f -o gpurun_out/prof_embed python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe.jsonl 2>&1
  head -c 1500 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
fi
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/train_bench.jsonl; tail -3 gpurun_out/train_bench.err; cat gpurun_out/train_e2e.jsonl; tail -3 gpurun_out/train_e2e.err
