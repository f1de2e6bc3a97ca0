#!/bin/bash
# One GPU trip for the training step: parity tests, step benchmark, per-kernel launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -q --tb=short -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_train.log
timeout 300 python tools/train_bench.py --steps 50 --warmup 5 --torch-baseline > gpurun_out/train_bench.jsonl 2> gpurun_out/train_bench.err
timeout 300 python tools/train_bench.py --steps 50 --warmup 5 --two-sided >> gpurun_out/train_bench.jsonl 2>> gpurun_out/train_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/train_launches.csv python tools/train_bench.py --steps 2 --warmup 2 > gpurun_out/ncu_train.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log

tail -15 gpurun_out/pytest_train.log; cat gpurun_out/train_bench.jsonl; tail -3 gpurun_out/train_bench.err; tail -4 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgpr_train_edge_fwd -s 1 -c 1 -f -o gpurun_out/prof_train_fwd python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgpr_train_edge_bwd -s 1 -c 1 -f -o gpurun_out/prof_train_bwd python tools/train_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/train_e2e_bench.py --graphs 1000 --pairs 2560 --cpu-steps 1 2> gpurun_out/train_e2e.err | grep -E "^\{" > gpurun_out/train_e2e.jsonl; cat gpurun_out/train_e2e.jsonl; tail -3 gpurun_out/train_e2e.err
