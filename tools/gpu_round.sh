#!/bin/bash
# One GPU trip: parity tests, smoke, sanitizer, microbench, bench, ncu launch list + full capture of the fused kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 120 ./tools/microbench_ffma > gpurun_out/microbench.json 2>&1
timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/perf_probe.py > gpurun_out/perf_probe.jsonl 2>&1
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgpr_embed -s 5 -c 2 -f -o gpurun_out/prof_embed python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/microbench.json; head -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/sanitizer.log
cat gpurun_out/perf_probe.jsonl
SGPR_B200_LIB=$PWD/tools/variants/lib_tl_base.so timeout 120 python tools/timeline.py 16 > gpurun_out/timeline.txt 2>&1; cat gpurun_out/timeline.txt | awk "NR<4 || /front start|select done|front done|barrier B|back done|layers done|attention|end/" | cut -c1-60
timeout 300 python tools/eval_batch_bench.py --graphs 1000 --pairs 12800 2>&1 | grep -E "^\{" > gpurun_out/eval_batch_bench.jsonl; cat gpurun_out/eval_batch_bench.jsonl | cut -c1-220
