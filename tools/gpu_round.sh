#!/bin/bash
# One GPU trip (round 2): parity tests, smoke, sanitizer, bench, ncu launch list + full capture of the fused kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 400 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgpr_embed -s 5 -c 2 -f -o gpurun_out/prof_embed python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/embed_probe.py > gpurun_out/embed_probe.json 2>&1
SGPR_SCOREMAT_V2=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:umma_kernel -s 1 -c 1 -f -o gpurun_out/prof_scoremat python tools/scoremat_profile.py 3 > gpurun_out/ncu_scoremat.log 2>&1
timeout 120 python tools/scoremat_check.py > gpurun_out/scoremat_check.jsonl 2>&1
timeout 300 python tools/train_e2e_bench.py > gpurun_out/train_e2e.jsonl 2> gpurun_out/train_e2e.err
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; head -c 1200 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err; head -c 600 gpurun_out/bench_reference.json; echo; tail -3 gpurun_out/sanitizer.log; cat gpurun_out/embed_probe.json; cat gpurun_out/train_e2e.jsonl
