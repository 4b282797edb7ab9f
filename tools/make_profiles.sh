#!/bin/bash
# Run on the GPU box (via gpurun): captures the evidence that tools/collect_profiles.py turns into profiles/.
#   tools/make_profiles.sh <round-tag>
tag=${1:-r02}
mkdir -p gpurun_out
export DGTTA_BENCH_NO_TTA=1 DGTTA_BENCH_NO_EAGER=1
# 1. every launch of the bench command with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench_under_ncu.log 2>&1
unset DGTTA_BENCH_NO_TTA DGTTA_BENCH_NO_EAGER
# 2. full sections for the hot kernels (one capture each; -lineinfo sources imported)
ncu --set full --clock-control none --import-source on -k regex:"mind_fast_kernel" -s 2 -c 1 -o gpurun_out/${tag}_mind_noise \
    python tools/prof_mind.py mind_noise > gpurun_out/${tag}_ncu_mind_noise.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mind_fast_kernel" -s 2 -c 1 -o gpurun_out/${tag}_mind_clean \
    python tools/prof_mind.py mind > gpurun_out/${tag}_ncu_mind_clean.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gin_stack_kernel" -s 4 -c 2 -o gpurun_out/${tag}_gin3333 \
    python tools/prof_mind.py gin3333 > gpurun_out/${tag}_ncu_gin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"affine_sample" -c 4 -o gpurun_out/${tag}_sampler \
    python tools/prof_mind.py sampler > gpurun_out/${tag}_ncu_sampler.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"normal_fill_kernel" -s 1 -c 1 -o gpurun_out/${tag}_philox \
    python tools/prof_mind.py philox > gpurun_out/${tag}_ncu_philox.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sums_kernel|grad_kernel" -s 2 -c 2 -o gpurun_out/${tag}_closs \
    python tools/prof_closs.py > gpurun_out/${tag}_ncu_closs.log 2>&1
# 3. the numbers themselves (never taken under a profiler)
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
python bench.py --workload tta --steps 16 --warmup 3 > gpurun_out/${tag}_bench_tta.json 2>> gpurun_out/${tag}_bench.err
python tools/kernel_times.py > gpurun_out/${tag}_kernel_times.txt 2>&1
python tests/perf_eager_gpu.py > gpurun_out/${tag}_eager_vs_ours.json 2>> gpurun_out/${tag}_bench.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${tag}_nvidia_smi.csv
# 4. compute-sanitizer over small-shape parity tests of every kernel family (memcheck) and the TMA / mbarrier MIND path (racecheck)
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_boundary_behaviour_gpu.py tests/test_resize_gpu.py tests/test_philox_gpu.py \
    "tests/test_gin_gpu.py::test_single_block_forward" "tests/test_sampler_gpu.py::test_deterministic_gather_backward" tests/test_host_pipeline_gpu.py -m gpu -q -x > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${tag}_sanitizer_racecheck.log
compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/${tag}_sanitizer_memcheck_small.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${tag}_sanitizer_memcheck_small.log
for f in gpurun_out/${tag}_sanitizer_*.log; do tail -n 3 $f; done
ls -la gpurun_out | tail -30
