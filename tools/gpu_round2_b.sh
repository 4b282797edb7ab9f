#!/bin/bash
# GPU call B of round 2: GIN tests + timings after the one-barrier / grid-plan change, MIND full ncu capture for the phase split
mkdir -p gpurun_out
python -m pytest tests/test_gin_gpu.py tests/test_multires_gpu.py tests/test_chain_gpu.py tests/test_mind_gpu.py tests/test_resize_gpu.py -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest.log
tail -15 gpurun_out/r02b_pytest.log
python tools/kernel_times.py > gpurun_out/r02b_kernel_times.txt 2>&1; grep -E "gin|mind_noise|mind_clean_d1_2x" gpurun_out/r02b_kernel_times.txt
ncu --set full --clock-control none --import-source on -k regex:"mind_fast_kernel" -s 2 -c 1 -o gpurun_out/r02b_mind_noise \
    python tools/prof_mind.py mind_noise > gpurun_out/r02b_ncu_mind_noise.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gin_stack_kernel" -s 4 -c 2 -o gpurun_out/r02b_gin3333 \
    python tools/prof_mind.py gin3333 > gpurun_out/r02b_ncu_gin.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['tta']['value'], d['roofline']['frac'], d['gpu_launches'])"
ls -la gpurun_out | tail
