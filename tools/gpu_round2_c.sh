#!/bin/bash
# GPU call C of round 2: full suite with the fused warp+loss kernels and the GIN micro-optimisations; timings; bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest.log
tail -25 gpurun_out/r02c_pytest.log
python tools/kernel_times.py > gpurun_out/r02c_kernel_times.txt 2>&1; grep -E "gin_k|gin_aug|sample_|closs|epilogue|get_batch|lowres|Error|error" gpurun_out/r02c_kernel_times.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['tta']['value'], d['tta']['transform_ms_per_step'], d['roofline']['frac'], d['gpu_launches'])"
tail -3 gpurun_out/r02c_bench.err
