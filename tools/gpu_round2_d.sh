#!/bin/bash
# GPU call D (2 GPUs): the bench contract under torchrun at N=2 (affinity, copy ceiling, TTA sub-record on both ranks) + the failing test re-run
mkdir -p gpurun_out
python -m pytest tests/test_consistency_gpu.py tests/test_sampler_gpu.py tests/test_tta_step_gpu.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
tail -3 gpurun_out/r02d_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_n2.json')); print('N2', d['ms_per_step'], d['value'], d['e2e'], d['tta'])"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench.json')); print('N1', d['ms_per_step'], d['value'], d['e2e'], d['tta']['value'])"
python tools/kernel_times.py 2>&1 | grep -E "sample_img|get_batch|epilogue"
