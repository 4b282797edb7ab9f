#!/bin/bash
# GPU call A of round 2: the whole -m gpu suite, the default bench line, MIND timings under the TMA L2-promotion knob.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a_pytest.log
tail -40 gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
for p in 0 1 2 3; do echo "promo $p"; DGTTA_TMA_PROMO=$p python tools/kernel_times.py mind 2>&1 | grep -E "noise_tensor|clean_d1_2x"; done > gpurun_out/r02a_promo.txt 2>&1
cat gpurun_out/r02a_promo.txt
python tools/kernel_times.py > gpurun_out/r02a_kernel_times.txt 2>&1; tail -30 gpurun_out/r02a_kernel_times.txt
