#!/bin/bash
# first GPU call for the plane-pipelined MIND kernel: parity tests, then A/B timing against the batch kernel and role variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mind_gpu.py tests/test_chain_gpu.py -m gpu -q -x > gpurun_out/pipe_a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pipe_a_pytest.log
tail -15 gpurun_out/pipe_a_pytest.log
echo "== pipe (default)"; timeout 300 python tools/kernel_times.py mind 2>&1 | grep -E "mind_" | tee gpurun_out/pipe_a_times_pipe.txt
echo "== batch kernel (DGTTA_MIND_NO_PIPE=1)"; DGTTA_MIND_NO_PIPE=1 timeout 300 python tools/kernel_times.py mind 2>&1 | grep -E "mind_" | tee gpurun_out/pipe_a_times_batch.txt
for v in roles1 roles2 roles3 ns4; do
  echo "== $v"; DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_$v.so timeout 300 python tools/kernel_times.py mind 2>&1 | grep -E "mind_(noise|clean_d1_2)" | tee gpurun_out/pipe_a_times_$v.txt
done
