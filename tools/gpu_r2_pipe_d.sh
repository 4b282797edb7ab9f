#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mind_gpu.py tests/test_chain_gpu.py tests/test_multires_gpu.py -m gpu -q -x > gpurun_out/pipe_d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pipe_d_pytest.log
tail -5 gpurun_out/pipe_d_pytest.log
export DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_dbg.so
echo "=== pipe CTA times"; python tools/dbg_pipe_times.py 2x1x192x192x192 2>&1 | tail -9
echo "=== batch CTA times"; DGTTA_MIND_NO_PIPE=1 python tools/dbg_pipe_times.py 2x1x192x192x192 2>&1 | tail -9
unset DGTTA_LIB_PATH
echo "== pipe (default)"; timeout 300 python tools/kernel_times.py mind 2>&1 | grep -E "mind_" | tee gpurun_out/pipe_d_times_pipe.txt
echo "== batch kernel (DGTTA_MIND_NO_PIPE=1)"; DGTTA_MIND_NO_PIPE=1 timeout 300 python tools/kernel_times.py mind 2>&1 | grep -E "mind_" | tee gpurun_out/pipe_d_times_batch.txt
