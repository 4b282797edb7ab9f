#!/bin/bash
# MIND-only GPU check: parity tests, per-CTA durations (needs gpurun_variants/lib_times.so), A/B timing rounds
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mind_gpu.py tests/test_chain_gpu.py tests/test_multires_gpu.py -m gpu -q -x > gpurun_out/mind_check_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/mind_check_pytest.log
tail -4 gpurun_out/mind_check_pytest.log
[ -f gpurun_variants/lib_times.so ] && DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_times.so python tools/dbg_cta_times.py 2x1x192x192x192 2>&1 | tail -9
for r in 1 2; do
  echo "== round $r  $(nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active --format=csv,noheader)"
  timeout 300 python tools/kernel_times.py mind 2>&1 | grep -E "mind_"
done
