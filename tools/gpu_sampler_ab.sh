#!/bin/bash
# sampler backward: scatter (default) vs the opt-in deterministic gather
export DGTTA_SAMPLE_BWD_DETERMINISTIC=1
timeout 900 python -m pytest tests/test_sampler_gpu.py tests/test_consistency_gpu.py tests/test_tta_step_gpu.py tests/test_boundary_behaviour_gpu.py -m gpu -q -x 2>&1 | tail -3
for r in 1 2; do
echo "== gather $r"; python tools/kernel_times.py sampler 2>&1 | grep -E "sample_logits_bwd"
[ -f gpurun_variants/lib_g3.so ] && { echo "== gather, 3 blocks/SM $r"; DGTTA_LIB_PATH=$PWD/gpurun_variants/lib_g3.so python tools/kernel_times.py sampler 2>&1 | grep -E "sample_logits_bwd"; }
done
unset DGTTA_SAMPLE_BWD_DETERMINISTIC
echo "== scatter"; python tools/kernel_times.py sampler 2>&1 | grep -E "sample_logits_bwd"
