#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest.log
tail -6 gpurun_out/r02e_pytest.log
python tools/kernel_times.py > gpurun_out/r02e_kernel_times.txt 2>&1; grep -E "gin_k|gin_aug|gin_mind_aug_2x19|sample_|closs|epilogue|get_batch|Error|error" gpurun_out/r02e_kernel_times.txt
