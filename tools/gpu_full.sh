#!/bin/bash
# full GPU check: the whole -m gpu suite, the kernel timing probe and the bench line
tag=${1:-r02x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python tools/kernel_times.py > gpurun_out/${tag}_kernel_times.txt 2>&1; cat gpurun_out/${tag}_kernel_times.txt | tail -60
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
