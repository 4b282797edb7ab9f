timeout 900 python -m pytest tests/test_gin_gpu.py tests/test_chain_gpu.py tests/test_multires_gpu.py -m gpu -q -x 2>&1 | tail -3
for r in 1 2; do for sk in 1 1.5 1.3 1.7; do echo "== skew $sk (round $r)"; DGTTA_GIN_SKEW=$sk python tools/kernel_times.py gin 2>&1 | grep -E "gin_k3333|gin_k3131|gin_aug_seeds"; done; done
