export LLKV_GPU_JIT_VERBOSE=1
timeout 300 python tools/exp_part.py 20000000 250000 2>&1 | tail -12
