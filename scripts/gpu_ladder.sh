#!/bin/bash
# Run the GPU test groups one process each under a hard timeout, so that a hung or trapped kernel in one group
# cannot hide the results of the others.  Logs go to gpurun_out/ladder_*.log.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, timeout, pytest args...
  name=$1; to=$2; shift 2
  timeout -s KILL $to python -m pytest -x -q -s -m gpu "$@" > gpurun_out/ladder_$name.log 2>&1
  echo "[$name] exit $?"; tail -n 4 gpurun_out/ladder_$name.log
}
run basic 300 tests/test_gpu_kernels.py -k "library_loads or split_planes or layernorm_embed or error_codes"
run gemm 300 tests/test_gpu_kernels.py -k "gemm_planes"
run sdpa 300 tests/test_gpu_kernels.py -k "sdpa"
run mha 300 tests/test_gpu_kernels.py -k "mha or ffn"
