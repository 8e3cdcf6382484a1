#!/bin/bash
# second, wider compute-sanitizer pass (after tools/sanitize_gpu.sh came back clean and cheap): racecheck / initcheck over the per-kernel tests and
# the tiny-model engine tests, memcheck over the whole engine file except the full-size models.
export PYTORCH_NO_CUDA_MEMORY_CACHING=1 OA_SKIP_SLOW_PARITY=1
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 86 --print-limit 30"
run() {
  local name=$1 secs=$2 tool=$3; shift 3
  timeout $secs $S --tool $tool --log-file gpurun_out/sanitize_${name}.log "$@" > gpurun_out/sanitize_${name}.out 2>&1
  echo "$name rc=$? | $(grep -E 'SUMMARY' gpurun_out/sanitize_${name}.log | tail -1) | $(tail -1 gpurun_out/sanitize_${name}.out)"
}
ENG='not full_size and not 1b and not llama_3_8b and not safetensors and not malformed'
run racecheck_kernels 400 racecheck python -m pytest tests/test_kernels_gpu.py -x -q
run initcheck_kernels 300 initcheck python -m pytest tests/test_kernels_gpu.py -x -q
run memcheck_engine_all 500 memcheck python -m pytest tests/test_engine_gpu.py -x -q -k "$ENG"
run racecheck_engine 400 racecheck python -m pytest tests/test_engine_gpu.py -x -q -k "greedy_generation or concurrent_requests or grammar_constrained or prefix_cache or long_context or cluster_splitk and not 1b"
run initcheck_engine 300 initcheck python -m pytest tests/test_engine_gpu.py -x -q -k "greedy_generation or concurrent_requests or grammar_constrained or prefix_cache or long_context"
