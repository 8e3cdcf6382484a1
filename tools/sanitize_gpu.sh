#!/bin/bash
# compute-sanitizer passes over the CUDA path.  Run on a GPU box from the repo root; logs go to gpurun_out/sanitize_*.log.
# Tiny presets only: the sanitizer serialises and instruments every launch (10-100x slower).
export PYTORCH_NO_CUDA_MEMORY_CACHING=1 OA_SKIP_SLOW_PARITY=1
mkdir -p gpurun_out
S="compute-sanitizer --error-exitcode 86 --print-limit 30"
run() {  # name, seconds, tool, command...
  local name=$1 secs=$2 tool=$3; shift 3
  timeout $secs $S --tool $tool --log-file gpurun_out/sanitize_${name}.log "$@" > gpurun_out/sanitize_${name}.out 2>&1
  echo "$name rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${name}.log) $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${name}.log | tail -1) | $(tail -1 gpurun_out/sanitize_${name}.out)"
}
run memcheck_kernels 420 memcheck python -m pytest tests/test_kernels_gpu.py -x -q
run memcheck_engine 420 memcheck python -m pytest tests/test_engine_gpu.py -x -q -k "greedy_generation or concurrent_requests or grammar_constrained or prefix_cache or preemption or native_front_routes or cluster_splitk and not 1b"
run racecheck_smoke 240 racecheck python -c "import __graft_entry__ as g; g.smoke()"
run synccheck_smoke 240 synccheck python -c "import __graft_entry__ as g; g.smoke()"
run initcheck_smoke 240 initcheck python -c "import __graft_entry__ as g; g.smoke()"
