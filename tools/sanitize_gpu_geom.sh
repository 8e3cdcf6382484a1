#!/bin/bash
# memcheck at the benchmarked geometry: the 8B head layout at B=128 (decode attention work plan, paged KV) and the cluster split-K projections at 1B scale
export PYTORCH_NO_CUDA_MEMORY_CACHING=1 OA_SKIP_SLOW_PARITY=1
mkdir -p gpurun_out
timeout 400 compute-sanitizer --error-exitcode 86 --print-limit 30 --tool memcheck --log-file gpurun_out/sanitize_memcheck_geom.log \
  python -m pytest tests/test_engine_gpu.py -q -k "benchmarked_batch_geometry or (cluster_splitk and 1b)" > gpurun_out/sanitize_memcheck_geom.out 2>&1
echo "memcheck_geom rc=$? | $(grep SUMMARY gpurun_out/sanitize_memcheck_geom.log | tail -1) | $(tail -1 gpurun_out/sanitize_memcheck_geom.out)"
