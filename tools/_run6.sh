export PYTORCH_NO_CUDA_MEMORY_CACHING=1 OA_SKIP_SLOW_PARITY=1
ENG='not full_size and not 1b and not llama_3_8b and not safetensors and not malformed'
timeout 600 compute-sanitizer --error-exitcode 86 --print-limit 30 --tool memcheck --log-file gpurun_out/sanitize_memcheck_engine_all2.log python -m pytest tests/test_engine_gpu.py -q -k "$ENG" > gpurun_out/sanitize_memcheck_engine_all2.out 2>&1
echo "memcheck_engine_all2 rc=$? | $(grep SUMMARY gpurun_out/sanitize_memcheck_engine_all2.log | tail -1) | $(tail -1 gpurun_out/sanitize_memcheck_engine_all2.out)"
unset PYTORCH_NO_CUDA_MEMORY_CACHING OA_SKIP_SLOW_PARITY
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/gpu_suite_r02j.log 2>&1; echo "suite rc=$? $(tail -1 gpurun_out/gpu_suite_r02j.log)"
