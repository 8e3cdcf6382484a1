OA_SKIP_SLOW_PARITY=1 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "cluster_splitk" > gpurun_out/pytest_ck.log 2>&1; tail -15 gpurun_out/pytest_ck.log
OA_SWEEP_KT=1 python tools/sweep_decode.py "{}" '{"sk_clusterk":1}' '{"sk_clusterk":2}' '{"sk_clusterk":4}' '{"sk_clusterk":7}' > gpurun_out/sweep_ck.log 2>&1; cat gpurun_out/sweep_ck.log
python tools/gemm_shape_sweep.py 256 > gpurun_out/gemm_shapes_m256.log 2>&1; cat gpurun_out/gemm_shapes_m256.log
for x in '{}' '{"mixed_steps":1}'; do python bench.py --react-only --react-tool-ms 100 --engine-extra "$x" 2>/dev/null | tail -1 | cut -c1-900; done > gpurun_out/react_mixed_ab.log; cat gpurun_out/react_mixed_ab.log
