# round-2 ncu evidence (one GPU).  Numbers printed by runs under ncu are never bench values.
set -x
OA_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-react > gpurun_out/bench_under_ncu_r2.log 2>&1
OA_CUDA_PROFILER=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"decode_attention|gemm_streamk|sk_resid|sk_rope|gemm_tcgen05" -c 12 \
    -f -o gpurun_out/prof_r2_decode python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/prof_r2_decode.log 2>&1
OA_CUDA_PROFILER_PREFILL=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"prefill_attention|gemm_persistent" -c 6 \
    -f -o gpurun_out/prof_r2_prefill python tools/profile_step.py --steps 1 --warmup 1 > gpurun_out/prof_r2_prefill.log 2>&1
OA_CUDA_PROFILER_PREFILL=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"prefill_attention" -c 3 \
    -f -o gpurun_out/prof_r2_prefill16k python tools/profile_step.py --batch 4 --ctx 16384 --steps 1 --warmup 1 --extra '{"max_seq_len":16640}' > gpurun_out/prof_r2_prefill16k.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2_bench.csv
