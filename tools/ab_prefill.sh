OA_SKIP_SLOW_PARITY=1 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_wg.log 2>&1; tail -4 gpurun_out/pytest_wg.log
for w in 2 1; do OA_PREFILL_WG=$w python tools/profile_step.py --steps 4 --warmup 2 > gpurun_out/pf_$w.log 2>&1; echo "WG=$w 8B B=128 ctx1664: $(tail -1 gpurun_out/pf_$w.log | cut -c1-200)"; done
for w in 2 1; do OA_PREFILL_WG=$w python tools/profile_step.py --model llama-3-8b --batch 8 --ctx 16384 --steps 2 --warmup 1 --extra '{"max_seq_len":16640}' > gpurun_out/pf16k_$w.log 2>&1; echo "WG=$w 8B B=8 ctx16k: $(tail -1 gpurun_out/pf16k_$w.log | cut -c1-200)"; done
