"""Tensor-parallel decode benchmark for BASELINE configs[3] (Qwen2.5-32B TP=4, B=256, ctx 2048) and configs[4]
(Llama-3-70B TP=8, B=64, ctx ~16.9k).  Run under torchrun with --nproc-per-node T; rank 0 prints one JSON line.
    python -m torch.distributed.run --nproc-per-node 4 tools/bench_tp.py --model qwen2.5-32b --batch 256 --ctx 2048
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="qwen2.5-32b")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--ctx", type=int, default=2048)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--step-tokens", type=int, default=8192)
ap.add_argument("--extra", default="{}")
a = ap.parse_args()
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
max_seq = (a.ctx + a.steps + a.warmup + 64 + 63) // 64 * 64
kvb = {"qwen2.5-32b": 262144, "llama-3-70b": 327680, "llama-3-8b": 131072}[a.model] // world
pages = (a.batch * ((max_seq + 63) // 64)) + 64
cfg = {"model": a.model, "device": rank, "tp": world, "tp_rank": rank, "tp_shm": f"/oa_tp_bench_{os.environ.get('MASTER_PORT', '0')}",
       "num_pages": pages, "max_batch": a.batch, "max_seq_len": max_seq, "max_step_tokens": a.step_tokens}
cfg.update(json.loads(a.extra))
eng = Engine(cfg)
if rank > 0:
    eng.serve(); eng.close(); sys.exit(0)
ctx0 = a.ctx - a.steps // 2 - a.warmup
r = eng.bench_decode(a.batch, ctx0, a.steps, a.warmup)
peak = 6576.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
bytes_gpu = r["algorithmic_bytes_per_step"]
line = {"workload": f"{a.model} TP={world} B={a.batch} mean ctx {r['mean_ctx']:.0f}", "ms_per_step": round(r["ms_per_step"], 3),
        "tokens_per_sec_group": round(a.batch / r["ms_per_step"] * 1e3, 1), "tokens_per_sec_per_gpu": round(a.batch / r["ms_per_step"] * 1e3 / world, 1),
        "algorithmic_bytes_per_gpu_per_step": bytes_gpu, "hbm_GBps_per_gpu": round(bytes_gpu / r["ms_per_step"] / 1e6, 1),
        "frac_of_measured_hbm_peak": round(bytes_gpu / r["ms_per_step"] / 1e6 / peak, 4), "prefill_ms": round(r["prefill_ms"], 1),
        "prefill_tokens_per_sec": round(a.batch * (ctx0 - 1) / r["prefill_ms"] * 1e3, 1), "launches_per_step": r["launches_per_step"],
        "kv_pages": pages, "n_gpus": world}
print(json.dumps(line), flush=True)
if os.environ.get("OA_SWEEP_KT"):      # in-situ per-class kernel times on the leader (events between launches: breaks PDL overlap, shows where the step goes)
    os.environ["OA_PROFILE_ALL"] = "1"
    eng.kernel_times(True)
    eng.bench_decode(a.batch, ctx0, 8, 2)
    kt = eng.kernel_times(True)
    print("in-situ us/launch (launches/step): " + ", ".join(f"{k}={v[0] * 1e3 / max(v[1], 1):.1f}({v[1] // 10})" for k, v in kt.items() if v[1]), flush=True)
    print("in-situ ms/step by class: " + ", ".join(f"{k}={v[0] / 10:.3f}" for k, v in kt.items() if v[1]), flush=True)
eng.close()
