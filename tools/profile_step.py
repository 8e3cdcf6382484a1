"""Runs the BASELINE configs[1] decode workload (Llama-3-8B, B=128, ctx~1664) for a few steps so that ncu can
capture exactly the timed decode steps:  OA_CUDA_PROFILER=1 ncu --profile-from-start off ... python tools/profile_step.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-3-8b")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--ctx", type=int, default=1664)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--extra", default="{}")
a = ap.parse_args()
cfg = {"model": a.model, "kv_gb": 60, "max_batch": a.batch, "max_seq_len": max(2048, (a.ctx + a.steps + a.warmup + 64) // 64 * 64 + 64),
       "max_step_tokens": 8192}
cfg.update(json.loads(a.extra))
eng = Engine(cfg)
r = eng.bench_decode(a.batch, a.ctx - a.steps // 2 - a.warmup, a.steps, a.warmup)
print(json.dumps(r))
eng.close()
