"""Runs the BASELINE configs[1] decode step under several engine configs and prints ms/step for each (one engine per config)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine  # noqa: E402

configs = [json.loads(a) for a in sys.argv[1:]] or [{}]
for extra in configs:
    cfg = {"model": "llama-3-8b", "kv_gb": 40, "max_batch": 128, "max_seq_len": 2048, "max_step_tokens": 8192}
    cfg.update(extra)
    eng = Engine(cfg)
    r = eng.bench_decode(128, 1664 - 16 - 4, 32, 4)
    print(json.dumps({"extra": extra, "ms_per_step": round(r["ms_per_step"], 3), "tok_s": round(128e3 / r["ms_per_step"], 1)}), flush=True)
    if os.environ.get("OA_SWEEP_KT"):
        os.environ["OA_PROFILE_ALL"] = "1"
        eng.kernel_times(True)
        eng.bench_decode(128, 1664 - 4 - 2, 8, 2)
        os.environ["OA_PROFILE_ALL"] = "0"
        kt = eng.kernel_times(True)
        steps = 10
        print("   in-situ us/launch (launches/step): " + ", ".join(f"{k}={v[0] * 1e3 / max(v[1], 1):.1f}({v[1] // steps})" for k, v in kt.items() if v[1]), flush=True)
        print("   in-situ ms/step by class: " + ", ".join(f"{k}={v[0] / steps:.3f}" for k, v in kt.items() if v[1]), flush=True)
    eng.close()
