"""Do two engines in ONE process (one per GPU, each with its own scheduler thread / stream) slow each other down?  Runs the BASELINE
configs[1] decode step on engine 0 alone, then on both engines concurrently from two host threads (the C ABI releases the GIL)."""
import json
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = {"model": "llama-3-8b", "kv_gb": 40, "max_batch": 128, "max_seq_len": 2048, "max_step_tokens": 8192}
engines = [Engine({**cfg, "device": d}) for d in range(n)]
alone = engines[0].bench_decode(128, 1640, 24, 4)
print(json.dumps({"mode": "engine 0 alone", "ms_per_step": round(alone["ms_per_step"], 3), "prefill_ms": round(alone["prefill_ms"], 1)}), flush=True)
res = [None] * n


def run(i):
    res[i] = engines[i].bench_decode(128, 1640, 24, 4)


th = [threading.Thread(target=run, args=(i,)) for i in range(n)]
[t.start() for t in th]; [t.join() for t in th]
print(json.dumps({"mode": f"{n} engines concurrently, one process", "ms_per_step": [round(r["ms_per_step"], 3) for r in res],
                  "prefill_ms": [round(r["prefill_ms"], 1) for r in res]}), flush=True)
[e.close() for e in engines]
