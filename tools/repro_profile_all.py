import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine
from oracle import oracle as O
name = sys.argv[1] if len(sys.argv) > 1 else "tiny-llama-d128"
extra = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
if name in O.PRESETS and name.startswith("tiny"):
    cfg = O.PRESETS[name].engine_json(num_pages=128, max_seq_len=512, max_batch=16, max_step_tokens=256)
    B, ctx = 16, 300
else:
    cfg = {"model": name, "kv_gb": 40, "max_batch": 128, "max_seq_len": 2048, "max_step_tokens": 8192}
    B, ctx = 128, 1644
cfg.update(extra)
eng = Engine(cfg)
print("call1", eng.bench_decode(B, ctx, 8, 2)["ms_per_step"], flush=True)
os.environ["OA_PROFILE_ALL"] = "1"
print("call2", eng.bench_decode(B, ctx + 14, 8, 2)["ms_per_step"], flush=True)
print(eng.kernel_times())
eng.close()
