"""Open arrival vs closed batch on ONE engine: the 40/30/30 mix of BASELINE configs[2] (128 requests, max_tokens 256) submitted (a) at once through
the C ABI, (b) by 128 concurrent HTTP clients through the native front.  Prints seconds and the engine's step counters for each, per engine-option
variant given on the command line as JSON objects (default: the shipped options, then larger admission batches)."""
import http.client
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine  # noqa: E402
from opsagent_b200 import workloads as WL  # noqa: E402
from opsagent_b200.native_front import NativeFront  # noqa: E402

TOK = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bpe_k8s_8k.json")
N = 128
variants = [json.loads(a) for a in sys.argv[1:]] or [{}, {"prefill_batch_tokens": 8192, "prefill_max_wait_ms": 40}, {"mixed_steps": 0}]
sys.setswitchinterval(2e-4)
for extra in variants:
    eng = Engine({"model": "llama-3-8b", "kv_gb": 60, "max_batch": N, "max_seq_len": 2048, "max_step_tokens": 8192, "tokenizer": TOK, "prefix_cache": 0, "max_queue": 4096, **extra})
    reqs = [WL.mixed_request(i, eng.count_tokens, p_analyze=1536)[1] for i in range(N)]
    msgs = [[(m.Role, m.Content) for m in r] for r in reqs]
    bodies = [json.dumps({"model": "llama-3-8b", "max_tokens": 256, "messages": [{"role": r, "content": c} for r, c in m]}).encode() for m in msgs]
    front = NativeFront([eng], max_inflight=4 * N)

    def direct():
        t0 = time.perf_counter()
        tickets = [eng.chat_submit("llama-3-8b", m, 256, flags=0) for m in msgs]
        [eng.wait(t) for t in tickets]
        return time.perf_counter() - t0

    def post(i):
        c = http.client.HTTPConnection("127.0.0.1", front.port, timeout=600)
        c.request("POST", "/v1/chat/completions", body=bodies[i], headers={"Content-Type": "application/json", "Authorization": "Bearer x"})
        c.getresponse().read(); c.close()

    def served():
        go = threading.Barrier(N + 1)

        def w(i):
            go.wait(); post(i)
        th = [threading.Thread(target=w, args=(i,)) for i in range(N)]
        [t.start() for t in th]; go.wait(); t0 = time.perf_counter(); [t.join() for t in th]
        return time.perf_counter() - t0

    direct()
    for name, fn in (("closed batch (submit all, wait all)", direct), ("128 HTTP clients -> native front", served), ("128 HTTP clients -> native front (again)", served)):
        s0 = eng.stats(); dt = fn(); s1 = eng.stats()
        d = {k: round(s1[k] - s0[k], 1) for k in ("prefill_steps", "decode_steps", "busy_ms", "prefill_tokens", "decode_tokens") if k in s0}
        print(json.dumps({"options": extra, "path": name, "seconds": round(dt, 3), "tokens_per_sec": round(N * 256 / dt, 1), **d}), flush=True)
    front.shutdown(); eng.close()
