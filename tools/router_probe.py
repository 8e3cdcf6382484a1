"""Where does the time go when requests come through the Python HTTP front?  One engine, 128 analyze requests (P=1536, G=256):
(a) direct submit/wait through the C ABI, (b) through Router.chat_complete from 128 threads, (c) through the HTTP front."""
import http.client
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine  # noqa: E402
from opsagent_b200 import workloads as WL  # noqa: E402
from opsagent_b200.http_front import serve  # noqa: E402
from opsagent_b200.router import Router  # noqa: E402

TOK = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bpe_k8s_8k.json")
N = 128
eng = Engine({"model": "llama-3-8b", "kv_gb": 60, "max_batch": N, "max_seq_len": 2048, "max_step_tokens": 8192, "tokenizer": TOK, "prefix_cache": 0, "max_queue": 4096})
reqs = [WL.fit_to_tokens(WL.analyze_messages, WL.synthetic_pod_yaml(i, 12 * 1536), 1536, eng.count_tokens) for i in range(N)]
msgs = [[(m.Role, m.Content) for m in r] for r in reqs]


def delta(s0, s1):
    return {k: round(s1[k] - s0[k], 1) for k in ("prefill_steps", "decode_steps", "busy_ms", "prefill_tokens")}


def direct():
    t0 = time.perf_counter()
    tickets = [eng.chat_submit("", m, 256, flags=1) for m in msgs]
    [eng.wait(t) for t in tickets]
    return time.perf_counter() - t0


def threaded(fn):
    go = threading.Barrier(N + 1)

    def w(i):
        go.wait(); fn(i)
    th = [threading.Thread(target=w, args=(i,)) for i in range(N)]
    [t.start() for t in th]; go.wait(); t0 = time.perf_counter(); [t.join() for t in th]
    return time.perf_counter() - t0


rt = Router([eng], max_inflight=4 * N)
srv, _ = serve(rt, port=0)
port = srv.server_address[1]
bodies = [json.dumps({"model": "", "max_tokens": 256, "messages": [{"role": r, "content": c} for r, c in m]}).encode() for m in msgs]


def post(i):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=600)
    c.request("POST", "/v1/chat/completions", body=bodies[i], headers={"Content-Type": "application/json", "Authorization": "Bearer x"})
    c.getresponse().read(); c.close()


direct()
for name, fn in (("direct submit/wait", direct), ("128 threads -> Router.chat_complete", lambda: threaded(lambda i: rt.chat_complete("", msgs[i], 256, flags=1))),
                 ("128 threads -> HTTP front", lambda: threaded(post))):
    s0 = eng.stats(); dt = fn(); s1 = eng.stats()
    print(json.dumps({"path": name, "seconds": round(dt, 3), **delta(s0, s1)}), flush=True)
if len(sys.argv) > 1:
    sys.setswitchinterval(float(sys.argv[1]))
    s0 = eng.stats(); dt = threaded(post); s1 = eng.stats()
    print(json.dumps({"path": f"HTTP front, switchinterval {sys.argv[1]}", "seconds": round(dt, 3), **delta(s0, s1)}), flush=True)
srv.shutdown(); eng.close()
