"""ReAct-loop benchmark (BASELINE configs[1]/[2] style, multi-step): N concurrent agent conversations, each driven by the mirror of the
reference's loop (opsagent_b200.assistants.AssistantWithConfig <-> pkg/assistants/simple.go:292) through LocalCUDAClient.Chat, with the
engine in json_mode (grammar-forced tools.ToolPrompt output) so that every reply parses and the loop really iterates:
`tool_steps` tool calls answered by a synthetic kubectl, then a final answer.  Reports ReAct steps/s = completed Chat calls per second.
    python tools/bench_react.py --agents 128 --tool-steps 3
"""
import argparse
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opsagent_b200 import Engine, LocalCUDAClient, ChatCompletionMessage  # noqa: E402
from opsagent_b200.assistants import AssistantWithConfig  # noqa: E402
from opsagent_b200.synthetic import copilot_tools  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama-3-8b")
ap.add_argument("--agents", type=int, default=128)
ap.add_argument("--tool-steps", type=int, default=3)
ap.add_argument("--prompt-bytes", type=int, default=1200)
ap.add_argument("--extra", default="{}", help="engine config overrides (JSON), e.g. {\"prefill_batch_tokens\": 0}")
a = ap.parse_args()

from opsagent_b200 import workloads as WL  # noqa: E402

eng = Engine({"model": a.model, "kv_gb": 60, "max_batch": a.agents, "max_seq_len": 8192, "max_step_tokens": 8192, "json_mode": 1,
              "react_tool_steps": a.tool_steps, **json.loads(a.extra)})
calls = [0] * a.agents
results = [None] * a.agents


class CountingClient(LocalCUDAClient):
    def __init__(self, engine, slot):
        super().__init__(engine); self.slot = slot

    def Chat(self, model, maxTokens, prompts):
        calls[self.slot] += 1
        return super().Chat(model, maxTokens, prompts)


def agent(i):
    msgs = WL.analyze_messages(WL.synthetic_pod_yaml(i, a.prompt_bytes)[: a.prompt_bytes])
    results[i] = AssistantWithConfig(a.model, msgs, 2048, True, False, a.tool_steps + 2, CountingClient(eng, i), copilot_tools(i),
                                     count_tokens=eng.count_tokens)


def round_():
    for i in range(a.agents):
        calls[i] = 0
    th = [threading.Thread(target=agent, args=(i,)) for i in range(a.agents)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    return time.perf_counter() - t0


round_()                      # warm-up
s0 = eng.stats()
dt = round_()
s1 = eng.stats()
n_calls = sum(calls)
ok = sum(1 for r in results if r and len(r[0]) >= 10)
print(json.dumps({"workload": f"{a.agents} concurrent ReAct conversations, {a.tool_steps} kubectl tool steps + final answer, {a.model}, json_mode",
                  "react_steps_per_sec": round(n_calls / dt, 2), "chat_calls": n_calls, "seconds": round(dt, 2), "conversations_with_final_answer": ok,
                  "completion_tokens_per_sec": round((s1["decode_tokens"] - s0["decode_tokens"]) / dt, 1),
                  "prefill_tokens": s1["prefill_tokens"] - s0["prefill_tokens"], "decode_steps": s1["decode_steps"] - s0["decode_steps"],
                  "preemptions": s1["preemptions"] - s0["preemptions"], "prefill_steps": s1["prefill_steps"] - s0["prefill_steps"],
                  "admissions_deferred": s1.get("admissions_deferred", 0) - s0.get("admissions_deferred", 0), "extra": json.loads(a.extra)}), flush=True)
eng.close()
