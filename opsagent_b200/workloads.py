"""workloads — the reference's request shapes, rebuilt byte for byte, for bench.py and the tests.

Each builder returns the `messages` the reference hands to the Chat seam for one of its entry points, using the reference's system
prompts VERBATIM (fixtures under tests/golden/prompts/, extracted by tests/golden/prompts/extract_prompts.py with file:line provenance):

  execute_messages   POST /api/execute            pkg/handlers/execute.go:172-199   system = executeSystemPrompt_cn, user = cleaned instructions
  diagnose_messages  `kube-copilot diagnose`      cmd/kube-copilot/diagnose.go:109-118
  analyze_messages   swarm analysis flow          pkg/workflows/analyze.go:47-66    (SimpleFlow{System, Steps[0].Instructions, Inputs{k8s_manifest}})
  audit_messages     swarm audit flow             pkg/workflows/audit.go:58-80      (SimpleFlow{System, Instructions, Inputs{pod, namespace}})

swarm-go v0.2.1 is not vendored in the reference tree (go.mod:10), so the exact rendering of a SimpleFlow step is restated as: system
message = `System`, user message = `Instructions` followed by the step inputs as "key: value" lines — the fields and their order are the
reference's, the joining punctuation is ours.

Synthetic payloads (Pod manifests, kubectl tables, trivy reports) are seeded and process-independent; `fit_to_tokens` pads or trims a
payload so that a request has an exact prompt-token count under the ENGINE's tokenizer (BASELINE configs quote P per class)."""
from __future__ import annotations

import os
import random

from .llms import ChatCompletionMessage

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROMPTS_DIR = os.environ.get("OA_PROMPTS_DIR", os.path.join(_ROOT, "tests", "golden", "prompts"))
_cache: dict = {}


def prompt(name: str) -> str:
    """the reference constant `name` (e.g. "executeSystemPrompt_cn"), verbatim"""
    if name not in _cache:
        with open(os.path.join(PROMPTS_DIR, name + ".txt"), "rb") as f:
            _cache[name] = f.read().decode("utf-8")
    return _cache[name]


def execute_messages(instructions: str, args: str = "") -> list:
    """pkg/handlers/execute.go:172-199: args appended unless already contained, the leading word "execute" trimmed, then TrimSpace"""
    ins = instructions
    if args != "" and args not in ins:
        ins = f"{instructions} {args}"
    if ins.startswith("execute"):
        ins = ins[len("execute"):]
    ins = ins.strip()
    return [ChatCompletionMessage("system", prompt("executeSystemPrompt_cn")), ChatCompletionMessage("user", ins)]


def diagnose_messages(name: str, namespace: str = "default") -> list:
    """cmd/kube-copilot/diagnose.go:109-118"""
    return [ChatCompletionMessage("system", prompt("diagnoseSystemPrompt")),
            ChatCompletionMessage("user", "Your goal is to ensure that both the issues and their solutions are communicated effectively and understandably. "
                                          f"As you diagnose issues for Pod {name} in namespace {namespace}, remember to avoid using any delete or edit commands.")]


def analyze_messages(manifest: str) -> list:
    """pkg/workflows/analyze.go:47-66: System line, analysisPrompt as the step instructions, input k8s_manifest"""
    return [ChatCompletionMessage("system", prompt("analysisSystem")),
            ChatCompletionMessage("user", prompt("analysisPrompt") + "\nk8s_manifest: " + manifest)]


def audit_messages(namespace: str, name: str) -> list:
    """pkg/workflows/audit.go:58-80: System line, auditPrompt as the step instructions, inputs pod_namespace / pod_name"""
    return [ChatCompletionMessage("system", prompt("auditSystem")),
            ChatCompletionMessage("user", prompt("auditPrompt") + f"\npod_namespace: {namespace}\npod_name: {name}")]


# ---- seeded synthetic payloads ---------------------------------------------------------------------------------------------------
def synthetic_pod_yaml(i: int, n_bytes: int, seed: int = 42) -> str:
    """Seeded synthetic Pod manifest of at least n_bytes ASCII bytes (1-3+ containers, env, probes, resources, status.conditions)"""
    r = random.Random(seed * 1000003 + i)
    parts = [f"apiVersion: v1\nkind: Pod\nmetadata:\n  name: app-{i:04d}-{r.randrange(16**6):06x}\n  namespace: ns-{r.randrange(40)}\n"
             f"  labels:\n    app: svc-{r.randrange(200)}\n    tier: {r.choice(['web', 'db', 'cache', 'batch'])}\nspec:\n  containers:\n"]
    while sum(map(len, parts)) < n_bytes:
        c = r.randrange(1000)
        parts.append(f"  - name: c{c}\n    image: registry.local/team{r.randrange(30)}/img{c}:{r.randrange(9)}.{r.randrange(20)}.{r.randrange(50)}\n"
                     f"    resources:\n      requests: {{cpu: {r.randrange(50, 2000)}m, memory: {r.randrange(64, 4096)}Mi}}\n"
                     f"      limits: {{cpu: {r.randrange(100, 4000)}m, memory: {r.randrange(128, 8192)}Mi}}\n"
                     f"    env:\n    - name: VAR_{r.randrange(100)}\n      value: \"{r.randrange(10**8)}\"\n"
                     f"    livenessProbe: {{httpGet: {{path: /healthz, port: {r.randrange(1024, 9999)}}}, periodSeconds: {r.randrange(5, 60)}}}\n"
                     f"status:\n  phase: {r.choice(['Running', 'Pending', 'CrashLoopBackOff', 'Failed'])}\n  conditions:\n"
                     f"  - type: Ready\n    status: \"{r.choice(['True', 'False'])}\"\n    reason: {r.choice(['ContainersNotReady', 'PodCompleted', 'Unschedulable', 'OK'])}\n")
    return "".join(parts)


EXECUTE_QUESTIONS = ["how many namespace in the cluster?", "查询 default 命名空间下所有 pod 的镜像版本", "which pods are in CrashLoopBackOff and why?",
                     "列出 kube-system 中重启次数最多的 5 个 pod", "show the nodes with the highest memory pressure", "查看 ingress-nginx 的 service 暴露了哪些端口"]


def fit_to_tokens(build, payload: str, target: int, count_tokens, filler: str = "\n# pad") -> list:
    """messages = build(payload[:k] + filler...) with exactly `target` prompt tokens under count_tokens (the engine's chat template +
    tokenizer).  Binary-searches the payload length, then tops up with filler characters one at a time."""
    def n_tok(text):
        return count_tokens([(m.Role, m.Content) for m in build(text)])
    if n_tok("") > target:
        raise ValueError(f"the fixed part of the request already has {n_tok('')} tokens, more than the target {target}")
    lo, hi = 0, len(payload)
    if n_tok(payload) < target:
        lo = hi
    else:
        while lo < hi:                               # largest prefix with <= target tokens
            mid = (lo + hi + 1) // 2
            if n_tok(payload[:mid]) <= target:
                lo = mid
            else:
                hi = mid - 1
    text = payload[:lo]
    n, k = n_tok(text), 0
    while n < target:                                # top up: every accepted character adds 0 or 1 token, never overshoots
        for ch in (filler[k % len(filler)], "x", " ", ".", "\n"):
            m = n_tok(text + ch)
            if n <= m <= min(target, n + 1):
                text += ch; n = m
                break
        else:
            raise ValueError("cannot reach the target token count exactly")
        k += 1
        if k > 16 * target:
            raise ValueError("cannot reach the target token count")
    return build(text)


def mixed_request(i: int, count_tokens, p_analyze: int = 1536, p_diagnose: int = 0, p_execute: int = 0, seed: int = 42):
    """request i of BASELINE configs[2]'s mix: 40 % analyze / 30 % diagnose / 30 % execute.  -> (kind, messages).
    analyze requests are padded with Pod YAML to p_analyze tokens; diagnose / execute carry no bulk payload in the reference, so their
    prompt length is whatever the verbatim prompt + the question tokenises to (p_* = 0), or padded with a trailing comment if p_* > 0."""
    r = random.Random(seed * 7919 + i)
    slot = i % 10
    if slot < 4:
        return "analyze", fit_to_tokens(analyze_messages, synthetic_pod_yaml(i, 12 * p_analyze, seed), p_analyze, count_tokens)
    if slot < 7:
        pod = f"web-{r.randrange(1000)}"
        build = lambda extra: diagnose_messages(pod + extra, f"ns-{i % 40}")      # noqa: E731
        return "diagnose", (fit_to_tokens(build, "", p_diagnose, count_tokens) if p_diagnose else build(""))
    q = EXECUTE_QUESTIONS[i % len(EXECUTE_QUESTIONS)]
    build = lambda extra: execute_messages("execute " + q + extra)                                       # noqa: E731
    return "execute", (fit_to_tokens(build, "", p_execute, count_tokens) if p_execute else build(""))
