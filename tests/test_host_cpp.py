"""Builds and runs the C++ twin of the Go provider (opsagent_b200/host/localcuda_client.hpp) against the real
shared library on the CPU: without a GPU the engine cannot be created, which exercises NewOpenAIClient's key check
and Chat's 500 -> backoff -> "throttled after retrying 5 times" path exactly as openai.go:77-103 specifies."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cstdio>
#include <vector>
#include "opsagent_b200/host/localcuda_client.hpp"
using namespace opsagent;
int main() {
    LocalCUDAClient c; Error e = LocalCUDAClient::New("", nullptr, &c);
    if (e.Message != "OPENAI_API_KEY is not set") { std::printf("FAIL key check\n"); return 1; }
    oa_engine* eng = nullptr;
    int rc = oa_engine_create("{\"model\": \"llama-3-8b\"}", &eng);
    std::printf("create rc=%d\n", rc);
    if (rc == 0) { std::printf("GPU present: skipping the failure path\n"); oa_engine_destroy(eng); return 0; }
    e = LocalCUDAClient::New("sk-local", eng, &c);
    if (!e.ok()) return 2;
    std::vector<long> sleeps; c.Sleep = [&](std::chrono::milliseconds d) { sleeps.push_back((long)d.count()); };
    Error err; std::string out = c.Chat("llama-3-8b", 8192, {{"system", "s"}, {"user", "how many namespace in the cluster?"}}, &err);
    std::printf("out='%s' err='%s' sleeps=%zu:", out.c_str(), err.Message.c_str(), sleeps.size());
    for (long s : sleeps) std::printf(" %ld", s);
    std::printf("\n");
    bool ok = out.empty() && err.Message == "OpenAI request throttled after retrying 5 times" && sleeps.size() == 5 &&
              sleeps[0] == 1000 && sleeps[1] == 2000 && sleeps[2] == 4000 && sleeps[3] == 8000 && sleeps[4] == 16000;
    return ok ? 0 : 3;
}
'''


SRC_GPU = r'''
#include <cstdio>
#include <vector>
#include "opsagent_b200/host/localcuda_client.hpp"
using namespace opsagent;
int main(int argc, char** argv) {
    oa_engine* eng = nullptr;
    if (argc < 2 || oa_engine_create(argv[1], &eng) != 0) { std::printf("create failed: %s\n", oa_last_error()); return 1; }
    LocalCUDAClient c; Error e = LocalCUDAClient::New("sk-local", eng, &c);
    if (!e.ok()) return 2;
    Error err; std::string out = c.Chat("", 24, {{"system", "You are a Kubernetes expert."}, {"user", "how many namespace in the cluster?"}}, &err);
    if (!err.ok()) { std::printf("chat failed: %d %s\n", err.HTTPStatusCode, err.Message.c_str()); return 3; }
    std::printf("hex:");
    for (unsigned char ch : out) std::printf("%02x", ch);
    std::printf("\n");
    // a model the engine does not serve: 400, returned at once, no retries (openai.go:95-97)
    int sleeps = 0; c.Sleep = [&](std::chrono::milliseconds) { ++sleeps; };
    out = c.Chat("gpt-4", 8, {{"user", "hi"}}, &err);
    if (err.HTTPStatusCode != 400 || sleeps != 0) { std::printf("expected a fast 400, got %d after %d sleeps\n", err.HTTPStatusCode, sleeps); return 4; }
    oa_engine_destroy(eng);
    return 0;
}
'''


@pytest.mark.gpu
def test_cpp_client_chat_on_the_gpu_matches_the_oracle(tmp_path):
    """the compiled C++ twin of the Go provider driving the engine for real: Chat() returns the bytes the oracle's greedy decode renders"""
    import json
    import numpy as np
    from oracle import oracle as O          # checker only
    spec = O.PRESETS["tiny-llama"]
    src = tmp_path / "g.cpp"; src.write_text(SRC_GPU)
    exe = tmp_path / "g"
    lib = os.path.join(ROOT, "opsagent_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}", str(src), "-o", str(exe), f"-L{lib}", "-lopsagent_b200",
                    f"-Wl,-rpath,{lib}", "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"], check=True)
    r = subprocess.run([str(exe), json.dumps(spec.engine_json(num_pages=32, max_seq_len=512, max_batch=4, max_step_tokens=256))],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    got = bytes.fromhex(r.stdout.split("hex:")[1].split()[0]) if r.stdout.split("hex:")[1].strip() else b""
    ids = O.apply_chat_template(spec, [("system", "You are a Kubernetes expert."), ("user", "how many namespace in the cluster?")])
    orc = O.Oracle(spec, max_pos=512, mode=1)
    ref, margins, _ = orc.generate(np.array(ids, np.int32), 24, eos=O.eos_ids(spec))
    orc.close()
    want = O.detokenize(ref)                  # the oracle stops BEFORE an EOS token, as the engine's content does
    k = 0
    while k < min(len(got), len(want)) and got[k] == want[k]:
        k += 1
    assert (got == want) or (k < len(margins) and margins[k] <= 5e-2), (k, got, want)


def test_cpp_client_semantics(tmp_path):
    src = tmp_path / "t.cpp"; src.write_text(SRC)
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "opsagent_b200", "lib")
    assert os.path.exists(os.path.join(lib, "libopsagent_b200.so")), "build the library first (python __graft_entry__.py)"
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}", str(src), "-o", str(exe), f"-L{lib}", "-lopsagent_b200",
                    f"-Wl,-rpath,{lib}", "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr


# ---- the C++ twin of the ReAct loop (host/assistants.hpp) against the Python mirror (assistants.py) on scripted scenarios ----------------
def _scenarios(n, seed):
    import json
    import random
    r = random.Random(seed)
    words = ["pods", "<none>", "a&b", "中文 节点", "CrashLoopBackOff", "tab\there", "quote\"q", "back\\slash", "bell\b\f", " sep", "emoji 😀", "x" * 40, "", "line1\nline2"]

    def text(k=3):
        return " ".join(r.choice(words) for _ in range(r.randrange(1, k + 1)))

    def reply():
        kind = r.randrange(12)
        if kind == 0:
            return "plain text, not JSON at all " + text()
        if kind == 1:
            return "!ERR:backend said no"
        if kind == 2:
            return json.dumps({"question": 5})                                # wrong type -> Unmarshal error
        if kind == 3:
            return json.dumps([1, 2])
        tp = {"question": text(), "thought": text(), "action": {"name": r.choice(["kubectl", "trivy", "nosuchtool", "", "failing"]), "input": text()},
              "observation": r.choice(["", text()]), "final_answer": r.choice(["", "", "<final_answer placeholder>", "short", "The deployment has " + text(4)])}
        if kind == 4:
            tp["action"] = None
        if kind == 5:
            tp = {("Question" if k == "question" else k.upper() if k == "thought" else k): v for k, v in tp.items()}      # case-insensitive keys
        if kind == 6:
            tp["extra"] = {"nested": [1, 2, {"a": None}]}; del tp["thought"]
        if kind == 7:
            tp["action"] = "kubectl"                                          # non-object action -> error
        if kind == 8:
            tp["final_answer"] = None
        return json.dumps(tp, ensure_ascii=r.random() < 0.5)
    out = []
    for _ in range(n):
        long_table = "\n".join(f"pod-{i} 1/1 Running {text(2)}" for i in range(r.randrange(1, 400)))
        out.append({"prompts": [["system", "You are " + text()], ["user", text()]] if r.random() < 0.95 else [],
                    "replies": [reply() for _ in range(r.randrange(0, 9))],
                    "tools": {"kubectl": {"outputs": ["  " + long_table + " \n", text(), ""]}, "trivy": {"outputs": [text()]}, "failing": {"error": "exit status 1: " + text()}},
                    "maxIterations": r.choice([0, 1, 2, 5, 7]), "count_div": r.choice([0, 1, 3, 4])})
    return out


def _python_mirror(sc):
    from opsagent_b200.assistants import AssistantWithConfig
    from opsagent_b200.llms import ChatCompletionMessage

    class Client:
        def __init__(self):
            self.i, self.calls = 0, 0

        def Chat(self, model, maxTokens, prompts):
            self.calls += 1
            if self.i >= len(sc["replies"]):
                raise RuntimeError("no more replies")
            rep = sc["replies"][self.i]; self.i += 1
            if rep.startswith("!ERR:"):
                raise RuntimeError(rep[5:])
            return rep
    cursor = {}

    def tool(name, spec):
        def run(inp):
            if "error" in spec:
                raise RuntimeError(spec["error"])
            k = cursor.get(name, 0); cursor[name] = k + 1
            return spec["outputs"][k % len(spec["outputs"])]
        return run
    tools = {k: tool(k, v) for k, v in sc["tools"].items()}
    div = sc["count_div"]
    count = (lambda ms: sum(4 + len(c.encode("utf-8")) // div for _, c in ms)) if div else None
    cl = Client()
    from opsagent_b200.perf import GetPerfStats
    GetPerfStats().Reset()
    try:
        res, hist = AssistantWithConfig("m", [ChatCompletionMessage(r_, c) for r_, c in sc["prompts"]], 256, False, False, sc["maxIterations"], cl, tools, count_tokens=count)
        return {"result": res, "error": "", "chat_calls": cl.calls, "history": [[m.Role, m.Content] for m in hist], "callCounts": GetPerfStats().GetStats()["callCounts"]}
    except Exception as e:      # noqa: BLE001
        return {"result": "", "error": str(e), "chat_calls": cl.calls, "history": None, "callCounts": GetPerfStats().GetStats()["callCounts"]}


def test_cpp_react_loop_equals_the_python_mirror_on_scripted_scenarios(tmp_path):
    """host/assistants.hpp (AssistantWithConfig, ToolPrompt Marshal / Unmarshal, isTemplateValue, ConstrictPrompt) against assistants.py / llms.py on
    400 seeded scenarios: same result, same error, same number of Chat calls and a byte-identical chat history (the history IS the next prompt)."""
    import json
    exe = tmp_path / "assistants_driver"
    b = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-I", ROOT, os.path.join(ROOT, "tests", "host", "assistants_driver.cpp"), "-o", str(exe),
                        "-L", os.path.join(ROOT, "opsagent_b200", "lib"), "-lopsagent_b200", f"-Wl,-rpath,{os.path.join(ROOT, 'opsagent_b200', 'lib')}"], capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-2000:]
    scs = _scenarios(400, seed=20260921)
    path = tmp_path / "scenarios.json"
    path.write_text(json.dumps(scs, ensure_ascii=False), encoding="utf-8")
    r = subprocess.run([str(exe), str(path)], capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.decode("utf-8").rstrip("\n").split("\n")          # not splitlines(): U+2028 inside a JSON string is not a line end
    assert len(lines) == len(scs)
    kinds = set()
    for sc, line in zip(scs, lines):
        got, want = json.loads(line), _python_mirror(sc)
        assert got["chat_calls"] == want["chat_calls"], (sc, got, want)
        assert got["callCounts"] == want["callCounts"], (got["callCounts"], want["callCounts"])      # the reference's operation names (simple.go:296-569), same counts
        if want["error"]:
            assert got["error"].endswith(want["error"].split("chat completion error: ")[-1]) and got["result"] == "", (got, want)
            kinds.add("error")
            continue
        assert got["error"] == "" and got["result"] == want["result"], (sc["replies"], got["result"], want["result"])
        assert got["history"] == want["history"]
        kinds.add("loop" if want["chat_calls"] > 1 else "first")
    assert kinds == {"error", "loop", "first"}


SRC_REACT_GPU = r'''
#include <cstdio>
#include "opsagent_b200/host/assistants.hpp"
using namespace opsagent;
int main(int argc, char** argv) {
    oa_engine* eng = nullptr;
    if (argc < 2 || oa_engine_create(argv[1], &eng) != 0) { std::printf("create failed: %s\n", oa_last_error()); return 1; }
    LocalCUDAClient c; if (!LocalCUDAClient::New("sk-local", eng, &c).ok()) return 2;
    ChatFn chat = [&](const std::string& model, int maxTokens, const std::vector<ChatCompletionMessage>& prompts, Error* err) { return c.Chat(model, maxTokens, prompts, err); };
    int calls = 0;
    std::map<std::string, Tool> tools;
    for (const char* name : {"kubectl", "python", "trivy", "jq", "search"})
        tools[name] = [&calls, name](const std::string& input, std::string*) { ++calls; return std::string("  NAME READY STATUS <none>\nweb-0 0/1 CrashLoopBackOff  # ") + name + " " + input + " \n"; };
    CountTokensFn count = [&](const std::vector<ChatCompletionMessage>& ms) {
        std::vector<oa_msg> m(ms.size()); for (size_t i = 0; i < ms.size(); ++i) { m[i].role = ms[i].Role.c_str(); m[i].content = ms[i].Content.c_str(); }
        int32_t n = 0; oa_count_tokens(eng, m.data(), (int32_t)m.size(), &n); return (int)n;
    };
    AssistantResult R = AssistantWithConfig("", {{"system", "You are a Kubernetes expert. Answer in JSON."}, {"user", "why is pod web-0 crashing?"}}, 600, false, false, 5, chat, tools, count);
    if (!R.Err.ok()) { std::printf("error: %s\n", R.Err.Message.c_str()); return 3; }
    std::string line = "{\"result\": " + oa::jstr(R.Result) + ", \"tool_calls\": " + std::to_string(calls) + ", \"history\": [";
    for (size_t i = 0; i < R.ChatHistory.size(); ++i) line += (i ? ", [" : "[") + oa::jstr(R.ChatHistory[i].Role) + ", " + oa::jstr(R.ChatHistory[i].Content) + "]";
    std::printf("%s]}\n", line.c_str());
    oa_engine_destroy(eng);
    return 0;
}
'''


@pytest.mark.gpu
def test_cpp_react_loop_drives_the_engine_like_the_python_mirror(tmp_path):
    """the compiled host side end to end: host/assistants.hpp + host/localcuda_client.hpp over the C ABI run a multi-step ReAct conversation on the GPU
    (json_mode: two grammar-forced tool calls, then a final answer); the Python mirror over the same engine configuration produces the same history"""
    import json
    from oracle import oracle as O          # presets only
    from opsagent_b200 import Engine, LocalCUDAClient
    from opsagent_b200.assistants import AssistantWithConfig
    from opsagent_b200.llms import ChatCompletionMessage
    spec = O.PRESETS["tiny-llama"]
    cfg = spec.engine_json(num_pages=160, max_seq_len=4096, max_batch=4, max_step_tokens=512, json_mode=1, react_tool_steps=2, prefix_cache=1)
    src = tmp_path / "react.cpp"; src.write_text(SRC_REACT_GPU)
    exe = tmp_path / "react"
    lib = os.path.join(ROOT, "opsagent_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", f"-I{ROOT}", str(src), "-o", str(exe), f"-L{lib}", "-lopsagent_b200", f"-Wl,-rpath,{lib}", "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"], check=True)
    r = subprocess.run([str(exe), json.dumps(cfg)], capture_output=True, timeout=240)
    assert r.returncode == 0, r.stdout[-600:] + r.stderr[-600:]
    got = json.loads(r.stdout.decode("utf-8").strip().split("\n")[-1])
    eng = Engine(cfg)
    calls = [0]

    def tool(name):
        def run(inp):
            calls[0] += 1
            return f"  NAME READY STATUS <none>\nweb-0 0/1 CrashLoopBackOff  # {name} {inp} \n"
        return run
    tools = {n: tool(n) for n in ("kubectl", "python", "trivy", "jq", "search")}
    res, hist = AssistantWithConfig("", [ChatCompletionMessage("system", "You are a Kubernetes expert. Answer in JSON."), ChatCompletionMessage("user", "why is pod web-0 crashing?")],
                                    600, False, False, 5, LocalCUDAClient(eng), tools, count_tokens=eng.count_tokens)
    eng.close()
    assert got["tool_calls"] == calls[0] == 2                      # react_tool_steps grammar-forced tool calls, then the final answer
    assert got["result"] == res and len(res.encode()) >= 10
    assert got["history"] == [[m.Role, m.Content] for m in hist]
    assert "\\u003cnone\\u003e" in got["history"][3][1]            # the observation went back Go-marshalled


def test_cpp_llms_helpers_on_the_reference_cases(tmp_path):
    """tests/host/llms_check.cpp: GetTokenLimits (reference-held cases, tokens_test.go:20-51), ConstrictMessages, ConstrictPrompt, TrimSpace, GoJSONString and
    isTemplateValue of host/assistants.hpp on the cases the Python mirrors are tested with, under ASAN + UBSAN"""
    exe = tmp_path / "llms_check"
    lib = os.path.join(ROOT, "opsagent_b200", "lib")
    b = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-I", ROOT, os.path.join(ROOT, "tests", "host", "llms_check.cpp"), "-o", str(exe),
                        "-L", lib, "-lopsagent_b200", f"-Wl,-rpath,{lib}"], capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.stdout, r.stderr[-1500:])


def test_cpp_toolprompt_unmarshal_equals_the_python_mirror_on_random_json(tmp_path):
    """ToolPrompt::Unmarshal + Marshal (host/assistants.hpp) against ToolPrompt.unmarshal / marshal (assistants.py) on 3000 seeded random documents: keys in
    random case, duplicate keys, null / number / list / object values where strings are expected, nested junk, broken JSON"""
    import json
    import random
    from opsagent_b200.assistants import ToolPrompt
    exe = tmp_path / "assistants_driver"
    lib = os.path.join(ROOT, "opsagent_b200", "lib")
    b = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-I", ROOT, os.path.join(ROOT, "tests", "host", "assistants_driver.cpp"), "-o", str(exe), "-L", lib, "-lopsagent_b200", f"-Wl,-rpath,{lib}"],
                       capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-2000:]
    r = random.Random(7)
    names = ["question", "thought", "action", "observation", "final_answer", "name", "input", "extra"]

    def key(k):
        return r.choice([k, k, k, k.upper(), k.capitalize(), k[:-1]])

    def val(depth=0):
        c = r.randrange(10)
        if c < 5:
            return r.choice(["", "x", "<none> &  ", "中文", "a\"b\\c", "tab\t\b\f", "😀"])
        if c == 5:
            return None
        if c == 6:
            return r.randrange(100)
        if c == 7:
            return [1, "a"]
        if c == 8 and depth < 2:
            return {key(n): val(depth + 1) for n in r.sample(names, r.randrange(0, 4))}
        return True

    texts = []
    for _ in range(3000):
        pairs = [(key(n), val()) for n in r.sample(names, r.randrange(0, 7))]
        if r.random() < 0.3:
            pairs += [(key("action"), {key("name"): val(1), key("input"): val(1)})]
        if r.random() < 0.2 and pairs:
            pairs.append((pairs[0][0], val()))                       # duplicate key: the last one wins
        body = "{" + ", ".join(json.dumps(k) + ": " + json.dumps(v, ensure_ascii=r.random() < 0.5) for k, v in pairs) + "}"
        if r.random() < 0.05:
            body = body[:r.randrange(len(body))]                      # broken JSON
        texts.append(body)
    path = tmp_path / "u.json"
    path.write_text(json.dumps([{"unmarshal": texts}], ensure_ascii=False), encoding="utf-8")
    out = subprocess.run([str(exe), str(path)], capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1500:]
    got = json.loads(out.stdout.decode("utf-8").rstrip("\n").split("\n")[0])
    n_ok = 0
    for text, g in zip(texts, got):
        try:
            want = ToolPrompt.unmarshal(text).marshal()
            n_ok += 1
        except Exception:      # noqa: BLE001
            want = None
        assert g == want, (text, g, want)
    assert 300 < n_ok < 2900          # both outcomes are well represented
