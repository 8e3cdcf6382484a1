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
