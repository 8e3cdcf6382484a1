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
