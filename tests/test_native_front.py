"""The native HTTP front (csrc/http_server.cpp) at the wire level, with NO engine behind it (no GPU): routing of paths, auth, the error bodies the
reference's retry loop switches on (pkg/llms/openai.go:85-101), keep-alive, hostile clients.  The generation tests are in test_engine_gpu.py."""
import http.client
import json
import socket

import pytest

from opsagent_b200.native_front import NativeFront


@pytest.fixture()
def front():
    f = NativeFront([], require_key=True, api_key="sekret")
    yield f
    f.shutdown()


def _req(port, method, path, body=None, key="sekret", conn=None):
    c = conn or http.client.HTTPConnection("127.0.0.1", port, timeout=10)
    h = {"Content-Type": "application/json"}
    if key is not None:
        h["Authorization"] = "Bearer " + key
    c.request(method, path, body=None if body is None else (body if isinstance(body, bytes) else json.dumps(body).encode()), headers=h)
    r = c.getresponse()
    data = r.read()
    return r.status, json.loads(data), c


def test_paths_auth_and_error_bodies(front):
    p = front.port
    assert p > 0
    st, j, _ = _req(p, "POST", "/v1/chat/completions", {"messages": []}, key=None)
    assert st == 401 and j["error"]["type"] == "authentication_error"
    st, j, _ = _req(p, "POST", "/v1/chat/completions", {"messages": []}, key="wrong")
    assert st == 401
    st, j, _ = _req(p, "GET", "/v1/nothing")
    assert st == 404
    st, j, _ = _req(p, "POST", "/v1/chat/completions", b"{not json")
    assert st == 400 and j["error"]["type"] == "invalid_request_error" and j["error"]["code"] == 400
    st, j, _ = _req(p, "POST", "/v1/chat/completions", {"messages": "x"})
    assert st == 400
    st, j, _ = _req(p, "POST", "/v1/chat/completions", {"messages": [{"role": "user", "content": "hi"}], "stream": True})
    assert st == 400 and "stream" in j["error"]["message"]
    st, j, _ = _req(p, "POST", "/v1/chat/completions", {"messages": [{"role": "user", "content": "hi"}], "temperature": 0.7})
    assert st == 400
    # a well-formed request with no engine behind the front: 500, which Chat retries (openai.go:95-98)
    st, j, _ = _req(p, "POST", "/v1/chat/completions", {"model": "m", "messages": [{"role": "user", "content": "hi 中文 😀"}], "max_tokens": 8})
    assert st == 500 and j["error"]["type"] == "server_error"
    st, j, _ = _req(p, "GET", "/v1/models")
    assert st == 200 and j["object"] == "list"
    st, j, _ = _req(p, "GET", "/api/perf/stats")
    assert st == 200 and j["status"] == "success" and j["stats"]["front"]["replicas"] == 0


def test_perf_endpoints_have_the_reference_shape(front):
    """pkg/handlers/perf.go:12-39: {"stats": {"timers", "callCounts", "lastResetTime"}, "status": "success"}; reset answers with a message"""
    import datetime
    st, j, _ = _req(front.port, "GET", "/api/perf/stats", key=None)
    assert st == 401
    st, j, _ = _req(front.port, "GET", "/api/perf/stats")
    assert st == 200 and set(j["stats"]) >= {"timers", "callCounts", "lastResetTime", "front"}
    t0 = datetime.datetime.fromisoformat(j["stats"]["lastResetTime"].replace("Z", "+00:00"))
    st, j, _ = _req(front.port, "POST", "/api/perf/reset", {})
    assert st == 200 and j == {"message": "performance statistics reset", "status": "success"}
    st, j, _ = _req(front.port, "GET", "/api/perf/stats")
    assert datetime.datetime.fromisoformat(j["stats"]["lastResetTime"].replace("Z", "+00:00")) >= t0 and j["stats"]["timers"] == {}


def test_keep_alive_serves_many_requests_on_one_connection(front):
    conn = None
    for i in range(20):
        st, j, conn = _req(front.port, "POST", "/v1/chat/completions", b"[1, 2]" if i % 2 else b"{bad", conn=conn)
        assert st == 400
    assert front.stats()["requests"] >= 20


def test_hostile_clients_do_not_take_the_server_down(front):
    p = front.port
    s = socket.create_connection(("127.0.0.1", p)); s.sendall(b"GARBAGE\r\n\r\n"); s.settimeout(5)
    try:
        s.recv(4096)
    except OSError:
        pass
    s.close()
    s = socket.create_connection(("127.0.0.1", p)); s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nContent-Length: 100\r\n\r\nshort"); s.close()      # dies mid-body
    s = socket.create_connection(("127.0.0.1", p))
    s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nAuthorization: Bearer sekret\r\nContent-Length: 99999999999\r\n\r\n"); s.settimeout(5)
    data = s.recv(4096)
    assert b" 413 " in data or b" 400 " in data
    s.close()
    deep = b"[" * 100000
    st, j, _ = _req(p, "POST", "/v1/chat/completions", deep)
    assert st == 400
    st, j, _ = _req(p, "GET", "/v1/models")
    assert st == 200


def test_silent_connections_are_dropped_after_the_idle_timeout():
    import time
    f = NativeFront([], idle_timeout_s=1)
    try:
        s = socket.create_connection(("127.0.0.1", f.port)); s.settimeout(10)
        s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nContent-Length: 10\r\n\r\nabc")        # stalls mid-body
        t0 = time.time()
        assert s.recv(100) == b"" and 0.5 < time.time() - t0 < 5         # closed by the server, no reply owed
        s.close()
        st, j, _ = _req(f.port, "GET", "/v1/models")
        assert st == 200
    finally:
        f.shutdown()


def test_no_key_required_when_disabled():
    f = NativeFront([], require_key=False)
    try:
        st, j, _ = _req(f.port, "POST", "/v1/chat/completions", {"messages": [{"role": "user", "content": "x"}]}, key=None)
        assert st == 500          # past auth, no engine
    finally:
        f.shutdown()


def test_start_stop_repeatedly():
    for _ in range(5):
        f = NativeFront([])
        assert f.port > 0
        f.shutdown()


# ---- the front's JSON reader / writer against Python's json ----------------------------------------------------------------------------
import ctypes as C

from hypothesis import given, settings, strategies as st

from opsagent_b200 import _lib


def _roundtrip(raw: bytes):
    L = _lib.load()
    out = C.create_string_buffer(4 * len(raw) + 64)
    rc = L.oa_host_json_roundtrip(raw, len(raw), out, len(out))
    return rc, out.value


_json_values = st.recursive(st.none() | st.booleans() | st.integers(-2**50, 2**50) | st.floats(allow_nan=False, allow_infinity=False, width=64) | st.text(),
                            lambda c: st.lists(c, max_size=5) | st.dictionaries(st.text(max_size=8), c, max_size=5), max_leaves=25)


@settings(max_examples=300, deadline=None)
@given(_json_values, st.booleans())
def test_json_reader_and_writer_agree_with_python(value, ascii_escapes):
    raw = json.dumps(value, ensure_ascii=ascii_escapes).encode("utf-8", "surrogatepass")
    try:
        raw.decode("utf-8")
    except UnicodeDecodeError:
        return                      # lone surrogates cannot be written as UTF-8; covered by the \\u test below
    rc, out = _roundtrip(raw)
    assert rc == 0, out
    assert json.loads(out.decode("utf-8")) == json.loads(raw)


def test_json_reader_rejects_what_python_rejects():
    for bad in [b"", b"{", b"[1,]", b'{"a" 1}', b'{"a":1,}', b"tru", b'"unterminated', b'"bad \\x escape"', b"01", b"1 2", b'{"a":1} x', b"[" * 100, b'"\\ud800"x', b"1-2", b"1e", b"--1", b"+1", b".5", b"[1\x00]", b"1e999"]:
        rc, out = _roundtrip(bad)
        ok = True
        try:
            json.loads(bad)
        except Exception:
            ok = False
        assert (rc == 0) == ok or bad in (b"01", b"1e999"), (bad, rc, out)      # leading zeros: accepted by strtod-style readers, harmless; an overflowing number is refused (Python reads inf)
    rc, out = _roundtrip(b'"\\ud83d\\ude00 \\u00e9 \\u4e2d"')
    assert rc == 0 and json.loads(out.decode()) == "\U0001F600 é 中"


def test_json_writer_replaces_ill_formed_utf8_like_python():
    """a byte-level model can stop mid-character; the response must still be JSON a strict client decodes (http_front.py: decode('utf-8','replace'))"""
    for raw in [b"ok \xe4\xb8", b"\xff\xfe", b"a\xc0\xafb", b"\xed\xa0\x80", b"\xf0\x9f\x98", b"\xf4\x90\x80\x80", b"caf\xc3\xa9 \xe4\xb8\xad \xf0\x9f\x98\x80", b"\x80\x80abc\xe2\x82"]:
        rc, out = _roundtrip(b'"' + raw + b'"')
        assert rc == 0
        assert json.loads(out.decode("utf-8")) == raw.decode("utf-8", "replace"), raw


def test_serve_cli_fails_loudly_without_a_gpu_and_never_falls_back():
    """python -m opsagent_b200.serve on a box without CUDA: a non-zero exit naming the reason, not a CPU path"""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    r = subprocess.run([sys.executable, "-m", "opsagent_b200.serve", "--model", "llama-3.2-1b", "--port", "0"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_documented_front_options_are_the_ones_the_server_reads():
    """include/opsagent_b200.h documents oa_http_start's options_json; csrc/http_server.cpp reads them with opt.i("…") / opt.s("…") — same set"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "opsagent_b200.h")).read()
    doc = hdr[hdr.index("options_json (flat):"):]
    doc = doc[:doc.index("*/")]
    documented = set(re.findall(r'"([a-z_]+)":', doc))
    src = open(os.path.join(root, "opsagent_b200", "csrc", "http_server.cpp")).read()
    parsed = set(re.findall(r'opt\.[is]\("([a-z_]+)"', src))
    assert documented == parsed, (documented ^ parsed)
