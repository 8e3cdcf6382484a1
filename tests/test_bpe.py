"""The C++ byte-level BPE tokenizer (opsagent_b200/csrc/bpe.hpp) against the Hugging Face `tokenizers` library itself:
tests/golden/gen_golden_bpe.py trained two small tokenizers in exactly the Llama-3 and Qwen2.5 tokenizer.json configuration and
recorded that library's encodings of the probe strings (contractions, digit grouping 1-3 vs 1, CJK / Greek / Cyrillic / Arabic,
emoji, whitespace runs with and without newlines, CRLF, JSON, YAML, empty input).  Bit-exact ids; decode round-trips."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from opsagent_b200 import _lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "bpe_cases.json")))


def encode(L, path, text):
    raw = text.encode("utf-8")
    out = np.zeros(max(16, 4 * len(raw) + 16), np.int32); n = C.c_int32()
    rc = L.oa_host_bpe_encode(path.encode(), raw, len(raw), out.ctypes.data, len(out), C.byref(n))
    assert rc == 0, _lib.last_error()
    return out[: n.value].tolist()


def decode(L, path, ids):
    arr = np.asarray(ids, np.int32); buf = C.create_string_buffer(max(16, 64 * len(ids) + 16)); n = C.c_int32()
    rc = L.oa_host_bpe_decode(path.encode(), arr.ctypes.data, len(arr), buf, len(buf), C.byref(n))
    assert rc == 0, _lib.last_error()
    return buf.raw[: n.value]


@pytest.mark.parametrize("name", ["llama3", "qwen2"])
def test_encode_matches_the_tokenizers_library_bit_for_bit(name):
    L = _lib.load()
    path = os.path.join(GOLDEN, f"bpe_{name}_tiny.json")
    for case in CASES[name]:
        got = encode(L, path, case["text"])
        assert got == case["ids"], (case["text"], got, case["ids"])
        assert decode(L, path, got) == case["text"].encode("utf-8")


def test_against_the_live_library_on_random_text():
    """beyond the committed probes: seeded random mixtures of scripts, digits, punctuation and whitespace, compared with the
    `tokenizers` library loaded here (skipped where it is not installed)"""
    tokenizers = pytest.importorskip("tokenizers")
    import random
    L = _lib.load()
    alphabet = (list("abcXYZ'stTreveLLdm ") + list("0123456789") + list("  \t\n\r") + list("!?.,;:{}[]()\"#$%&*+-/<=>@\\^_`|~") +
                list("集群命名空间节点") + list("éüïßñ") + list("λμπ") + list("жщю") + list("مرحب") + ["🙂", "→", "½", "①", " ", "　", " ", "ſ"])
    for name in ("llama3", "qwen2"):
        path = os.path.join(GOLDEN, f"bpe_{name}_tiny.json")
        ref = tokenizers.Tokenizer.from_file(path)
        r = random.Random(1234)
        for _ in range(300):
            text = "".join(r.choice(alphabet) for _ in range(r.randrange(0, 60)))
            assert encode(L, path, text) == ref.encode(text, add_special_tokens=False).ids, repr(text)


def test_control_tokens_and_error_paths(tmp_path):
    L = _lib.load()
    path = os.path.join(GOLDEN, "bpe_llama3_tiny.json")
    ids = [CASES["llama3_specials"]["<|begin_of_text|>"]] + encode(L, path, "hi") + [CASES["llama3_specials"]["<|eot_id|>"]]
    assert decode(L, path, ids) == b"<|begin_of_text|>hi<|eot_id|>"
    assert CASES["llama3_specials"]["<|eot_id|>"] not in encode(L, path, "<|eot_id|>")          # control tokens are never produced from text
    n = C.c_int32()
    assert L.oa_host_bpe_encode(str(tmp_path / "missing.json").encode(), b"x", 1, None, 0, C.byref(n)) == 400 and "not found" in _lib.last_error()
    bad = json.load(open(path)); bad["pre_tokenizer"]["pretokenizers"][0]["pattern"]["Regex"] = r"\w+|\s+"
    (tmp_path / "bad.json").write_text(json.dumps(bad))
    assert L.oa_host_bpe_encode(str(tmp_path / "bad.json").encode(), b"x", 1, None, 0, C.byref(n)) == 400 and "unsupported split pattern" in _lib.last_error()
    out = np.zeros(1, np.int32)
    assert L.oa_host_bpe_encode(path.encode(), b"hello world", 11, out.ctypes.data, 1, C.byref(n)) == 400 and n.value > 1      # size query


def test_chat_template_uses_the_bpe_vocabulary_when_configured():
    L = _lib.load()
    path = os.path.join(GOLDEN, "bpe_llama3_tiny.json")
    sp = CASES["llama3_specials"]
    cfg = json.dumps({"model": "custom", "hidden": 64, "n_layers": 1, "n_heads": 2, "n_kv_heads": 1, "head_dim": 64, "ffn": 128, "vocab": 704,
                      "template": "llama3", "tokenizer": path}).encode()
    msgs = (_lib.OaMsg * 2)(); msgs[0].role = b"system"; msgs[0].content = b"you are a kubectl expert"; msgs[1].role = b"user"; msgs[1].content = "集群 has 5 namespaces".encode()
    out = np.zeros(512, np.int32); n = C.c_int32()
    assert L.oa_host_apply_chat_template(cfg, msgs, 2, out.ctypes.data, len(out), C.byref(n)) == 0, _lib.last_error()
    ids = out[: n.value].tolist()
    expect = [sp["<|begin_of_text|>"]]
    for role, content in (("system", "you are a kubectl expert"), ("user", "集群 has 5 namespaces")):
        expect += [sp["<|start_header_id|>"]] + encode(L, path, role) + [sp["<|end_header_id|>"]] + encode(L, path, "\n\n") + encode(L, path, content) + [sp["<|eot_id|>"]]
    expect += [sp["<|start_header_id|>"]] + encode(L, path, "assistant") + [sp["<|end_header_id|>"]] + encode(L, path, "\n\n")
    assert ids == expect


def test_nfc_normaliser_is_exact_or_refused():
    """Qwen2.5's tokenizer.json normalises to NFC.  The C++ side does not implement normalisation: text that is already NFC passes
    (and must encode exactly as the library does), anything that could change under NFC is refused with 400 — never silently wrong."""
    tokenizers = pytest.importorskip("tokenizers")
    import random
    import unicodedata
    L = _lib.load()
    qwen = os.path.join(GOLDEN, "bpe_qwen2_tiny.json"); llama = os.path.join(GOLDEN, "bpe_llama3_tiny.json")
    ref_q = tokenizers.Tokenizer.from_file(qwen); ref_l = tokenizers.Tokenizer.from_file(llama)

    def try_encode(path, text):
        raw = text.encode("utf-8"); out = np.zeros(4 * len(raw) + 16, np.int32); n = C.c_int32()
        rc = L.oa_host_bpe_encode(path.encode(), raw, len(raw), out.ctypes.data, len(out), C.byref(n))
        return (out[: n.value].tolist() if rc == 0 else None), rc

    ids, rc = try_encode(qwen, "café")                       # e + combining acute: NFC would compose it
    assert ids is None and rc == 400 and "NFC" in _lib.last_error()
    assert try_encode(qwen, "café")[0] == ref_q.encode("café", add_special_tokens=False).ids
    assert try_encode(llama, "café")[0] == ref_l.encode("café", add_special_tokens=False).ids       # Llama-3: no normaliser
    pool = list("ab 1.\n") + ["é", "́", "̧", "ᄀ", "ᅡ", "ᆨ", "가", "Å", "क़", "া", "̈́", "中", "\U0001f642", "Ω", "̀"]
    r = random.Random(5)
    n_ref = n_ok = 0
    for _ in range(3000):
        text = "".join(r.choice(pool) for _ in range(r.randrange(1, 10)))
        ids, rc = try_encode(qwen, text)
        if ids is None:
            n_ref += 1
            assert rc == 400
        else:
            n_ok += 1
            assert unicodedata.is_normalized("NFC", text), repr(text)                 # accepted => really was NFC
            assert ids == ref_q.encode(text, add_special_tokens=False).ids, repr(text)
    assert n_ref > 100 and n_ok > 100


def test_legacy_merge_strings_and_rejections(tmp_path):
    """older tokenizer.json files write merges as "a b" strings; unsupported files are refused with a reason, not mis-tokenised"""
    L = _lib.load()
    src = json.load(open(os.path.join(GOLDEN, "bpe_llama3_tiny.json")))
    legacy = json.loads(json.dumps(src)); legacy["model"]["merges"] = [" ".join(m) for m in src["model"]["merges"]]
    (tmp_path / "legacy.json").write_text(json.dumps(legacy, ensure_ascii=True))      # \\uXXXX escapes exercise the JSON reader too
    for case in CASES["llama3"]:
        assert encode(L, str(tmp_path / "legacy.json"), case["text"]) == case["ids"]
    n = C.c_int32()
    for name, mutate, needle in (
            ("wordpiece.json", lambda j: j["model"].__setitem__("type", "WordPiece"), "not BPE"),
            ("nobytelevel.json", lambda j: j["pre_tokenizer"].__setitem__("pretokenizers", j["pre_tokenizer"]["pretokenizers"][:1]), "byte-level"),
            ("nfkc.json", lambda j: j.__setitem__("normalizer", {"type": "NFKC"}), "unsupported normalizer"),
            ("prefixspace.json", lambda j: j["pre_tokenizer"]["pretokenizers"][1].__setitem__("add_prefix_space", True), "not supported")):
        j = json.loads(json.dumps(src)); mutate(j)
        (tmp_path / name).write_text(json.dumps(j))
        assert L.oa_host_bpe_encode(str(tmp_path / name).encode(), b"x", 1, None, 0, C.byref(n)) == 400 and needle in _lib.last_error(), (name, _lib.last_error())
    (tmp_path / "broken.json").write_text("{\"model\": ")
    assert L.oa_host_bpe_encode(str(tmp_path / "broken.json").encode(), b"x", 1, None, 0, C.byref(n)) == 400 and "tokenizer.json" in _lib.last_error()


def test_k8s_tokenizer_and_long_runs_match_the_live_library():
    """the 8k-vocabulary tokenizer the benchmark's end-to-end leg uses (tests/golden/gen_k8s_bpe.py) on the reference's verbatim prompts,
    synthetic manifests and kubectl/trivy tables, plus long single-character runs that go through the O(n log n) merge path"""
    tokenizers = pytest.importorskip("tokenizers")
    from opsagent_b200.synthetic import copilot_tools
    from opsagent_b200.workloads import prompt, synthetic_pod_yaml
    L = _lib.load()
    path = os.path.join(GOLDEN, "bpe_k8s_8k.json")
    ref = tokenizers.Tokenizer.from_file(path)
    tools = copilot_tools(3)
    texts = [prompt(n) for n in ("executeSystemPrompt_cn", "diagnoseSystemPrompt", "analysisPrompt", "auditPrompt", "analysisSystem")]
    texts += [synthetic_pod_yaml(i, 3000) for i in range(3)] + [tools["kubectl"]("get pods -A"), tools["trivy"]("nginx:1.25")]
    texts += ["=" * 5000, "namespace" * 700, "ab" * 3000 + "kubectl" * 400, "集群命名空间" * 500, "-" * 200 + "\n" * 50 + " " * 300 + "x"]
    for t in texts:
        ids = encode(L, path, t)
        assert ids == ref.encode(t, add_special_tokens=False).ids, t[:40]
        assert decode(L, path, ids) == t.encode("utf-8")
