"""ctypes front-end of the CPU oracle (oracle/llama_ref.c) plus pure-Python restatements of the
host-side string logic (synthetic byte-level tokenizer, chat templates).

TEST INFRASTRUCTURE ONLY — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import this module.  The product (opsagent_b200) never does.

PARITY UNPINNED BY THE REFERENCE (see llama_ref.c header): the reference's Chat seam
(reference pkg/llms/openai.go:69-104) carries no arithmetic; this oracle is pinned against
HF transformers fp32 via tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, asdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class RefConfig(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32), ("ffn", C.c_int32), ("vocab", C.c_int32),
        ("tie_embeddings", C.c_int32), ("qkv_bias", C.c_int32), ("rope_scaling", C.c_int32),
        ("norm_random", C.c_int32),
        ("rope_theta", C.c_float), ("rms_eps", C.c_float),
        ("rope_factor", C.c_float), ("rope_low_freq", C.c_float), ("rope_high_freq", C.c_float),
        ("rope_orig_ctx", C.c_int32),
        ("init_std", C.c_float),
        ("seed", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc + OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "llama_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def default_threads() -> int:
    """Threads the oracle may use: the CPUs this process is allowed on, capped by the cgroup CPU quota and by 64
    (beyond that the memory-bound mat-vec loops only add barrier contention: 128 threads measured 70x slower than 64)."""
    env = os.environ.get("OA_ORACLE_THREADS")
    if env:
        return max(1, int(env))
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        pass
    return max(1, min(n, 64))


def lib():
    global _LIB
    if _LIB is None:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        L = C.CDLL(build())
        L.oa_ref_create.restype = C.c_void_p
        L.oa_ref_create.argtypes = [C.POINTER(RefConfig), C.c_int32, C.c_int32, C.c_int32]
        L.oa_ref_destroy.argtypes = [C.c_void_p]
        L.oa_ref_tensor.restype = C.c_void_p
        L.oa_ref_tensor.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.oa_ref_forward.restype = C.c_int
        L.oa_ref_forward.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]
        L.oa_ref_generate.restype = C.c_int32
        L.oa_ref_generate.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oa_ref_gen_bf16.restype = C.c_uint16
        L.oa_ref_gen_bf16.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
        L.oa_ref_rope_table.argtypes = [C.POINTER(RefConfig), C.c_int32, C.c_void_p, C.c_void_p]
        L.oa_ref_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_int32]
        L.oa_ref_linear.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.oa_ref_argmax.restype = C.c_int32
        L.oa_ref_argmax.argtypes = [C.c_void_p, C.c_int32]
        L.oa_ref_num_threads.restype = C.c_int32
        L.oa_ref_set_threads.argtypes = [C.c_int32]
        L.oa_ref_set_threads(default_threads())
        _LIB = L
    return _LIB


# ------------------------------------------------------------------------------------------------
# Model presets — public architectures.  The engine's csrc/model_config.cpp holds
# the same table; tests/test_host_logic.py checks they agree through oa_model_info().
# ------------------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    name: str
    hidden: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    head_dim: int
    ffn: int
    vocab: int
    tie_embeddings: int = 0
    qkv_bias: int = 0
    rope_scaling: int = 0
    norm_random: int = 0
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5
    rope_factor: float = 32.0
    rope_low_freq: float = 1.0
    rope_high_freq: float = 4.0
    rope_orig_ctx: int = 8192
    init_std: float = 0.02
    seed: int = 1234
    template: str = "llama3"      # or "chatml"

    def to_c(self) -> RefConfig:
        d = asdict(self)
        d.pop("name"); d.pop("template")
        return RefConfig(**d)

    def engine_json(self, **extra) -> dict:
        d = asdict(self)
        d["model"] = d.pop("name")
        d.update(extra)
        return d


PRESETS = {
    "llama-3.2-1b": ModelSpec("llama-3.2-1b", 2048, 16, 32, 8, 64, 8192, 128256, tie_embeddings=1, rope_scaling=1),
    "llama-3-8b": ModelSpec("llama-3-8b", 4096, 32, 32, 8, 128, 14336, 128256),
    "qwen2.5-32b": ModelSpec("qwen2.5-32b", 5120, 64, 40, 8, 128, 27648, 152064, qkv_bias=1, rope_theta=1e6,
                             rms_eps=1e-6, template="chatml"),
    "llama-3-70b": ModelSpec("llama-3-70b", 8192, 80, 64, 8, 128, 28672, 128256),
    # tiny configs for parity tests (same code paths, seconds on CPU)
    "tiny-llama": ModelSpec("tiny-llama", 256, 2, 4, 2, 64, 512, 2304, norm_random=1),
    "tiny-llama-d128": ModelSpec("tiny-llama-d128", 512, 3, 4, 1, 128, 1024, 1024, norm_random=1, rope_scaling=1,
                                 tie_embeddings=1),
    # GQA group 8 (what one rank of Llama-3-70B TP=8 runs: 8 query heads on 1 kv head) and group 1 (MHA)
    "tiny-llama-g8": ModelSpec("tiny-llama-g8", 512, 2, 8, 1, 64, 512, 1024, norm_random=1),
    "tiny-llama-mha": ModelSpec("tiny-llama-mha", 256, 2, 4, 4, 64, 512, 1024, norm_random=1),
    # tensor-parallel-able tiny configs (kv heads divisible by the TP degree)
    "tiny-llama-tp": ModelSpec("tiny-llama-tp", 512, 2, 8, 4, 64, 1024, 2048, norm_random=1),
    "tiny-qwen-tp": ModelSpec("tiny-qwen-tp", 512, 2, 8, 2, 128, 512, 1280, qkv_bias=1, rope_theta=1e6, rms_eps=1e-6, norm_random=1,
                              template="chatml"),
    # the per-rank layouts of the two BASELINE tensor-parallel configs: Qwen2.5-32B TP=4 (40/8 heads -> 10 query heads on 2 kv heads per
    # rank, qkv bias) and Llama-3-70B TP=8 (64/8 heads -> 8 query heads on ONE kv head per rank)
    "tiny-qwen-tp4": ModelSpec("tiny-qwen-tp4", 512, 2, 40, 8, 128, 1024, 1280, qkv_bias=1, rope_theta=1e6, rms_eps=1e-6, norm_random=1,
                               template="chatml"),
    "tiny-llama-tp8": ModelSpec("tiny-llama-tp8", 512, 2, 64, 8, 64, 1024, 2048, norm_random=1),
    # Llama-3-8B's attention geometry (32 query heads on 8 kv heads, head_dim 128, llama3 RoPE scaling) on tiny other dims: what the decode
    # attention work plan of the BENCHMARKED batch (128 sequences x ctx ~1700 -> 296 CTAs, cut items, in-kernel merges) depends on
    "tiny-llama-8bheads": ModelSpec("tiny-llama-8bheads", 256, 1, 32, 8, 128, 256, 512, norm_random=1, rope_scaling=1),
    "tiny-qwen": ModelSpec("tiny-qwen", 320, 2, 5, 1, 64, 768, 1280, qkv_bias=1, rope_theta=1e6, rms_eps=1e-6,
                           norm_random=1, template="chatml"),
}


def _np_bf16_to_f32(a: np.ndarray) -> np.ndarray:
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = np.uint32(0x7FFF) + ((u >> 16) & 1)
    return ((u + r) >> 16).astype(np.uint16)


class Oracle:
    """One random-init model on the CPU.  mode: 0 = fp32 activations, 1 = bf16-faithful."""

    KINDS = {"wq": 0, "wk": 1, "wv": 2, "wo": 3, "wg": 4, "wu": 5, "wd": 6, "ln1": 7, "ln2": 8, "bq": 9, "bk": 10, "bv": 11}

    def __init__(self, spec: ModelSpec, max_pos: int = 512, n_slots: int = 1, mode: int = 1):
        self.spec, self.max_pos, self.n_slots, self.mode = spec, max_pos, n_slots, mode
        self._cfg = spec.to_c()
        self._h = lib().oa_ref_create(C.byref(self._cfg), max_pos, n_slots, mode)
        if not self._h:
            raise MemoryError("oa_ref_create failed")

    def close(self):
        if self._h:
            lib().oa_ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tensor(self, layer: int, kind, shape) -> np.ndarray:
        """bf16 bits of a weight tensor as uint16 ndarray (copy). layer=-1: kind 0 embed / 1 norm / 2 lm_head."""
        k = self.KINDS[kind] if isinstance(kind, str) else kind
        p = lib().oa_ref_tensor(self._h, layer, k)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint16)), (n,)).reshape(shape).copy()

    def tensor_f32(self, layer, kind, shape) -> np.ndarray:
        return _np_bf16_to_f32(self.tensor(layer, kind, shape))

    def forward(self, tokens, pos0: int = 0, slot: int = 0, all_logits: bool = False, want_hidden: bool = False):
        toks = np.ascontiguousarray(tokens, dtype=np.int32)
        T = len(toks)
        V = self.spec.vocab
        logits = np.empty((T if all_logits else 1, V), dtype=np.float32)
        hidden = np.empty((self.spec.hidden,), dtype=np.float32) if want_hidden else None
        rc = lib().oa_ref_forward(self._h, slot, toks.ctypes.data, T, pos0, int(all_logits), logits.ctypes.data,
                                  hidden.ctypes.data if want_hidden else None)
        if rc != 0:
            raise ValueError("oa_ref_forward: bad slot or position")
        return (logits, hidden) if want_hidden else logits

    def fill_kv(self, n: int, slot: int = 0, seed: int = 99) -> None:
        """timing aid (bench.py cpu_baseline only): seeded cache contents for positions [0, n) instead of a CPU prefill"""
        lib().oa_ref_fill_kv.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64]; lib().oa_ref_fill_kv.restype = C.c_int
        if lib().oa_ref_fill_kv(self._h, slot, n, seed) != 0:
            raise ValueError("oa_ref_fill_kv: bad slot or length")

    def generate(self, prompt, max_new: int, eos=(), slot: int = 0):
        """-> (tokens[int32], margins[float32], first_logits[float32 V])"""
        p = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty((max_new,), dtype=np.int32)
        mar = np.empty((max_new,), dtype=np.float32)
        fl = np.empty((self.spec.vocab,), dtype=np.float32)
        e = np.ascontiguousarray(list(eos), dtype=np.int32)
        n = lib().oa_ref_generate(self._h, slot, p.ctypes.data, len(p), max_new, e.ctypes.data if len(e) else None,
                                  len(e), out.ctypes.data, mar.ctypes.data, fl.ctypes.data)
        if n < 0:
            raise ValueError("oa_ref_generate failed")
        return out[:n].copy(), mar[:max(n, 1)].copy(), fl


# ------------------------------------------------------------------------------------------------
# Synthetic byte-level tokenizer + chat templates (restated; the product's is csrc/tokenizer.cpp).
# No tokenizer files exist offline, so text maps to ids 0..255 (one per UTF-8
# byte) and the chat-format control tokens keep their real ids inside the real vocab size.
# ------------------------------------------------------------------------------------------------
LLAMA3_SPECIALS = {"<|begin_of_text|>": 128000, "<|end_of_text|>": 128001, "<|start_header_id|>": 128006,
                   "<|end_header_id|>": 128007, "<|eot_id|>": 128009}
CHATML_SPECIALS = {"<|endoftext|>": 151643, "<|im_start|>": 151644, "<|im_end|>": 151645}


def specials_for(spec: ModelSpec) -> dict:
    """Control-token ids; tiny test vocabularies place them at the top of the vocab."""
    base = LLAMA3_SPECIALS if spec.template == "llama3" else CHATML_SPECIALS
    if max(base.values()) < spec.vocab:
        return dict(base)
    names = list(base.keys())
    return {n: spec.vocab - len(names) + i for i, n in enumerate(names)}


def apply_chat_template(spec: ModelSpec, messages) -> list[int]:
    """messages: [(role, content)] -> prompt token ids ending with the assistant generation header.
    Roles/order are those the ReAct loop produces (reference pkg/assistants/simple.go:358,496-501;
    seeds at pkg/handlers/execute.go:190-199)."""
    sp = specials_for(spec)
    ids: list[int] = []
    by = lambda s: list(s.encode("utf-8"))
    if spec.template == "llama3":
        ids.append(sp["<|begin_of_text|>"])
        for role, content in messages:
            ids += [sp["<|start_header_id|>"]] + by(role) + [sp["<|end_header_id|>"]] + by("\n\n") + by(content) + [sp["<|eot_id|>"]]
        ids += [sp["<|start_header_id|>"]] + by("assistant") + [sp["<|end_header_id|>"]] + by("\n\n")
    else:
        for role, content in messages:
            ids += [sp["<|im_start|>"]] + by(role + "\n") + by(content) + [sp["<|im_end|>"]] + by("\n")
        ids += [sp["<|im_start|>"]] + by("assistant\n")
    return ids


def eos_ids(spec: ModelSpec) -> list[int]:
    sp = specials_for(spec)
    return [sp["<|eot_id|>"], sp["<|end_of_text|>"]] if spec.template == "llama3" else [sp["<|im_end|>"], sp["<|endoftext|>"]]


def detokenize(ids) -> bytes:
    """Generated ids -> bytes.  The synthetic vocabulary is many-to-one on decode: every non-control
    id t renders as the single byte (t & 0xFF), so random-init models still produce text whose length
    equals the completion length (encode is the identity on bytes 0..255)."""
    return bytes(int(i) & 0xFF for i in ids)


# ------------------------------------------------------------------------------------------------
# Grammar that forces a completion to parse as tools.ToolPrompt (reference pkg/tools/tool.go:29-38).
# Restatement of opsagent_b200/csrc/grammar.hpp; tests walk both and compare every mask.
# ------------------------------------------------------------------------------------------------
TOOLS = ["kubectl", "python", "trivy", "jq", "search"]          # reference pkg/tools/tool.go:20-26
GRAMMAR_TOOLCALL, GRAMMAR_FINAL, GRAMMAR_FUNCTION, GRAMMAR_TEXT = 1, 2, 3, 4


class ToolPromptGrammar:
    def __init__(self, kind: int, functions: str = ""):
        L = lambda s: ("lit", s.encode(), 0, 0)
        S = lambda lo, hi: ("str", b"", lo, hi)
        self.opts = [t.encode() for t in TOOLS]
        self.close = 0x22
        if kind == GRAMMAR_TOOLCALL:
            self.segs = [L('{"question":"'), S(1, 64), L('","thought":"'), S(1, 96), L('","action":{"name":"'), ("enum", b"", 0, 0),
                         L('","input":"'), S(1, 96), L('"},"observation":"","final_answer":""}')]
        elif kind == GRAMMAR_FINAL:
            self.segs = [L('{"question":"'), S(1, 64), L('","thought":"'), S(1, 96),
                         L('","action":{"name":"","input":""},"observation":"","final_answer":"'), S(10, 160), L('"}')]
        elif kind == GRAMMAR_FUNCTION:       # OpenAI function calling: {"name":"<fn>","arguments":{"<param>":"..."}}
            self.opts = [(n + '","arguments":{"' + p + '":"').encode() for n, p in (it.split(":", 1) for it in functions.split(",") if ":" in it)]
            self.segs = [L('{"name":"'), ("enum", b"", 0, 0), S(1, 96), L('"}}')]
        elif kind == GRAMMAR_TEXT:
            self.close = 0x0A
            self.segs = [S(10, 200), L("\n")]
        else:
            raise ValueError("kind")
        self.seg, self.off, self.cand = 0, 0, set(range(len(self.opts)))

    @staticmethod
    def string_byte(b: int) -> bool:
        return 0x20 <= b <= 0x7E and b not in (0x22, 0x5C)

    def done(self) -> bool:
        return self.seg >= len(self.segs)

    def allowed(self) -> set:
        if self.done():
            return set()
        t, lit, lo, hi = self.segs[self.seg]
        if t == "lit":
            return {lit[self.off]}
        if t == "str":
            a = {b for b in range(0x20, 0x7F) if self.string_byte(b)} if self.off < hi else set()
            if self.off >= lo:
                a.add(self.close)
            return a
        return {self.opts[i][self.off] for i in self.cand}

    def _next(self):
        self.seg += 1; self.off = 0; self.cand = set(range(len(self.opts)))

    def advance(self, b: int) -> bool:
        if self.done() or b not in self.allowed():
            return False
        t, lit, lo, hi = self.segs[self.seg]
        if t == "lit":
            self.off += 1
            if self.off == len(lit):
                self._next()
        elif t == "str":
            if b == self.close:
                self._next(); self.off = 1
                if self.off == len(self.segs[self.seg][1]):
                    self._next()
            else:
                self.off += 1
        else:
            self.cand = {i for i in self.cand if self.opts[i][self.off] == b}
            self.off += 1
            if any(self.off == len(self.opts[i]) for i in self.cand):
                self._next()
        return True

    def mask_words(self) -> list:
        w = [0] * 8
        for b in self.allowed():
            w[b >> 5] |= 1 << (b & 31)
        return w


GRAMMAR_MAX_TOKEN_BYTES = 32      # restated from opsagent_b200/csrc/token_mask.hpp


def byte_level_token_bytes() -> list:
    """token id -> bytes for the synthetic byte-level vocabulary: ids 0..255 are one byte each, every other id is not a text token"""
    return [bytes([b]) for b in range(256)]


def bpe_token_bytes(tokenizer_json_path: str) -> list:
    """token id -> raw bytes of a Hugging Face byte-level BPE tokenizer.json (added/control tokens: b"") — parsed from the file itself
    (GPT-2 byte<->unicode alphabet), independent of opsagent_b200/csrc/bpe.hpp"""
    import json as _json
    d = _json.load(open(tokenizer_json_path, encoding="utf-8"))
    bs = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256)); cs = bs[:]; n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b); cs.append(256 + n); n += 1
    u2b = {chr(c): b for b, c in zip(bs, cs)}
    vocab = d["model"]["vocab"]
    added = {a["id"] for a in d.get("added_tokens", [])}
    size = max(max(vocab.values()), max(added) if added else 0) + 1
    out = [b""] * size
    for tok, i in vocab.items():
        if i not in added:
            out[i] = bytes(u2b[ch] for ch in tok)
    return out


def allowed_tokens(g: "ToolPromptGrammar", token_bytes) -> list:
    """Token ids allowed in g's current state, by brute force: a token is allowed iff ALL its bytes walk the automaton from here (the JSON
    may complete exactly at its last byte, never earlier), it is a text token, and it is at most GRAMMAR_MAX_TOKEN_BYTES long.
    Restates opsagent_b200/csrc/token_mask.hpp (trie walk + canonicalised cache) the slow, obvious way."""
    import copy
    out = []
    first = g.allowed()
    for tid, bs in enumerate(token_bytes):
        if not bs or len(bs) > GRAMMAR_MAX_TOKEN_BYTES or bs[0] not in first:
            continue
        h = copy.copy(g); h.cand = set(g.cand)
        if all(h.advance(b) for b in bs):
            out.append(tid)
    return out


def generate_constrained(orc: "Oracle", prompt, kind: int, max_new: int = 600, slot: int = 0, functions: str = "", token_bytes=None):
    """Greedy decoding under the ToolPrompt grammar: arg-max over the allowed TOKENS only (ties -> lowest id).  token_bytes: id -> bytes
    of the engine's tokenizer (default: the synthetic byte-level vocabulary, where tokens are bytes).
    -> (bytes, margins among the allowed set, token ids)"""
    token_bytes = token_bytes if token_bytes is not None else byte_level_token_bytes()
    g = ToolPromptGrammar(kind, functions)
    out, margins, ids = [], [], []
    logits = orc.forward(np.ascontiguousarray(prompt, dtype=np.int32), slot=slot)[0]
    pos = len(prompt)
    while not g.done() and len(ids) < max_new:
        allowed = allowed_tokens(g, token_bytes)
        vals = logits[allowed]
        k = int(np.argmax(vals))                 # first maximum = lowest id
        tok = allowed[k]
        rest = np.delete(vals, k)
        margins.append(float(vals[k] - rest.max()) if len(rest) else float("inf"))
        ids.append(tok)
        for b in token_bytes[tok]:
            out.append(b); g.advance(b)
        if g.done():
            break
        logits = orc.forward(np.array([tok], np.int32), pos0=pos, slot=slot)[0]
        pos += 1
    return bytes(out), margins, ids


def write_safetensors(orc: "Oracle", path: str, dtype: str = "BF16") -> None:
    """Dump the oracle's tensors as a Hugging Face style *.safetensors checkpoint (test fixture for the engine's loader)."""
    import json as _json
    import struct
    sp = orc.spec
    H, qd, kd, F, V = sp.hidden, sp.n_heads * sp.head_dim, sp.n_kv_heads * sp.head_dim, sp.ffn, sp.vocab
    items = [("model.embed_tokens.weight", orc.tensor(-1, 0, (V, H))), ("model.norm.weight", orc.tensor(-1, 1, (H,)))]
    if not sp.tie_embeddings:
        items.append(("lm_head.weight", orc.tensor(-1, 2, (V, H))))
    for l in range(sp.n_layers):
        p = f"model.layers.{l}."
        items += [(p + "self_attn.q_proj.weight", orc.tensor(l, "wq", (qd, H))), (p + "self_attn.k_proj.weight", orc.tensor(l, "wk", (kd, H))),
                  (p + "self_attn.v_proj.weight", orc.tensor(l, "wv", (kd, H))), (p + "self_attn.o_proj.weight", orc.tensor(l, "wo", (H, qd))),
                  (p + "mlp.gate_proj.weight", orc.tensor(l, "wg", (F, H))), (p + "mlp.up_proj.weight", orc.tensor(l, "wu", (F, H))),
                  (p + "mlp.down_proj.weight", orc.tensor(l, "wd", (H, F))), (p + "input_layernorm.weight", orc.tensor(l, "ln1", (H,))),
                  (p + "post_attention_layernorm.weight", orc.tensor(l, "ln2", (H,)))]
        if sp.qkv_bias:
            items += [(p + "self_attn.q_proj.bias", orc.tensor(l, "bq", (qd,))), (p + "self_attn.k_proj.bias", orc.tensor(l, "bk", (kd,))),
                      (p + "self_attn.v_proj.bias", orc.tensor(l, "bv", (kd,)))]
    header, blobs, off = {"__metadata__": {"format": "pt"}}, [], 0
    for name, bits in items:
        if dtype == "BF16":
            raw = np.ascontiguousarray(bits).tobytes()
        elif dtype == "F32":
            raw = _np_bf16_to_f32(bits).astype(np.float32).tobytes()
        else:
            raw = _np_bf16_to_f32(bits).astype(np.float16).tobytes()
        header[name] = {"dtype": dtype, "shape": list(bits.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw); off += len(raw)
    hj = _json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj))); f.write(hj)
        for b in blobs:
            f.write(b)
