"""Pins the CPU oracle (oracle/llama_ref.c) against the committed HF-transformers fp32 fixtures
(tests/golden/hf_*.npz, written by tests/golden/gen_golden_hf.py).  The reference holds no golden
vectors for this path — these are the pin."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["tiny-llama", "tiny-llama-d128", "tiny-qwen", "tiny-llama-g8", "tiny-llama-mha", "tiny-qwen-tp4", "tiny-llama-tp8", "tiny-llama-8bheads"]
FP32_TOL = 2e-4      # abs, logits are O(1); observed 1.5e-6


@pytest.mark.parametrize("name", CASES)
def test_fp32_logits_match_hf(name):
    g = np.load(os.path.join(GOLD, f"hf_{name}.npz"))
    m = O.Oracle(O.PRESETS[name], max_pos=128, mode=0)
    lg = m.forward(g["prompt"], all_logits=True)
    assert lg.shape == g["logits"].shape
    assert np.abs(lg - g["logits"]).max() < FP32_TOL


@pytest.mark.parametrize("name", CASES)
def test_greedy_tokens_match_hf(name):
    g = np.load(os.path.join(GOLD, f"hf_{name}.npz"))
    m = O.Oracle(O.PRESETS[name], max_pos=128, mode=0)
    toks, margins, _ = m.generate(g["prompt"], len(g["gen"]))
    # token equality is required wherever HF's own top1-top2 margin exceeds the fp32 tolerance
    for i, (a, b) in enumerate(zip(toks, g["gen"])):
        if g["margins"][i] > 10 * FP32_TOL:
            assert a == b, f"step {i}"
        if a != b:
            break


@pytest.mark.parametrize("name", CASES)
def test_incremental_decode_equals_full_prefill(name):
    """KV-cache path: prefill T-3 then 3 single-token steps == one full forward (same arithmetic)."""
    g = np.load(os.path.join(GOLD, f"hf_{name}.npz"))
    p = g["prompt"]
    m = O.Oracle(O.PRESETS[name], max_pos=128, mode=0)
    m.forward(p[:-3])
    for k in range(3, 0, -1):
        lg = m.forward(p[len(p) - k:len(p) - k + 1], pos0=len(p) - k)
        assert np.abs(lg[0] - g["logits"][len(p) - k]).max() < FP32_TOL


@pytest.mark.parametrize("name", CASES)
def test_bf16_mode_close_to_fp32(name):
    """bf16-faithful mode differs from fp32 only by bf16 rounding noise (bound is loose; it documents
    the noise floor that the GPU-vs-oracle tolerance must stay well under)."""
    g = np.load(os.path.join(GOLD, f"hf_{name}.npz"))
    m = O.Oracle(O.PRESETS[name], max_pos=128, mode=1)
    lg = m.forward(g["prompt"], all_logits=True)
    err = np.abs(lg - g["logits"]).max()
    assert 0 < err < 0.08, err


def test_weight_generator_statistics_and_determinism():
    L = O.lib()
    a = np.array([L.oa_ref_gen_bf16(1234, 5, i, 0.02, 0.0) for i in range(20000)], dtype=np.uint16)
    b = np.array([L.oa_ref_gen_bf16(1234, 5, i, 0.02, 0.0) for i in range(20000)], dtype=np.uint16)
    assert (a == b).all()
    f = (a.astype(np.uint32) << 16).view(np.float32)
    assert abs(f.mean()) < 1e-3 and abs(f.std() - 0.02) < 1e-3
    c = np.array([L.oa_ref_gen_bf16(1235, 5, i, 0.02, 0.0) for i in range(100)], dtype=np.uint16)
    assert (a[:100] != c).any()
