"""GPU probes for configurations added WITHOUT a GPU run (round 1 ended with the GPU budget spent): GQA group 8 — what one rank of
Llama-3-70B TP=8 runs — and group 1, an engine whose text goes through a trained byte-level BPE tokenizer, and the opt-in scheduler
mode `mixed_steps=1` (decoding sequences ride along in prefill steps: both attention kernels in one forward).  Their CPU halves are
pinned (oracle vs HF fixtures, BPE vs the tokenizers library); these run the GPU half.  Each probe runs in its own process
(tests/probe_worker.py) and is a NON-strict xfail: green shows up as XPASS, a failure as xfail with the reason — the main suite stays
meaningful either way.  Promote to hard tests once seen green."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.xfail(reason="added without a GPU run in round 1; promote to a hard test once seen green", strict=False)
@pytest.mark.parametrize("probe", ["tiny-llama-g8", "tiny-llama-mha", "bpe", "mixed"])
def test_probe(probe):
    r = subprocess.run([sys.executable, os.path.join(HERE, "probe_worker.py"), probe], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
