"""Tensor-parallel parity worker — run as `python -m torch.distributed.run --nproc-per-node T tests/tp_worker.py`.
Rank 0 drives the engine and checks it against the CPU oracle; ranks > 0 serve.  Prints TP_OK on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opsagent_b200 import Engine  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)

LOGIT_TOL = 2.5e-2
rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
port = os.environ.get("MASTER_PORT", "0")
cases = [c for c in sys.argv[1:]] or (["tiny-llama-tp", "tiny-qwen-tp"] if world == 2 else ["tiny-llama-tp"])

for name in cases:
    spec = O.PRESETS[name]
    cfg = spec.engine_json(num_pages=64, max_seq_len=512, max_batch=16, max_step_tokens=256, device=rank, tp=world, tp_rank=rank,
                           tp_shm=f"/oa_tp_{port}_{name}")
    eng = Engine(cfg)
    if rank > 0:
        eng.serve(); eng.close()
        continue
    orc = O.Oracle(spec, max_pos=512, mode=1)
    rng = np.random.default_rng(1)
    for n in (1, 17, 64, 150, 200, 256):                 # <=256: stream-K + fp32 all-reduce; larger prefill chunks: tile GEMM + bf16 all-reduce
        toks = rng.integers(0, spec.vocab, size=n).astype(np.int32)
        got = eng.debug_prefill_logits(toks)
        ref = orc.forward(toks, all_logits=True)
        err = float(np.abs(got - ref).max())
        assert np.isfinite(got).all() and err < LOGIT_TOL, (name, n, err)
        print(f"[tp{world}] {name} prefill n={n} max|dlogit|={err:.3e}", flush=True)
    for n, g in ((5, 40), (130, 30)):
        prompt = rng.integers(0, spec.vocab, size=n).astype(np.int32)
        ref, margins, _ = orc.generate(prompt, g)
        out = eng.generate(prompt.tolist(), g, flags=1)
        k = 0
        while k < g and out.token_ids[k] == ref[k]:
            k += 1
        assert k == g or margins[k] <= 2 * LOGIT_TOL, (name, n, k, margins[k])
        print(f"[tp{world}] {name} generate n={n}: {k}/{g} tokens identical", flush=True)
    # batched requests through the scheduler
    prompts = [rng.integers(0, spec.vocab, size=int(m)).astype(np.int32) for m in rng.integers(3, 150, size=10)]
    tickets = [eng.tokens_submit(p.tolist(), 12, flags=1) for p in prompts]
    outs = [eng.wait(t) for t in tickets]
    for p, o in zip(prompts, outs):
        ref, margins, _ = orc.generate(p, 12)
        k = 0
        while k < 12 and o.token_ids[k] == ref[k]:
            k += 1
        assert k == 12 or margins[k] <= 2 * LOGIT_TOL
    print(f"[tp{world}] {name} batched ok; stats={eng.stats()['decode_steps']} decode steps", flush=True)
    eng.close(); orc.close()
if rank == 0:
    print("TP_OK", flush=True)
