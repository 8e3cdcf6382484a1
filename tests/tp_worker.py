"""Tensor-parallel parity worker — run as `python -m torch.distributed.run --nproc-per-node T tests/tp_worker.py [case ...]`.
Rank 0 drives the engine and checks it against the CPU oracle; ranks > 0 serve.  Prints TP_OK on success.

`run_cases()` is also what `bench.py --gpus N` calls after its data-parallel headline (N >= 2), so that the driver's own scaling runs
carry a tensor-parallel parity verdict at t = N (the GPU test tier runs on a 1-GPU box, where tests/test_tp_gpu.py skips).
Default cases per degree: t=2 tiny-llama-tp + tiny-qwen-tp; t=4 tiny-llama-tp + tiny-qwen-tp4 (Qwen2.5-32B TP=4's per-rank layout:
10 query heads on 2 kv heads, qkv bias); t=8 tiny-llama-tp8 (Llama-3-70B TP=8's per-rank layout: 8 query heads on ONE kv head)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LOGIT_TOL = 2.5e-2
DEFAULT_CASES = {2: ["tiny-llama-tp", "tiny-qwen-tp"], 4: ["tiny-llama-tp", "tiny-qwen-tp4"], 8: ["tiny-llama-tp8"]}


def run_cases(rank: int, world: int, cases=None, tag: str = "0", log=print, extra: dict | None = None) -> dict:
    """-> on rank 0: {"t", "cases", "max_dlogit", "tokens_identical", "tokens_compared", "ok"}; on followers: {} (they only serve)."""
    from opsagent_b200 import Engine
    cases = list(cases or DEFAULT_CASES.get(world, ["tiny-llama-tp"]))
    out = {"t": world, "cases": cases, "max_dlogit": 0.0, "tokens_identical": 0, "tokens_compared": 0, "near_tie_flips": 0, "ok": True}
    for name in cases:
        if rank == 0:
            from oracle import oracle as O                      # checker only (rank 0)
            spec = O.PRESETS[name]
            cfg = spec.engine_json()
        else:
            from opsagent_b200.presets_tiny import TINY_TP      # followers never touch the oracle package
            cfg = dict(TINY_TP[name])
        cfg.update(num_pages=64, max_seq_len=512, max_batch=16, max_step_tokens=256, device=rank, tp=world, tp_rank=rank,
                   tp_two_shot_rows=160,        # prefill chunks of 200 / 256 rows: two-shot all-reduce; 150 rows: one-shot (both paths covered)
                   tp_shm=f"/oa_tp_{tag}_{name}", tp_nonce=int(tag) if str(tag).isdigit() else 0)
        cfg.update(extra or {})
        eng = Engine(cfg)
        if rank == 0:
            out.setdefault("nvls", []).append(int(eng.info.get("tp_nvls", 0)))
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
            out["max_dlogit"] = max(out["max_dlogit"], err)
            if not (np.isfinite(got).all() and err < LOGIT_TOL):
                out["ok"] = False
            log(f"[tp{world}] {name} prefill n={n} max|dlogit|={err:.3e}")

        def compare(ids, ref, margins, g):
            k = 0
            while k < g and ids[k] == ref[k]:
                k += 1
            out["tokens_identical"] += k; out["tokens_compared"] += g if k == g else k + 1
            if k < g:
                out["near_tie_flips"] += 1
                if margins[k] > 2 * LOGIT_TOL:
                    out["ok"] = False
            return k
        for n, g in ((5, 40), (130, 30)):
            prompt = rng.integers(0, spec.vocab, size=n).astype(np.int32)
            ref, margins, _ = orc.generate(prompt, g)
            res = eng.generate(prompt.tolist(), g, flags=1)
            k = compare(res.token_ids, ref, margins, g)
            log(f"[tp{world}] {name} generate n={n}: {k}/{g} tokens identical")
        # batched requests through the scheduler
        prompts = [rng.integers(0, spec.vocab, size=int(m)).astype(np.int32) for m in rng.integers(3, 150, size=10)]
        tickets = [eng.tokens_submit(p.tolist(), 12, flags=1) for p in prompts]
        outs = [eng.wait(t) for t in tickets]
        for p, o in zip(prompts, outs):
            ref, margins, _ = orc.generate(p, 12)
            compare(o.token_ids, ref, margins, 12)
        log(f"[tp{world}] {name} batched ok={out['ok']}; stats={eng.stats()['decode_steps']} decode steps")
        eng.close(); orc.close()
    out["max_dlogit"] = round(out["max_dlogit"], 6)
    return out if rank == 0 else {}


if __name__ == "__main__":
    rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    import json
    res = run_cases(rank, world, [a for a in sys.argv[1:] if not a.startswith("{")] or None, tag=os.environ.get("MASTER_PORT", "0"), log=lambda m: print(m, flush=True),
                    extra=next((json.loads(a) for a in sys.argv[1:] if a.startswith("{")), None))
    if rank == 0:
        print(res, flush=True)
        assert res["ok"], res
        print("TP_OK", flush=True)
