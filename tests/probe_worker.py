"""Child process of tests/test_probe_gpu.py: one configuration per process, so that a device fault in one of them cannot poison the CUDA
context of the main test run.  Exit code 0 = probe passed; anything else = failed (reason on stdout)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opsagent_b200 import Engine, _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)

LOGIT_TOL = 2.5e-2


def greedy_ok(ref, margins, got):
    for i, (a, b) in enumerate(zip(got, ref)):
        if a != b:
            return margins[i] <= 2 * LOGIT_TOL          # a legitimate near-tie flip; later tokens differ by construction
    return len(got) == len(ref)


def parity(name):
    spec = O.PRESETS[name]
    eng = Engine(spec.engine_json(num_pages=64, max_seq_len=512, max_batch=16, max_step_tokens=256))
    orc = O.Oracle(spec, max_pos=512, n_slots=1, mode=1)
    rng = np.random.default_rng(21)
    for n in (1, 17, 65, 150):
        toks = rng.integers(0, spec.vocab, size=n).astype(np.int32)
        err = float(np.abs(eng.debug_prefill_logits(toks) - orc.forward(toks, all_logits=True)).max())
        if not err < LOGIT_TOL:
            print(f"{name}: prefill logits n={n} err={err}"); return 1
    for n, g in ((5, 40), (130, 48)):
        prompt = rng.integers(0, spec.vocab, size=n).astype(np.int32)
        ref, margins, _ = orc.generate(prompt, g)
        out = eng.generate(prompt.tolist(), g, flags=1)
        if not greedy_ok(list(ref), list(margins), list(out.token_ids)):
            print(f"{name}: greedy decode diverges beyond a near tie (prompt {n})"); return 1
    eng.close(); orc.close()
    print(f"{name}: ok"); return 0


def bpe():
    tok = os.path.join(ROOT, "tests", "golden", "bpe_llama3_tiny.json")
    spec = O.ModelSpec("bpe-probe", 256, 2, 4, 2, 64, 512, 704, norm_random=1)
    eng = Engine(spec.engine_json(num_pages=64, max_seq_len=512, max_batch=4, max_step_tokens=256, tokenizer=tok))
    orc = O.Oracle(spec, max_pos=512, n_slots=1, mode=1)
    msgs = [("system", "you are a kubectl expert"), ("user", "集群 has 5 namespaces, list them")]
    ids = eng.apply_chat_template(msgs)
    if eng.count_tokens(msgs) != len(ids):
        print("bpe: count_tokens disagrees with the template"); return 1
    out = eng.chat_complete("", msgs, 24, flags=1)
    ref, margins, _ = orc.generate(np.asarray(ids, np.int32), 24)
    if not greedy_ok(list(ref), list(margins), list(out.token_ids)):
        print("bpe: generated ids diverge from the oracle beyond a near tie"); return 1
    import ctypes as C
    L = _lib.load(); arr = np.asarray(out.token_ids, np.int32); buf = C.create_string_buffer(4096); n = C.c_int32()
    if L.oa_host_bpe_decode(tok.encode(), arr.ctypes.data, len(arr), buf, len(buf), C.byref(n)) != 0 or buf.raw[: n.value] != bytes(out.content):
        print("bpe: completion text is not the BPE decoding of the generated ids"); return 1
    # schema-constrained decoding over the BPE vocabulary (token masks: csrc/token_mask.hpp): the completion must parse as tools.ToolPrompt
    # and equal the oracle's constrained greedy decode over the same token set, token for token (up to a near tie)
    import json as _json
    tb = O.bpe_token_bytes(tok)
    for kind, flag in ((O.GRAMMAR_TOOLCALL, 2), (O.GRAMMAR_FINAL, 4)):
        out = eng.chat_complete("", msgs, 400, flags=flag)
        doc = _json.loads(bytes(out.content).decode("utf-8"))
        if list(doc.keys()) != ["question", "thought", "action", "observation", "final_answer"] or out.finish_reason != "stop":
            print("bpe: constrained completion is not ToolPrompt JSON:", bytes(out.content)[:200]); return 1
        _text, margins, ref_ids = O.generate_constrained(orc, np.asarray(ids, np.int32), kind, token_bytes=tb)
        got = list(out.token_ids)
        k = 0
        while k < min(len(ref_ids), len(got)) and ref_ids[k] == got[k]:
            k += 1
        if not (k == len(ref_ids) == len(got) or margins[k] <= 2 * LOGIT_TOL):
            print(f"bpe: constrained decode (kind {kind}) diverges from the oracle at token {k} with margin {margins[k]}"); return 1
    if eng.stats().get("grammar_states_computed", 0) <= 0:
        print("bpe: no token mask was computed"); return 1
    eng.close(); orc.close()
    print("bpe: ok"); return 0


def mixed():
    """decode rows riding along in prefill steps (engine option mixed_steps=1): every request must still match the oracle"""
    import threading
    import time
    spec = O.PRESETS["tiny-llama-d128"]
    eng = Engine(spec.engine_json(num_pages=96, max_seq_len=512, max_batch=8, max_step_tokens=256, mixed_steps=1, prefill_batch_tokens=0))
    orc = O.Oracle(spec, max_pos=512, n_slots=1, mode=1)
    rng = np.random.default_rng(33)
    jobs = [(20, 420)] + [(int(n), 30) for n in (40, 90, 150, 200, 64, 129)]
    prompts = [rng.integers(0, spec.vocab, size=n).astype(np.int32) for n, _ in jobs]
    results = [None] * len(jobs)

    def run(i):
        results[i] = eng.generate(prompts[i].tolist(), jobs[i][1], flags=1)

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
    th[0].start()
    t_end = time.time() + 30
    while eng.stats()["decode_steps"] < 3 and time.time() < t_end:      # the long request is provably decoding when the others arrive
        time.sleep(0.0005)
    for t in th[1:]:
        t.start(); time.sleep(0.002)
    [t.join() for t in th]
    st = eng.stats()
    if st.get("mixed_steps", 0) <= 0:
        print("mixed: no mixed step was scheduled (inconclusive)"); return 1
    for i, (n, g) in enumerate(jobs):
        ref, margins, _ = orc.generate(prompts[i], g)
        if results[i].completion_tokens != g or not greedy_ok(list(ref), list(margins), list(results[i].token_ids)):
            print(f"mixed: request {i} (prompt {n}) diverges from the oracle beyond a near tie"); return 1
    if st["pages_free"] != st["pages_total"]:
        print("mixed: pages leaked"); return 1
    eng.close(); orc.close()
    print(f"mixed: ok ({st['mixed_steps']} mixed steps)"); return 0


if __name__ == "__main__":
    import faulthandler
    faulthandler.dump_traceback_later(200, exit=True)        # a hung probe ends itself (the parent's own timeout is 240 s)
    sys.exit({"bpe": bpe, "mixed": mixed}.get(sys.argv[1], lambda: parity(sys.argv[1]))())
