"""router — the data-parallel front: ONE process that owns N engines (one per GPU) and routes live requests to them.

The reference serves every request from one process (pkg/api/router.go:95 — gin, one goroutine per request); with the model replicated
on N GPUs (BASELINE configs[2]) that one front has to spread independent requests over N engines with no data-path collective:

  * sticky per conversation — the ReAct loop resends its whole history on every step (pkg/assistants/simple.go:498-501), so a
    conversation is keyed by its first two messages (system prompt + first user turn) and always goes to the replica that already holds
    the prefix KV pages of its earlier steps;
  * new conversations go to the replica with the fewest requests in flight;
  * per-replica admission control: more than `max_inflight` requests on the chosen replica -> 429 (the caller's retry loop,
    pkg/llms/openai.go:91-94, backs off), instead of an unbounded queue on one GPU while others idle.

`Router` has the Engine surface http_front needs (chat_complete / info / stats / count_tokens), so `http_front.serve(Router([...]))` is the
single OpenAI-compatible endpoint of an N-GPU box.  Engines live in this process (one scheduler thread and CUDA stream each, all C++);
caller threads block inside the C ABI with the GIL released."""
from __future__ import annotations

import hashlib
import threading
from collections import OrderedDict

from .engine import Engine, EngineError


def conversation_key(messages) -> bytes:
    """first two messages of the history (role + content): constant across the steps of one ReAct conversation"""
    h = hashlib.blake2b(digest_size=16)
    for role, content in list(messages)[:2]:
        r = role if isinstance(role, bytes) else str(role).encode("utf-8")
        c = content if isinstance(content, bytes) else str(content).encode("utf-8")
        h.update(len(r).to_bytes(4, "little")); h.update(r); h.update(len(c).to_bytes(8, "little")); h.update(c)
    return h.digest()


class Router:
    def __init__(self, engines, max_inflight: int = 256, max_conversations: int = 65536):
        if not engines:
            raise ValueError("Router needs at least one engine")
        self.engines = list(engines)
        self.max_inflight = max_inflight
        self._mu = threading.Lock()
        self._inflight = [0] * len(self.engines)
        self._routed = [0] * len(self.engines)
        self._home: "OrderedDict[bytes, int]" = OrderedDict()      # conversation -> replica (LRU-bounded)
        self._max_conv = max_conversations
        self._rejected = 0
        self._sticky_hits = 0
        self.info = dict(self.engines[0].info)
        self.info["replicas"] = len(self.engines)

    @classmethod
    def create(cls, config: dict, devices, **kw) -> "Router":
        """one engine per device from the same config (engine creation is parallel: weights are generated / loaded on each GPU)"""
        engines = [None] * len(devices)
        errors = []

        def make(i, dev):
            try:
                engines[i] = Engine({**config, "device": dev})
            except Exception as e:      # noqa: BLE001
                errors.append(e)
        th = [threading.Thread(target=make, args=(i, d)) for i, d in enumerate(devices)]
        [t.start() for t in th]; [t.join() for t in th]
        if errors:
            for e in engines:
                if e is not None:
                    e.close()
            raise errors[0]
        return cls(engines, **kw)

    # ---- routing ----
    def _acquire(self, messages) -> int:
        key = conversation_key(messages)
        with self._mu:
            r = self._home.get(key)
            if r is not None:
                self._home.move_to_end(key); self._sticky_hits += 1
            else:
                r = min(range(len(self.engines)), key=lambda i: (self._inflight[i], self._routed[i]))
                self._home[key] = r
                if len(self._home) > self._max_conv:
                    self._home.popitem(last=False)
            if self._inflight[r] >= self.max_inflight:
                self._rejected += 1
                raise EngineError(429, f"replica {r} has {self._inflight[r]} requests in flight (limit {self.max_inflight})")
            self._inflight[r] += 1; self._routed[r] += 1
            return r

    def _release(self, r: int) -> None:
        with self._mu:
            self._inflight[r] -= 1

    def replica_of(self, messages):
        """where this conversation's next step would go (None: not seen yet)"""
        with self._mu:
            return self._home.get(conversation_key(messages))

    # ---- Engine surface ----
    def chat_complete(self, model, messages, max_tokens, flags: int = 0, functions=None):
        r = self._acquire(messages)
        try:
            return self.engines[r].chat_complete(model, messages, max_tokens, flags=flags, functions=functions)
        finally:
            self._release(r)

    def count_tokens(self, messages) -> int:
        return self.engines[0].count_tokens(messages)

    def apply_chat_template(self, messages):
        return self.engines[0].apply_chat_template(messages)

    def stats(self) -> dict:
        per = [e.stats() for e in self.engines]
        agg = {}
        for k in per[0]:
            if isinstance(per[0][k], (int, float)):
                agg[k] = sum(p[k] for p in per)
        with self._mu:
            agg.update({"replicas": len(self.engines), "inflight": list(self._inflight), "routed": list(self._routed), "rejected_429": self._rejected,
                        "sticky_hits": self._sticky_hits, "conversations_tracked": len(self._home)})
        agg["per_replica"] = per
        return agg

    def close(self) -> None:
        for e in self.engines:
            e.close()
