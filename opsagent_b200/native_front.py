"""native_front — the C++ OpenAI-compatible HTTP front (csrc/http_server.cpp) from Python: `NativeFront(engines, ...)` starts it on the engines of
this process (one: a single GPU; several: the data-parallel router of BASELINE configs[2] — sticky least-loaded routing inside the server) and
`.port` / `.stats()` / `.shutdown()` manage it.  Same wire behaviour as http_front.py, no interpreter between the socket and the engines."""
from __future__ import annotations

import ctypes as C
import json

from . import _lib
from .engine import EngineError


class NativeFront:
    def __init__(self, engines, host: str = "127.0.0.1", port: int = 0, require_key: bool = True, api_key: str = "", tool_steps: int = 3,
                 max_inflight: int = 256, **options):
        self._L = _lib.load()
        self.engines = list(engines)               # keep them alive as long as the front runs
        arr = (C.c_void_p * max(1, len(self.engines)))(*[e._h for e in self.engines])
        opts = json.dumps({"host": host, "port": port, "require_key": int(require_key), "api_key": api_key, "tool_steps": tool_steps, "max_inflight": max_inflight,
                           **options})          # further flat options of oa_http_start: max_connections, max_body_bytes, idle_timeout_s
        h = C.c_void_p()
        rc = self._L.oa_http_start(arr, len(self.engines), opts.encode(), C.byref(h))
        if rc != 0:
            raise EngineError(rc, (self._L.oa_http_last_error() or b"").decode("utf-8", "replace"))
        self._h = h
        self.port = int(self._L.oa_http_port(h))

    def stats(self) -> dict:
        buf = C.create_string_buffer(1 << 16)
        self._L.oa_http_stats(self._h, buf, len(buf))
        return json.loads(buf.value.decode())

    def shutdown(self) -> None:
        if getattr(self, "_h", None):
            self._L.oa_http_stop(self._h)
            self._h = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass
