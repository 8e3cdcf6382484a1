"""http_front — OpenAI-compatible /v1/chat/completions in front of the engine.

Lets the UNMODIFIED reference binary use the engine through its existing knobs: `baseUrl` in POST /api/execute
(reference pkg/handlers/execute.go:21,205) or OPENAI_API_BASE for the swarm flows (pkg/workflows/swarm.go:83).
Wire format = what go-openai v1.38.0 sends/expects at pkg/llms/openai.go:70-82: request {model, messages[{role,content}],
max_tokens, temperature}; response {choices[{index, message{role,content}, finish_reason}], usage}; errors
{error{message,type,code}} with the HTTP status the retry loop switches on (openai.go:85-101)."""
from __future__ import annotations

import json
import threading
import time
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from .engine import EngineError
from .perf import GetPerfStats

_STATUS_TYPE = {400: "invalid_request_error", 401: "authentication_error", 429: "rate_limit_error", 500: "server_error"}


MAX_BODY_BYTES = 64 << 20        # a chat request is a few KB (16k-token observations: ~100 KB); anything near this is not a chat request


def make_handler(engine, require_key: bool = True, tool_steps: int = 3, api_key: str | None = None):
    """`api_key`: when given, the bearer token must equal it; otherwise any non-empty token is accepted — the reference forwards
    whatever X-API-Key the caller sent (openai.go:40-44) and a local engine has no account to check it against."""
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _read_body(self):
            """-> bytes, or None after answering 400/413 (connection closed: the unread body must not be parsed as the next request)"""
            try:
                n = int(self.headers.get("Content-Length", "0"))
            except ValueError:
                n = -1
            if n < 0 or n > MAX_BODY_BYTES:
                self.close_connection = True
                self._error(400 if n < 0 else 413, "bad Content-Length")
                return None
            return self.rfile.read(n) if n else b""

        def log_message(self, *a):      # quiet
            pass

        def _send(self, status: int, body: dict):
            data = json.dumps(body, ensure_ascii=False).encode("utf-8")
            self.send_response(status)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(data)))
            self.end_headers()
            self.wfile.write(data)

        def _authorised(self) -> bool:
            auth = self.headers.get("Authorization", "")
            if not require_key:
                return True
            if not (auth.startswith("Bearer ") and len(auth) > 7):
                return False
            return api_key is None or auth[7:] == api_key

        def _error(self, status: int, message: str):
            self._send(status, {"error": {"message": message, "type": _STATUS_TYPE.get(status, "server_error"), "code": status}})

        def do_GET(self):
            if self.path.rstrip("/").endswith("/models"):
                return self._send(200, {"object": "list", "data": [{"id": engine.info["model"], "object": "model", "owned_by": "opsagent_b200"}]})
            if self.path.rstrip("/").endswith("/perf/stats"):
                if not self._authorised():
                    return self._error(401, "missing bearer token")        # the reference serves it from the authenticated group (router.go:91-105)
                # the reference's GET /api/perf/stats (pkg/api/router.go:104, pkg/handlers/perf.go:12-25) plus the engine's own
                # counters (steps, tokens, bytes moved, kernel launches) under "engine" — steps/sec from the product's endpoint
                stats = GetPerfStats().GetStats()
                if hasattr(engine, "stats"):
                    stats["engine"] = engine.stats()
                return self._send(200, {"stats": stats, "status": "success"})
            self._error(404, "not found")

        def do_POST(self):
            body = self._read_body()          # always drained first: early 401/404 replies keep the keep-alive connection in sync
            if body is None:
                return
            if self.path.rstrip("/").endswith("/perf/reset"):          # pkg/api/router.go:105, pkg/handlers/perf.go:28-39
                if not self._authorised():
                    return self._error(401, "missing bearer token")
                GetPerfStats().Reset()
                return self._send(200, {"message": "performance statistics reset", "status": "success"})
            if not self.path.rstrip("/").endswith("/chat/completions"):
                return self._error(404, "not found")
            if not self._authorised():
                return self._error(401, "missing bearer token")        # the reference always sends its apiKey (openai.go:44)
            try:
                req = json.loads(body or b"{}")
                msgs = []
                for m in req["messages"]:            # function-calling turns are flattened into text the byte-level template can carry
                    if m.get("tool_calls"):
                        f = m["tool_calls"][0]["function"]
                        msgs.append((m["role"], '{"name":%s,"arguments":%s}' % (json.dumps(f["name"]), f.get("arguments") or "{}")))
                    else:
                        msgs.append((m["role"], m.get("content") or ""))
                if req.get("stream"):
                    return self._error(400, "streaming is not implemented (the reference does not request it)")
                if float(req.get("temperature", 0.0) or 0.0) > 1e-3:
                    return self._error(400, "only greedy decoding is implemented")
                max_tokens = int(req.get("max_tokens") or req.get("max_completion_tokens") or 1024)
            except Exception as e:
                return self._error(400, f"bad request: {e}")
            # OpenAI function calling (swarm-go flows, reference pkg/workflows/swarm.go:14-78, analyze.go:47-75): while fewer than
            # `react_tool_steps` tool results are in the history the reply is a grammar-forced call of one offered function,
            # afterwards one line of text.
            flags, functions = 0, None
            tools = req.get("tools") or []
            if tools:
                specs = []
                for t in tools:
                    f = t.get("function", {})
                    props = list((f.get("parameters") or {}).get("properties", {}).keys()) or ["input"]
                    specs.append(f"{f.get('name', 'fn')}:{props[0]}")
                n_results = sum(1 for m in req["messages"] if m.get("role") == "tool")
                if n_results < tool_steps:
                    flags, functions = 8, ",".join(specs)
                else:
                    flags = 16
            done = GetPerfStats().TraceFunc("chat_completion")
            try:
                out = engine.chat_complete(req.get("model", ""), msgs, max_tokens, flags=flags, functions=functions)
            except EngineError as e:
                GetPerfStats().RecordMetric("chat_completion_failed", done())
                return self._error(e.code if e.code in (400, 401, 429, 500) else 500, e.message)
            done()
            if flags == 8:
                try:
                    call = json.loads(out.content.decode("utf-8", "replace"))
                    call["name"], call["arguments"]
                except Exception:
                    # the grammar-forced call was cut off (max_tokens, or max_seq_len minus a long history): nothing parseable to return.
                    # 400 is what the API answers when the context budget cannot hold the completion; the reference does not retry it.
                    return self._error(400, f"function call truncated after {out.completion_tokens} tokens (finish_reason={out.finish_reason}): "
                                            "raise max_tokens or shorten the history")
                message = {"role": "assistant", "content": None,
                           "tool_calls": [{"id": f"call_{int(time.time() * 1e6):x}", "type": "function",
                                           "function": {"name": call["name"], "arguments": json.dumps(call["arguments"])}}]}
                return self._send(200, {"id": f"chatcmpl-{int(time.time() * 1e6):x}", "object": "chat.completion", "created": int(time.time()),
                                        "model": req.get("model", ""), "choices": [{"index": 0, "message": message, "finish_reason": "tool_calls"}],
                                        "usage": {"prompt_tokens": out.prompt_tokens, "completion_tokens": out.completion_tokens,
                                                  "total_tokens": out.prompt_tokens + out.completion_tokens}})
            self._send(200, {"id": f"chatcmpl-{int(time.time() * 1e6):x}", "object": "chat.completion", "created": int(time.time()),
                             "model": req.get("model", ""),
                             "choices": [{"index": 0, "message": {"role": "assistant", "content": out.content.decode("utf-8", "replace").rstrip("\n")},
                                          "finish_reason": out.finish_reason}],
                             "usage": {"prompt_tokens": out.prompt_tokens, "completion_tokens": out.completion_tokens,
                                       "total_tokens": out.prompt_tokens + out.completion_tokens}})
    return Handler


def serve(engine, host: str = "127.0.0.1", port: int = 8000, require_key: bool = True, tool_steps: int = 3, api_key: str | None = None):
    """-> (server, thread).  One OS thread per in-flight request, each blocking in the engine — the same concurrency shape as
    gin's goroutine-per-request; batching happens inside the engine."""
    class Server(ThreadingHTTPServer):
        request_queue_size = 4096          # hundreds of agents connect at once (BASELINE configs[2]: 1024 concurrent requests); the default backlog is 5
    srv = Server((host, port), make_handler(engine, require_key, tool_steps, api_key))
    srv.daemon_threads = True
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    return srv, th
