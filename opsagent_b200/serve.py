"""python -m opsagent_b200.serve — start the engine(s) and the native OpenAI-compatible endpoint the unmodified reference binary is pointed at
(`baseUrl` of POST /api/execute, pkg/handlers/execute.go:21,205; OPENAI_API_BASE for the swarm-go flows, pkg/workflows/swarm.go:80-89).

    python -m opsagent_b200.serve --model llama-3-8b --weights /ckpt/Meta-Llama-3-8B-Instruct --tokenizer /ckpt/.../tokenizer.json \\
        --devices 0,1,2,3 --port 8000 --json-mode

One engine per device (data-parallel replicas behind one endpoint, BASELINE configs[2]); `--engine '{"kv_gb": 100, ...}'` passes any other engine
option.  Fails loudly if the CUDA library or a GPU is missing — there is no CPU path."""
from __future__ import annotations

import argparse
import json
import signal
import threading


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m opsagent_b200.serve", description=__doc__.split("\n\n")[0])
    ap.add_argument("--model", default="llama-3-8b", help="preset (llama-3.2-1b, llama-3-8b, qwen2.5-32b, llama-3-70b) or 'custom' with the dimensions in --engine")
    ap.add_argument("--weights", default="", help="*.safetensors file or shard directory (empty: seeded random weights)")
    ap.add_argument("--tokenizer", default="", help="tokenizer.json (empty: byte-level ids)")
    ap.add_argument("--devices", default="0", help="comma separated CUDA devices, one replica each")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--api-key", default="", help="required bearer token (empty: any non-empty token, as the reference sends its apiKey)")
    ap.add_argument("--json-mode", action="store_true", help="grammar-force every completion into tools.ToolPrompt (the ReAct loop of POST /api/execute)")
    ap.add_argument("--tool-steps", type=int, default=3, help="`tools` requests: forced function calls before the text answer")
    ap.add_argument("--max-inflight", type=int, default=256, help="requests per replica before 429")
    ap.add_argument("--engine", default="{}", help="JSON object of further engine options")
    args = ap.parse_args(argv)

    from .native_front import NativeFront
    from .router import Router
    cfg = {"model": args.model, "weights": args.weights, "tokenizer": args.tokenizer, "model_aliases": "*", "json_mode": int(args.json_mode), "react_tool_steps": args.tool_steps,
           **json.loads(args.engine)}
    devices = [int(d) for d in args.devices.split(",") if d != ""]
    replicas = Router.create(cfg, devices)                    # engines are created in parallel, one per device
    front = NativeFront(replicas.engines, host=args.host, port=args.port, api_key=args.api_key, tool_steps=args.tool_steps, max_inflight=args.max_inflight)
    print(json.dumps({"listening": f"http://{args.host}:{front.port}/v1", "model": replicas.info.get("model"), "replicas": len(devices)}), flush=True)
    stop = threading.Event()
    for sig in (signal.SIGINT, signal.SIGTERM):
        signal.signal(sig, lambda *_: stop.set())
    stop.wait()
    front.shutdown()
    replicas.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
