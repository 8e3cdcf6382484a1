"""synthetic — seeded stand-ins for what the reference shells out to (kubectl/trivy via bash, pkg/tools/*.go) so that ReAct loops can
run on-box without a cluster: BASELINE.json configs describe 'synthetic Pod YAML', 'kubectl tool-calls' and 'synthetic trivy-scan
observations'.  Not on the hot path; used by bench.py / tests as the `tools` argument of assistants.AssistantWithConfig."""
from __future__ import annotations

import hashlib
import random


def _seed(seed: int, text: str) -> int:
    """process-independent seed (the built-in hash() of a str is salted per process)"""
    return int.from_bytes(hashlib.blake2b(f"{seed}\x00{text}".encode("utf-8"), digest_size=8).digest(), "little")


def fake_kubectl(seed: int = 0, rows: int = 12):
    def kubectl(inp: str) -> str:
        r = random.Random(_seed(seed, inp))
        lines = ["NAME                      READY   STATUS             RESTARTS   AGE    IP            NODE"]
        for i in range(rows):
            st = r.choice(["Running", "Running", "Running", "CrashLoopBackOff", "Pending", "Completed"])
            lines.append(f"app-{r.randrange(16**5):05x}-{i:02d}          {r.randrange(0, 2)}/1     {st:<18} {r.randrange(0, 40):<10} {r.randrange(1, 90)}d    "
                         f"10.{r.randrange(256)}.{r.randrange(256)}.{r.randrange(256)}   node-{r.randrange(32)}")
        return "\n".join(lines)
    return kubectl


def fake_trivy(seed: int = 0, rows: int = 40):
    def trivy(image: str) -> str:
        r = random.Random(_seed(seed, image))
        lines = [f"{image} (debian 12.5)", "Total: %d (HIGH: %d, CRITICAL: %d)" % (rows, rows * 2 // 3, rows // 3), "LIBRARY | VULNERABILITY | SEVERITY | INSTALLED | FIXED | TITLE"]
        for _ in range(rows):
            lines.append(f"lib{r.randrange(400)} | CVE-20{r.randrange(15, 26)}-{r.randrange(1000, 60000)} | {r.choice(['HIGH', 'CRITICAL', 'MEDIUM'])} | "
                         f"{r.randrange(1, 9)}.{r.randrange(30)}.{r.randrange(30)} | {r.randrange(1, 9)}.{r.randrange(30)}.{r.randrange(40)} | overflow in parser {r.randrange(10**6)}")
        return "\n".join(lines)
    return trivy


def copilot_tools(seed: int = 0) -> dict:
    """same keys as tools.CopilotTools (reference pkg/tools/tool.go:20-26)"""
    k = fake_kubectl(seed)
    return {"kubectl": k, "trivy": fake_trivy(seed), "python": lambda s: "ok", "jq": lambda s: "{}", "search": lambda s: "no results"}
