"""The boundary is a C ABI: include/opsagent_b200.h must compile as plain C99 (no C++ types, no default arguments) and a C program
linked against the library must be able to drive it.  No compute calls: engine creation is only asked to reject a bad config."""
import os
import shutil
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "opsagent_b200", "lib")

SRC = textwrap.dedent(r'''
    #include <stdio.h>
    #include <string.h>
    #include "opsagent_b200.h"

    int main(int argc, char** argv) {
        oa_engine* e = (oa_engine*)0;
        int rc = oa_engine_create("{\"model\": \"no-such-model\"}", &e);
        if (rc != 400 || e != (oa_engine*)0) { printf("create rc=%d\n", rc); return 1; }
        if (!strstr(oa_last_error(), "unknown model")) { printf("error text: %s\n", oa_last_error()); return 2; }
        if (!strstr(oa_version(), "sm_100a")) return 3;
        {   /* request struct is plain data */
            oa_msg m[2]; oa_chat_req r; oa_chat_resp out;
            memset(&r, 0, sizeof r); memset(&out, 0, sizeof out);
            m[0].role = "system"; m[0].content = "s"; m[1].role = "user"; m[1].content = "u";
            r.model = "m"; r.msgs = m; r.n_msgs = 2; r.max_tokens = 8; r.temperature = 0.0f; r.flags = OA_FLAG_IGNORE_EOS;
            if (oa_chat_complete((oa_engine*)0, &r, &out) != 400) return 4;      /* null engine is a bad request, not a crash */
        }
        if (argc > 1) {      /* byte-level BPE through the C ABI */
            int32_t ids[64]; int32_t n = 0; char buf[64]; int32_t nb = 0;
            if (oa_host_bpe_encode(argv[1], "kubectl get pods", 16, ids, 64, &n) != 0 || n <= 0) { printf("encode: %s\n", oa_last_error()); return 5; }
            if (oa_host_bpe_decode(argv[1], ids, n, buf, 64, &nb) != 0 || nb != 16 || memcmp(buf, "kubectl get pods", 16)) return 6;
        }
        printf("ok\n");
        return 0;
    }
''')


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_header_is_c99_and_a_c_program_links_and_runs(tmp_path):
    if not os.path.exists(os.path.join(LIBDIR, "libopsagent_b200.so")):
        pytest.skip("library not built")
    src = tmp_path / "abi.c"; src.write_text(SRC)
    exe = tmp_path / "abi"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", LIBDIR, "-lopsagent_b200", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    tok = os.path.join(ROOT, "tests", "golden", "bpe_llama3_tiny.json")
    r = subprocess.run([str(exe), tok], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", (r.returncode, r.stdout, r.stderr)
