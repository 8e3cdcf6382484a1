/* opsagent_serve — a stand-alone server in strict C99 on top of the C ABI (include/opsagent_b200.h): one engine per listed GPU behind the library's
 * own OpenAI-compatible endpoint.  This is the whole host side a deployment needs when the reference binary stays unmodified and is pointed at the
 * engine through `baseUrl` (pkg/handlers/execute.go:21,205) / OPENAI_API_BASE (pkg/workflows/swarm.go:80-89):
 *
 *   gcc -std=c99 -O2 -I include integration/c/opsagent_serve.c -L opsagent_b200/lib -lopsagent_b200 -Wl,-rpath,$PWD/opsagent_b200/lib -o opsagent_serve
 *   ./opsagent_serve --port 8000 --devices 0,1,2,3 --engine '{"model": "llama-3-8b", "weights": "/ckpt", "tokenizer": "/ckpt/tokenizer.json", "json_mode": 1}'
 *
 * Exit codes: 0 after SIGINT / SIGTERM, 2 usage, 1 an engine or the front could not be started (message on stderr; there is no CPU fallback). */
#define _POSIX_C_SOURCE 200809L
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "opsagent_b200.h"

#define MAX_REPLICAS 16

static volatile sig_atomic_t stopping = 0;
static void on_signal(int sig) { (void)sig; stopping = 1; }

int main(int argc, char** argv) {
    const char* engine_json = "{\"model\": \"llama-3-8b\"}";
    const char* devices = "0";
    const char* host = "127.0.0.1";
    const char* api_key = "";
    int port = 8000, tool_steps = 3, max_inflight = 256, i, n = 0;
    oa_engine* engines[MAX_REPLICAS];
    oa_http* front = NULL;
    char cfg[8192], opts[1024];
    sigset_t block, old;
    struct sigaction sa;

    /* SIGINT / SIGTERM are blocked before the library starts its threads (they inherit the mask), so only this thread, inside sigsuspend, takes them */
    sigemptyset(&block); sigaddset(&block, SIGINT); sigaddset(&block, SIGTERM);
    sigprocmask(SIG_BLOCK, &block, &old);
    memset(&sa, 0, sizeof sa); sa.sa_handler = on_signal; sigemptyset(&sa.sa_mask);
    sigaction(SIGINT, &sa, NULL); sigaction(SIGTERM, &sa, NULL);

    for (i = 1; i < argc; ++i) {
        const char* a = argv[i];
        const char* v = i + 1 < argc ? argv[i + 1] : NULL;
        if (!strcmp(a, "--engine") && v) { engine_json = v; ++i; }
        else if (!strcmp(a, "--devices") && v) { devices = v; ++i; }
        else if (!strcmp(a, "--host") && v) { host = v; ++i; }
        else if (!strcmp(a, "--port") && v) { port = atoi(v); ++i; }
        else if (!strcmp(a, "--api-key") && v) { api_key = v; ++i; }
        else if (!strcmp(a, "--tool-steps") && v) { tool_steps = atoi(v); ++i; }
        else if (!strcmp(a, "--max-inflight") && v) { max_inflight = atoi(v); ++i; }
        else {
            fprintf(stderr, "usage: %s [--engine JSON] [--devices 0,1,..] [--host H] [--port P] [--api-key K] [--tool-steps N] [--max-inflight N]\n", argv[0]);
            return 2;
        }
    }
    {   /* the engine options with "device" and the catch-all model alias spliced in front of the user's keys (later keys win in the flat reader) */
        const char* open = strchr(engine_json, '{');
        const char* p = devices;
        if (!open) { fprintf(stderr, "--engine must be a JSON object\n"); return 2; }
        while (*p && n < MAX_REPLICAS) {
            char* end;
            const long dev = strtol(p, &end, 10);
            const char* rest = open + 1;
            int empty;
            if (end == p) { fprintf(stderr, "--devices must be a comma separated list of integers\n"); return 2; }
            while (*rest == ' ') ++rest;
            empty = *rest == '}';
            if (snprintf(cfg, sizeof cfg, "{\"device\": %ld, \"model_aliases\": \"*\"%s%s", dev, empty ? "" : ", ", open + 1) >= (int)sizeof cfg) { fprintf(stderr, "--engine too long\n"); return 2; }
            if (oa_engine_create(cfg, &engines[n]) != OA_OK) {
                fprintf(stderr, "engine on device %ld: %s\n", dev, oa_last_error());
                while (n > 0) oa_engine_destroy(engines[--n]);
                return 1;
            }
            ++n;
            p = *end == ',' ? end + 1 : end;
        }
    }
    if (n == 0) { fprintf(stderr, "no devices given\n"); return 2; }
    if (strpbrk(host, "\"\\") || strpbrk(api_key, "\"\\")) { fprintf(stderr, "--host / --api-key must not contain quotes or backslashes\n"); while (n > 0) oa_engine_destroy(engines[--n]); return 2; }
    snprintf(opts, sizeof opts, "{\"host\": \"%s\", \"port\": %d, \"require_key\": 1, \"api_key\": \"%s\", \"tool_steps\": %d, \"max_inflight\": %d}", host, port, api_key, tool_steps, max_inflight);
    if (oa_http_start(engines, n, opts, &front) != OA_OK) {
        fprintf(stderr, "front: %s\n", oa_http_last_error());
        while (n > 0) oa_engine_destroy(engines[--n]);
        return 1;
    }
    printf("{\"listening\": \"http://%s:%d/v1\", \"replicas\": %d}\n", host, (int)oa_http_port(front), n);
    fflush(stdout);
    while (!stopping) sigsuspend(&old);
    oa_http_stop(front);
    while (n > 0) oa_engine_destroy(engines[--n]);
    return 0;
}
