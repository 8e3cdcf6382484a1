/*
 * oracle/llama_ref.c — CPU restatement of the decoder-only transformer chat-completion path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.
 *
 * PARITY UNPINNED BY THE REFERENCE: myysophia/OpsAgent contains no model arithmetic at all.  Its
 * "LLM path" is llms.OpenAIClient.Chat (reference pkg/llms/openai.go:69-104), an HTTP call to
 * whichever OpenAI-compatible server the operator runs; no test, fixture or golden vector in the
 * reference constrains logits or token ids.  The arithmetic restated here is
 * therefore the published Llama-3 / Qwen2.5 decoder definition, and it is pinned against an
 * independent implementation instead: HF transformers 5.5 LlamaForCausalLM / Qwen2ForCausalLM in
 * fp32 on CPU (tests/golden/gen_golden_hf.py writes the fixtures, tests/test_oracle_golden.py
 * checks them).  From the reference the oracle follows only the *contract*:
 *   - greedy decoding: Temperature = math.SmallestNonzeroFloat32  (pkg/llms/openai.go:73)
 *   - first choice's content is the result                      (pkg/llms/openai.go:82)
 *   - max_tokens bounds the completion                          (pkg/llms/openai.go:72)
 *   - message order/roles produced by the ReAct loop            (pkg/assistants/simple.go:358,496-501)
 *
 * Two arithmetic modes:
 *   mode 0 "fp32"  — bf16-valued weights, every activation kept in fp32.  This is what is compared
 *                    with HF fp32 (tolerance 2e-4 abs on logits of the tiny golden configs).
 *   mode 1 "bf16"  — same weights; activations rounded to bf16 at exactly the points where the
 *                    CUDA path stores bf16 tensors (listed at ref_layer()).  This is what the GPU
 *                    engine is compared with; remaining differences are fp32 accumulation order.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* bf16 helpers (round-to-nearest-even, identical bit logic to csrc/common.cuh)               */
/* ------------------------------------------------------------------------------------------ */
static inline uint16_t f32_to_bf16(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* NaN */
    uint32_t r = 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)((u + r) >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16; float f; memcpy(&f, &u, 4); return f;
}
static inline float rbf(float f) { return bf16_to_f32(f32_to_bf16(f)); }

/* ------------------------------------------------------------------------------------------ */
/* Deterministic weight generator.  The SAME function is implemented in csrc/weights.cu; both  */
/* use only integer arithmetic plus one exact int->float conversion and one fp32 multiply, so  */
/* CPU and GPU produce bit-identical bf16 tensors from (seed, tensor_id, logical index).       */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* Irwin-Hall(4) of 16-bit uniforms: mean 131070, std 65536*sqrt(4/12)=37837.23 -> ~N(0,1) */
static inline uint64_t gen_key(uint64_t seed, uint64_t tensor_id) { return mix64(seed ^ (tensor_id * 0xD6E8FEB86659FD93ull)); }
static inline float gen_unit_k(uint64_t key, uint64_t idx) {
    uint64_t h = mix64(key + idx);
    int32_t s = (int32_t)(h & 0xffff) + (int32_t)((h >> 16) & 0xffff) +
                (int32_t)((h >> 32) & 0xffff) + (int32_t)((h >> 48) & 0xffff) - 131070;
    return (float)s * (1.0f / 37837.227f);
}
REF_API uint16_t oa_ref_gen_bf16(uint64_t seed, uint64_t tensor_id, uint64_t idx, float std, float mean) {
    return f32_to_bf16(gen_unit_k(gen_key(seed, tensor_id), idx) * std + mean);
}

/* tensor ids: layer*16 + kind; globals use layer index = n_layers */
enum { T_WQ = 0, T_WK, T_WV, T_WO, T_WG, T_WU, T_WD, T_LN1, T_LN2, T_BQ, T_BK, T_BV };
enum { TG_EMBED = 0, TG_NORM = 1, TG_LMHEAD = 2 };

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t hidden, n_layers, n_heads, n_kv_heads, head_dim, ffn, vocab;
    int32_t tie_embeddings;     /* lm_head = embed (Llama-3.2-1B) */
    int32_t qkv_bias;           /* Qwen2.5 */
    int32_t rope_scaling;       /* 0 none, 1 llama3 */
    int32_t norm_random;        /* 0: norm gains = 1 (HF init), 1: 1 + 0.1*N(0,1) to exercise the multiply */
    float   rope_theta, rms_eps;
    float   rope_factor, rope_low_freq, rope_high_freq; int32_t rope_orig_ctx;
    float   init_std;           /* 0.02 */
    uint64_t seed;
} ref_config;

typedef struct {
    ref_config c;
    uint16_t *embed, *lm_head, *norm;
    uint16_t **wq, **wk, **wv, **wo, **wg, **wu, **wd, **ln1, **ln2, **bq, **bk, **bv;
    float *rope_cos, *rope_sin;      /* [max_pos][head_dim/2] */
    int32_t max_pos;
    /* KV cache for n_slots independent sequences, fp32 storage (values already bf16-rounded in mode 1) */
    int32_t n_slots; float *kcache, *vcache; /* [slot][layer][pos][n_kv*head_dim] */
    int32_t mode;
} ref_model;

static uint16_t *gen_tensor(const ref_config *c, uint64_t tid, size_t n, float std, float mean) {
    uint16_t *p = (uint16_t *)malloc(n * sizeof(uint16_t));
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu elems)\n", n); abort(); }
    const uint64_t key = gen_key(c->seed, tid);
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) p[i] = f32_to_bf16(gen_unit_k(key, (uint64_t)i) * std + mean);
    return p;
}

/* RoPE table, computed in double then cast; csrc/engine.cpp builds the identical table on the host. */
REF_API void oa_ref_rope_table(const ref_config *c, int32_t max_pos, float *cosv, float *sinv) {
    int half = c->head_dim / 2;
    for (int i = 0; i < half; ++i) {
        double inv = pow((double)c->rope_theta, -2.0 * i / (double)c->head_dim);
        if (c->rope_scaling == 1) {   /* llama3 frequency scaling (HF _compute_llama3_parameters) */
            double wavelen = 2.0 * M_PI / inv;
            double low_wl = (double)c->rope_orig_ctx / c->rope_low_freq;
            double high_wl = (double)c->rope_orig_ctx / c->rope_high_freq;
            if (wavelen > low_wl) inv = inv / c->rope_factor;
            else if (wavelen >= high_wl) {
                double smooth = ((double)c->rope_orig_ctx / wavelen - c->rope_low_freq) /
                                (c->rope_high_freq - c->rope_low_freq);
                inv = (1.0 - smooth) * inv / c->rope_factor + smooth * inv;
            }
        }
        for (int p = 0; p < max_pos; ++p) {
            double a = (double)p * inv;
            cosv[(size_t)p * half + i] = (float)cos(a);
            sinv[(size_t)p * half + i] = (float)sin(a);
        }
    }
}

REF_API ref_model *oa_ref_create(const ref_config *cfg, int32_t max_pos, int32_t n_slots, int32_t mode) {
    ref_model *m = (ref_model *)calloc(1, sizeof(ref_model));
    m->c = *cfg; m->mode = mode; m->max_pos = max_pos; m->n_slots = n_slots;
    const ref_config *c = &m->c;
    int L = c->n_layers; size_t h = c->hidden, qd = (size_t)c->n_heads * c->head_dim, kd = (size_t)c->n_kv_heads * c->head_dim;
    float nstd = c->norm_random ? 0.1f : 0.0f;
    m->embed = gen_tensor(c, (uint64_t)L * 16 + TG_EMBED, (size_t)c->vocab * h, c->init_std, 0.f);
    m->norm  = gen_tensor(c, (uint64_t)L * 16 + TG_NORM, h, nstd, 1.f);
    m->lm_head = c->tie_embeddings ? m->embed : gen_tensor(c, (uint64_t)L * 16 + TG_LMHEAD, (size_t)c->vocab * h, c->init_std, 0.f);
    uint16_t ***arrs[] = { &m->wq, &m->wk, &m->wv, &m->wo, &m->wg, &m->wu, &m->wd, &m->ln1, &m->ln2, &m->bq, &m->bk, &m->bv };
    for (int a = 0; a < 12; ++a) *arrs[a] = (uint16_t **)calloc(L, sizeof(uint16_t *));
    for (int l = 0; l < L; ++l) {
        uint64_t b = (uint64_t)l * 16;
        m->wq[l] = gen_tensor(c, b + T_WQ, qd * h, c->init_std, 0.f);
        m->wk[l] = gen_tensor(c, b + T_WK, kd * h, c->init_std, 0.f);
        m->wv[l] = gen_tensor(c, b + T_WV, kd * h, c->init_std, 0.f);
        m->wo[l] = gen_tensor(c, b + T_WO, h * qd, c->init_std, 0.f);
        m->wg[l] = gen_tensor(c, b + T_WG, (size_t)c->ffn * h, c->init_std, 0.f);
        m->wu[l] = gen_tensor(c, b + T_WU, (size_t)c->ffn * h, c->init_std, 0.f);
        m->wd[l] = gen_tensor(c, b + T_WD, h * (size_t)c->ffn, c->init_std, 0.f);
        m->ln1[l] = gen_tensor(c, b + T_LN1, h, nstd, 1.f);
        m->ln2[l] = gen_tensor(c, b + T_LN2, h, nstd, 1.f);
        if (c->qkv_bias) {
            m->bq[l] = gen_tensor(c, b + T_BQ, qd, c->init_std, 0.f);
            m->bk[l] = gen_tensor(c, b + T_BK, kd, c->init_std, 0.f);
            m->bv[l] = gen_tensor(c, b + T_BV, kd, c->init_std, 0.f);
        }
    }
    int half = c->head_dim / 2;
    m->rope_cos = (float *)malloc((size_t)max_pos * half * sizeof(float));
    m->rope_sin = (float *)malloc((size_t)max_pos * half * sizeof(float));
    oa_ref_rope_table(c, max_pos, m->rope_cos, m->rope_sin);
    size_t kvn = (size_t)n_slots * L * max_pos * kd;
    m->kcache = (float *)calloc(kvn, sizeof(float));
    m->vcache = (float *)calloc(kvn, sizeof(float));
    return m;
}

REF_API void oa_ref_destroy(ref_model *m) {
    if (!m) return;
    int L = m->c.n_layers;
    uint16_t **arrs[] = { m->wq, m->wk, m->wv, m->wo, m->wg, m->wu, m->wd, m->ln1, m->ln2, m->bq, m->bk, m->bv };
    for (int a = 0; a < 12; ++a) { for (int l = 0; l < L; ++l) free(arrs[a][l]); free(arrs[a]); }
    if (m->lm_head != m->embed) free(m->lm_head);
    free(m->embed); free(m->norm); free(m->rope_cos); free(m->rope_sin); free(m->kcache); free(m->vcache); free(m);
}

/* TIMING AID for bench.py's cpu_baseline leg only (never used by a parity test): fills cache positions [0, n) of `slot` with seeded
 * values of a realistic scale, so that the CPU decode step can be timed at the benchmark's context length (ctx 1664: attention over
 * 1,664 cached tokens per layer) without first spending a minute prefilling them on the host cores. */
REF_API int oa_ref_fill_kv(ref_model *m, int32_t slot, int32_t n, uint64_t seed) {
    if (!m || slot < 0 || slot >= m->n_slots || n < 0 || n > m->max_pos) return -1;
    const size_t kd = (size_t)m->c.n_kv_heads * m->c.head_dim;
    for (int l = 0; l < m->c.n_layers; ++l) {
        float *k = m->kcache + ((size_t)slot * m->c.n_layers + l) * m->max_pos * kd;
        float *v = m->vcache + ((size_t)slot * m->c.n_layers + l) * m->max_pos * kd;
        const uint64_t key = gen_key(seed, 0x4b56ull + (uint64_t)l);
        #pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)((size_t)n * kd); ++i) {
            k[i] = rbf(gen_unit_k(key, (uint64_t)i) * 0.5f);
            v[i] = rbf(gen_unit_k(key, (uint64_t)i + (1ull << 40)) * 0.5f);
        }
    }
    return 0;
}

/* expose weight pointers so the HF cross-check can load the very same tensors */
REF_API const uint16_t *oa_ref_tensor(ref_model *m, int32_t layer, int32_t kind) {
    if (layer < 0) { return kind == TG_EMBED ? m->embed : kind == TG_NORM ? m->norm : m->lm_head; }
    uint16_t **arrs[] = { m->wq, m->wk, m->wv, m->wo, m->wg, m->wu, m->wd, m->ln1, m->ln2, m->bq, m->bk, m->bv };
    return arrs[kind][layer];
}

/* ------------------------------------------------------------------------------------------ */
/* y[T,N] = x[T,K] . W[N,K]^T (+bias), W bf16, fp32 accumulate (SIMD lanes reduced at the end)                   */
static void linear(const float *x, const uint16_t *W, const uint16_t *bias, float *y, int T, int K, int N) {
    #pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const uint16_t *w = W + (size_t)n * K;
        float b = bias ? bf16_to_f32(bias[n]) : 0.f;
        for (int t = 0; t < T; ++t) {
            const float *xr = x + (size_t)t * K;
            float acc = 0.f;
            #pragma omp simd reduction(+:acc)
            for (int k = 0; k < K; ++k) acc += xr[k] * bf16_to_f32(w[k]);
            y[(size_t)t * N + n] = acc + b;
        }
    }
}

static void rmsnorm(const float *x, const uint16_t *g, float *y, int T, int H, float eps, int mode) {
    #pragma omp parallel for schedule(static)
    for (int t = 0; t < T; ++t) {
        const float *xr = x + (size_t)t * H; float *yr = y + (size_t)t * H;
        double ss = 0.0; for (int i = 0; i < H; ++i) ss += (double)xr[i] * xr[i];
        float r = 1.0f / sqrtf((float)(ss / H) + eps);
        for (int i = 0; i < H; ++i) { float v = xr[i] * r * bf16_to_f32(g[i]); yr[i] = mode ? rbf(v) : v; }
    }
}

/* HF rotate_half convention: pairs (i, i+d/2) */
static void rope_inplace(float *v, int n_heads, int d, const float *cosr, const float *sinr, int mode) {
    int half = d / 2;
    for (int hh = 0; hh < n_heads; ++hh) {
        float *p = v + (size_t)hh * d;
        for (int i = 0; i < half; ++i) {
            float a = p[i], b = p[i + half];
            float ra = a * cosr[i] - b * sinr[i], rb = b * cosr[i] + a * sinr[i];
            p[i] = mode ? rbf(ra) : ra; p[i + half] = mode ? rbf(rb) : rb;
        }
    }
}

/*
 * One forward pass over T new tokens of ONE sequence (slot), positions pos0..pos0+T-1, appended to
 * the slot's KV cache.  logits_out: [T_out, vocab] where T_out = all_logits ? T : 1 (last token).
 * hidden_out (optional): final-norm output of the last token [hidden].
 *
 * bf16 rounding points in mode 1 (mirrors the CUDA path's stored tensors):
 *   embedding row; rmsnorm output; q/k/v projections (+bias); RoPE outputs; softmax probabilities
 *   fed to P·V (row sum kept from unrounded p); attention output; residual stream after o_proj and
 *   after down_proj; silu(gate)*up; final norm output.  Logits stay fp32.
 */
REF_API int oa_ref_forward(ref_model *m, int32_t slot, const int32_t *tokens, int32_t T, int32_t pos0,
                           int32_t all_logits, float *logits_out, float *hidden_out) {
    const ref_config *c = &m->c; const int mode = m->mode;
    const int H = c->hidden, L = c->n_layers, nh = c->n_heads, nkv = c->n_kv_heads, d = c->head_dim, F = c->ffn;
    const int qd = nh * d, kd = nkv * d, grp = nh / nkv, half = d / 2;
    if (slot < 0 || slot >= m->n_slots || pos0 + T > m->max_pos) return -1;
    float *x = (float *)malloc((size_t)T * H * 4), *xn = (float *)malloc((size_t)T * H * 4);
    float *q = (float *)malloc((size_t)T * qd * 4), *k = (float *)malloc((size_t)T * kd * 4), *v = (float *)malloc((size_t)T * kd * 4);
    float *att = (float *)malloc((size_t)T * qd * 4), *tmp = (float *)malloc((size_t)T * H * 4);
    float *g = (float *)malloc((size_t)T * F * 4), *u = (float *)malloc((size_t)T * F * 4);
    for (int t = 0; t < T; ++t) {
        int id = tokens[t]; if (id < 0 || id >= c->vocab) { id = 0; }
        for (int i = 0; i < H; ++i) x[(size_t)t * H + i] = bf16_to_f32(m->embed[(size_t)id * H + i]);
    }
    const float scale = 1.0f / sqrtf((float)d);
    for (int l = 0; l < L; ++l) {
        float *kc = m->kcache + ((size_t)slot * L + l) * m->max_pos * kd;
        float *vc = m->vcache + ((size_t)slot * L + l) * m->max_pos * kd;
        rmsnorm(x, m->ln1[l], xn, T, H, c->rms_eps, mode);
        linear(xn, m->wq[l], c->qkv_bias ? m->bq[l] : NULL, q, T, H, qd);
        linear(xn, m->wk[l], c->qkv_bias ? m->bk[l] : NULL, k, T, H, kd);
        linear(xn, m->wv[l], c->qkv_bias ? m->bv[l] : NULL, v, T, H, kd);
        if (mode) { for (size_t i = 0; i < (size_t)T * qd; ++i) q[i] = rbf(q[i]);
                    for (size_t i = 0; i < (size_t)T * kd; ++i) { k[i] = rbf(k[i]); v[i] = rbf(v[i]); } }
        for (int t = 0; t < T; ++t) {
            const float *cr = m->rope_cos + (size_t)(pos0 + t) * half, *sr = m->rope_sin + (size_t)(pos0 + t) * half;
            rope_inplace(q + (size_t)t * qd, nh, d, cr, sr, mode);
            rope_inplace(k + (size_t)t * kd, nkv, d, cr, sr, mode);
            memcpy(kc + (size_t)(pos0 + t) * kd, k + (size_t)t * kd, (size_t)kd * 4);
            memcpy(vc + (size_t)(pos0 + t) * kd, v + (size_t)t * kd, (size_t)kd * 4);
        }
        #pragma omp parallel for collapse(2) schedule(dynamic)
        for (int t = 0; t < T; ++t) for (int hh = 0; hh < nh; ++hh) {
            int n_ctx = pos0 + t + 1, kvh = hh / grp;
            const float *qr = q + (size_t)t * qd + (size_t)hh * d;
            float *s = (float *)malloc((size_t)n_ctx * 4);
            float mx = -INFINITY;
            for (int j = 0; j < n_ctx; ++j) {
                const float *kr = kc + (size_t)j * kd + (size_t)kvh * d; float acc = 0.f;
                for (int i = 0; i < d; ++i) acc += qr[i] * kr[i];
                s[j] = acc * scale; if (s[j] > mx) mx = s[j];
            }
            float sum = 0.f; float *o = att + (size_t)t * qd + (size_t)hh * d;
            for (int i = 0; i < d; ++i) o[i] = 0.f;
            for (int j = 0; j < n_ctx; ++j) {
                float p = expf(s[j] - mx); sum += p; float pr = mode ? rbf(p) : p;
                const float *vr = vc + (size_t)j * kd + (size_t)kvh * d;
                for (int i = 0; i < d; ++i) o[i] += pr * vr[i];
            }
            float inv = 1.0f / sum;
            for (int i = 0; i < d; ++i) { float y = o[i] * inv; o[i] = mode ? rbf(y) : y; }
            free(s);
        }
        linear(att, m->wo[l], NULL, tmp, T, qd, H);
        for (size_t i = 0; i < (size_t)T * H; ++i) { float y = x[i] + tmp[i]; x[i] = mode ? rbf(y) : y; }
        rmsnorm(x, m->ln2[l], xn, T, H, c->rms_eps, mode);
        linear(xn, m->wg[l], NULL, g, T, H, F);
        linear(xn, m->wu[l], NULL, u, T, H, F);
        for (size_t i = 0; i < (size_t)T * F; ++i) {
            float gv = g[i]; float sv = gv / (1.0f + expf(-gv)); float y = sv * u[i]; g[i] = mode ? rbf(y) : y;
        }
        linear(g, m->wd[l], NULL, tmp, T, F, H);
        for (size_t i = 0; i < (size_t)T * H; ++i) { float y = x[i] + tmp[i]; x[i] = mode ? rbf(y) : y; }
    }
    rmsnorm(x, m->norm, xn, T, H, c->rms_eps, mode);
    if (hidden_out) memcpy(hidden_out, xn + (size_t)(T - 1) * H, (size_t)H * 4);
    if (logits_out) {
        if (all_logits) linear(xn, m->lm_head, NULL, logits_out, T, H, c->vocab);
        else linear(xn + (size_t)(T - 1) * H, m->lm_head, NULL, logits_out, 1, H, c->vocab);
    }
    free(x); free(xn); free(q); free(k); free(v); free(att); free(tmp); free(g); free(u);
    return 0;
}

/* greedy argmax, ties -> lowest id (what the engine's fused argmax implements) */
REF_API int32_t oa_ref_argmax(const float *logits, int32_t n) {
    int32_t best = 0; float bv = logits[0];
    for (int32_t i = 1; i < n; ++i) if (logits[i] > bv) { bv = logits[i]; best = i; }
    return best;
}

/*
 * Greedy generation for one sequence: prefill `prompt`, then decode up to max_new tokens, stopping
 * at any id in eos[] (the EOS token is not emitted).  Returns number of generated tokens.
 * margins_out (optional) receives top1-top2 logit margin per generated token, so parity tests can
 * require token equality only where the margin exceeds the stated tolerance.
 */
REF_API int32_t oa_ref_generate(ref_model *m, int32_t slot, const int32_t *prompt, int32_t n_prompt, int32_t max_new,
                                const int32_t *eos, int32_t n_eos, int32_t *out_tokens, float *margins_out,
                                float *first_logits_out) {
    int V = m->c.vocab; float *logits = (float *)malloc((size_t)V * 4);
    int pos = 0, n_out = 0;
    if (oa_ref_forward(m, slot, prompt, n_prompt, 0, 0, logits, NULL) != 0) { free(logits); return -1; }
    pos = n_prompt;
    if (first_logits_out) memcpy(first_logits_out, logits, (size_t)V * 4);
    while (n_out < max_new) {
        int32_t tok = oa_ref_argmax(logits, V);
        if (margins_out) { float second = -INFINITY; for (int i = 0; i < V; ++i) if (i != tok && logits[i] > second) second = logits[i];
                           margins_out[n_out] = logits[tok] - second; }
        int stop = 0; for (int e = 0; e < n_eos; ++e) if (tok == eos[e]) stop = 1;
        if (stop) break;
        out_tokens[n_out++] = tok;
        if (n_out == max_new || pos + 1 > m->max_pos - 1) break;
        if (oa_ref_forward(m, slot, &tok, 1, pos, 0, logits, NULL) != 0) break;
        pos += 1;
    }
    free(logits);
    return n_out;
}

/* ---- standalone kernels' restatements, used by kernel-level parity tests ---- */
REF_API void oa_ref_rmsnorm(const float *x, const uint16_t *g, float *y, int32_t T, int32_t H, float eps, int32_t mode) { rmsnorm(x, g, y, T, H, eps, mode); }
REF_API void oa_ref_linear(const float *x, const uint16_t *W, const uint16_t *bias, float *y, int32_t T, int32_t K, int32_t N) { linear(x, W, bias, y, T, K, N); }
REF_API void oa_ref_set_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
REF_API int32_t oa_ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
