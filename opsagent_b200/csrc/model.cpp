// model.cpp — see model.hpp.
#include "model.hpp"
#include "safetensors.hpp"

#include <algorithm>
#include <cmath>
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace oa {

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// tensor ids of the seeded generator: layer*16 + kind; globals at layer index n_layers (oracle/llama_ref.c)
enum { T_WQ = 0, T_WK, T_WV, T_WO, T_WG, T_WU, T_WD, T_LN1, T_LN2, T_BQ, T_BK, T_BV };
enum { TG_EMBED = 0, TG_NORM = 1, TG_LMHEAD = 2 };

void* DeviceModel::dmalloc(size_t bytes) {
    void* p = nullptr;
    cuda_check(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc");
    allocs_.push_back(p);
    return p;
}

void DeviceModel::make_weight_maps(WeightMat& w) {
    const int bns[4] = {32, 64, 128, 256};
    for (int i = 0; i < 4; ++i) {
        int r = make_tmap_bf16_2d(&w.tm[i], w.ptr, (uint64_t)w.N, (uint64_t)w.K, (uint64_t)w.K, (uint32_t)bns[i], 64);
        if (r != 0) throw std::runtime_error("cuTensorMapEncodeTiled failed for a weight matrix (code " + std::to_string(r) + ")");
    }
}
void DeviceModel::alloc_weight(WeightMat& w, int N, int K) {
    w.N = N; w.K = K; w.ptr = dmalloc((size_t)N * K * 2); weight_bytes += (size_t)N * K * 2;
    make_weight_maps(w);
}

// Build the RoPE table exactly as oracle/llama_ref.c:oa_ref_rope_table does (double math, then cast).
static void rope_table(const ModelConfig& c, int max_pos, std::vector<float>& cosv, std::vector<float>& sinv) {
    const int half = c.head_dim / 2;
    cosv.resize((size_t)max_pos * half); sinv.resize((size_t)max_pos * half);
    for (int i = 0; i < half; ++i) {
        double inv = std::pow((double)c.rope_theta, -2.0 * i / (double)c.head_dim);
        if (c.rope_scaling == 1) {
            double wavelen = 2.0 * M_PI / inv;
            double low_wl = (double)c.rope_orig_ctx / c.rope_low_freq, high_wl = (double)c.rope_orig_ctx / c.rope_high_freq;
            if (wavelen > low_wl) inv = inv / c.rope_factor;
            else if (wavelen >= high_wl) {
                double smooth = ((double)c.rope_orig_ctx / wavelen - c.rope_low_freq) / (c.rope_high_freq - c.rope_low_freq);
                inv = (1.0 - smooth) * inv / c.rope_factor + smooth * inv;
            }
        }
        for (int p = 0; p < max_pos; ++p) {
            double a = (double)p * inv;
            cosv[(size_t)p * half + i] = (float)std::cos(a); sinv[(size_t)p * half + i] = (float)std::sin(a);
        }
    }
}

// The CUDA "current device" is per host thread.  Creating or destroying an engine must not leave the CALLER's thread on another device (a
// router process owns one engine per GPU; torch / NCCL in the same process assume their own device stays current), so both restore it.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

DeviceModel::DeviceModel(const ModelConfig& c, const EngineOptions& o) : cfg(c), opt(o) {
    DeviceGuard guard(opt.device);
    cuda_check(cudaSetDevice(opt.device), "cudaSetDevice");
    cudaDeviceProp prop;
    cuda_check(cudaGetDeviceProperties(&prop, opt.device), "cudaGetDeviceProperties");
    if (prop.major != 10) throw std::runtime_error("opsagent_b200 requires an sm_100 (Blackwell B200) device; found sm_" +
                                                   std::to_string(prop.major) + std::to_string(prop.minor));
    sm_count = prop.multiProcessorCount;
    cuda_check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
    tp = opt.tp; tp_rank = opt.tp_rank;
    nh_l = cfg.n_heads / tp; nkv_l = cfg.n_kv_heads / tp; F_l = cfg.ffn / tp; V_l = cfg.vocab / tp;
    const int H = cfg.hidden, L = cfg.n_layers, F = F_l, V = cfg.vocab, D = cfg.head_dim;
    const int qd = nh_l * D, kd = nkv_l * D, qkvd = qd + 2 * kd;              // this rank's shard
    const int qd_g = cfg.q_dim(), F_g = cfg.ffn;                              // logical (global) sizes for the weight generator
    const int64_t r = tp_rank;
    const float nstd = cfg.norm_random ? 0.1f : 0.0f;

    // ---- weights: seeded init directly in HBM, bit-identical to the oracle's generator ----
    embed = dmalloc((size_t)V * H * 2); weight_bytes += (size_t)V * H * 2;
    cuda_check(launch_init_weight(embed, cfg.seed, (uint64_t)L * 16 + TG_EMBED, -1, V, H, cfg.init_std, 0.f, stream), "init embed");
    final_norm = dmalloc((size_t)H * 2);
    cuda_check(launch_init_weight(final_norm, cfg.seed, (uint64_t)L * 16 + TG_NORM, -1, 1, H, nstd, 1.f, stream), "init norm");
    // LM head is vocab-parallel: this rank owns rows [r*V_l, (r+1)*V_l)
    if (cfg.tie_embeddings) { lm_head.ptr = reinterpret_cast<uint16_t*>(embed) + (size_t)r * V_l * H; lm_head.N = V_l; lm_head.K = H; make_weight_maps(lm_head); }
    else { alloc_weight(lm_head, V_l, H); cuda_check(launch_init_weight(lm_head.ptr, cfg.seed, (uint64_t)L * 16 + TG_LMHEAD, -1, V_l, H, cfg.init_std, 0.f, stream, r * V_l), "init lm_head"); }
    layers.resize(L);
    for (int l = 0; l < L; ++l) {
        Layer& ly = layers[l]; const uint64_t b = (uint64_t)l * 16;
        alloc_weight(ly.qkv, qkvd, H); alloc_weight(ly.o, H, qd); alloc_weight(ly.gu, 2 * F, H); alloc_weight(ly.down, H, F);
        uint16_t* wq = reinterpret_cast<uint16_t*>(ly.qkv.ptr);
        // column-parallel q|k|v and gate|up (row shards), row-parallel o and down (column shards of the logical tensors)
        cuda_check(launch_init_weight(wq, cfg.seed, b + T_WQ, -1, qd, H, cfg.init_std, 0.f, stream, r * qd), "init wq");
        cuda_check(launch_init_weight(wq + (size_t)qd * H, cfg.seed, b + T_WK, -1, kd, H, cfg.init_std, 0.f, stream, r * kd), "init wk");
        cuda_check(launch_init_weight(wq + (size_t)(qd + kd) * H, cfg.seed, b + T_WV, -1, kd, H, cfg.init_std, 0.f, stream, r * kd), "init wv");
        cuda_check(launch_init_weight(ly.o.ptr, cfg.seed, b + T_WO, -1, H, qd, cfg.init_std, 0.f, stream, 0, r * qd, qd_g), "init wo");
        cuda_check(launch_init_weight(ly.gu.ptr, cfg.seed, b + T_WG, (int64_t)(b + T_WU), 2 * (int64_t)F, H, cfg.init_std, 0.f, stream, r * F), "init wgu");
        cuda_check(launch_init_weight(ly.down.ptr, cfg.seed, b + T_WD, -1, H, F, cfg.init_std, 0.f, stream, 0, r * F, F_g), "init wd");
        ly.ln1 = dmalloc((size_t)H * 2); ly.ln2 = dmalloc((size_t)H * 2);
        cuda_check(launch_init_weight(ly.ln1, cfg.seed, b + T_LN1, -1, 1, H, nstd, 1.f, stream), "init ln1");
        cuda_check(launch_init_weight(ly.ln2, cfg.seed, b + T_LN2, -1, 1, H, nstd, 1.f, stream), "init ln2");
        if (cfg.qkv_bias) {
            ly.bqkv = dmalloc((size_t)qkvd * 2);
            uint16_t* bq = reinterpret_cast<uint16_t*>(ly.bqkv);
            cuda_check(launch_init_weight(bq, cfg.seed, b + T_BQ, -1, 1, qd, cfg.init_std, 0.f, stream, 0, r * qd, qd_g), "init bq");
            cuda_check(launch_init_weight(bq + qd, cfg.seed, b + T_BK, -1, 1, kd, cfg.init_std, 0.f, stream, 0, r * kd, cfg.kv_dim()), "init bk");
            cuda_check(launch_init_weight(bq + qd + kd, cfg.seed, b + T_BV, -1, 1, kd, cfg.init_std, 0.f, stream, 0, r * kd, cfg.kv_dim()), "init bv");
        }
    }

    if (!opt.weights.empty()) load_checkpoint(opt.weights);

    // ---- RoPE table ----
    {
        std::vector<float> cs, sn; rope_table(cfg, opt.max_seq_len, cs, sn);
        rope_cos = reinterpret_cast<float*>(dmalloc(cs.size() * 4)); rope_sin = reinterpret_cast<float*>(dmalloc(sn.size() * 4));
        cuda_check(cudaMemcpyAsync(rope_cos, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice, stream), "rope cos");
        cuda_check(cudaMemcpyAsync(rope_sin, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice, stream), "rope sin");
        cuda_check(cudaStreamSynchronize(stream), "rope sync");
    }

    // ---- activations ----
    max_rows_ = std::max(opt.max_step_tokens, opt.max_batch);
    max_sample_ = std::max(opt.max_batch, std::min(max_rows_, 2048));
    const int MR = max_rows_;
    x_ = dmalloc((size_t)MR * H * 2); xn_ = dmalloc((size_t)MR * H * 2); qkv_ = dmalloc((size_t)MR * qkvd * 2);
    q_ = dmalloc((size_t)MR * qd * 2); attn_ = dmalloc((size_t)MR * qd * 2); act_ = dmalloc((size_t)MR * F * 2);
    cuda_check(cudaMemsetAsync(xn_, 0, (size_t)MR * H * 2, stream), "memset");
    cuda_check(cudaMemsetAsync(attn_, 0, (size_t)MR * qd * 2, stream), "memset");
    cuda_check(cudaMemsetAsync(act_, 0, (size_t)MR * F * 2, stream), "memset");
    auto amap = [&](CUtensorMap* tm, void* p, int rows, int cols) {
        int r = make_tmap_bf16_2d(tm, p, (uint64_t)rows, (uint64_t)cols, (uint64_t)cols, 128, 64);
        if (r != 0) throw std::runtime_error("cuTensorMapEncodeTiled failed for an activation buffer (code " + std::to_string(r) + ")");
    };
    amap(&tm_xn_, xn_, MR, H); amap(&tm_attn_, attn_, MR, qd); amap(&tm_act_, act_, MR, F); amap(&tm_q_, q_, MR, qd);
    cuda_check(cudaMemsetAsync(q_, 0, (size_t)MR * qd * 2, stream), "memset");
    // sampled rows (debug logits sample every row of a short prefill: up to 2048)
    const int MS = max_sample_;
    xs_ = dmalloc((size_t)MS * H * 2); xsn_ = dmalloc((size_t)MS * H * 2);
    cuda_check(cudaMemsetAsync(xsn_, 0, (size_t)MS * H * 2, stream), "memset");
    amap(&tm_xsn_, xsn_, MS, H);
    const int lm_tiles_max = gemm_n_tiles(V_l, 32);
    amax_val_ = reinterpret_cast<float*>(dmalloc((size_t)MS * lm_tiles_max * 4));
    amax_idx_ = reinterpret_cast<int*>(dmalloc((size_t)MS * lm_tiles_max * 4));
    d_out_ids_ = reinterpret_cast<int32_t*>(dmalloc((size_t)MS * 4));
    // token-mask table of the grammar-constrained decoder: one bitset over the whole vocabulary per cached automaton state
    mask_words = (cfg.vocab + 31) / 32;
    mask_slots = std::max(1024, 2 * opt.max_batch);
    max_mask_updates = opt.max_batch;               // a scheduler step samples at most one row per running sequence
    mask_table_ = reinterpret_cast<uint32_t*>(dmalloc((size_t)mask_slots * mask_words * 4));
    cuda_check(cudaMemsetAsync(mask_table_, 0, (size_t)mask_slots * mask_words * 4, stream), "mask table memset");
    cuda_check(cudaMallocHost(&h_out_ids, (size_t)MS * 4), "cudaMallocHost");

    // ---- paged KV pool ----
    const size_t page_bytes = (size_t)2 * L * nkv_l * 64 * D * 2;    // all layers, K and V, of 64 tokens (this rank's kv heads)
    if (opt.num_pages > 0) num_pages = opt.num_pages;
    else {
        size_t free_b = 0, total_b = 0;
        cuda_check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        double budget = opt.kv_gb > 0 ? opt.kv_gb * 1e9 : (double)free_b - 6e9;
        budget = std::min(budget, (double)free_b - 2e9);
        num_pages = (int)std::max(0.0, budget / (double)page_bytes);
    }
    max_pages_per_seq = opt.max_seq_len / 64;
    if (num_pages < max_pages_per_seq) throw std::runtime_error("KV pool too small: " + std::to_string(num_pages) + " pages < one max-length sequence");
    const int64_t plane_rows = (int64_t)num_pages * nkv_l * 64;
    if ((double)plane_rows * 2 * L >= 2147483647.0) { num_pages = (int)(2147483647.0 / (2.0 * L * nkv_l * 64)) - 1; }
    kv.page_size = 64; kv.n_kv = nkv_l; kv.head_dim = D; kv.num_pages = num_pages;
    kv.kv_stride_rows = (int64_t)num_pages * nkv_l * 64; kv.layer_stride_rows = 2 * kv.kv_stride_rows;
    kv_pool_bytes = (size_t)num_pages * page_bytes;
    kv.base = dmalloc(kv_pool_bytes);
    cuda_check(cudaMemsetAsync(kv.base, 0, kv_pool_bytes, stream), "kv memset");   // stale pages must hold finite values (0*NaN)
    {
        int r = make_tmap_bf16_2d(&tm_kv, kv.base, (uint64_t)kv.layer_stride_rows * L, (uint64_t)D, (uint64_t)D, 64, 64);
        if (r != 0) throw std::runtime_error("cuTensorMapEncodeTiled failed for the KV pool (code " + std::to_string(r) + ")");
    }

    // ---- decode-attention partial workspace + per-step metadata arena ----
    const int n_ctas = opt.attn_ctas > 0 ? opt.attn_ctas : 2 * sm_count;
    max_part_slots_ = 2 * n_ctas + 2 * opt.max_batch * nkv_l + 16;
    const int grp = cfg.n_heads / cfg.n_kv_heads;
    part_o_ = reinterpret_cast<float*>(dmalloc((size_t)max_part_slots_ * grp * D * 4));
    part_ml_ = reinterpret_cast<float*>(dmalloc((size_t)max_part_slots_ * grp * 2 * 4));
    meta_cap_words_ = (size_t)4 * MR + (size_t)opt.max_batch * (max_pages_per_seq + 2) + (size_t)MR / 64 * 4 + 4 * opt.max_batch +
                      (size_t)(opt.max_batch * nkv_l + n_ctas + 8) * 8 * 2 + (size_t)n_ctas + 4096 + (size_t)max_sample_ + (size_t)max_mask_updates * (mask_words + 4);
    for (int i = 0; i < 2; ++i) {
        cuda_check(cudaMallocHost(&h_meta_buf_[i], meta_cap_words_ * 4), "cudaMallocHost meta");
        cuda_check(cudaEventCreateWithFlags(&meta_ev_[i], cudaEventDisableTiming), "cudaEventCreate meta");
    }
    h_meta_ = h_meta_buf_[0];
    d_meta_ = reinterpret_cast<int32_t*>(dmalloc(meta_cap_words_ * 4));
    sk_bn_ = (opt.sk_bn == 256) ? 256 : 128;
    sk_G_ = opt.sk_ctas > 0 ? opt.sk_ctas : sm_count;
    {
        size_t wsb = 0;
        for (int n : {qkvd, H, 2 * F}) { wsb = std::max(wsb, streamk_ws_bytes(n, 256, sk_G_, 128)); wsb = std::max(wsb, streamk_ws_bytes(n, 128, sk_G_, 256)); wsb = std::max(wsb, streamk_ws_bytes(n, 64, sk_G_, 128)); }
        sk_ws_ = reinterpret_cast<float*>(dmalloc(wsb));
        tile_flags_ = reinterpret_cast<unsigned int*>(dmalloc((size_t)4 * SK_MAX_FLAG_TILES * 4));
        cuda_check(cudaMemset(tile_flags_, 0, (size_t)4 * SK_MAX_FLAG_TILES * 4), "stream-K tile flags memset");
        // 32-bit index arithmetic in the consumers: (units of the largest shape + kb) * G must stay below 2^32
        const double worst = ((double)((2 * F + sk_bn_ - 1) / sk_bn_) * ((std::max(H, F) + 63) / 64) + (std::max(H, F) + 63) / 64) * sk_G_;
        if (worst >= 4.0e9) throw std::runtime_error("stream-K index range exceeds 32 bits for this model");
    }
    ev.resize(2 * (size_t)L);
    for (auto& e : ev) cuda_check(cudaEventCreate(&e), "cudaEventCreate");
    if (tp > 1) {
        if (opt.num_pages <= 0) throw std::runtime_error("tensor parallel engines need an explicit num_pages (all ranks must agree on the pool size)");
        // bf16 prefill partials | fp32 decode partials | up to 256 rows of this rank's fp32 logits shard (debug parity hook)
        const size_t sym_bytes = std::max({(size_t)MR * H * 2, (size_t)256 * H * 4, (size_t)256 * V_l * 4});
        // NVLS buffer: two partial + two reduced regions of 128 rows x H fp32 (decode batches: <= 128 rows on the stream-K path)
        nvls_region_ = (size_t)128 * H * 4;
        comm.reset(new TpComm(tp, tp_rank, opt.tp_shm, sym_bytes, max_sample_, opt.tp_nonce, opt.tp_nvls ? 4 * nvls_region_ + 4096 : 0));      // + per-row flags: [2 buffers][128 rows] uint32 epochs
    }
    cuda_check(cudaStreamSynchronize(stream), "init sync");
}

// ---- Hugging Face checkpoint loading (Llama / Qwen2 naming) --------------------------------------------------------
static void to_bf16_rows(const StTensor& t, int64_t C, int64_t row, int64_t col0, int64_t cols, uint16_t* dst) {
    if (t.dtype == "BF16") { std::memcpy(dst, reinterpret_cast<const uint16_t*>(t.data) + row * C + col0, (size_t)cols * 2); return; }
    if (t.dtype == "F32") { const float* s = reinterpret_cast<const float*>(t.data) + row * C + col0; for (int64_t i = 0; i < cols; ++i) dst[i] = f32_to_bf16_bits(s[i]); return; }
    if (t.dtype == "F16") {
        const uint16_t* s = reinterpret_cast<const uint16_t*>(t.data) + row * C + col0;
        for (int64_t i = 0; i < cols; ++i) {
            const uint32_t h = s[i], sign = (h & 0x8000u) << 16, e = (h >> 10) & 31, m = h & 1023; uint32_t f;
            if (e == 0) { if (m == 0) f = sign; else { int sh = 0; uint32_t mm = m; while (!(mm & 1024)) { mm <<= 1; ++sh; } f = sign | ((uint32_t)(113 - sh) << 23) | ((mm & 1023) << 13); } }
            else if (e == 31) f = sign | 0x7f800000u | (m << 13);
            else f = sign | ((e + 112) << 23) | (m << 13);
            float x; std::memcpy(&x, &f, 4); dst[i] = f32_to_bf16_bits(x);
        }
        return;
    }
    throw std::runtime_error("unsupported checkpoint dtype " + t.dtype);
}

void DeviceModel::load_checkpoint(const std::string& path) {
    SafeTensors st(path);
    const int H = cfg.hidden, L = cfg.n_layers, D = cfg.head_dim, V = cfg.vocab;
    const int qd = nh_l * D, kd = nkv_l * D;
    const int64_t r = tp_rank;
    std::vector<uint16_t> buf;
    // every tensor must have exactly the logical shape [R, C] (vectors: [C]) this model configuration implies — a transposed, truncated or
    // wrong-config (hidden / ffn / heads / vocab) checkpoint is refused with a message instead of being read out of bounds
    auto checked = [&](const std::string& name, int64_t R, int64_t C) -> const StTensor& {
        const StTensor& t = st.get(name);
        const bool ok = (t.shape.size() == 2 && t.shape[0] == R && t.shape[1] == C) || (R == 1 && t.shape.size() == 1 && t.shape[0] == C);
        if (!ok) {
            std::string got = "["; for (size_t i = 0; i < t.shape.size(); ++i) got += (i ? ", " : "") + std::to_string(t.shape[i]); got += "]";
            throw std::runtime_error("checkpoint tensor " + name + " has shape " + got + ", this model needs [" + (R == 1 ? "" : std::to_string(R) + ", ") + std::to_string(C) + "]");
        }
        if (t.dtype != "BF16" && t.dtype != "F16" && t.dtype != "F32") throw std::runtime_error("checkpoint tensor " + name + " has unsupported dtype " + t.dtype);
        return t;       // SafeTensors already verified bytes == prod(shape) * sizeof(dtype)
    };
    // logical [R, C] tensor `name`: rows [row0, row0+rows) x cols [col0, col0+cols) -> device dst (row-major [rows, cols])
    auto slice = [&](const std::string& name, int64_t R, int64_t C, int64_t row0, int64_t rows, int64_t col0, int64_t cols, void* dst) {
        const StTensor& t = checked(name, R, C);
        if (row0 < 0 || col0 < 0 || row0 + rows > R || col0 + cols > C) throw std::runtime_error("checkpoint slice out of range for " + name);
        buf.resize((size_t)rows * cols);
        for (int64_t i = 0; i < rows; ++i) to_bf16_rows(t, C, row0 + i, col0, cols, buf.data() + (size_t)i * cols);
        cuda_check(cudaMemcpy(dst, buf.data(), buf.size() * 2, cudaMemcpyHostToDevice), "checkpoint H2D");
    };
    slice("model.embed_tokens.weight", V, H, 0, V, 0, H, embed);
    slice("model.norm.weight", 1, H, 0, 1, 0, H, final_norm);
    if (!cfg.tie_embeddings) slice(st.has("lm_head.weight") ? "lm_head.weight" : "model.embed_tokens.weight", V, H, r * V_l, V_l, 0, H, lm_head.ptr);
    for (int l = 0; l < L; ++l) {
        Layer& ly = layers[l]; const std::string p = "model.layers." + std::to_string(l) + ".";
        uint16_t* wq = reinterpret_cast<uint16_t*>(ly.qkv.ptr);
        slice(p + "self_attn.q_proj.weight", cfg.q_dim(), H, r * qd, qd, 0, H, wq);
        slice(p + "self_attn.k_proj.weight", cfg.kv_dim(), H, r * kd, kd, 0, H, wq + (size_t)qd * H);
        slice(p + "self_attn.v_proj.weight", cfg.kv_dim(), H, r * kd, kd, 0, H, wq + (size_t)(qd + kd) * H);
        slice(p + "self_attn.o_proj.weight", H, cfg.q_dim(), 0, H, r * qd, qd, ly.o.ptr);
        slice(p + "mlp.down_proj.weight", H, cfg.ffn, 0, H, r * F_l, F_l, ly.down.ptr);
        slice(p + "input_layernorm.weight", 1, H, 0, 1, 0, H, ly.ln1);
        slice(p + "post_attention_layernorm.weight", 1, H, 0, 1, 0, H, ly.ln2);
        if (cfg.qkv_bias) {
            uint16_t* bq = reinterpret_cast<uint16_t*>(ly.bqkv);
            slice(p + "self_attn.q_proj.bias", 1, cfg.q_dim(), 0, 1, r * qd, qd, bq);
            slice(p + "self_attn.k_proj.bias", 1, cfg.kv_dim(), 0, 1, r * kd, kd, bq + qd);
            slice(p + "self_attn.v_proj.bias", 1, cfg.kv_dim(), 0, 1, r * kd, kd, bq + qd + kd);
        }
        // gate/up: physical rows interleaved in blocks of 16 (16 gate rows, 16 up rows)
        const StTensor& tg = checked(p + "mlp.gate_proj.weight", cfg.ffn, H); const StTensor& tu = checked(p + "mlp.up_proj.weight", cfg.ffn, H);
        buf.resize((size_t)2 * F_l * H);
        for (int64_t blk = 0; blk < F_l / 16; ++blk)
            for (int64_t w = 0; w < 32; ++w)
                to_bf16_rows(w < 16 ? tg : tu, H, r * F_l + blk * 16 + (w & 15), 0, H, buf.data() + (size_t)(blk * 32 + w) * H);
        cuda_check(cudaMemcpy(ly.gu.ptr, buf.data(), buf.size() * 2, cudaMemcpyHostToDevice), "checkpoint H2D");
    }
}

DeviceModel::~DeviceModel() {
    DeviceGuard guard(opt.device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto& e : ev) cudaEventDestroy(e);
    for (void* p : allocs_) cudaFree(p);
    if (h_out_ids) cudaFreeHost(h_out_ids);
    for (int i = 0; i < 2; ++i) { if (h_meta_buf_[i]) cudaFreeHost(h_meta_buf_[i]); if (meta_ev_[i]) cudaEventDestroy(meta_ev_[i]); }
    comm.reset();                 // peer mappings and symmetric buffers: released while this engine's device is still current
    if (stream) cudaStreamDestroy(stream);
}

const char* DeviceModel::kt_name(int id) {
    static const char* n[16] = {"h2d+embed", "rmsnorm", "qkv_gemm", "rope_kv", "attention", "attn_merge", "o_gemm", "resid_rmsnorm", "gate_up_gemm",
                                "swiglu", "down_gemm", "gather+norm", "lm_head", "argmax+d2h", "", ""};
    return n[id & 15];
}
void DeviceModel::sync() { cuda_check(cudaStreamSynchronize(stream), "stream sync"); }

// N-tile heuristic.  Decode (M <= 128) streams weights once: prefer the largest tile that still yields >= 1 CTA
// per SM; prefill has M/128 row tiles as well, so the widest tile (least A re-reads) wins.
int DeviceModel::pick_bn(int M, int N, int override_bn, bool swiglu) const {
    if (override_bn == 32 || override_bn == 64 || override_bn == 128 || override_bn == 256) return override_bn;
    const int m_tiles = (M + 127) / 128;
    const int cands[4] = {256, 128, 64, 32};
    // one or two row tiles (decode-sized batches off the stream-K path, e.g. Qwen2.5-32B TP=4 at B=256): a CTA's k-loop is latency-bound
    // (~0.3 us per k-block), so wider tiles with fewer, longer-lived CTAs win as long as about half the SMs stream — measured per shape in
    // profiles/r02e_gemm_shapes_m256.md (gate_up 51 -> 39 us at BN=256, down 36 -> 33 us at BN=128); many row tiles: fill every SM
    const long long want = m_tiles <= 2 ? sm_count / 2 : sm_count;
    for (int bn : cands) { if ((long long)((N + bn - 1) / bn) * m_tiles >= want) return bn; }
    (void)swiglu;
    return 32;
}

void build_decode_plan(const int32_t* ctx_lens, int n_seqs, int n_kv, int n_ctas_target, int force_splits, DecodePlan& out) {
    out.segs.clear(); out.cta_ptr.clear(); out.merges.clear(); out.n_slots = 0;
    long long total = 0;
    for (int i = 0; i < n_seqs; ++i) total += (long long)((ctx_lens[i] + 63) / 64) * n_kv;
    if (total <= 0) { out.cta_ptr.push_back(0); return; }
    if (force_splits > 0) {
        // test mode: every (seq, kvh) item is cut into `force_splits` pieces, one CTA per piece
        out.cta_ptr.push_back(0);
        for (int s = 0; s < n_seqs; ++s) for (int h = 0; h < n_kv; ++h) {
            const int nch = (ctx_lens[s] + 63) / 64; if (nch == 0) continue;
            const int pieces = std::min(force_splits, nch), per = (nch + pieces - 1) / pieces;
            const int real = (nch + per - 1) / per;
            MergeItem mi{s, h, out.n_slots, real};
            const int slot0 = out.n_slots, item = (int)out.merges.size();
            for (int pz = 0; pz < real; ++pz) {
                DecodeSeg sg{s, h, pz * per, std::min(nch, (pz + 1) * per), real > 1 ? out.n_slots++ : -1, slot0, real, item};
                out.segs.push_back(sg); out.cta_ptr.push_back((int32_t)out.segs.size());
            }
            if (real > 1) out.merges.push_back(mi);
        }
        return;
    }
    long long n_ctas = std::min<long long>(n_ctas_target, std::max<long long>(1, total / 2));
    const long long per = (total + n_ctas - 1) / n_ctas;
    n_ctas = (total + per - 1) / per;
    out.cta_ptr.assign((size_t)n_ctas + 1, 0);
    long long gpos = 0;   // global chunk cursor
    std::vector<int> cta_of_seg;
    for (int s = 0; s < n_seqs; ++s) for (int h = 0; h < n_kv; ++h) {
        const int nch = (ctx_lens[s] + 63) / 64; if (nch == 0) continue;
        const long long g0 = gpos, g1 = gpos + nch;
        const long long c_first = g0 / per, c_last = (g1 - 1) / per;
        const int pieces = (int)(c_last - c_first + 1);
        MergeItem mi{s, h, out.n_slots, pieces};
        const int slot0 = out.n_slots, item = (int)out.merges.size();
        for (long long c = c_first; c <= c_last; ++c) {
            const long long a = std::max(g0, c * per), b = std::min(g1, (c + 1) * per);
            DecodeSeg sg{s, h, (int32_t)(a - g0), (int32_t)(b - g0), pieces > 1 ? out.n_slots++ : -1, slot0, pieces, item};
            out.segs.push_back(sg); cta_of_seg.push_back((int)c);
        }
        if (pieces > 1) out.merges.push_back(mi);
        gpos = g1;
    }
    // segs are already ordered by CTA; build the CSR pointer
    size_t k = 0;
    for (long long c = 0; c < n_ctas; ++c) {
        out.cta_ptr[(size_t)c] = (int32_t)k;
        while (k < cta_of_seg.size() && cta_of_seg[k] == c) ++k;
    }
    out.cta_ptr[(size_t)n_ctas] = (int32_t)k;
}

void DeviceModel::forward(const StepInput& in, float* logits_out) {
    const int T = (int)in.tokens.size(), S = (int)in.sample_rows.size();
    // all projection sizes below are this rank's shard (== the global sizes when tp == 1)
    const int H = cfg.hidden, L = cfg.n_layers, F = F_l, V = V_l, D = cfg.head_dim;
    const int nh = nh_l, nkv = nkv_l, qd = nh * D, qkvd = qd + 2 * nkv * D;
    const bool use_tp = tp > 1;
    if (T <= 0 || T > max_rows_ || S > max_sample_) throw std::runtime_error("forward: bad batch size");
    if (in.decode && in.n_seqs != T) throw std::runtime_error("forward: decode needs one token per sequence");

    // ---- pack per-step metadata into one pinned arena, one H2D copy ----
    meta_idx_ ^= 1; h_meta_ = h_meta_buf_[meta_idx_];
    cuda_check(cudaEventSynchronize(meta_ev_[meta_idx_]), "meta staging fence");     // its previous H2D copy has been consumed
    size_t w = 0;
    auto put = [&](const void* src, size_t words) -> size_t {
        size_t at = w; if (at + words > meta_cap_words_) throw std::runtime_error("forward: metadata arena overflow");
        if (words) std::memcpy(h_meta_ + at, src, words * 4);
        w = (at + words + 3) & ~size_t(3); return at;
    };
    const size_t o_tok = put(in.tokens.data(), T), o_pos = put(in.positions.data(), T), o_slot = put(in.slots.data(), T);
    const size_t o_samp = put(in.sample_rows.data(), S);
    const size_t o_bt = put(in.block_tables.data(), in.block_tables.size());
    const size_t o_ctx = put(in.ctx_lens.data(), in.ctx_lens.size());
    if (!in.mask_slots.empty() && (int)in.mask_slots.size() != S) throw std::runtime_error("forward: mask_slots must be [n_sample]");
    const size_t rec = (size_t)mask_words + 1, n_upd = in.mask_updates.size() / rec;
    if (in.mask_updates.size() % rec != 0 || (int)n_upd > max_mask_updates) throw std::runtime_error("forward: malformed or too many token-mask updates");
    for (int ms : in.mask_slots) if (ms >= mask_slots) throw std::runtime_error("forward: token-mask slot out of range");
    const size_t o_mslot = put(in.mask_slots.data(), in.mask_slots.size());
    const size_t o_mupd = put(in.mask_updates.data(), in.mask_updates.size());
    size_t o_segs = 0, o_ptr = 0, o_tiles = 0, o_cnt = 0;
    const int n_dec = in.decode ? in.n_seqs : in.n_decode;          // sequences served by the decode attention kernel (rows 0..n_dec-1)
    if (n_dec < 0 || n_dec > in.n_seqs || n_dec > T) throw std::runtime_error("forward: bad n_decode");
    if (n_dec > 0) {
        const int n_ctas = opt.attn_ctas > 0 ? opt.attn_ctas : 2 * sm_count;
        build_decode_plan(in.ctx_lens.data(), n_dec, nkv, n_ctas, 0, plan_);
        if (plan_.n_slots > max_part_slots_) throw std::runtime_error("forward: partial workspace overflow");
        o_segs = put(plan_.segs.data(), plan_.segs.size() * 8);
        o_ptr = put(plan_.cta_ptr.data(), plan_.cta_ptr.size());
        zero_counters_.assign(plan_.merges.size() + 1, 0);
        o_cnt = put(zero_counters_.data(), zero_counters_.size());     // merge counters start every step at zero
    }
    if (!in.decode) {
        // heaviest query tiles (most visible keys) first: the causal triangle otherwise leaves a long tail of big CTAs
        sorted_tiles_ = in.tiles;
        std::stable_sort(sorted_tiles_.begin(), sorted_tiles_.end(), [](const PrefillTile& a, const PrefillTile& b) { return a.pos0 + a.n_rows > b.pos0 + b.n_rows; });
        o_tiles = put(sorted_tiles_.data(), sorted_tiles_.size() * 4);
    }
    cuda_check(cudaMemcpyAsync(d_meta_, h_meta_, w * 4, cudaMemcpyHostToDevice, stream), "meta H2D");
    for (size_t u = 0; u < n_upd; ++u) {           // new automaton states: bitset rows of the device table (read by the LM-head epilogue below)
        const uint32_t slot = in.mask_updates[u * rec];
        if ((int)slot >= mask_slots) throw std::runtime_error("forward: token-mask update slot out of range");
        cuda_check(cudaMemcpyAsync(mask_table_ + (size_t)slot * mask_words, d_meta_ + o_mupd + u * rec + 1, (size_t)mask_words * 4, cudaMemcpyDeviceToDevice, stream), "token-mask update");
    }
    cuda_check(cudaEventRecord(meta_ev_[meta_idx_], stream), "meta event");
    h2d_bytes += w * 4;
    const int32_t* d_tok = d_meta_ + o_tok; const int32_t* d_pos = d_meta_ + o_pos; const int32_t* d_slot = d_meta_ + o_slot;
    const int32_t* d_samp = d_meta_ + o_samp; const int32_t* d_bt = d_meta_ + o_bt; const int32_t* d_ctx = d_meta_ + o_ctx;

    size_t ek = 0;
    if (profile_all) {
        if (ev_all.size() < (size_t)L * 12 + 16) { size_t o = ev_all.size(); ev_all.resize((size_t)L * 12 + 16); for (size_t i = o; i < ev_all.size(); ++i) cudaEventCreate(&ev_all[i]); }
        ev_ids.assign(ev_all.size(), 0);
        cudaEventRecord(ev_all[ek++], stream);
    }
    auto MARK = [&](int id) { if (profile_all && ek < ev_all.size()) { ev_ids[ek] = id; cudaEventRecord(ev_all[ek++], stream); } };
    const float scale_log2e = (1.0f / std::sqrt((float)D)) * 1.4426950408889634f;
    cuda_check(launch_embed_gather(d_tok, embed, x_, T, H, cfg.vocab, stream), "embed"); MARK(0);
    const bool use_sk = opt.streamk && T <= std::min(256, std::max(128, opt.sk_max_rows));   // decode-sized batches (one or, opt-in, two 128-row tiles)
    const int sk_rows = T > 128 ? 256 : 128;
    auto attention = [&](int l) {
        if (profile_attn) cudaEventRecord(ev[2 * l], stream);
        if (n_dec > 0) {
            DecodeAttnParams a{}; a.q = q_; a.out = attn_; a.block_tables = d_bt; a.ctx_lens = d_ctx; a.max_pages_per_seq = max_pages_per_seq;
            a.segs = reinterpret_cast<const DecodeSeg*>(d_meta_ + o_segs); a.cta_seg_ptr = d_meta_ + o_ptr; a.n_ctas = (int)plan_.cta_ptr.size() - 1;
            a.part_o = part_o_; a.part_ml = part_ml_; a.layer = l; a.n_heads = nh; a.n_kv = nkv; a.scale_log2e = scale_log2e;
            a.merge_counters = d_meta_ + o_cnt;      // cut items are merged inside the kernel by the CTA finishing their last piece
            cuda_check(launch_decode_attention(&tm_kv, kv, a, stream), "decode attention"); MARK(4);
        }
        if (!in.decode && !in.tiles.empty()) {     // mixed step: both kernels, disjoint rows of q_/attn_
            PrefillAttnParams a{}; a.q = q_; a.out = attn_; a.block_tables = d_bt; a.max_pages_per_seq = max_pages_per_seq;
            a.tiles = reinterpret_cast<const PrefillTile*>(d_meta_ + o_tiles); a.n_tiles = (int)in.tiles.size();
            a.layer = l; a.n_heads = nh; a.n_kv = nkv; a.scale_log2e = scale_log2e;
            static const bool legacy = [] { const char* e = std::getenv("OA_PREFILL_ATTN"); return e && std::string(e) == "legacy"; }();
            if (legacy) cuda_check(launch_prefill_attention(&tm_kv, kv, a, stream), "prefill attention (mma.sync)");
            else cuda_check(launch_prefill_attention_tc(&tm_q_, &tm_kv, kv, a, stream), "prefill attention (tcgen05)");
            MARK(4);
        }
        if (profile_attn) cudaEventRecord(ev[2 * l + 1], stream);
    };
    if (use_sk) {
        // decode-sized batch: persistent stream-K projections (fp32 partials) + fused consumers
        auto bn_of = [&](int o) { return sk_rows == 256 ? 128 : ((o == 64 || o == 128 || o == 256) ? o : sk_bn_); };   // two row tiles need BN=128 (TMEM)
        auto with_pf = [&](StreamK k) { k.l2_prefetch_units = std::max(0, opt.sk_l2_prefetch_kb * 1024 / (k.bn * 128)); return k; };
        const StreamK sk_qkv = make_streamk(sk_ws_, qkvd, H, bn_of(opt.sk_bn_qkv), sk_G_, sk_rows), sk_o = make_streamk(sk_ws_, H, qd, bn_of(opt.sk_bn_o), sk_G_, sk_rows);
        const StreamK sk_gu = make_streamk(sk_ws_, 2 * F, H, bn_of(opt.sk_bn_gu), sk_G_, sk_rows), sk_dn = make_streamk(sk_ws_, H, F, bn_of(opt.sk_bn_down), sk_G_, sk_rows);
        cuda_check(launch_rmsnorm(x_, layers[0].ln1, xn_, T, H, cfg.rms_eps, stream), "rmsnorm1"); MARK(1);
        const StreamK pf_qkv = with_pf(sk_qkv), pf_o = with_pf(sk_o), pf_gu = with_pf(sk_gu), pf_dn = with_pf(sk_dn);
        // qkv / o / down epilogues finished inside the GEMM (same conditions as the fused SwiGLU; every projection at BN = 128, tile counts within the flag arrays)
        const bool fuse_ok = sk_rows == 128 && sk_G_ <= sm_count && sk_qkv.bn == 128 && sk_o.bn == 128 && sk_dn.bn == 128 &&
                              sk_qkv.n_tiles <= SK_MAX_FLAG_TILES && sk_o.n_tiles <= SK_MAX_FLAG_TILES && (D == 64 || D == 128);
        const bool fuse_rope = fuse_ok && (opt.sk_fuse_epi & 1), fuse_o = fuse_ok && !use_tp && (opt.sk_fuse_epi & 2), fuse_dn = fuse_ok && !use_tp && (opt.sk_fuse_epi & 4);
        // cluster split-K (few-tile projections): cluster size per shape, 0 = keep stream-K.  o / down finish the residual add themselves, so
        // they are single-GPU only (a tensor-parallel rank must hand its partial to the all-reduce instead)
        const int ck_qkv = (sk_rows == 128 && (opt.sk_clusterk & 1) && (D == 64 || D == 128)) ? clusterk_pick(qkvd, H, sm_count, opt.sk_clusterk_min_fill) : 0;
        const int ck_o = (sk_rows == 128 && (opt.sk_clusterk & 2) && !use_tp) ? clusterk_pick(H, qd, sm_count, opt.sk_clusterk_min_fill) : 0;
        const int ck_dn = (sk_rows == 128 && (opt.sk_clusterk & 4) && !use_tp) ? clusterk_pick(H, F, sm_count, opt.sk_clusterk_min_fill) : 0;
        const bool fuse_swiglu = opt.sk_fuse_swiglu && sk_rows == 128 && sk_gu.bn == 128 && sk_G_ <= sm_count && sk_gu.n_tiles <= SK_MAX_FLAG_TILES;
        for (int l = 0; l < L; ++l) {
            const Layer& ly = layers[l];
            if (ck_qkv) {
                cuda_check(launch_gemm_clusterk_rope(&tm_xn_, ly.qkv.map(128), T, qkvd, H, ck_qkv, make_sk_rope_args(ly.bqkv, d_pos, d_slot, rope_cos, rope_sin, q_, kv, l, nh), stream),
                           "qkv gemm + RoPE + KV write (cluster split-K)"); MARK(2);
            } else if (fuse_rope) {
                cuda_check(launch_gemm_streamk_rope(&tm_xn_, ly.qkv.map(128), T, qkvd, H, pf_qkv, make_sk_rope_args(ly.bqkv, d_pos, d_slot, rope_cos, rope_sin, q_, kv, l, nh),
                                                    tile_flags_ + 0 * SK_MAX_FLAG_TILES, stream), "qkv gemm + RoPE + KV write (stream-K)"); MARK(2);
            } else {
                cuda_check(launch_gemm_streamk(&tm_xn_, ly.qkv.map(sk_qkv.bn), T, qkvd, H, pf_qkv, stream), "qkv gemm (stream-K)"); MARK(2);
                cuda_check(launch_sk_rope_kv_write(sk_qkv, ly.bqkv, d_pos, d_slot, rope_cos, rope_sin, q_, kv, l, T, nh, stream), "rope (stream-K)"); MARK(3);
            }
            attention(l);
            if (ck_o) {
                cuda_check(launch_gemm_clusterk_resid(&tm_attn_, ly.o.map(128), T, H, qd, ck_o, x_, H, stream), "o gemm + residual (cluster split-K)"); MARK(6);
                cuda_check(launch_rmsnorm_wide(x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream), "rmsnorm2"); MARK(7);
            } else if (fuse_o) {
                cuda_check(launch_gemm_streamk_resid(&tm_attn_, ly.o.map(128), T, H, qd, pf_o, x_, H, tile_flags_ + 1 * SK_MAX_FLAG_TILES, stream), "o gemm + residual (stream-K)"); MARK(6);
                cuda_check(launch_rmsnorm_wide(x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream), "rmsnorm2"); MARK(7);
            } else {
            cuda_check(launch_gemm_streamk(&tm_attn_, ly.o.map(sk_o.bn), T, H, qd, pf_o, stream), "o gemm (stream-K)"); MARK(6);
            if (use_tp) {      // row-parallel projection: all-reduce the rank partials over NVLink peer memory, then residual + norm
                const int b = comm->next_buffer();
                const TpComm::Signal sg = comm->next_signal();     // "buffer written" handshake rides on the two kernels: no barrier launch
                if (comm->nvls() && T <= 128) {
                    char* loc = reinterpret_cast<char*>(comm->nvls_local());
                    cuda_check(launch_sk_reduce_f32(sk_o, reinterpret_cast<float*>(loc + (size_t)b * nvls_region_), T, H, stream, &sg), "o partial -> multicast buffer");
                    cuda_check(launch_ar_nvls_resid_rmsnorm(comm->nvls_multicast(), loc, (size_t)b * nvls_region_, (size_t)(2 + b) * nvls_region_, 4 * nvls_region_ + (size_t)b * 512, tp, tp_rank,
                                                            x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream, sg), "in-switch allreduce+resid+rmsnorm2");
                } else if (opt.tp_ar_bf16) {
                    cuda_check(launch_sk_reduce_bf16(sk_o, comm->sym(b), T, H, stream, &sg), "o partial -> symmetric buffer (bf16)");
                    cuda_check(launch_ar_resid_rmsnorm_bf16in(comm->d_peer_sym(b), tp, x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream, &sg), "allreduce+resid+rmsnorm2");
                } else {
                    cuda_check(launch_sk_reduce_f32(sk_o, reinterpret_cast<float*>(comm->sym(b)), T, H, stream, &sg), "o partial -> symmetric buffer");
                    cuda_check(launch_ar_resid_rmsnorm(comm->d_peer_sym(b), tp, x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream, &sg), "allreduce+resid+rmsnorm2");
                }
            } else {
                cuda_check(launch_sk_resid_rmsnorm(sk_o, x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream), "resid+rmsnorm2");
            }
            MARK(7);
            }
            if (fuse_swiglu) {
                cuda_check(launch_gemm_streamk_swiglu(&tm_xn_, ly.gu.map(128), T, F, H, pf_gu, act_, tile_flags_ + 2 * SK_MAX_FLAG_TILES, stream), "gate_up gemm + SwiGLU (stream-K)"); MARK(8);
            } else {
                cuda_check(launch_gemm_streamk(&tm_xn_, ly.gu.map(sk_gu.bn), T, 2 * F, H, pf_gu, stream), "gate_up gemm (stream-K)"); MARK(8);
                cuda_check(launch_sk_swiglu(sk_gu, act_, T, F, stream), "swiglu"); MARK(9);
            }
            const void* next_gain = (l + 1 < L) ? layers[l + 1].ln1 : final_norm;
            if (ck_dn) {
                cuda_check(launch_gemm_clusterk_resid(&tm_act_, ly.down.map(128), T, H, F, ck_dn, x_, H, stream), "down gemm + residual (cluster split-K)"); MARK(10);
                cuda_check(launch_rmsnorm_wide(x_, next_gain, xn_, T, H, cfg.rms_eps, stream), "rmsnorm1"); MARK(7);
                continue;
            }
            if (fuse_dn) {
                cuda_check(launch_gemm_streamk_resid(&tm_act_, ly.down.map(128), T, H, F, pf_dn, x_, H, tile_flags_ + 3 * SK_MAX_FLAG_TILES, stream), "down gemm + residual (stream-K)"); MARK(10);
                cuda_check(launch_rmsnorm_wide(x_, next_gain, xn_, T, H, cfg.rms_eps, stream), "rmsnorm1"); MARK(7);
                continue;
            }
            cuda_check(launch_gemm_streamk(&tm_act_, ly.down.map(sk_dn.bn), T, H, F, pf_dn, stream), "down gemm (stream-K)"); MARK(10);
            if (use_tp) {
                const int b = comm->next_buffer();
                const TpComm::Signal sg = comm->next_signal();
                if (comm->nvls() && T <= 128) {
                    char* loc = reinterpret_cast<char*>(comm->nvls_local());
                    cuda_check(launch_sk_reduce_f32(sk_dn, reinterpret_cast<float*>(loc + (size_t)b * nvls_region_), T, H, stream, &sg), "down partial -> multicast buffer");
                    cuda_check(launch_ar_nvls_resid_rmsnorm(comm->nvls_multicast(), loc, (size_t)b * nvls_region_, (size_t)(2 + b) * nvls_region_, 4 * nvls_region_ + (size_t)b * 512, tp, tp_rank,
                                                            x_, next_gain, xn_, T, H, cfg.rms_eps, stream, sg), "in-switch allreduce+resid+rmsnorm1");
                } else if (opt.tp_ar_bf16) {
                    cuda_check(launch_sk_reduce_bf16(sk_dn, comm->sym(b), T, H, stream, &sg), "down partial -> symmetric buffer (bf16)");
                    cuda_check(launch_ar_resid_rmsnorm_bf16in(comm->d_peer_sym(b), tp, x_, next_gain, xn_, T, H, cfg.rms_eps, stream, &sg), "allreduce+resid+rmsnorm1");
                } else {
                    cuda_check(launch_sk_reduce_f32(sk_dn, reinterpret_cast<float*>(comm->sym(b)), T, H, stream, &sg), "down partial -> symmetric buffer");
                    cuda_check(launch_ar_resid_rmsnorm(comm->d_peer_sym(b), tp, x_, next_gain, xn_, T, H, cfg.rms_eps, stream, &sg), "allreduce+resid+rmsnorm1");
                }
            } else {
                cuda_check(launch_sk_resid_rmsnorm(sk_dn, x_, next_gain, xn_, T, H, cfg.rms_eps, stream), "resid+rmsnorm1");
            }
            MARK(7);
        }
    } else {
        // x += sum over ranks of the bf16 partials in symmetric buffer b (every rank's partial is complete: barrier passed)
        auto allreduce_resid_bf16 = [&](int b) {
            if (T >= opt.tp_two_shot_rows) {
                cuda_check(launch_ar2_reduce_scatter(comm->d_peer_sym(b), tp, tp_rank, x_, comm->sym(TpComm::GATHER), T, H, stream), "reduce-scatter+resid");
                cuda_check(comm->barrier(stream), "xgpu barrier");
                cuda_check(launch_ar2_all_gather(comm->d_peer_sym(TpComm::GATHER), tp, tp_rank, x_, T, H, stream), "all-gather");
            } else {
                cuda_check(launch_ar_resid_bf16(comm->d_peer_sym(b), tp, x_, T, H, stream), "allreduce+resid");
            }
        };
        const int bn_qkv = pick_bn(T, qkvd, opt.bn_qkv, false), bn_o = pick_bn(T, H, opt.bn_o, false);
        const int bn_gu = pick_bn(T, 2 * F, opt.bn_gu, true), bn_down = pick_bn(T, H, opt.bn_down, false);
        for (int l = 0; l < L; ++l) {
            const Layer& ly = layers[l];
            cuda_check(launch_rmsnorm(x_, ly.ln1, xn_, T, H, cfg.rms_eps, stream), "rmsnorm1"); MARK(1);
            GemmParams g{}; g.M = T; g.N = qkvd; g.K = H; g.out = qkv_; g.ldo = qkvd; g.bias = ly.bqkv;
            cuda_check(launch_gemm(&tm_xn_, ly.qkv.map(bn_qkv), g, EPI_STORE, bn_qkv, stream), "qkv gemm"); MARK(2);
            cuda_check(launch_rope_kv_write(qkv_, d_pos, d_slot, rope_cos, rope_sin, q_, kv, l, T, nh, stream), "rope"); MARK(3);
            attention(l);
            if (use_tp) {      // bf16 partial -> symmetric buffer, barrier, x += sum over ranks (rank order)
                const int b = comm->next_buffer();
                GemmParams go{}; go.M = T; go.N = H; go.K = qd; go.out = comm->sym(b); go.ldo = H;
                cuda_check(launch_gemm(&tm_attn_, ly.o.map(bn_o), go, EPI_STORE, bn_o, stream), "o gemm"); MARK(6);
                cuda_check(comm->barrier(stream), "xgpu barrier");
                allreduce_resid_bf16(b);
            } else {
                GemmParams go{}; go.M = T; go.N = H; go.K = qd; go.out = x_; go.ldo = H; go.resid = x_; go.ldr = H;
                cuda_check(launch_gemm(&tm_attn_, ly.o.map(bn_o), go, EPI_RESID, bn_o, stream), "o gemm"); MARK(6);
            }
            cuda_check(launch_rmsnorm(x_, ly.ln2, xn_, T, H, cfg.rms_eps, stream), "rmsnorm2"); MARK(1);
            GemmParams gg{}; gg.M = T; gg.N = 2 * F; gg.K = H; gg.out = act_; gg.ldo = F;
            cuda_check(launch_gemm(&tm_xn_, ly.gu.map(bn_gu), gg, EPI_SWIGLU, bn_gu, stream), "gate_up gemm"); MARK(8);
            if (use_tp) {
                const int b = comm->next_buffer();
                GemmParams gd{}; gd.M = T; gd.N = H; gd.K = F; gd.out = comm->sym(b); gd.ldo = H;
                cuda_check(launch_gemm(&tm_act_, ly.down.map(bn_down), gd, EPI_STORE, bn_down, stream), "down gemm"); MARK(10);
                cuda_check(comm->barrier(stream), "xgpu barrier");
                allreduce_resid_bf16(b);
            } else {
                GemmParams gd{}; gd.M = T; gd.N = H; gd.K = F; gd.out = x_; gd.ldo = H; gd.resid = x_; gd.ldr = H;
                cuda_check(launch_gemm(&tm_act_, ly.down.map(bn_down), gd, EPI_RESID, bn_down, stream), "down gemm"); MARK(10);
            }
        }
    }
    if (S > 0) {
        if (use_sk) {      // xn_ already holds final_norm(x): just pick the sampled rows
            cuda_check(launch_gather_rows(xn_, d_samp, xsn_, S, H, stream), "gather");
        } else {
            cuda_check(launch_gather_rows(x_, d_samp, xs_, S, H, stream), "gather");
            cuda_check(launch_rmsnorm(xs_, final_norm, xsn_, S, H, cfg.rms_eps, stream), "final norm");
        }
        MARK(11);
        const int bn_lm = pick_bn(S, V, opt.bn_lm, false);
        const int n_tiles = gemm_n_tiles(V, bn_lm);
        GemmParams gl{}; gl.M = S; gl.N = V; gl.K = H; gl.logits = logits_out; gl.ldl = V; gl.amax_val = amax_val_; gl.amax_idx = amax_idx_;
        if (!in.mask_slots.empty()) {          // constrained rows: the arg-max runs over their state's allowed tokens, on every vocabulary shard
            gl.mask_table = mask_table_; gl.mask_slot = d_meta_ + o_mslot; gl.mask_words = mask_words; gl.col_offset = tp_rank * V;
        }
        int lb = -1;
        if (use_tp) {          // vocab-parallel head: each rank's fp32 logits shard is staged in its symmetric buffer when asked for
            gl.logits = nullptr;
            if (in.want_logits) {
                if ((size_t)S * V * 4 > comm->sym_bytes()) throw std::runtime_error("debug logits do not fit the symmetric buffer");
                lb = comm->next_buffer(); gl.logits = reinterpret_cast<float*>(comm->sym(lb)); gl.ldl = V;
            }
        }
        cuda_check(launch_gemm(&tm_xsn_, lm_head.map(bn_lm), gl, EPI_LOGITS, bn_lm, stream), "lm_head gemm"); MARK(12);
        if (use_tp) {
            const int b = comm->next_buffer();
            cuda_check(launch_argmax_reduce_pair(amax_val_, amax_idx_, S, n_tiles, tp_rank * V, comm->arg(b), stream), "argmax (local shard)");
            cuda_check(comm->barrier(stream), "xgpu barrier");
            cuda_check(launch_ar_argmax(comm->d_peer_arg(b), tp, S, d_out_ids_, stream), "argmax (all ranks)");
            if (lb >= 0 && logits_out) {   // leader assembles the full [S, vocab] logits from the peers' shards over NVLink
                for (int p = 0; p < tp; ++p)
                    cuda_check(cudaMemcpy2DAsync(logits_out + (size_t)p * V, (size_t)cfg.vocab * 4, comm->peer_sym_host(lb, p), (size_t)V * 4, (size_t)V * 4, S,
                                                 cudaMemcpyDeviceToDevice, stream), "logits gather");
            }
            if (lb >= 0) cuda_check(comm->barrier(stream), "xgpu barrier");     // peers must not reuse the staged shard before the leader has read it
        } else {
            cuda_check(launch_argmax_reduce(amax_val_, amax_idx_, S, n_tiles, d_out_ids_, nullptr, stream), "argmax");
        }
        cuda_check(cudaMemcpyAsync(h_out_ids, d_out_ids_, (size_t)S * 4, cudaMemcpyDeviceToHost, stream), "ids D2H"); MARK(13);
        d2h_bytes += (size_t)S * 4;
    }
    if (profile_all) {
        cuda_check(cudaStreamSynchronize(stream), "profile sync");
        for (size_t i = 1; i < ek; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev_all[i - 1], ev_all[i]); kt_ms[ev_ids[i] & 15] += ms; kt_n[ev_ids[i] & 15]++; }
    }
    if (profile_attn) {
        cuda_check(cudaStreamSynchronize(stream), "profile sync");
        for (int l = 0; l < L; ++l) { float ms = 0; cudaEventElapsedTime(&ms, ev[2 * l], ev[2 * l + 1]); attn_ms_accum += ms; }
    }
}

}  // namespace oa
