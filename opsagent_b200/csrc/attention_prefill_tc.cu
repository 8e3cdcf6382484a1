// attention_prefill_tc.cu — causal prefill attention over the paged KV cache on the 5th-generation tensor cores.
//
// One CTA = 128 query rows of one (sequence, head); KV is consumed in tiles of 128 tokens (two 64-token pages):
//   warp 0    : TMA producers — lane 0: Q tile once, then the K pages of each KV tile; lane 1: the V pages.  K and V have SEPARATE
//               2-stage rings: K_j is released as soon as S_j has been computed, V_j only after P_j·V_j, so the K load of tile
//               j+2 gets a full tile period of slack instead of being exposed behind the second MMA
//   warp 1    : tcgen05.mma issuer —  S_j = Q · K_j^T   (M=128, N=128, K=D;   A, B K-major)        -> TMEM S[j&1]
//                                     O  += P_j · V_j   (M=128, N=D,   K=128; A K-major, B MN-major: V[kv, d] as stored)  -> TMEM O
//   warps 2-5 : one query row per thread.  tcgen05.ld S_j once (128 values), causal mask, P_j = 2^(s - m_ref) rounded to bf16 into
//               swizzled shared memory (the A operand of the second MMA).  O stays in TMEM for the whole KV sweep: the reference
//               maximum m_ref is only raised — and O, l rescaled via tcgen05.ld/st — when a row's maximum exceeds it by more than
//               2^8 (lazy rescaling), so the common tile costs one TMEM read, 128 exp2 and 16 shared-memory stores per thread and
//               overlaps the tensor-core work of its neighbours (S and P are double-buffered).
// Row sums use the unrounded p, P·V the bf16-rounded p — the rounding points the oracle mirrors.  TMEM: 2x128 (S) + D (O).
#include <cstdlib>

#include "common.cuh"
#include "kernels.hpp"

namespace oa {

static constexpr int PTC_MAX_THREADS = 64 + 2 * 128;
static constexpr int KV_TILE = 128;

template <int D>
struct PtcCfg {
    static constexpr int HALVES = D / 64;
    static constexpr int HALF_BYTES = 128 * 128;                 // 128 rows x 64 bf16 columns, one swizzle-atom column
    static constexpr int Q_BYTES = HALVES * HALF_BYTES;
    static constexpr int KV_BYTES = HALVES * HALF_BYTES;         // K (or V) of one 128-token tile
    static constexpr int STAGE_BYTES = 2 * KV_BYTES;
    static constexpr int P_BYTES = 2 * HALF_BYTES;               // [128 q, 128 kv] bf16
    static constexpr int XCH_BYTES = 2 * 2 * 128 * 4;            // row-statistic exchange between the two softmax warpgroups: [tile parity][warpgroup][row]
    static constexpr int SMEM_BYTES = Q_BYTES + 2 * STAGE_BYTES + 2 * P_BYTES + 256 + XCH_BYTES;     // D=128: 231,680 B of the 232,448 available — the base is declared 1024-aligned, no slack
    static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
    static constexpr uint32_t TMEM_COLS = 512;
    static constexpr uint32_t T_COL = 256;                       // S[0] at 0, S[1] at 128, T at 256
};

// NWG = softmax warpgroups.  With one, each SM sub-partition holds a single softmax warp and the 128 exp2 + pack + store chain of a tile
// (~1.1 us) is longer than its two MMAs (~0.75 us): the tensor pipe idles 70 % of the time (profiles/r01f).  With two, the warps w and w+4
// share the same 32 query rows (same TMEM lanes) and split every S tile by KEY columns — 64 scores, 64 exp2, 8 P stores per thread, two
// warps per sub-partition to hide latencies; the halves' row maxima meet through shared memory once per tile, the row sums once at the end.
template <int D, int NWG>
__global__ void __launch_bounds__(64 + 128 * NWG, 1) prefill_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                              const __grid_constant__ CUtensorMap tm_kv,
                                                                              const PrefillAttnParams p, const int64_t layer_row0,
                                                                              const int64_t kv_stride_rows) {
    using Cfg = PtcCfg<D>;
    constexpr int HALVES = Cfg::HALVES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];       // 128-byte-swizzled TMA / UMMA tiles need 1024-byte alignment
    uint8_t* smem = smem_raw;
    uint8_t* q_s = smem;
    uint8_t* kv_s = q_s + Cfg::Q_BYTES;                          // [2 stages][K | V]
    uint8_t* p_s = kv_s + 2 * Cfg::STAGE_BYTES;                  // [2][P]
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 2 * Cfg::P_BYTES);
    uint64_t* q_full = bars;                  // 1
    uint64_t* k_full = bars + 1;              // [2]
    uint64_t* k_empty = bars + 3;             // [2]
    uint64_t* v_full = bars + 5;              // [2]
    uint64_t* v_empty = bars + 7;             // [2]
    uint64_t* s_full = bars + 9;              // [2]
    uint64_t* s_empty = bars + 11;            // [2]
    uint64_t* p_full = bars + 13;             // [2]
    uint64_t* t_full = bars + 15;             // 1 (one completion per P·V)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
    float* xch = reinterpret_cast<float*>(p_s + 2 * Cfg::P_BYTES + 256);       // [2][2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PrefillTile tile = p.tiles[blockIdx.x];
    const int head = blockIdx.y;
    const int grp = p.n_heads / p.n_kv, kvh = head / grp;
    const int n_kvt = (tile.pos0 + tile.n_rows + KV_TILE - 1) / KV_TILE;       // KV tiles 0..n_kvt-1 cover every visible key

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
                                      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4 * NWG); mbar_init(&p_full[i], 4 * NWG); }
        mbar_init(t_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();

    if (warp == 0) {
        if (lane < 2) {
            const bool is_v = lane == 1;
            griddep_wait();
            if (!is_v) {
                mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
                for (int h = 0; h < HALVES; ++h) tma_load_2d(q_s + h * Cfg::HALF_BYTES, &tm_q, q_full, head * D + h * 64, tile.q_row0, kEvictFirst);
            }
            const int32_t* bt = p.block_tables + (size_t)tile.seq * p.max_pages_per_seq;
            uint64_t* full = is_v ? v_full : k_full;
            uint64_t* empty = is_v ? v_empty : k_empty;
            for (int j = 0; j < n_kvt; ++j) {
                const int st = j & 1;
                mbar_wait(&empty[st], (((uint32_t)j >> 1) & 1) ^ 1);
                uint8_t* dst = kv_s + st * Cfg::STAGE_BYTES + (is_v ? Cfg::KV_BYTES : 0);
                mbar_expect_tx(&full[st], Cfg::KV_BYTES);
#pragma unroll
                for (int pg = 0; pg < 2; ++pg) {
                    const int pi = 2 * j + pg;
                    const int page = bt[pi < p.max_pages_per_seq ? pi : 2 * j];      // a missing second page is fully masked anyway
                    const int64_t row = layer_row0 + ((int64_t)page * p.n_kv + kvh) * 64 + (is_v ? kv_stride_rows : 0);
#pragma unroll
                    for (int h = 0; h < HALVES; ++h)
                        tma_load_2d(dst + h * Cfg::HALF_BYTES + pg * 8192, &tm_kv, &full[st], h * 64, (int32_t)row, kEvictLast);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, KV_TILE);
            constexpr uint32_t idesc_t = umma_idesc_bf16_bmn(128, D);
            const uint32_t q_addr = smem_u32(q_s);
            auto issue_s = [&](int j) {
                const int st = j & 1, sb = j & 1;
                mbar_wait(&k_full[st], ((uint32_t)j >> 1) & 1);
                mbar_wait(&s_empty[sb], (((uint32_t)j >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                const uint32_t k_addr = smem_u32(kv_s + st * Cfg::STAGE_BYTES);
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t off = (uint32_t)((ks >> 2) * Cfg::HALF_BYTES + (ks & 3) * 32);
                    umma_bf16(tmem_base + (uint32_t)(sb * 128), umma_desc_sw128(q_addr + off), umma_desc_sw128(k_addr + off), idesc_s, ks > 0 ? 1u : 0u);
                }
                umma_commit(&s_full[sb]);
                umma_commit(&k_empty[st]);                                 // K_j is free once S_j exists
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kvt; ++j) {
                if (j + 1 < n_kvt) issue_s(j + 1);                         // keep the tensor pipe ahead of the softmax warps
                const int st = j & 1, pb = j & 1;
                mbar_wait(&p_full[pb], ((uint32_t)j >> 1) & 1);            // P_j is in shared memory
                mbar_wait(&v_full[st], ((uint32_t)j >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t p_addr = smem_u32(p_s + pb * Cfg::P_BYTES);
                const uint32_t v_addr = smem_u32(kv_s + st * Cfg::STAGE_BYTES + Cfg::KV_BYTES);
#pragma unroll
                for (int ks = 0; ks < KV_TILE / 16; ++ks) {
                    const uint64_t a_desc = umma_desc_sw128(p_addr + (uint32_t)((ks >> 2) * Cfg::HALF_BYTES + (ks & 3) * 32));
                    const uint64_t b_desc = umma_desc_sw128_mn(v_addr + (uint32_t)(ks * 2048), (uint32_t)Cfg::HALF_BYTES, 1024u);
                    umma_bf16(tmem_base + Cfg::T_COL, a_desc, b_desc, idesc_t, (j > 0 || ks > 0) ? 1u : 0u);      // O accumulates over all KV tiles
                }
                umma_commit(t_full);
                umma_commit(&v_empty[st]);
            }
        }
    } else {
        const int q = warp & 3, wg = (warp - 2) >> 2;               // warps w and w + 4 own the same rows (TMEM lanes 32 * (w % 4) ...)
        const int row = q * 32 + lane;                              // query row of this thread inside the tile
        const int qpos = tile.pos0 + row;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        constexpr int CH = 4 / NWG;                                 // 32-key chunks of a tile this thread handles
        const int key_base = wg * (KV_TILE / NWG);
        constexpr int OC = D / NWG;                                 // O columns this thread rescales / writes
        const int o_base = wg * OC;
        constexpr float kRescaleThreshold = 8.0f;                   // log2 units: p may reach 2^8 before O is rescaled
        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < n_kvt; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], ((uint32_t)j >> 1) & 1);
            tcgen05_fence_after();
            uint32_t v[CH][32];
#pragma unroll
            for (int c = 0; c < CH; ++c) tmem_ld_32x32b_x32(tmem_base + lane_addr + (uint32_t)(sb * 128 + key_base + c * 32), v[c]);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[sb]);               // S_j is in registers: the tensor core may overwrite the buffer
            const bool diag = (j * KV_TILE + KV_TILE - 1) > tile.pos0;          // some (row, key) pairs of this tile are masked
            if (diag) {                                                         // only the last tile(s) of a row block pay for the mask
#pragma unroll
                for (int c = 0; c < CH; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if ((j * KV_TILE + key_base + c * 32 + i) > qpos) v[c][i] = 0xff800000u;   // -inf
            }
            float mraw = -INFINITY;                                             // max of the raw scores; the scale is positive
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) mraw = fmaxf(mraw, __uint_as_float(v[c][i]));
            if constexpr (NWG == 2) {                                           // the row's other half lives in the partner warpgroup
                xch[(sb * 2 + wg) * 128 + row] = mraw;
                asm volatile("bar.sync 2, 256;" ::: "memory");
                mraw = fmaxf(mraw, xch[(sb * 2 + (wg ^ 1)) * 128 + row]);
            }
            const float mx = mraw * p.scale_log2e;
            // lazy rescaling: raise the reference maximum only when a row would overshoot it by more than 2^8
            const bool raise = mx > m_ref + kRescaleThreshold;      // always true on tile 0 (key 0 is visible to every row)
            if (__any_sync(0xffffffffu, raise)) {                   // both warps of a row pair see the same mx, hence take the same branch
                if (j > 0) {
                    mbar_wait(t_full, (uint32_t)(j - 1) & 1);       // O += P_{j-1}·V_{j-1} has landed; O += P_j·V_j is not issued before p_full
                    tcgen05_fence_after();
                    const float alpha = raise ? exp2f(m_ref - mx) : 1.0f;
#pragma unroll 1
                    for (int c = 0; c < OC; c += 32) {
                        uint32_t o[32];
                        tmem_ld_32x32b_x32(tmem_base + lane_addr + Cfg::T_COL + (uint32_t)(o_base + c), o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32b_x32(tmem_base + lane_addr + Cfg::T_COL + (uint32_t)(o_base + c), o);
                    }
                    tmem_st_wait();
                    tcgen05_fence_before();
                    l_run *= alpha;
                }
                if (raise) m_ref = mx;
            }
            uint8_t* prow = p_s + sb * Cfg::P_BYTES + row * 128;
            float psum = 0.f;
            const float neg_ref = -m_ref;                                       // p = 2^(s*scale - m_ref): one FFMA + one SFU op per score
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                float pr[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { pr[i] = ex2_approx(fmaf(__uint_as_float(v[c][i]), p.scale_log2e, neg_ref)); psum += pr[i]; }
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {                    // four 16-byte chunks (8 keys each)
                    const int key0 = key_base + c * 32 + g8 * 8, hh = key0 >> 6, chunk = (key0 & 63) >> 3;
                    uint4 w;
                    w.x = pack_bf16x2(pr[g8 * 8 + 0], pr[g8 * 8 + 1]); w.y = pack_bf16x2(pr[g8 * 8 + 2], pr[g8 * 8 + 3]);
                    w.z = pack_bf16x2(pr[g8 * 8 + 4], pr[g8 * 8 + 5]); w.w = pack_bf16x2(pr[g8 * 8 + 6], pr[g8 * 8 + 7]);
                    *reinterpret_cast<uint4*>(prow + hh * Cfg::HALF_BYTES + ((chunk ^ (row & 7)) << 4)) = w;
                }
            }
            l_run += psum;
            fence_proxy_async();                                    // P stores must be visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[sb]);
        }
        mbar_wait(t_full, (uint32_t)(n_kvt - 1) & 1);
        tcgen05_fence_after();
        if constexpr (NWG == 2) {                                   // row sum = the two key halves' sums (fixed order: half 0 + half 1)
            const int xb = n_kvt & 1;                               // the buffer the last tile did not use
            xch[(xb * 2 + wg) * 128 + row] = l_run;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            l_run = xch[(xb * 2 + 0) * 128 + row] + xch[(xb * 2 + 1) * 128 + row];
        }
        const float inv = 1.0f / l_run;
        uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + (size_t)(tile.q_row0 + row) * p.n_heads * D + (size_t)head * D;
#pragma unroll 1
        for (int c = 0; c < OC; c += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_base + lane_addr + Cfg::T_COL + (uint32_t)(o_base + c), o);
            tmem_ld_wait();
            if (row < tile.n_rows) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    uint4 w;
                    w.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
                    w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
                    w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
                    w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + o_base + c + i) = w;
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<Cfg::TMEM_COLS>(tmem_base); }
}

template <int D, int NWG>
static cudaError_t launch_ptc(const CUtensorMap* tm_q, const CUtensorMap* tm_kv, const KvLayout& kv, const PrefillAttnParams& p, cudaStream_t s) {
    auto kern = prefill_attention_tc_kernel<D, NWG>;
    static bool attr_done[16] = {};
    { cudaError_t e = ensure_dynamic_smem(kern, PtcCfg<D>::SMEM_BYTES, attr_done); if (e != cudaSuccess) return e; }
    return launch_k(kern, dim3(p.n_tiles, p.n_heads), dim3(64 + 128 * NWG), PtcCfg<D>::SMEM_BYTES, s, *tm_q, *tm_kv, p,
                    (int64_t)p.layer * kv.layer_stride_rows, kv.kv_stride_rows);
}

cudaError_t launch_prefill_attention_tc(const CUtensorMap* tm_q, const CUtensorMap* tm_kv, const KvLayout& kv, const PrefillAttnParams& p, cudaStream_t s) {
    if (p.n_tiles <= 0) return cudaSuccess;
    if (kv.page_size != 64 || p.n_heads % p.n_kv != 0) return cudaErrorInvalidValue;
    static const int nwg = [] { const char* e = std::getenv("OA_PREFILL_WG"); return (e && e[0] == '1') ? 1 : 2; }();      // OA_PREFILL_WG=1: the one-warpgroup kernel of round 1 (A/B)
    if (kv.head_dim == 128) return nwg == 2 ? launch_ptc<128, 2>(tm_q, tm_kv, kv, p, s) : launch_ptc<128, 1>(tm_q, tm_kv, kv, p, s);
    if (kv.head_dim == 64) return nwg == 2 ? launch_ptc<64, 2>(tm_q, tm_kv, kv, p, s) : launch_ptc<64, 1>(tm_q, tm_kv, kv, p, s);
    return cudaErrorInvalidValue;
}

}  // namespace oa
