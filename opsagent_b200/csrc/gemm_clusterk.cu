// gemm_clusterk.cu — decode projections with FEW output tiles (o, down, qkv: 32-48 tiles of 128 columns on 148 SMs) as a
// cluster split-K GEMM: the S CTAs of one thread-block cluster own ONE 128 x 128 output tile and a 1/S slice of K each, reduce their
// TMEM accumulators through distributed shared memory and finish the epilogue together — no fp32 partial round trip through L2, no
// consumer launch, no serial tail of a few finishing CTAs (what made the stream-K finisher fusion slower, profiles/r02a).
//
//   grid = n_tiles x S CTAs, cluster (S,1,1); CTA (tile, r) streams k-blocks [r*kb/S, (r+1)*kb/S) of the tile's weights
//   warp 0   : TMA producer (weights issued BEFORE griddepcontrol.wait: they depend on no predecessor)
//   warp 1   : tcgen05.mma 128 x 128 x 16 into TMEM
//   warps 2-5: TMEM -> registers -> st.shared::cluster into the OWNER CTA of each 32 x 64 "pair unit" (8 per tile, owner = unit % S);
//              after one cluster barrier every CTA sums the S partials of its own units IN RANK ORDER (deterministic) and applies the
//              epilogue with all 128 threads: residual add into the bf16 stream (o / down) or bias + RoPE + paged-KV write (qkv).
// A pair unit = rows [32q, 32q+32) x the two 32-column chunks that RoPE rotates together (columns i and i + D/2 of one head), so every
// epilogue is thread-local.  The K split differs from the stream-K path, so results are equal to it only up to fp32 summation order
// (parity against the oracle holds with the same tolerance; the bit-identity test covers the stream-K variants among themselves).
#include <algorithm>

#include "common.cuh"
#include "kernels.hpp"

namespace oa {

static constexpr int CK_BM = 128, CK_BN = 128, CK_BK = 64, CK_THREADS = 192, CK_STAGES = 4;
static constexpr int CK_A_BYTES = CK_BM * CK_BK * 2, CK_B_BYTES = CK_BN * CK_BK * 2, CK_STAGE_BYTES = CK_A_BYTES + CK_B_BYTES;
static constexpr int CK_ROW_F = 36;                              // 32 + 4 pad floats: conflict-free 16-byte accesses
static constexpr int CK_BUF_F = 32 * CK_ROW_F;                   // one 32 x 32 partial chunk

struct CkEpi {
    int kind;                           // 2: x[M, ldx] += acc (residual stream, bf16, in place)   3: bias + RoPE + paged-KV write
    uint16_t* x; int ldx; int N;
    SkRopeArgs rope;
};

OA_DEVINL uint32_t ck_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
OA_DEVINL uint32_t ck_mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
OA_DEVINL void ck_st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
OA_DEVINL void ck_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// chunk c (32 columns) of a tile -> (pair index within the row block, member 0 = low half / 1 = high half of the rotation)
template <int D> OA_DEVINL void ck_pair_of(int c, int& pidx, int& member) { if (D == 128) { pidx = c & 1; member = c >> 1; } else { pidx = c >> 1; member = c & 1; } }
template <int D> OA_DEVINL int ck_chunk_of(int pidx, int member) { return D == 128 ? pidx + 2 * member : 2 * pidx + member; }

template <int S, int D>
__global__ void __launch_bounds__(CK_THREADS, 1) gemm_clusterk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                      const int M, const int K, const CkEpi ep) {
    constexpr int SLOTS = (8 + S - 1) / S;                                   // pair units a CTA may own
    extern __shared__ __align__(1024) uint8_t smem[];
    float* recv = reinterpret_cast<float*>(smem + CK_STAGES * CK_STAGE_BYTES);          // [SLOTS][2 members][S partials][32][36]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(recv + SLOTS * 2 * S * CK_BUF_F);
    uint64_t* empty_bar = full_bar + CK_STAGES;
    uint64_t* acc_bar = empty_bar + CK_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ck_cluster_rank();
    const int tile = (int)blockIdx.x / S;
    const int kb_total = (K + CK_BK - 1) / CK_BK;
    const int kb0 = (int)((long long)rank * kb_total / S), kb1 = (int)((long long)(rank + 1) * kb_total / S), n_kb = kb1 - kb0;   // host: kb_total >= S

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
        for (int s = 0; s < CK_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<128>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();

    if (warp == 0) {
        if (lane == 0) {
            const int pre = n_kb < CK_STAGES ? n_kb : CK_STAGES;
            for (int i = 0; i < pre; ++i) {          // weights first: the HBM stream starts while the predecessor drains
                mbar_expect_tx(&full_bar[i], CK_STAGE_BYTES);
                tma_load_2d(smem + i * CK_STAGE_BYTES + CK_A_BYTES, &tmB, &full_bar[i], (kb0 + i) * CK_BK, tile * CK_BN, kEvictFirst);
            }
            griddep_wait();
            for (int i = 0; i < pre; ++i) tma_load_2d(smem + i * CK_STAGE_BYTES, &tmA, &full_bar[i], (kb0 + i) * CK_BK, 0, kEvictLast);
            int s = pre == CK_STAGES ? 0 : pre; uint32_t ph = pre == CK_STAGES ? 1 : 0;
            for (int i = pre; i < n_kb; ++i) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* dst = smem + s * CK_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], CK_STAGE_BYTES);
                tma_load_2d(dst, &tmA, &full_bar[s], (kb0 + i) * CK_BK, 0, kEvictLast);
                tma_load_2d(dst + CK_A_BYTES, &tmB, &full_bar[s], (kb0 + i) * CK_BK, tile * CK_BN, kEvictFirst);
                if (++s == CK_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(CK_BM, CK_BN);
            int s = 0; uint32_t ph = 0;
            for (int i = 0; i < n_kb; ++i) {
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * CK_STAGE_BYTES);
                const uint64_t a_desc = umma_desc_sw128(a_addr), b_desc = umma_desc_sw128(a_addr + CK_A_BYTES);
#pragma unroll
                for (int k = 0; k < CK_BK / 16; ++k) umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (i | k) != 0);
                umma_commit(&empty_bar[s]);
                if (++s == CK_STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(acc_bar);
        }
    }
    // every CTA of the cluster is running (co-scheduled) and past its setup before anybody writes into a peer's shared memory
    ck_cluster_sync();
    if (warp >= 2) {
        const int q = warp & 3;                                   // TMEM lanes / tile rows [32q, 32q + 32)
        griddep_wait();
        mbar_wait(acc_bar, 0);
        tcgen05_fence_after();
        const uint32_t recv_base = smem_u32(recv);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            tmem_ld_wait();
            int pidx, member; ck_pair_of<D>(c, pidx, member);
            const int pu = q * 2 + pidx, owner = pu % S, slot = pu / S;
            const uint32_t local = recv_base + (uint32_t)((((slot * 2 + member) * S + (int)rank) * CK_BUF_F + lane * CK_ROW_F) * 4);
            const uint32_t remote = ck_mapa(local, (uint32_t)owner);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                ck_st_cluster_f4(remote + (uint32_t)(j * 4), __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        tcgen05_fence_before();
    }
    ck_cluster_sync();                                            // all partials of my units have landed in my shared memory
    if (warp >= 2) {
        const int t = (int)threadIdx.x - 64;                      // 0..127
        const int r32 = t >> 2, col8 = (t & 3) * 8;
#pragma unroll 1
        for (int slot = 0; slot < SLOTS; ++slot) {
            const int pu = (int)rank + slot * S;
            if (pu >= 8) break;
            const int q = pu >> 1, pidx = pu & 1;
            const int row = q * 32 + r32;
            float acc[2][8];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[m][i] = 0.f;
#pragma unroll
                for (int p = 0; p < S; ++p) {                     // cluster-rank order = ascending k: a fixed order
                    const float* b = recv + ((slot * 2 + m) * S + p) * CK_BUF_F + r32 * CK_ROW_F + col8;
                    const float4 u0 = *reinterpret_cast<const float4*>(b), u1 = *reinterpret_cast<const float4*>(b + 4);
                    acc[m][0] += u0.x; acc[m][1] += u0.y; acc[m][2] += u0.z; acc[m][3] += u0.w;
                    acc[m][4] += u1.x; acc[m][5] += u1.y; acc[m][6] += u1.z; acc[m][7] += u1.w;
                }
            }
            if (row >= M) continue;
            const int col_lo = tile * CK_BN + ck_chunk_of<D>(pidx, 0) * 32 + col8, col_hi = tile * CK_BN + ck_chunk_of<D>(pidx, 1) * 32 + col8;
            if (ep.kind == 2) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int col = m ? col_hi : col_lo;
                    if (col >= ep.N) continue;
                    uint4* xp = reinterpret_cast<uint4*>(ep.x + (size_t)row * ep.ldx + col);
                    const uint4 xo = *xp; uint4 xn;
                    xn.x = pack_bf16x2(bf16lo(xo.x) + acc[m][0], bf16hi(xo.x) + acc[m][1]); xn.y = pack_bf16x2(bf16lo(xo.y) + acc[m][2], bf16hi(xo.y) + acc[m][3]);
                    xn.z = pack_bf16x2(bf16lo(xo.z) + acc[m][4], bf16hi(xo.z) + acc[m][5]); xn.w = pack_bf16x2(bf16lo(xo.w) + acc[m][6], bf16hi(xo.w) + acc[m][7]);
                    *xp = xn;
                }
            } else {
                if (col_lo >= ep.N) continue;
                const SkRopeArgs& a = ep.rope;
                const int half = D >> 1, head = col_lo / D, i0 = col_lo - head * D;          // i0 in [0, D/2): the low half of a rotation pair
                const int slot_tok = a.slots[row], page = slot_tok / a.page_size, off = slot_tok - page * a.page_size;
#pragma unroll
                for (int m = 0; m < 2; ++m) {                     // + bias, round: the projection output is a bf16 tensor
                    const int col = m ? col_hi : col_lo;
                    if (a.bias) {
                        const uint4 bb = *reinterpret_cast<const uint4*>(a.bias + col);
                        acc[m][0] += bf16lo(bb.x); acc[m][1] += bf16hi(bb.x); acc[m][2] += bf16lo(bb.y); acc[m][3] += bf16hi(bb.y);
                        acc[m][4] += bf16lo(bb.z); acc[m][5] += bf16hi(bb.z); acc[m][6] += bf16lo(bb.w); acc[m][7] += bf16hi(bb.w);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[m][i] = bf16_bits_to_f32(f32_to_bf16_bits(acc[m][i]));
                }
                auto pack8 = [](const float (&v)[8]) { uint4 o; o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]); return o; };
                if (head >= a.nh + a.nkv) {                       // V: straight copy of both halves into the V plane
                    uint16_t* dst = a.kv_base + (size_t)(a.v_plane_row0 + ((int64_t)page * a.nkv + (head - a.nh - a.nkv)) * a.page_size + off) * D;
                    *reinterpret_cast<uint4*>(dst + i0) = pack8(acc[0]);
                    *reinterpret_cast<uint4*>(dst + i0 + half) = pack8(acc[1]);
                } else {
                    const int pos = a.positions[row];
                    const float* cr = a.rope_cos + (size_t)pos * half + i0;
                    const float* sr = a.rope_sin + (size_t)pos * half + i0;
                    const float4 c0 = *reinterpret_cast<const float4*>(cr), c1 = *reinterpret_cast<const float4*>(cr + 4);
                    const float4 s0 = *reinterpret_cast<const float4*>(sr), s1 = *reinterpret_cast<const float4*>(sr + 4);
                    const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    float ra[8], rb[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { ra[i] = acc[0][i] * cv[i] - acc[1][i] * sv[i]; rb[i] = acc[1][i] * cv[i] + acc[0][i] * sv[i]; }
                    uint16_t* dst;
                    if (head < a.nh) dst = a.q_out + (size_t)row * a.nh * D + (size_t)head * D;
                    else dst = a.kv_base + (size_t)(a.k_plane_row0 + ((int64_t)page * a.nkv + (head - a.nh)) * a.page_size + off) * D;
                    *reinterpret_cast<uint4*>(dst + i0) = pack8(ra);
                    *reinterpret_cast<uint4*>(dst + i0 + half) = pack8(rb);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc<128>(tmem_base); }
}

template <int S>
static constexpr int ck_smem_bytes() { return CK_STAGES * CK_STAGE_BYTES + ((8 + S - 1) / S) * 2 * S * CK_BUF_F * 4 + 256; }

template <int S, int D>
static cudaError_t launch_ck(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, const CkEpi& ep, cudaStream_t stream) {
    auto kern = gemm_clusterk_kernel<S, D>;
    constexpr int SMEM = ck_smem_bytes<S>();
    static_assert(SMEM <= 232448, "shared memory budget of one CTA");
    static bool attr_done[16] = {};
    { cudaError_t e = ensure_dynamic_smem(kern, SMEM, attr_done); if (e != cudaSuccess) return e; }
    const int n_tiles = (N + CK_BN - 1) / CK_BN;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles * S); cfg.blockDim = dim3(CK_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kern, *tmA, *tmB, M, K, ep);
}

template <int D>
static cudaError_t launch_ck_s(int S, const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, const CkEpi& ep, cudaStream_t st) {
    switch (S) {
        case 2: return launch_ck<2, D>(tmA, tmB, M, N, K, ep, st);
        case 3: return launch_ck<3, D>(tmA, tmB, M, N, K, ep, st);
        case 4: return launch_ck<4, D>(tmA, tmB, M, N, K, ep, st);
        case 8: return launch_ck<8, D>(tmA, tmB, M, N, K, ep, st);
    }
    return cudaErrorInvalidValue;
}

// cluster size for an N x K projection on `sms` SMs: the largest of {8, 4, 3, 2} with n_tiles * S <= sms and >= 2 k-blocks per CTA
// (0: not worth it — fewer than min_fill_pct % of the SMs would stream)
int clusterk_pick(int N, int K, int sms, int min_fill_pct) {
    const int n_tiles = (N + CK_BN - 1) / CK_BN, kb = (K + CK_BK - 1) / CK_BK;
    for (int S : {8, 4, 3, 2})
        if (n_tiles * S <= sms && kb >= 2 * S) return (n_tiles * S * 100 >= sms * min_fill_pct) ? S : 0;
    return 0;
}

cudaError_t launch_gemm_clusterk_resid(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, int S, void* x, int ldx, cudaStream_t stream) {
    if (M <= 0 || M > CK_BM || (N % 8) != 0 || (K % 8) != 0 || !x || (ldx % 8) != 0) return cudaErrorInvalidValue;
    CkEpi ep{}; ep.kind = 2; ep.x = reinterpret_cast<uint16_t*>(x); ep.ldx = ldx; ep.N = N;
    return launch_ck_s<128>(S, tmA, tmB, M, N, K, ep, stream);          // D only fixes the unit pairing: any pairing suits the residual epilogue
}
cudaError_t launch_gemm_clusterk_rope(const CUtensorMap* tmA, const CUtensorMap* tmB, int M, int N, int K, int S, const SkRopeArgs& rope, cudaStream_t stream) {
    if (M <= 0 || M > CK_BM || (K % 8) != 0 || N != (rope.nh + 2 * rope.nkv) * rope.D) return cudaErrorInvalidValue;
    CkEpi ep{}; ep.kind = 3; ep.N = N; ep.rope = rope;
    if (rope.D == 128) return launch_ck_s<128>(S, tmA, tmB, M, N, K, ep, stream);
    if (rope.D == 64) return launch_ck_s<64>(S, tmA, tmB, M, N, K, ep, stream);
    return cudaErrorInvalidValue;
}

}  // namespace oa
