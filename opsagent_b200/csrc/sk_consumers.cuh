// sk_consumers.cuh — the consumers of stream-K partials as device functions: each adds a tile's fp32 partials in CTA
// order (deterministic) and finishes the op.  Shared by the stand-alone consumer kernels (elementwise.cu) and by the
// GEMM epilogues that finish an op themselves (gemm_tcgen05.cu FUSE=2/3, gemm_clusterk.cu), so every path produces bit-identical results.
#pragma once
#include "common.cuh"
#include "kernels.hpp"

namespace oa {

OA_DEVINL void named_bar_sync(int id, int n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }

// CTAs c_first..c_last own pieces of `tile` (sk_tile_ctas, kernels.hpp); their partials are added in that order.
OA_DEVINL void sk_sum8(const StreamK& sk, int row, int col, float (&acc)[8]) {
    const uint32_t tile = (uint32_t)col / (uint32_t)sk.bn, cc = (uint32_t)col - tile * (uint32_t)sk.bn;
    uint32_t c_first, c_last;
    sk_tile_ctas(sk, tile, c_first, c_last);
    const int n = (int)(c_last - c_first) + 1;
    const size_t slot_stride = (size_t)sk.rows * sk.bn;
    const float* p = sk.ws + ((size_t)(c_first + tile) * sk.rows + row) * sk.bn + cc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // issue up to 6 partials' loads back to back (independent L2 round trips), then add them in CTA order
    for (int i = 0; i < n; i += 6) {
        float4 a[6], b[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (i + j < n) {
                const float4* q4 = reinterpret_cast<const float4*>(p + (size_t)(i + j) * slot_stride);
                a[j] = __ldcg(q4); b[j] = __ldcg(q4 + 1);
            } else { a[j] = make_float4(0.f, 0.f, 0.f, 0.f); b[j] = a[j]; }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (i + j < n) {
                acc[0] += a[j].x; acc[1] += a[j].y; acc[2] += a[j].z; acc[3] += a[j].w;
                acc[4] += b[j].x; acc[5] += b[j].y; acc[6] += b[j].z; acc[7] += b[j].w;
            }
        }
    }
}

// Two 8-column groups of the SAME tile (a RoPE pair: columns c and c + D/2): one index computation, both groups' partial loads in
// flight together, each summed in CTA order exactly as sk_sum8 does.
OA_DEVINL void sk_sum8_pair(const StreamK& sk, int row, int col_a, int col_b, float (&acc_a)[8], float (&acc_b)[8]) {
    const uint32_t tile = (uint32_t)col_a / (uint32_t)sk.bn;
    if ((uint32_t)col_b / (uint32_t)sk.bn != tile) { sk_sum8(sk, row, col_a, acc_a); sk_sum8(sk, row, col_b, acc_b); return; }
    const uint32_t ca = (uint32_t)col_a - tile * (uint32_t)sk.bn, cb = (uint32_t)col_b - tile * (uint32_t)sk.bn;
    uint32_t c_first, c_last;
    sk_tile_ctas(sk, tile, c_first, c_last);
    const int n = (int)(c_last - c_first) + 1;
    const size_t slot_stride = (size_t)sk.rows * sk.bn;
    const float* p = sk.ws + ((size_t)(c_first + tile) * sk.rows + row) * sk.bn;
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc_a[i] = 0.f; acc_b[i] = 0.f; }
    for (int i = 0; i < n; i += 4) {
        float4 a0[4], a1[4], b0[4], b1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j < n) {
                const float* q = p + (size_t)(i + j) * slot_stride;
                a0[j] = __ldcg(reinterpret_cast<const float4*>(q + ca)); a1[j] = __ldcg(reinterpret_cast<const float4*>(q + ca) + 1);
                b0[j] = __ldcg(reinterpret_cast<const float4*>(q + cb)); b1[j] = __ldcg(reinterpret_cast<const float4*>(q + cb) + 1);
            } else { a0[j] = make_float4(0.f, 0.f, 0.f, 0.f); a1[j] = a0[j]; b0[j] = a0[j]; b1[j] = a0[j]; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j < n) {
                acc_a[0] += a0[j].x; acc_a[1] += a0[j].y; acc_a[2] += a0[j].z; acc_a[3] += a0[j].w;
                acc_a[4] += a1[j].x; acc_a[5] += a1[j].y; acc_a[6] += a1[j].z; acc_a[7] += a1[j].w;
                acc_b[0] += b0[j].x; acc_b[1] += b0[j].y; acc_b[2] += b0[j].z; acc_b[3] += b0[j].w;
                acc_b[4] += b1[j].x; acc_b[5] += b1[j].y; acc_b[6] += b1[j].z; acc_b[7] += b1[j].w;
            }
        }
    }
}

// x[t,:] += sum of partials (bf16 residual stream), y[t,:] = rmsnorm(x[t,:]) * g.  SK_RESID_THREADS threads (tid) per row, so that
// at H <= 4096 every thread owns ONE 16-byte item and all its partial loads are in flight together; `red` = 16 floats of
// shared memory, `bar_id` = a hardware barrier those threads own.
constexpr int SK_RESID_THREADS = 512;
template <int VPT, bool SUM = true>     // SUM == false: the residual stream already holds x + projection (fused GEMM epilogue); only normalise
OA_DEVINL void sk_resid_rmsnorm_row(const StreamK& sk, int t, int tid, float* red, int bar_id, uint4* __restrict__ x,
                                    const uint4* __restrict__ g, uint4* __restrict__ y, int H8, float inv_h, float eps) {
    uint4* xr = x + (size_t)t * H8;
    uint4 v[VPT], gg[VPT];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int i = tid + k * SK_RESID_THREADS;
        if (i < H8) {
            const uint4 xo = __ldcg(xr + i);
            gg[k] = g[i];                              // needed only after the reduction: its latency hides behind the partial loads
            uint4 xn = xo;
            if constexpr (SUM) {
                float acc[8];
                sk_sum8(sk, t, i * 8, acc);
                xn.x = pack_bf16x2(bf16lo(xo.x) + acc[0], bf16hi(xo.x) + acc[1]); xn.y = pack_bf16x2(bf16lo(xo.y) + acc[2], bf16hi(xo.y) + acc[3]);
                xn.z = pack_bf16x2(bf16lo(xo.z) + acc[4], bf16hi(xo.z) + acc[5]); xn.w = pack_bf16x2(bf16lo(xo.w) + acc[6], bf16hi(xo.w) + acc[7]);
                xr[i] = xn;
            }
            v[k] = xn;
            float a;
            a = bf16lo(xn.x); ss += a * a; a = bf16hi(xn.x); ss += a * a; a = bf16lo(xn.y); ss += a * a; a = bf16hi(xn.y); ss += a * a;
            a = bf16lo(xn.z); ss += a * a; a = bf16hi(xn.z); ss += a * a; a = bf16lo(xn.w); ss += a * a; a = bf16hi(xn.w); ss += a * a;
        }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    named_bar_sync(bar_id, SK_RESID_THREADS);
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < SK_RESID_THREADS / 32; ++w) tot += red[w];
    const float r = 1.0f / sqrtf(tot * inv_h + eps);
    uint4* yr = y + (size_t)t * H8;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int i = tid + k * SK_RESID_THREADS;
        if (i < H8) {
            uint4 o;
            o.x = pack_bf16x2(bf16lo(v[k].x) * r * bf16lo(gg[k].x), bf16hi(v[k].x) * r * bf16hi(gg[k].x));
            o.y = pack_bf16x2(bf16lo(v[k].y) * r * bf16lo(gg[k].y), bf16hi(v[k].y) * r * bf16hi(gg[k].y));
            o.z = pack_bf16x2(bf16lo(v[k].z) * r * bf16lo(gg[k].z), bf16hi(v[k].z) * r * bf16hi(gg[k].z));
            o.w = pack_bf16x2(bf16lo(v[k].w) * r * bf16lo(gg[k].w), bf16hi(v[k].w) * r * bf16hi(gg[k].w));
            yr[i] = o;
        }
    }
}

// act[t, i*8 .. i*8+8) = silu(gate) * up.  Physical columns of the fused gate/up GEMM: per 32-column block, 16 gate then 16 up.
OA_DEVINL void sk_swiglu_item(const StreamK& sk, int t, int i, uint4* __restrict__ act, int F8) {
    const int f0 = i * 8, blk = f0 >> 4, o = f0 & 15;
    float gt[8], up[8];
    sk_sum8(sk, t, blk * 32 + o, gt);
    sk_sum8(sk, t, blk * 32 + 16 + o, up);
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __fdividef(gt[k], 1.0f + __expf(-gt[k])) * up[k];
    uint4 ov;
    ov.x = pack_bf16x2(f[0], f[1]); ov.y = pack_bf16x2(f[2], f[3]); ov.z = pack_bf16x2(f[4], f[5]); ov.w = pack_bf16x2(f[6], f[7]);
    act[(size_t)t * F8 + i] = ov;
}

// + bias, then round: the projection output is a bf16 tensor
OA_DEVINL void sk_finish8_bf16(const uint16_t* bias, int col, float (&v)[8]) {
    if (bias) {
        const uint4 bb = *reinterpret_cast<const uint4*>(bias + col);
        v[0] += bf16lo(bb.x); v[1] += bf16hi(bb.x); v[2] += bf16lo(bb.y); v[3] += bf16hi(bb.y);
        v[4] += bf16lo(bb.z); v[5] += bf16hi(bb.z); v[6] += bf16lo(bb.w); v[7] += bf16hi(bb.w);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = bf16_bits_to_f32(f32_to_bf16_bits(v[k]));
}
OA_DEVINL void sk_load8_bf16(const StreamK& sk, const uint16_t* bias, int row, int col, float (&v)[8]) {
    sk_sum8(sk, row, col, v);
    if (bias) {
        const uint4 bb = *reinterpret_cast<const uint4*>(bias + col);
        v[0] += bf16lo(bb.x); v[1] += bf16hi(bb.x); v[2] += bf16lo(bb.y); v[3] += bf16hi(bb.y);
        v[4] += bf16lo(bb.z); v[5] += bf16hi(bb.z); v[6] += bf16lo(bb.w); v[7] += bf16hi(bb.w);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = bf16_bits_to_f32(f32_to_bf16_bits(v[k]));     // the projection output is a bf16 tensor
}

OA_DEVINL int sk_rope_items_per_row(const SkRopeArgs& a) { return (a.nh + a.nkv) * (a.D >> 4) + a.nkv * (a.D >> 3); }
// item w of token row t: w < (nh+nkv)*D/16 rotates 8 (lo, hi) pairs of one q or k head; the rest copy 8 elements of v
OA_DEVINL void sk_rope_item(const StreamK& sk, const SkRopeArgs& a, int t, int w) {
    const int D = a.D, half = D >> 1, vec_per_head = half >> 3;
    const int slot = a.slots[t];
    const int page = slot / a.page_size, off = slot - page * a.page_size;
    const int n_rot = (a.nh + a.nkv) * vec_per_head;
    if (w < n_rot) {
        const int pos = a.positions[t];
        const float* cr = a.rope_cos + (size_t)pos * half;
        const float* sr = a.rope_sin + (size_t)pos * half;
        const int head = w / vec_per_head, i0 = (w - head * vec_per_head) * 8;
        float av[8], bv[8];
        sk_sum8_pair(sk, t, head * D + i0, head * D + i0 + half, av, bv);
        sk_finish8_bf16(a.bias, head * D + i0, av);
        sk_finish8_bf16(a.bias, head * D + i0 + half, bv);
        const float4 c0 = *reinterpret_cast<const float4*>(cr + i0), c1 = *reinterpret_cast<const float4*>(cr + i0 + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(sr + i0), s1 = *reinterpret_cast<const float4*>(sr + i0 + 4);
        const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        float ra[8], rb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { ra[k] = av[k] * cv[k] - bv[k] * sv[k]; rb[k] = bv[k] * cv[k] + av[k] * sv[k]; }
        uint4 oa, ob;
        oa.x = pack_bf16x2(ra[0], ra[1]); oa.y = pack_bf16x2(ra[2], ra[3]); oa.z = pack_bf16x2(ra[4], ra[5]); oa.w = pack_bf16x2(ra[6], ra[7]);
        ob.x = pack_bf16x2(rb[0], rb[1]); ob.y = pack_bf16x2(rb[2], rb[3]); ob.z = pack_bf16x2(rb[4], rb[5]); ob.w = pack_bf16x2(rb[6], rb[7]);
        uint16_t* dst;
        if (head < a.nh) dst = a.q_out + (size_t)t * a.nh * D + (size_t)head * D;
        else dst = a.kv_base + (size_t)(a.k_plane_row0 + ((int64_t)page * a.nkv + (head - a.nh)) * a.page_size + off) * D;
        *reinterpret_cast<uint4*>(dst + i0) = oa;
        *reinterpret_cast<uint4*>(dst + i0 + half) = ob;
    } else {
        const int wv = w - n_rot;
        const int head = wv / (D >> 3), i0 = (wv - head * (D >> 3)) * 8;
        float vv[8];
        sk_load8_bf16(sk, a.bias, t, (a.nh + a.nkv + head) * D + i0, vv);
        uint4 o;
        o.x = pack_bf16x2(vv[0], vv[1]); o.y = pack_bf16x2(vv[2], vv[3]); o.z = pack_bf16x2(vv[4], vv[5]); o.w = pack_bf16x2(vv[6], vv[7]);
        uint16_t* dst = a.kv_base + (size_t)(a.v_plane_row0 + ((int64_t)page * a.nkv + head) * a.page_size + off) * D;
        *reinterpret_cast<uint4*>(dst + i0) = o;
    }
}

}  // namespace oa
