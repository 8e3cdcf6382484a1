// elementwise.cu — the HBM-bound glue kernels of the decode path: embedding gather, RMSNorm, row gather,
// RoPE + paged-KV write.  All use 16-byte vector accesses; one CTA per token row.
#include "common.cuh"
#include "kernels.hpp"
#include "sk_consumers.cuh"

namespace oa {

// out[t,:] = table[ids[t],:]   (ids outside [0,vocab) read row 0, as the oracle does)
__global__ void embed_gather_kernel(const int32_t* __restrict__ ids, const uint4* __restrict__ table, uint4* __restrict__ out,
                                    int H8, int vocab) {
    griddep_launch(); griddep_wait();
    const int t = blockIdx.x;
    int id = ids[t];
    if (id < 0 || id >= vocab) id = 0;
    const uint4* src = table + (size_t)id * H8;
    uint4* dst = out + (size_t)t * H8;
    for (int i = threadIdx.x; i < H8; i += blockDim.x) dst[i] = src[i];
}
cudaError_t launch_embed_gather(const int32_t* ids, const void* table, void* out, int T, int H, int vocab, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    return launch_k(embed_gather_kernel, dim3(T), dim3(128), 0, s, ids, reinterpret_cast<const uint4*>(table), reinterpret_cast<uint4*>(out), H / 8, vocab);
}

__global__ void gather_rows_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ rows, uint4* __restrict__ out, int H8) {
    griddep_launch(); griddep_wait();
    const uint4* src = x + (size_t)rows[blockIdx.x] * H8;
    uint4* dst = out + (size_t)blockIdx.x * H8;
    for (int i = threadIdx.x; i < H8; i += blockDim.x) dst[i] = src[i];
}
cudaError_t launch_gather_rows(const void* x, const int32_t* rows, void* out, int n, int H, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    return launch_k(gather_rows_kernel, dim3(n), dim3(128), 0, s, reinterpret_cast<const uint4*>(x), rows, reinterpret_cast<uint4*>(out), H / 8);
}

// y = x * rsqrt(mean(x^2) + eps) * g ; fp32 statistics, bf16 in/out.  Row cached in registers (H <= 8192).
template <int VPT>   // uint4 vectors per thread, 256 threads
__global__ void __launch_bounds__(256) rmsnorm_kernel(const uint4* __restrict__ x, const uint4* __restrict__ g, uint4* __restrict__ y,
                                                      int H8, float inv_h, float eps) {
    griddep_launch(); griddep_wait();
    const int t = blockIdx.x;
    const uint4* xr = x + (size_t)t * H8;
    uint4 v[VPT];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        int i = threadIdx.x + k * 256;
        if (i < H8) {
            v[k] = xr[i];
            float a;
            a = bf16lo(v[k].x); ss += a * a; a = bf16hi(v[k].x); ss += a * a;
            a = bf16lo(v[k].y); ss += a * a; a = bf16hi(v[k].y); ss += a * a;
            a = bf16lo(v[k].z); ss += a * a; a = bf16hi(v[k].z); ss += a * a;
            a = bf16lo(v[k].w); ss += a * a; a = bf16hi(v[k].w); ss += a * a;
        }
    }
    ss = warp_sum(ss);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float r = 1.0f / sqrtf(tot * inv_h + eps);
    uint4* yr = y + (size_t)t * H8;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        int i = threadIdx.x + k * 256;
        if (i < H8) {
            uint4 gg = g[i], o;
            o.x = pack_bf16x2(bf16lo(v[k].x) * r * bf16lo(gg.x), bf16hi(v[k].x) * r * bf16hi(gg.x));
            o.y = pack_bf16x2(bf16lo(v[k].y) * r * bf16lo(gg.y), bf16hi(v[k].y) * r * bf16hi(gg.y));
            o.z = pack_bf16x2(bf16lo(v[k].z) * r * bf16lo(gg.z), bf16hi(v[k].z) * r * bf16hi(gg.z));
            o.w = pack_bf16x2(bf16lo(v[k].w) * r * bf16lo(gg.w), bf16hi(v[k].w) * r * bf16hi(gg.w));
            yr[i] = o;
        }
    }
}
cudaError_t launch_rmsnorm(const void* x, const void* gain, void* y, int T, int H, float eps, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (H % 8 != 0 || H > 8192) return cudaErrorInvalidValue;
    const int H8 = H / 8;
    auto X = reinterpret_cast<const uint4*>(x); auto G = reinterpret_cast<const uint4*>(gain); auto Y = reinterpret_cast<uint4*>(y);
    if (H8 <= 256) return launch_k(rmsnorm_kernel<1>, dim3(T), dim3(256), 0, s, X, G, Y, H8, 1.0f / H, eps);
    if (H8 <= 512) return launch_k(rmsnorm_kernel<2>, dim3(T), dim3(256), 0, s, X, G, Y, H8, 1.0f / H, eps);
    return launch_k(rmsnorm_kernel<4>, dim3(T), dim3(256), 0, s, X, G, Y, H8, 1.0f / H, eps);
}

// One CTA per token.  Thread i handles the rotate-half pair group (8 consecutive i in [0,D/2) and the
// matching 8 in [D/2,D)) of one head: two 16-byte loads, two 16-byte stores.
__global__ void rope_kv_write_kernel(const uint16_t* __restrict__ qkv, const int32_t* __restrict__ positions,
                                     const int32_t* __restrict__ slots, const float* __restrict__ rope_cos,
                                     const float* __restrict__ rope_sin, uint16_t* __restrict__ q_out, uint16_t* __restrict__ kv_base,
                                     int64_t k_plane_row0, int64_t v_plane_row0, int page_size, int nh, int nkv, int D) {
    griddep_launch(); griddep_wait();
    const int t = blockIdx.x;
    const int half = D >> 1, vec_per_head = half >> 3;        // 8 pairs per thread
    const int qkv_w = (nh + 2 * nkv) * D;
    const uint16_t* row = qkv + (size_t)t * qkv_w;
    const int pos = positions[t];
    const int slot = slots[t];
    const int page = slot / page_size, off = slot - page * page_size;
    const float* cr = rope_cos + (size_t)pos * half;
    const float* sr = rope_sin + (size_t)pos * half;
    const int n_rot = (nh + nkv) * vec_per_head;
    for (int w = threadIdx.x; w < n_rot; w += blockDim.x) {
        const int head = w / vec_per_head, i0 = (w - head * vec_per_head) * 8;
        const uint16_t* src = row + (size_t)head * D;
        uint4 a = *reinterpret_cast<const uint4*>(src + i0);
        uint4 b = *reinterpret_cast<const uint4*>(src + i0 + half);
        float4 c0 = *reinterpret_cast<const float4*>(cr + i0), c1 = *reinterpret_cast<const float4*>(cr + i0 + 4);
        float4 s0 = *reinterpret_cast<const float4*>(sr + i0), s1 = *reinterpret_cast<const float4*>(sr + i0 + 4);
        float av[8] = {bf16lo(a.x), bf16hi(a.x), bf16lo(a.y), bf16hi(a.y), bf16lo(a.z), bf16hi(a.z), bf16lo(a.w), bf16hi(a.w)};
        float bv[8] = {bf16lo(b.x), bf16hi(b.x), bf16lo(b.y), bf16hi(b.y), bf16lo(b.z), bf16hi(b.z), bf16lo(b.w), bf16hi(b.w)};
        float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        float ra[8], rb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { ra[k] = av[k] * cv[k] - bv[k] * sv[k]; rb[k] = bv[k] * cv[k] + av[k] * sv[k]; }
        uint4 oa, ob;
        oa.x = pack_bf16x2(ra[0], ra[1]); oa.y = pack_bf16x2(ra[2], ra[3]); oa.z = pack_bf16x2(ra[4], ra[5]); oa.w = pack_bf16x2(ra[6], ra[7]);
        ob.x = pack_bf16x2(rb[0], rb[1]); ob.y = pack_bf16x2(rb[2], rb[3]); ob.z = pack_bf16x2(rb[4], rb[5]); ob.w = pack_bf16x2(rb[6], rb[7]);
        uint16_t* dst;
        if (head < nh) dst = q_out + (size_t)t * nh * D + (size_t)head * D;
        else dst = kv_base + (size_t)(k_plane_row0 + ((int64_t)page * nkv + (head - nh)) * page_size + off) * D;
        *reinterpret_cast<uint4*>(dst + i0) = oa;
        *reinterpret_cast<uint4*>(dst + i0 + half) = ob;
    }
    // V: straight copy into the V plane
    const int n_v = nkv * (D >> 3);
    for (int w = threadIdx.x; w < n_v; w += blockDim.x) {
        const int head = w / (D >> 3), i0 = (w - head * (D >> 3)) * 8;
        uint4 vv = *reinterpret_cast<const uint4*>(row + (size_t)(nh + nkv + head) * D + i0);
        uint16_t* dst = kv_base + (size_t)(v_plane_row0 + ((int64_t)page * nkv + head) * page_size + off) * D;
        *reinterpret_cast<uint4*>(dst + i0) = vv;
    }
}
cudaError_t launch_rope_kv_write(const void* qkv, const int32_t* positions, const int32_t* slots, const float* rope_cos,
                                 const float* rope_sin, void* q_out, const KvLayout& kv, int layer, int T, int nh, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (kv.head_dim % 16 != 0) return cudaErrorInvalidValue;
    const int64_t k0 = (int64_t)layer * kv.layer_stride_rows, v0 = k0 + kv.kv_stride_rows;
    return launch_k(rope_kv_write_kernel, dim3(T), dim3(256), 0, s, reinterpret_cast<const uint16_t*>(qkv), positions, slots, rope_cos, rope_sin,
                    reinterpret_cast<uint16_t*>(q_out), reinterpret_cast<uint16_t*>(kv.base), k0, v0, kv.page_size, nh, kv.n_kv, kv.head_dim);
}


// =============================================================================================
// stream-K consumers (stand-alone kernels; the bodies live in sk_consumers.cuh)
// =============================================================================================
template <int VPT>
__global__ void __launch_bounds__(SK_RESID_THREADS) sk_resid_rmsnorm_kernel(const StreamK sk, uint4* __restrict__ x, const uint4* __restrict__ g,
                                                                            uint4* __restrict__ y, int H8, float inv_h, float eps) {
    griddep_launch(); griddep_wait();
    __shared__ float red[SK_RESID_THREADS / 32];
    sk_resid_rmsnorm_row<VPT>(sk, blockIdx.x, threadIdx.x, red, 0, x, g, y, H8, inv_h, eps);
}
cudaError_t launch_sk_resid_rmsnorm(const StreamK& sk, void* x, const void* gain, void* xn, int T, int H, float eps, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (H % 8 != 0 || H > 8192 || T > sk.rows) return cudaErrorInvalidValue;
    const int H8 = H / 8;
    auto X = reinterpret_cast<uint4*>(x); auto G = reinterpret_cast<const uint4*>(gain); auto Y = reinterpret_cast<uint4*>(xn);
    if (H8 <= SK_RESID_THREADS) return launch_k(sk_resid_rmsnorm_kernel<1>, dim3(T), dim3(SK_RESID_THREADS), 0, s, sk, X, G, Y, H8, 1.0f / H, eps);
    return launch_k(sk_resid_rmsnorm_kernel<2>, dim3(T), dim3(SK_RESID_THREADS), 0, s, sk, X, G, Y, H8, 1.0f / H, eps);
}

template <int VPT>
__global__ void __launch_bounds__(SK_RESID_THREADS) rmsnorm_wide_kernel(uint4* __restrict__ x, const uint4* __restrict__ g, uint4* __restrict__ y, int H8, float inv_h, float eps) {
    griddep_launch(); griddep_wait();
    __shared__ float red[SK_RESID_THREADS / 32];
    sk_resid_rmsnorm_row<VPT, false>(StreamK{}, blockIdx.x, threadIdx.x, red, 0, x, g, y, H8, inv_h, eps);
}
cudaError_t launch_rmsnorm_wide(const void* x, const void* gain, void* xn, int T, int H, float eps, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (H % 8 != 0 || H > 8192) return cudaErrorInvalidValue;
    const int H8 = H / 8;
    auto X = reinterpret_cast<uint4*>(const_cast<void*>(x)); auto G = reinterpret_cast<const uint4*>(gain); auto Y = reinterpret_cast<uint4*>(xn);
    if (H8 <= SK_RESID_THREADS) return launch_k(rmsnorm_wide_kernel<1>, dim3(T), dim3(SK_RESID_THREADS), 0, s, X, G, Y, H8, 1.0f / H, eps);
    return launch_k(rmsnorm_wide_kernel<2>, dim3(T), dim3(SK_RESID_THREADS), 0, s, X, G, Y, H8, 1.0f / H, eps);
}

__global__ void sk_swiglu_kernel(const StreamK sk, uint4* __restrict__ act, int F8) {
    griddep_launch(); griddep_wait();
    const int t = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F8; i += gridDim.x * blockDim.x) sk_swiglu_item(sk, t, i, act, F8);
}
cudaError_t launch_sk_swiglu(const StreamK& sk, void* act, int T, int F, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (F % 16 != 0 || T > sk.rows) return cudaErrorInvalidValue;
    return launch_k(sk_swiglu_kernel, dim3((F / 8 + 255) / 256, T), dim3(256), 0, s, sk, reinterpret_cast<uint4*>(act), F / 8);
}

__global__ void sk_rope_kv_write_kernel(const StreamK sk, const SkRopeArgs a) {
    griddep_launch(); griddep_wait();
    const int t = blockIdx.y, items = sk_rope_items_per_row(a);
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < items; w += gridDim.x * blockDim.x) sk_rope_item(sk, a, t, w);
}
SkRopeArgs make_sk_rope_args(const void* bias, const int32_t* positions, const int32_t* slots, const float* rope_cos, const float* rope_sin,
                             void* q_out, const KvLayout& kv, int layer, int nh) {
    SkRopeArgs a{};
    a.bias = reinterpret_cast<const uint16_t*>(bias); a.positions = positions; a.slots = slots; a.rope_cos = rope_cos; a.rope_sin = rope_sin;
    a.q_out = reinterpret_cast<uint16_t*>(q_out); a.kv_base = reinterpret_cast<uint16_t*>(kv.base);
    a.k_plane_row0 = (int64_t)layer * kv.layer_stride_rows; a.v_plane_row0 = a.k_plane_row0 + kv.kv_stride_rows;
    a.page_size = kv.page_size; a.nh = nh; a.nkv = kv.n_kv; a.D = kv.head_dim;
    return a;
}
cudaError_t launch_sk_rope_kv_write(const StreamK& sk, const void* bias, const int32_t* positions, const int32_t* slots,
                                    const float* rope_cos, const float* rope_sin, void* q_out, const KvLayout& kv, int layer, int T,
                                    int nh, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (kv.head_dim % 16 != 0 || T > sk.rows) return cudaErrorInvalidValue;
    const SkRopeArgs a = make_sk_rope_args(bias, positions, slots, rope_cos, rope_sin, q_out, kv, layer, nh);
    const int items = (nh + kv.n_kv) * (kv.head_dim / 16) + kv.n_kv * (kv.head_dim / 8);
    return launch_k(sk_rope_kv_write_kernel, dim3((items + 127) / 128, T), dim3(128), 0, s, sk, a);
}

}  // namespace oa
