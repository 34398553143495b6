// a1: channel L2-normalisation fused with the NCHW -> [N, D] rearrangement, and its backward.
// ref: model.py:40-41, :210-211 (F.normalize(p=2, dim=1) then 'b c h w -> (b h w) c').
//
// HBM-bound: reads x once (4*N*D bytes), writes xhat once (+ optional NCHW copy).  A CTA
// owns 32 consecutive patches of one image; a warp reads 32 consecutive hw of one channel
// (one 128 B line), the tile is transposed through shared memory (pitch 33) and each patch
// row is written with D contiguous floats.
#include "mgp_common.cuh"
#include <cuda_fp16.h>

// logprob_tc.cu: where the tensor-core log-likelihood kernels expect the patch-side operands inside their workspace
bool mgp_logprob_tc_stage_ptrs(void* ws, size_t ws_bytes, long long N, int P, int D, __half** ah, __half** al, float** sn);

namespace {

constexpr int NT = 32;  // patches per CTA

__global__ void __launch_bounds__(256) normalize_fwd_kernel(const float* __restrict__ x, float* __restrict__ xhat,
                                                            float* __restrict__ inv_norm,
                                                            float* __restrict__ xhat_nchw, int D, int HW,
                                                            __half* __restrict__ ah, __half* __restrict__ al,
                                                            float* __restrict__ sn, int stage_aniso) {
    extern __shared__ float tile[];  // [D][NT+1]
    __shared__ float red[8][NT];
    __shared__ float s_inv[NT];
    const int b = blockIdx.y;
    const int hw0 = blockIdx.x * NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hw = hw0 + lane;
    const bool ok = hw < HW;
    const float* xb = x + (size_t)b * D * HW;
    float ss = 0.f;
    for (int d = warp; d < D; d += 8) {
        float v = ok ? __ldg(xb + (size_t)d * HW + hw) : 0.f;
        tile[d * (NT + 1) + lane] = v;
        ss += v * v;
    }
    red[warp][lane] = ss;
    __syncthreads();
    if (warp == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][lane];
        float inv = 1.0f / fmaxf(sqrtf(t), 1e-12f);
        s_inv[lane] = inv;
        if (ok) inv_norm[(size_t)b * HW + hw] = inv;
    }
    __syncthreads();
    // [N, D] rows: warp w writes patches w, w+8, ...; lanes run over d (conflict-free: pitch 33)
    for (int r = warp; r < NT; r += 8) {
        if (hw0 + r >= HW) break;
        const float inv = s_inv[r];
        const size_t n = (size_t)b * HW + hw0 + r;
        float* dst = xhat + n * D;
        if (ah == nullptr) {
            for (int d = lane; d < D; d += 32) dst[d] = tile[d * (NT + 1) + r] * inv;
        } else {
            // ... and the operands the tensor-core log-likelihood kernels read (csrc/logprob_tc.cu tc_x_prep_kernel: rows
            // of [N, 2D] fp16, hi / lo split of 256 * [ x^2 | x ], the x^2 half only on request; sn = |xhat|^2)
            __half* hr = ah + n * 2 * D;
            __half* lr = al + n * 2 * D;
            float ss = 0.f;
            for (int d = lane; d < D; d += 32) {
                const float a = tile[d * (NT + 1) + r] * inv;
                dst[d] = a;
                ss = fmaf(a, a, ss);
                const float s1 = a * 256.0f;
                const __half h = __float2half_rn(s1);
                hr[D + d] = h;
                lr[D + d] = __float2half_rn(s1 - __half2float(h));
                if (stage_aniso) {
                    const float s2 = a * a * 256.0f;
                    const __half h2 = __float2half_rn(s2);
                    hr[d] = h2;
                    lr[d] = __float2half_rn(s2 - __half2float(h2));
                }
            }
            ss = warp_sum(ss);
            if (lane == 0) sn[n] = ss;
        }
    }
    if (xhat_nchw != nullptr && ok) {
        const float inv = s_inv[lane];
        float* dst = xhat_nchw + (size_t)b * D * HW + hw;
        for (int d = warp; d < D; d += 8) dst[(size_t)d * HW] = tile[d * (NT + 1) + lane] * inv;
    }
}

// g_x = (g - xhat <xhat, g>) * inv_norm, written back in NCHW.
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xhat,
                                                            const float* __restrict__ inv_norm,
                                                            float* __restrict__ gx, int D, int HW) {
    extern __shared__ float tile[];  // [D][NT+1] holds g_x rows
    const int b = blockIdx.y;
    const int hw0 = blockIdx.x * NT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < NT; r += 8) {
        if (hw0 + r >= HW) break;
        const size_t n = (size_t)b * HW + hw0 + r;
        const float* gr = g + n * D;
        const float* xr = xhat + n * D;
        float dot = 0.f;
        for (int d = lane; d < D; d += 32) dot += gr[d] * xr[d];
        dot = warp_sum(dot);
        const float inv = inv_norm[n];
        for (int d = lane; d < D; d += 32) tile[d * (NT + 1) + r] = (gr[d] - xr[d] * dot) * inv;
    }
    __syncthreads();
    const int hw = hw0 + lane;
    if (hw < HW) {
        float* dst = gx + (size_t)b * D * HW + hw;
        for (int d = warp; d < D; d += 8) dst[(size_t)d * HW] = tile[d * (NT + 1) + lane];
    }
}

}  // namespace

extern "C" int mgp_normalize_fwd(const float* x_nchw, float* xhat_nd, float* inv_norm, float* xhat_nchw, int B,
                                 int D, int HW, void* stream) {
    if (!x_nchw || !xhat_nd || !inv_norm || B <= 0 || D <= 0 || HW <= 0) return MGP_ERR_INVALID;
    size_t smem = (size_t)D * (NT + 1) * sizeof(float);
    if (smem > 200 * 1024) return MGP_ERR_UNSUPPORTED;
    MGP_CUDA(cudaFuncSetAttribute(normalize_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((HW + NT - 1) / NT, B);
    normalize_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x_nchw, xhat_nd, inv_norm, xhat_nchw, D, HW, nullptr,
                                                                    nullptr, nullptr, 0);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}

extern "C" int mgp_normalize_fwd_stage(const float* x_nchw, float* xhat_nd, float* inv_norm, float* xhat_nchw, void* ws,
                                       size_t ws_bytes, int B, int D, int HW, int P, int stage_aniso, void* stream) {
    if (!x_nchw || !xhat_nd || !inv_norm || !ws || B <= 0 || D <= 0 || HW <= 0 || P <= 0) return MGP_ERR_INVALID;
#ifdef MGP_WITH_TC
    __half *ah = nullptr, *al = nullptr;
    float* sn = nullptr;
    if (!mgp_logprob_tc_stage_ptrs(ws, ws_bytes, (long long)B * HW, P, D, &ah, &al, &sn)) return MGP_ERR_WORKSPACE;
    size_t smem = (size_t)D * (NT + 1) * sizeof(float);
    if (smem > 200 * 1024) return MGP_ERR_UNSUPPORTED;
    MGP_CUDA(cudaFuncSetAttribute(normalize_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((HW + NT - 1) / NT, B);
    normalize_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x_nchw, xhat_nd, inv_norm, xhat_nchw, D, HW, ah, al, sn,
                                                                    stage_aniso ? 1 : 0);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
#else
    (void)xhat_nchw; (void)ws_bytes; (void)stage_aniso; (void)stream;
    return MGP_ERR_UNSUPPORTED;
#endif
}

extern "C" int mgp_normalize_bwd(const float* g_xhat_nd, const float* xhat_nd, const float* inv_norm,
                                 float* g_x_nchw, int B, int D, int HW, void* stream) {
    if (!g_xhat_nd || !xhat_nd || !inv_norm || !g_x_nchw || B <= 0 || D <= 0 || HW <= 0) return MGP_ERR_INVALID;
    size_t smem = (size_t)D * (NT + 1) * sizeof(float);
    if (smem > 200 * 1024) return MGP_ERR_UNSUPPORTED;
    MGP_CUDA(cudaFuncSetAttribute(normalize_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((HW + NT - 1) / NT, B);
    normalize_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(g_xhat_nd, xhat_nd, inv_norm, g_x_nchw, D, HW);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}
