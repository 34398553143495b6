// a4-a7: top-T prototype mining over the patches of each image, wrong-class rule, block-diagonal
// pi mix and log; its backward; and the push-projection argmin (f1).
// ref: model.py:188-206, :214-222, :254, :54-74; push.py:125-158.
#include "mgp_common.cuh"

namespace {

__device__ __forceinline__ unsigned f2key(float f) {  // monotone float -> uint (larger float = larger key)
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// One warp selects the T largest of one [HW] row (descending, ties -> smaller index).
// Each lane keeps R = ceil(HW/32) keys in registers; per level: warp REDUX.max on the lanes'
// local maxima, REDUX.min on the index among equal maxima, winner removed from its lane.
template <int R>
__device__ __forceinline__ void warp_topT(const float* __restrict__ row, int HW, int T, int lane, float& out_v,
                                          int& out_i) {
    unsigned key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = lane + 32 * r;
        key[r] = (i < HW) ? f2key(__ldg(row + i)) : 0u;  // 0 sorts below every real float (incl. -inf)
    }
    out_v = 0.f;
    out_i = 0;
    for (int t = 0; t < T; ++t) {
        unsigned lm = key[0];
#pragma unroll
        for (int r = 1; r < R; ++r) lm = max(lm, key[r]);
        int li = 0x7fffffff;
#pragma unroll
        for (int r = R - 1; r >= 0; --r)
            if (key[r] == lm) li = lane + 32 * r;
        const unsigned best = __reduce_max_sync(0xffffffffu, lm);
        const int bi = __reduce_min_sync(0xffffffffu, (lm == best) ? li : 0x7fffffff);
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (bi == lane + 32 * r) key[r] = 0u;
        if (lane == t) {
            out_v = key2f(best);
            out_i = bi;
        }
    }
}

template <int R>
__global__ void __launch_bounds__(256)
head_select_kernel(const float* __restrict__ logp, const float* __restrict__ weight, const int64_t* __restrict__ gt,
                   float* __restrict__ logits, float* __restrict__ vals, int32_t* __restrict__ idx, int HW, int C,
                   int K, int T, int CT) {
    extern __shared__ float win[];  // [CT*K][T] exp(log p) of the winners
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * CT;
    const int nc = min(CT, C - c0);
    const int P = C * K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npl = nc * K;
    for (int pl = warp; pl < npl; pl += 8) {
        const int p = c0 * K + pl;
        float v;
        int i;
        warp_topT<R>(logp + ((size_t)b * P + p) * HW, HW, T, lane, v, i);
        if (lane < T) {
            const float e = expf(v);  // ref model.py:215
            win[pl * T + lane] = e;
            vals[((size_t)b * P + p) * T + lane] = e;
            idx[((size_t)b * P + p) * T + lane] = i;
        }
    }
    __syncthreads();
    const long long g = (gt != nullptr) ? (long long)gt[b] : -1;
    for (int e = threadIdx.x; e < nc * T; e += blockDim.x) {
        const int cl = e / T, t = e - cl * T;
        const int c = c0 + cl;
        const bool fold = (gt != nullptr) && ((long long)c != g) && (t > 0);  // ref model.py:218-221
        const float* wrow = weight + (size_t)c * P + (size_t)c * K;            // class-diagonal block of last_layer.weight
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(__ldg(wrow + k), win[(cl * K + k) * T + (fold ? 0 : t)], s);
        logits[((size_t)b * C + c) * T + t] = logf(s);                         // ref model.py:222, :254
    }
}

// ------------------------------------------------------------------------------------------
// Backward: for image b accumulate  G[n,:] = sum_{(p,t)->n} a_bpt * (w_p*mu_p - w_p*xhat_n)
//   a_bpt = gl[b,c,t] * pi_p * v[b,p,t] / exp(logits[b,c,t])      (wrong-class levels fold onto t = 0)
// One CTA per (image, D-chunk).  Non-zero entries are compacted in a fixed order into a
// shared-memory list; warp w owns rows n with (n & 7) == w and walks the list, so the
// accumulation is atomics-free and deterministic.
__global__ void proto_weight_kernel(const float* __restrict__ mu, const float* __restrict__ sigma,
                                    float* __restrict__ w, float* __restrict__ wm, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float s = sigma[i];
        float ww = 1.0f / (s * s);
        w[i] = ww;
        wm[i] = ww * mu[i];
    }
}

constexpr int LCAP = 4096;

__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ gl, const float* __restrict__ logits, const float* __restrict__ vals,
                const int32_t* __restrict__ idx, const float* __restrict__ weight, const int64_t* __restrict__ gt,
                const float* __restrict__ xhat, const float* __restrict__ w, const float* __restrict__ wm,
                float* __restrict__ g_xhat, int HW, int C, int K, int D, int T, int DC) {
    extern __shared__ float smem[];
    float* G = smem;                                        // [HW][DC+1]
    unsigned* lkey = reinterpret_cast<unsigned*>(G + (size_t)HW * (DC + 1));  // [LCAP] p*1024 + n
    float* lval = reinterpret_cast<float*>(lkey + LCAP);    // [LCAP]
    __shared__ int wcount[8];
    __shared__ int lcount;

    const int b = blockIdx.x;
    const int d0 = blockIdx.y * DC;
    const int dc = min(DC, D - d0);
    const int P = C * K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long g = (gt != nullptr) ? (long long)gt[b] : -1;

    for (int i = threadIdx.x; i < HW * (DC + 1); i += 256) G[i] = 0.f;
    if (threadIdx.x == 0) lcount = 0;
    __syncthreads();

    const int E = P * T;
    for (int e0 = 0; e0 < E; e0 += 256) {
        const int e = e0 + threadIdx.x;
        float a = 0.f;
        unsigned key = 0;
        if (e < E) {
            const int p = e / T, t = e - p * T;
            const int c = p / K;
            const bool wrong = (gt != nullptr) && ((long long)c != g);
            const float pi = __ldg(weight + (size_t)c * P + p);
            const size_t lo = ((size_t)b * C + c) * T;
            if (!wrong) {
                const float v = vals[((size_t)b * P + p) * T + t];
                a = gl[lo + t] * pi * v / expf(logits[lo + t]);
                key = (unsigned)p * 1024u + (unsigned)idx[((size_t)b * P + p) * T + t];
            } else if (t == 0) {
                float q = 0.f;
                for (int tt = 0; tt < T; ++tt) q += gl[lo + tt] / expf(logits[lo + tt]);
                a = q * pi * vals[((size_t)b * P + p) * T];
                key = (unsigned)p * 1024u + (unsigned)idx[((size_t)b * P + p) * T];
            }
        }
        const bool keep = (a != 0.f);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        int base = lcount;
        for (int wv = 0; wv < warp; ++wv) base += wcount[wv];
        if (keep) {
            const int pos = base + __popc(bal & ((1u << lane) - 1u));
            lkey[pos] = key;
            lval[pos] = a;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int wv = 0; wv < 8; ++wv) tot += wcount[wv];
            lcount += tot;
        }
        __syncthreads();
        const int cnt = lcount;
        const bool last = (e0 + 256 >= E);
        if (cnt + 256 > LCAP || last) {
            // drain: warp `warp` accumulates the rows it owns
            for (int l0 = 0; l0 < cnt; l0 += 32) {
                const int li = l0 + lane;
                unsigned k2 = 0;
                float a2 = 0.f;
                bool mine = false;
                if (li < cnt) {
                    k2 = lkey[li];
                    a2 = lval[li];
                    mine = ((k2 & 7u) == (unsigned)warp);
                }
                unsigned m = __ballot_sync(0xffffffffu, mine);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned kk = __shfl_sync(0xffffffffu, k2, src);
                    const float aa = __shfl_sync(0xffffffffu, a2, src);
                    const int p = kk >> 10, n = kk & 1023u;
                    const float* wr = w + (size_t)p * D + d0;
                    const float* wmr = wm + (size_t)p * D + d0;
                    const float* xr = xhat + ((size_t)b * HW + n) * D + d0;
                    float* gr = G + (size_t)n * (DC + 1);
                    for (int d = lane; d < dc; d += 32) gr[d] += aa * (__ldg(wmr + d) - __ldg(wr + d) * __ldg(xr + d));
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) lcount = 0;
            __syncthreads();
        }
    }
    __syncthreads();
    for (int n = warp; n < HW; n += 8) {
        float* dst = g_xhat + ((size_t)b * HW + n) * D + d0;
        const float* gr = G + (size_t)n * (DC + 1);
        for (int d = lane; d < dc; d += 32) dst[d] = gr[d];
    }
}

// f1: per (image, prototype of the image's class): argmax_hw log p and -exp(log p) there.
__global__ void push_argmin_kernel(const float* __restrict__ logp, const int64_t* __restrict__ labels,
                                   int32_t* __restrict__ arg, float* __restrict__ val, int HW, int C, int K, int B) {
    const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wg >= B * K) return;
    const int b = wg / K, k = wg - b * K;
    const long long c = labels[b];
    if (c < 0 || c >= C) {
        if (lane == 0) { arg[wg] = -1; val[wg] = 0.f; }
        return;
    }
    const float* row = logp + ((size_t)b * C * K + (size_t)c * K + k) * HW;
    unsigned best = 0;
    int bi = 0x7fffffff;
    for (int i = lane; i < HW; i += 32) {
        unsigned kk = f2key(row[i]);
        if (kk > best) { best = kk; bi = i; }
    }
    const unsigned wb = __reduce_max_sync(0xffffffffu, best);
    const int wi = __reduce_min_sync(0xffffffffu, (best == wb) ? bi : 0x7fffffff);
    if (lane == 0) {
        arg[wg] = wi;
        val[wg] = -expf(key2f(wb));
    }
}

}  // namespace

extern "C" int mgp_head_select(const float* logp_bphw, const float* weight_cp, const int64_t* gt, float* logits,
                               float* vals, int32_t* idx, int B, int HW, int C, int K, int T, void* stream) {
    if (!logp_bphw || !weight_cp || !logits || !vals || !idx) return MGP_ERR_INVALID;
    if (B <= 0 || HW <= 0 || C <= 0 || K <= 0 || T <= 0) return MGP_ERR_INVALID;
    if (T > 32 || T > HW || HW > 1024) return MGP_ERR_UNSUPPORTED;
    int CT = 64 / K;
    if (CT < 1) CT = 1;
    if (CT > C) CT = C;
    size_t smem = (size_t)CT * K * T * sizeof(float);
    if (smem > 160 * 1024) return MGP_ERR_UNSUPPORTED;
    dim3 grid((C + CT - 1) / CT, B);
    cudaStream_t st = (cudaStream_t)stream;
    const int R = (HW + 31) / 32;
#define MGP_LAUNCH_SEL(RR)                                                                                           \
    do {                                                                                                             \
        MGP_CUDA(cudaFuncSetAttribute(head_select_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        head_select_kernel<RR><<<grid, 256, smem, st>>>(logp_bphw, weight_cp, gt, logits, vals, idx, HW, C, K, T, CT); \
    } while (0)
    if (R <= 2) MGP_LAUNCH_SEL(2);
    else if (R <= 7) MGP_LAUNCH_SEL(7);
    else if (R <= 13) MGP_LAUNCH_SEL(13);
    else if (R <= 25) MGP_LAUNCH_SEL(25);
    else MGP_LAUNCH_SEL(32);
#undef MGP_LAUNCH_SEL
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}

extern "C" size_t mgp_head_bwd_ws_bytes(int B, int HW, int P, int D) {
    return ((size_t)2 * P * D + (size_t)B * HW * D) * sizeof(float);
}

extern "C" int mgp_head_bwd(const float* grad_logits, const float* logits, const float* vals, const int32_t* idx,
                            const float* weight_cp, const int64_t* gt, const float* xhat_nd, const float* inv_norm,
                            const float* mu, const float* sigma, void* ws, size_t ws_bytes, float* g_x_nchw, int B,
                            int HW, int C, int K, int D, int T, void* stream) {
    if (!grad_logits || !logits || !vals || !idx || !weight_cp || !xhat_nd || !inv_norm || !mu || !sigma || !ws ||
        !g_x_nchw)
        return MGP_ERR_INVALID;
    if (B <= 0 || HW <= 0 || C <= 0 || K <= 0 || D <= 0 || T <= 0) return MGP_ERR_INVALID;
    if (HW > 1024 || (size_t)C * K >= (1u << 22)) return MGP_ERR_UNSUPPORTED;
    const int P = C * K;
    if (ws_bytes < mgp_head_bwd_ws_bytes(B, HW, P, D)) return MGP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* w = reinterpret_cast<float*>(ws);
    float* wm = w + (size_t)P * D;
    float* g_xhat = wm + (size_t)P * D;
    const size_t npd = (size_t)P * D;
    proto_weight_kernel<<<(unsigned)((npd + 255) / 256), 256, 0, st>>>(mu, sigma, w, wm, npd);
    MGP_CHECK_LAUNCH();
    int DC = 128;
    while (DC > 32 && (size_t)HW * (DC + 1) * 4 > 110 * 1024) DC >>= 1;
    if (DC > D) DC = ((D + 31) / 32) * 32;
    size_t smem = (size_t)HW * (DC + 1) * 4 + (size_t)LCAP * 8;
    if (smem > 220 * 1024) return MGP_ERR_UNSUPPORTED;
    MGP_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(B, (D + DC - 1) / DC);
    head_bwd_kernel<<<grid, 256, smem, st>>>(grad_logits, logits, vals, idx, weight_cp, gt, xhat_nd, w, wm, g_xhat, HW,
                                             C, K, D, T, DC);
    MGP_CHECK_LAUNCH();
    return mgp_normalize_bwd(g_xhat, xhat_nd, inv_norm, g_x_nchw, B, D, HW, stream);
}

extern "C" int mgp_push_argmin(const float* logp_bphw, const int64_t* labels, int32_t* arg, float* val, int B, int HW,
                               int C, int K, void* stream) {
    if (!logp_bphw || !labels || !arg || !val || B <= 0 || HW <= 0 || C <= 0 || K <= 0) return MGP_ERR_INVALID;
    const int warps = B * K;
    push_argmin_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(logp_bphw, labels, arg, val, HW, C, K, B);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}
