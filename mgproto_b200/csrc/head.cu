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

// One warp selects the T largest of NR [HW] rows at once (descending, ties -> smaller index).
// Each lane keeps R = ceil(HW/32) keys per row in registers; per level: warp REDUX.max on the lanes'
// local maxima, REDUX.min on the index among equal maxima, winner removed from its lane.  The NR rows
// are independent dependency chains, interleaved to hide the REDUX latency.
template <int R, int NR>
__device__ __forceinline__ void warp_topT(const float* const (&rows)[NR], int HW, int T, int lane, float (&out_v)[NR],
                                          int (&out_i)[NR]) {
    unsigned key[NR][R];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = lane + 32 * r;
            key[i][r] = (j < HW) ? f2key(__ldg(rows[i] + j)) : 0u;  // 0 sorts below every real float (incl. -inf)
        }
        out_v[i] = 0.f;
        out_i[i] = 0;
    }
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            unsigned lm = key[i][0];
#pragma unroll
            for (int r = 1; r < R; ++r) lm = max(lm, key[i][r]);
            int li = 0x7fffffff;
#pragma unroll
            for (int r = R - 1; r >= 0; --r)
                if (key[i][r] == lm) li = lane + 32 * r;
            const unsigned best = __reduce_max_sync(0xffffffffu, lm);
            const int bi = __reduce_min_sync(0xffffffffu, (lm == best) ? li : 0x7fffffff);
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (bi == lane + 32 * r) key[i][r] = 0u;
            if (lane == t) {
                out_v[i] = key2f(best);
                out_i[i] = bi;
            }
        }
    }
}

// Faster variant for R <= 8 (HW <= 256): every lane first sorts its R keys (descending, compile-time
// compare-exchange network), so a level costs one REDUX.max over the lanes' heads, a ballot to find the
// owner (lowest lane among equal heads) and a predicated pop of the owner's list -- ~15 instructions instead
// of ~45.  Indices are recovered at the end from the owner's unsorted copy: lane t fetches the owner's R
// original keys by shuffle and takes the position of its value; equal values picked twice from one lane
// are disambiguated by their rank among earlier identical picks (MATCH.ANY).
template <int R, int NR>
__device__ __forceinline__ void warp_topT_sorted(const float* const (&rows)[NR], int HW, int T, int lane,
                                                 float (&out_v)[NR], int (&out_i)[NR]) {
    unsigned orig[NR][R], key[NR][R];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = lane + 32 * r;
            orig[i][r] = (j < HW) ? f2key(__ldg(rows[i] + j)) : 0u;
            key[i][r] = orig[i][r];
        }
#pragma unroll
        for (int a = 1; a < R; ++a)
#pragma unroll
            for (int b = a; b >= 1; --b) {
                const unsigned hi = max(key[i][b - 1], key[i][b]), lo = min(key[i][b - 1], key[i][b]);
                key[i][b - 1] = hi;
                key[i][b] = lo;
            }
    }
    unsigned my_key[NR];
    int my_owner[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) { my_key[i] = 0u; my_owner[i] = 0; }
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const unsigned best = __reduce_max_sync(0xffffffffu, key[i][0]);
            const unsigned m = __ballot_sync(0xffffffffu, key[i][0] == best);
            const int owner = __ffs(m) - 1;
            if (lane == owner) {
#pragma unroll
                for (int r = 0; r + 1 < R; ++r) key[i][r] = key[i][r + 1];
                key[i][R - 1] = 0u;
            }
            if (lane == t) { my_key[i] = best; my_owner[i] = owner; }
        }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        // rank of this pick among earlier picks of the same (value, lane)
        const unsigned long long tag = ((unsigned long long)my_key[i] << 8) | (unsigned)my_owner[i];
        const unsigned same = __match_any_sync(0xffffffffu, (lane < T) ? tag : (0xffffffffffffff00ull | (unsigned)lane));
        int skip = __popc(same & ((1u << lane) - 1u));
        int rr = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned o = __shfl_sync(0xffffffffu, orig[i][r], my_owner[i]);
            const bool hit = (o == my_key[i]);
            if (hit && skip == 0) rr = r;
            if (hit) --skip;
        }
        out_v[i] = key2f(my_key[i]);
        out_i[i] = my_owner[i] + 32 * rr;
    }
}

template <int R, int NR>
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
    for (int pl0 = warp * NR; pl0 < npl; pl0 += 8 * NR) {
        const float* rows[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int pl = min(pl0 + i, npl - 1);              // surplus slots recompute the last row, not stored
            rows[i] = logp + ((size_t)b * P + c0 * K + pl) * HW;
        }
        float v[NR];
        int ix[NR];
        if (R <= 8) warp_topT_sorted<(R <= 8 ? R : 1), NR>(rows, HW, T, lane, v, ix);
        else warp_topT<R, NR>(rows, HW, T, lane, v, ix);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int pl = pl0 + i;
            if (pl < npl && lane < T) {
                const int p = c0 * K + pl;
                const float e = expf(v[i]);  // ref model.py:215
                win[pl * T + lane] = e;
                vals[((size_t)b * P + p) * T + lane] = e;
                idx[((size_t)b * P + p) * T + lane] = ix[i];
            }
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

constexpr int BCAP = 640;   // entries per owner-warp bucket

__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ gl, const float* __restrict__ logits, const float* __restrict__ vals,
                const int32_t* __restrict__ idx, const float* __restrict__ weight, const int64_t* __restrict__ gt,
                const float* __restrict__ xhat, const float* __restrict__ w, const float* __restrict__ wm,
                float* __restrict__ g_xhat, int HW, int C, int K, int D, int T, int DC) {
    extern __shared__ float smem[];
    const int pitch = DC + 2;                               // even: float2 accesses stay 8 B aligned
    float* G = smem;                                        // [HW][pitch]
    unsigned* lkey = reinterpret_cast<unsigned*>(G + (size_t)HW * pitch);     // [8][BCAP] p*1024 + n
    float* lval = reinterpret_cast<float*>(lkey + 8 * BCAP);                  // [8][BCAP]
    float* Qs = lval + 8 * BCAP;                            // [C]  sum_t gl/exp(logit)      (wrong-class fold)
    float* qg = Qs + C;                                     // [T]  gl/exp(logit) of the GT class
    __shared__ int wcount[8][8];                            // [producer warp][owner bucket]
    __shared__ int bcount[8];

    const int b = blockIdx.x;
    const int d0 = blockIdx.y * DC;
    const int dc = min(DC, D - d0);
    const int P = C * K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool has_gt = (gt != nullptr);
    const long long g = has_gt ? (long long)gt[b] : -1;

    for (int i = threadIdx.x; i < HW * pitch; i += 256) G[i] = 0.f;
    if (has_gt) {
        for (int c = threadIdx.x; c < C; c += 256) {
            const size_t lo = ((size_t)b * C + c) * T;
            float q = 0.f;
            for (int t = 0; t < T; ++t) q += gl[lo + t] / expf(logits[lo + t]);
            Qs[c] = q;
        }
        if (g >= 0 && g < C)
            for (int t = threadIdx.x; t < T; t += 256) {
                const size_t lo = ((size_t)b * C + (size_t)g) * T;
                qg[t] = gl[lo + t] / expf(logits[lo + t]);
            }
    }
    if (threadIdx.x < 8) bcount[threadIdx.x] = 0;
    __syncthreads();

    // entry space: with labels only level 0 of every prototype plus levels 1..T-1 of the GT class carry
    // gradient (wrong-class levels alias level 0, ref model.py:221); without labels all P*T entries
    const bool gvalid = has_gt && g >= 0 && g < C;
    const int E = has_gt ? (P + (gvalid ? K * (T - 1) : 0)) : P * T;
    for (int e0 = 0; e0 < E; e0 += 256) {
        const int e = e0 + threadIdx.x;
        float a = 0.f;
        unsigned key = 0;
        if (e < E) {
            int p, t;
            float qv;
            if (has_gt) {
                if (e < P) {
                    p = e; t = 0;
                    const int c = p / K;
                    qv = ((long long)c == g) ? qg[0] : Qs[c];
                } else {
                    const int r = e - P;
                    const int k = r / (T - 1);
                    t = 1 + (r - k * (T - 1));
                    p = (int)g * K + k;
                    qv = qg[t];
                }
            } else {
                p = e / T; t = e - p * T;
                const size_t lo = ((size_t)b * C + p / K) * T + t;
                qv = gl[lo] / expf(logits[lo]);
            }
            const int c = p / K;
            const size_t vi = ((size_t)b * P + p) * T + t;
            a = qv * __ldg(weight + (size_t)c * P + p) * vals[vi];
            key = (unsigned)p * 1024u + (unsigned)idx[vi];
        }
        // ordered compaction into the bucket of the warp that owns row n ((n & 7) == owner)
        const bool keep = (a != 0.f);
        const int owner = key & 7u;
        unsigned mybal = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const unsigned bal = __ballot_sync(0xffffffffu, keep && owner == o);
            if (lane == 0) wcount[warp][o] = __popc(bal);
            if (owner == o) mybal = bal;
        }
        __syncthreads();
        if (keep) {
            int pos = bcount[owner];
            for (int wv = 0; wv < warp; ++wv) pos += wcount[wv][owner];
            pos += __popc(mybal & ((1u << lane) - 1u));
            lkey[owner * BCAP + pos] = key;
            lval[owner * BCAP + pos] = a;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            int tot = 0;
            for (int wv = 0; wv < 8; ++wv) tot += wcount[wv][threadIdx.x];
            bcount[threadIdx.x] += tot;
        }
        __syncthreads();
        int mx = 0;
#pragma unroll
        for (int o = 0; o < 8; ++o) mx = max(mx, bcount[o]);
        const bool last = (e0 + 256 >= E);
        if (mx + 256 > BCAP || last) {
            // drain: warp `warp` accumulates its own rows -- no atomics, fixed order.  Four entries are
            // fetched together so that their (L2-latency) prototype / patch row loads overlap.
            const int cnt = bcount[warp];
            const unsigned* mk = lkey + warp * BCAP;
            const float* mv = lval + warp * BCAP;
            for (int i0 = 0; i0 < cnt; i0 += 4) {
                float2 fw[4], fm[4], fx[4];
                float aa[4];
                int nn[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = min(i0 + j, cnt - 1);
                    const unsigned kk = mk[i];
                    aa[j] = (i0 + j < cnt) ? mv[i] : 0.f;
                    const int p = kk >> 10;
                    nn[j] = kk & 1023u;
                    fw[j] = fm[j] = fx[j] = make_float2(0.f, 0.f);
                    if (2 * lane < dc) {
                        fw[j] = __ldg(reinterpret_cast<const float2*>(w + (size_t)p * D + d0) + lane);
                        fm[j] = __ldg(reinterpret_cast<const float2*>(wm + (size_t)p * D + d0) + lane);
                        fx[j] = __ldg(reinterpret_cast<const float2*>(xhat + ((size_t)b * HW + nn[j]) * D + d0) + lane);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2* gr = reinterpret_cast<float2*>(G + (size_t)nn[j] * pitch) + lane;
                    if (2 * lane < dc) {
                        float2 acc = *gr;
                        acc.x += aa[j] * (fm[j].x - fw[j].x * fx[j].x);
                        acc.y += aa[j] * (fm[j].y - fw[j].y * fx[j].y);
                        *gr = acc;
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x < 8) bcount[threadIdx.x] = 0;
            __syncthreads();
        }
    }
    __syncthreads();
    for (int n = warp; n < HW; n += 8) {
        float* dst = g_xhat + ((size_t)b * HW + n) * D + d0;
        const float* gr = G + (size_t)n * pitch;
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
#define MGP_LAUNCH_SEL(RR, NRR)                                                                                      \
    do {                                                                                                             \
        MGP_CUDA(cudaFuncSetAttribute(head_select_kernel<RR, NRR>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                      (int)smem));                                                                   \
        head_select_kernel<RR, NRR><<<grid, 256, smem, st>>>(logp_bphw, weight_cp, gt, logits, vals, idx, HW, C, K,  \
                                                             T, CT);                                                 \
    } while (0)
    if (R <= 2) MGP_LAUNCH_SEL(2, 4);
    else if (R <= 4) MGP_LAUNCH_SEL(4, 4);
    else if (R <= 7) MGP_LAUNCH_SEL(7, 4);
    else if (R <= 13) MGP_LAUNCH_SEL(13, 2);
    else if (R <= 25) MGP_LAUNCH_SEL(25, 1);
    else MGP_LAUNCH_SEL(32, 1);
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
    int DC = 64;                                             // D-chunk per CTA (one float2 per lane): 2 CTAs per SM
    while (DC > 32 && (size_t)HW * (DC + 2) * 4 > 64 * 1024) DC >>= 1;
    if (DC > D) DC = ((D + 31) / 32) * 32;
    if (DC > 64 || (D & 1)) return MGP_ERR_UNSUPPORTED;
    size_t smem = (size_t)HW * (DC + 2) * 4 + (size_t)8 * BCAP * 8 + (size_t)(C + T) * 4;
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
