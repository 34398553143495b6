// a8/a9: enqueue of the mined top-1 patches into the per-class FIFO memory bank.
// ref: model.py:225-250 (dedupe per image, class-ascending, image order), utils/memory.py:31-73
// (FIFO with eviction of the oldest rows).
//
// The reference keeps each class buffer physically ordered oldest->newest by shifting it on
// every push and runs ~B*K host-synchronising torch.unique / torch.where calls; here the bank
// is one [C, cap, D] ring (head[c], mem_len[c]) and an iteration's enqueue is two launches with
// no host synchronisation: a single-CTA planner (dedupe, per-class offsets, ring arithmetic)
// and a row scatter.  Row order inside a class does not change any EM result except through
// fp32 summation order; mgp_bank_linearize reproduces the reference layout on demand.
#include <cuda_fp16.h>

#include "mgp_common.cuh"

namespace {

// fp16 hi / lo split of 256 x (22 mantissa bits; the factor keeps the lo part in fp16's normal range for unit-norm
// rows) + |x|^2: the tensor-core EM kernel (em_tc.cu) TMA-loads these tiles instead of converting fp32 rows on chip.
__device__ __forceinline__ void shadow_store_row(const float* __restrict__ src, __half* __restrict__ xh,
                                                 __half* __restrict__ xl, float* __restrict__ xx, size_t row, int D,
                                                 int lane) {
    float ss = 0.f;
    for (int d4 = lane; d4 < D / 4; d4 += 32) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * d4);
        const float a[4] = {v.x, v.y, v.z, v.w};
        __align__(8) __half h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ss = fmaf(a[i], a[i], ss);
            const float s1 = a[i] * 256.0f;
            h[i] = __float2half_rn(s1);
            l[i] = __float2half_rn(s1 - __half2float(h[i]));
        }
        *reinterpret_cast<uint2*>(xh + row * D + 4 * d4) = *reinterpret_cast<uint2*>(h);
        *reinterpret_cast<uint2*>(xl + row * D + 4 * d4) = *reinterpret_cast<uint2*>(l);
    }
    ss = warp_sum(ss);
    if (lane == 0) xx[row] = ss;
}

__global__ void bank_shadow_kernel(const float* __restrict__ bank, __half* __restrict__ xh, __half* __restrict__ xl,
                                   float* __restrict__ xx, long long rows, int D) {
    const long long wg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wg >= rows) return;
    shadow_store_row(bank + (size_t)wg * D, xh, xl, xx, (size_t)wg, D, threadIdx.x & 31);
}

__global__ void __launch_bounds__(1024)
enqueue_plan_kernel(int64_t* __restrict__ mem_len, int32_t* __restrict__ head, uint8_t* __restrict__ updated,
                    const int32_t* __restrict__ top1_bk, const int64_t* __restrict__ gt, int32_t* __restrict__ plan,
                    int B, int C, int K, int cap, int top1_stride, int gt_stride) {
    // top1_stride / gt_stride: elements between consecutive images (K / 1 when dense; the record stride when the
    // inputs are views into the all-gathered packed records of a batch-sharded run, parallel.py)
    extern __shared__ int sm[];
    int* ucount = sm;       // [B] unique rows of image b
    int* cls = sm + B;      // [B] class or -1
    int* wr_m = sm + 2 * B; // [B] rows accepted for the class if b is the class's first image, else -1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (K <= 32) {
        // warp per image, lane per prototype: duplicates by MATCH.ANY, rank among the distinct values by K shuffles
        for (int b = warp; b < B; b += nwarp) {
            const long long c = gt[(size_t)b * gt_stride];
            const bool okc = (c >= 0 && c < C);
            const int v = (okc && lane < K) ? top1_bk[(size_t)b * top1_stride + lane] : (-1 - lane);
            const unsigned m = __match_any_sync(0xffffffffu, v);
            const bool first = okc && lane < K && (__ffs(m) - 1 == lane);
            const unsigned fm = __ballot_sync(0xffffffffu, first);
            int rank = 0;                                     // ascending position among the distinct values (torch.unique order)
            for (int j = 0; j < K; ++j) {
                const int vj = __shfl_sync(0xffffffffu, v, j);
                rank += (((fm >> j) & 1u) && vj < v) ? 1 : 0;
            }
            if (lane < K) plan[b * K + lane] = first ? rank : -1;
            if (lane == 0) { cls[b] = okc ? (int)c : -1; ucount[b] = __popc(fm); }
        }
    } else {
        for (int b = threadIdx.x; b < B; b += blockDim.x) {
            const long long c = gt[(size_t)b * gt_stride];
            if (c < 0 || c >= C) {
                cls[b] = -1;
                ucount[b] = 0;
                for (int k = 0; k < K; ++k) plan[b * K + k] = -1;
                continue;
            }
            cls[b] = (int)c;
            const int32_t* top1 = top1_bk + (size_t)b * top1_stride;
            int u = 0;
            for (int k = 0; k < K; ++k) {
                const int v = top1[k];
                bool first = true;
                for (int k2 = 0; k2 < k; ++k2) first = first && (top1[k2] != v);
                int rank = -1;
                if (first) {
                    rank = 0;
                    for (int k2 = 0; k2 < K; ++k2) {
                        const int v2 = top1[k2];
                        if (v2 < v) {
                            bool f2 = true;
                            for (int k3 = 0; k3 < k2; ++k3) f2 = f2 && (top1[k3] != v2);
                            rank += f2 ? 1 : 0;
                        }
                    }
                    ++u;
                }
                plan[b * K + k] = rank;
            }
            ucount[b] = u;
        }
    }
    __syncthreads();
    // Offsets inside a class follow image order.  Two equivalent schedules, O(min(C, B) * B / 32) warp steps:
    //   class-major (large batches, e.g. the all-gathered global batch): a warp scans the batch for its class and
    //                hands out running offsets (prefix scan), ~45 instructions per (class, 32 images);
    //   image-major: a warp sums, for its image, the rows of the same class in earlier images, ~10 per (image, 32 images).
    // offs[b] = rows of cls[b] accepted before image b; ctot[b] = rows of cls[b] in the whole batch, stored as
    // -(total) - 1 for the first image of its class (that image publishes the new mem_len / head).
    int* offs = wr_m;            // [B]
    int* ctot = sm + 3 * B;      // [B]
    if (9 * C <= 2 * B) {
        for (int c = warp; c < C; c += nwarp) {
            int off = 0, firstb = -1;
            for (int b0 = 0; b0 < B; b0 += 32) {
                const int b = b0 + lane;
                const bool mine = (b < B) && (cls[b] == c);
                const unsigned mm = __ballot_sync(0xffffffffu, mine);
                if (firstb < 0 && mm) firstb = b0 + __ffs(mm) - 1;
                int u = mine ? ucount[b] : 0;
                int incl = u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                if (mine) offs[b] = off + incl - u;
                off += __shfl_sync(0xffffffffu, incl, 31);
            }
            for (int b0 = 0; b0 < B; b0 += 32) {
                const int b = b0 + lane;
                if (b < B && cls[b] == c) ctot[b] = (b == firstb) ? -off - 1 : off;
            }
        }
    } else {
        for (int b = warp; b < B; b += nwarp) {
            const int c = cls[b];
            if (c < 0) continue;
            int off = 0, tot = 0, earlier = 0;
            for (int b2 = lane; b2 < B; b2 += 32) {
                if (cls[b2] == c) {
                    const int u = ucount[b2];
                    tot += u;
                    if (b2 < b) { off += u; earlier = 1; }
                }
            }
            off = __reduce_add_sync(0xffffffffu, off);
            tot = __reduce_add_sync(0xffffffffu, tot);
            earlier = __reduce_add_sync(0xffffffffu, earlier);
            if (lane == 0) { offs[b] = off; ctot[b] = earlier ? tot : -tot - 1; }
        }
    }
    __syncthreads();
    for (int b = warp; b < B; b += nwarp) {
        const int c = cls[b];
        if (c < 0) continue;
        const int off = offs[b];
        const int ct = ctot[b];
        const int m = min(ct < 0 ? -ct - 1 : ct, cap);
        const int len = (int)mem_len[c];
        const int hd = head[c];
        for (int k = lane; k < K; k += 32) {
            const int r = plan[b * K + k];
            int slot = -1;
            if (r >= 0 && off + r < m) slot = (hd + len + off + r) % cap;
            plan[b * K + k] = slot;
        }
    }
    __syncthreads();       // every reader of mem_len / head has seen the old values
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int c = cls[b];
        if (c < 0) continue;
        const int ct = ctot[b];
        if (ct >= 0) continue;                                             // not the first image of its class
        const int m = min(-ct - 1, cap);
        if (m <= 0) continue;
        const int len = (int)mem_len[c];
        const int hd = head[c];
        if (len + m <= cap) {
            mem_len[c] = len + m;
        } else {
            head[c] = (hd + (len + m - cap)) % cap;
            mem_len[c] = cap;
        }
        updated[c] = 1;                                                     // ref model.py:250
    }
}

// top-1 patch of each of the GT class's K prototypes: spatial index and feature row
__global__ void mined_gather_kernel(const float* __restrict__ xhat, const int32_t* __restrict__ idx,
                                    const int64_t* __restrict__ gt, int32_t* __restrict__ top1,
                                    float* __restrict__ rows, int B, int HW, int C, int K, int D, int T, int rows_stride,
                                    int top1_stride) {
    const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wg >= B * K) return;
    const int b = wg / K, k = wg - b * K;
    const long long c = gt[b];
    float4* dst = reinterpret_cast<float4*>(rows + (size_t)b * rows_stride + (size_t)k * D);
    if (c < 0 || c >= C) {
        if (lane == 0) top1[(size_t)b * top1_stride + k] = -1;
        for (int d = lane; d < D / 4; d += 32) dst[d] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int n = idx[((size_t)b * C * K + (size_t)c * K + k) * T];     // level 0 (ref model.py:225-226)
    if (lane == 0) top1[(size_t)b * top1_stride + k] = n;
    const float4* src = reinterpret_cast<const float4*>(xhat + ((size_t)b * HW + n) * D);
    for (int d = lane; d < D / 4; d += 32) dst[d] = src[d];
}

__global__ void enqueue_scatter_kernel(float* __restrict__ bank, const float* __restrict__ rows,
                                       const int64_t* __restrict__ gt, const int32_t* __restrict__ plan,
                                       __half* __restrict__ xh, __half* __restrict__ xl, float* __restrict__ xx, int B,
                                       int K, int D, int cap, int rows_stride, int gt_stride) {
    const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wg >= B * K) return;
    const int slot = plan[wg];
    if (slot < 0) return;
    const int sb = wg / K, sk = wg - sb * K;
    const long long c = gt[(size_t)sb * gt_stride];
    const float* srow = rows + (size_t)sb * rows_stride + (size_t)sk * D;
    const float4* src = reinterpret_cast<const float4*>(srow);
    float4* dst = reinterpret_cast<float4*>(bank + ((size_t)c * cap + slot) * D);
    for (int d = lane; d < D / 4; d += 32) dst[d] = src[d];
    if (xh) shadow_store_row(srow, xh, xl, xx, (size_t)c * cap + slot, D, lane);   // keep the shadow in step
}

__global__ void bank_linearize_kernel(const float* __restrict__ bank, const int64_t* __restrict__ mem_len,
                                      const int32_t* __restrict__ head, float* __restrict__ lin, int C, int cap,
                                      int D) {
    const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wg >= C * cap) return;
    const int c = wg / cap, r = wg - c * cap;
    float* dst = lin + (size_t)wg * D;
    if (r < (int)mem_len[c]) {
        const float* src = bank + ((size_t)c * cap + (head[c] + r) % cap) * D;
        for (int d = lane; d < D; d += 32) dst[d] = src[d];
    } else {
        for (int d = lane; d < D; d += 32) dst[d] = 0.f;
    }
}

}  // namespace

extern "C" int mgp_mined_gather(const float* xhat_nd, const int32_t* idx, const int64_t* gt, int32_t* top1,
                               float* rows, int rows_stride, int top1_stride, int B, int HW, int C, int K, int D, int T,
                               void* stream) {
    if (!xhat_nd || !idx || !gt || !top1 || !rows) return MGP_ERR_INVALID;
    if (B <= 0 || HW <= 0 || C <= 0 || K <= 0 || D <= 0 || T <= 0 || (D & 3)) return MGP_ERR_INVALID;
    if (rows_stride == 0) rows_stride = K * D;
    if (top1_stride == 0) top1_stride = K;
    if (rows_stride < K * D || (rows_stride & 3) || top1_stride < K || !mgp_aligned16(rows)) return MGP_ERR_INVALID;
    const int warps = B * K;
    mined_gather_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(xhat_nd, idx, gt, top1, rows, B, HW, C, K, D, T,
                                                                          rows_stride, top1_stride);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}

extern "C" int mgp_bank_enqueue(float* bank, int64_t* mem_len, int32_t* head, uint8_t* updated, const float* rows,
                                const int32_t* top1, const int64_t* gt, int rows_stride, int top1_stride, int gt_stride,
                                int32_t* plan, void* shadow_h, void* shadow_l, float* shadow_xx, int B, int C, int K, int D,
                                int cap, void* stream) {
    if (!bank || !mem_len || !head || !updated || !rows || !top1 || !gt || !plan) return MGP_ERR_INVALID;
    if ((shadow_h != nullptr) != (shadow_l != nullptr) || (shadow_h != nullptr) != (shadow_xx != nullptr)) return MGP_ERR_INVALID;
    if (B <= 0 || C <= 0 || K <= 0 || D <= 0 || cap <= 0 || (D & 3)) return MGP_ERR_INVALID;
    if (B > 8192 || K > 64) return MGP_ERR_UNSUPPORTED;
    if (rows_stride == 0) rows_stride = K * D;
    if (top1_stride == 0) top1_stride = K;
    if (gt_stride == 0) gt_stride = 1;
    if (rows_stride < K * D || (rows_stride & 3) || top1_stride < K || gt_stride < 1 || !mgp_aligned16(rows)) return MGP_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    size_t smem = (size_t)4 * B * sizeof(int);
    MGP_CUDA(cudaFuncSetAttribute(enqueue_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enqueue_plan_kernel<<<1, 1024, smem, st>>>(mem_len, head, updated, top1, gt, plan, B, C, K, cap, top1_stride, gt_stride);
    MGP_CHECK_LAUNCH();
    const int warps = B * K;
    enqueue_scatter_kernel<<<(warps + 7) / 8, 256, 0, st>>>(bank, rows, gt, plan, reinterpret_cast<__half*>(shadow_h),
                                                            reinterpret_cast<__half*>(shadow_l), shadow_xx, B, K, D, cap,
                                                            rows_stride, gt_stride);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}

extern "C" int mgp_bank_shadow_sync(const float* bank, void* shadow_h, void* shadow_l, float* shadow_xx, int C, int cap,
                                    int D, void* stream) {
    if (!bank || !shadow_h || !shadow_l || !shadow_xx || C <= 0 || cap <= 0 || D <= 0 || (D & 3)) return MGP_ERR_INVALID;
    const long long rows = (long long)C * cap;
    bank_shadow_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        bank, reinterpret_cast<__half*>(shadow_h), reinterpret_cast<__half*>(shadow_l), shadow_xx, rows, D);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}

extern "C" int mgp_bank_linearize(const float* bank, const int64_t* mem_len, const int32_t* head, float* lin, int C,
                                  int cap, int D, void* stream) {
    if (!bank || !mem_len || !head || !lin || C <= 0 || cap <= 0 || D <= 0) return MGP_ERR_INVALID;
    const int warps = C * cap;
    bank_linearize_kernel<<<(warps + 7) / 8, 256, 0, (cudaStream_t)stream>>>(bank, mem_len, head, lin, C, cap, D);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}
