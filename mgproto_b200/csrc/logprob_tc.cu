// a2/a16 on the 5th-gen tensor cores (MGP_MATH_TC): the squared Mahalanobis distance as a GEMM
//
//   q[n,p] = sum_d w_pd x_nd^2 - 2 sum_d (w mu)_pd x_nd + sum_d w_pd mu_pd^2 ,   w = 1/(sigma+eps)^2
//          = [x^2 | x]_n . [w | -2 w mu]_p + c2_p                       (inner dim 2D, general diagonal)
//          = w_p |x_n|^2 + x_n . (-2 w mu)_p + c2_p                     (inner dim  D, sigma constant over d:
//                                                                        every state the shipped loop reaches)
//   log p = e0_p + e1_p * acc[n,p] + e2_p * |x_n|^2
//
// Precision: operands are split into fp16 hi + lo (22 mantissa bits) and accumulated as
// hi*hi + lo*hi + hi*lo in fp32 TMEM accumulators -- three kind::f16 passes instead of one TF32
// pass at half rate; |error| on q ~1e-6, inside the 1e-4 bar on logits (a single bf16 or tf32
// pass is not).  Power-of-two scalings keep the lo parts in fp16's normal range and are undone
// exactly in the epilogue.
//
// Structure (one persistent CTA per SM, 12 warps):
//   warp 0   TMA producer: prototype (A) K-blocks through an S-stage mbarrier ring
//   warp 3   TMA producer of the x tile (B, 128 patches x Kg), resident per n-tile, double-buffered when
//            sigma is isotropic so the next n-tile is prefetched under the current one's MMAs
//   warp 1   one thread issues tcgen05.mma (M=128 prototypes x N=128 patches x K=16), 2 TMEM accumulators
//   warp 2   TMEM allocator
//   warps 4-11 epilogue: tcgen05.ld 32 lanes x 32 columns, affine fix-up, stores.  TMEM lane = prototype,
//            column = patch, so for the [N,P] layout the 32 lanes of a warp write 32 consecutive floats
//            of one output row -- fully coalesced straight from registers, no staging pass.
// HBM traffic per launch: 4*N*P (output) + 8*N*Kg (fp16 hi/lo operand written by the prep pass and
// read once) + 4*N*D (x) -- the output dominates; the kernel is bound by the HBM write stream.
#include <cuda.h>
#include <cstdlib>
#include <cuda_fp16.h>

#include "mgp_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int LAYOUT_NP_TMA = 3; // internal: [N,P] output written by TMA bulk tensor stores from a shared-memory stage
constexpr int LAYOUT_BPHW_TMA = 4;   // internal: [B,P,HW] log p through a 3-D tensor map (boxes clipped at image ends)
constexpr int LAYOUT_NEGP_TMA = 5;   // internal: [B,P,HW] -exp(log p), same
constexpr int LAYOUT_TOP1 = 6;       // internal: no log p output at all -- per (image, prototype) max / arg-max (MGP_OUT_TOP1_BP)
constexpr int STAGING_BYTES = 8 * 32 * 32 * 4;   // one [32 patches x 32 prototypes] fp32 block per epilogue warp
constexpr int PT = 128;          // prototypes per tile (UMMA M)
constexpr int KB = 64;           // K elements per smem block (128 B rows, SWIZZLE_128B)
constexpr int SUB_BYTES = 128 * KB * 2;   // one [128 x 64] fp16 block = 16 KiB
constexpr float X_SCALE = 256.0f;

// ------------------------------------------------------------------------------------------ PTX (tc_ptx.cuh)
using namespace mgp_tc;

// kind::f16 instruction descriptor of the [PT prototypes x n_tile patches] tile, both operands K-major
__device__ __forceinline__ uint32_t make_idesc(int n_tile) { return umma_idesc_f16(PT, n_tile); }

// ------------------------------------------------------------------------------------------ prep
// Prototype side: Bh/Bl [P, 2D] fp16 = split of scale_p * [ w | -2 w mu ]; e0,e1,e2 [P]; noniso flag.
__global__ void tc_proto_prep_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, float eps,
                                     float eps_log, __half* __restrict__ bh, __half* __restrict__ bl,
                                     float* __restrict__ e0, float* __restrict__ e1, float* __restrict__ e2,
                                     int* __restrict__ noniso, int P, int D) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a dependent launched programmatically may start its prologue
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    const float* mr = mu + (size_t)p * D;
    const float* sr = sigma + (size_t)p * D;
    const float s0 = sr[0];
    float ls = 0.f, c2 = 0.f, mx = 0.f;
    bool same = true;
    for (int d = lane; d < D; d += 32) {
        const float s = sr[d];
        same = same && (s == s0);
        const float r = 1.0f / (s + eps);
        const float w = r * r;
        const float m = mr[d];
        ls += logf(s + eps_log) + 0.5f * MGP_LOG_2PI;   // per-dim terms (they cancel for sigma = 1/sqrt(2 pi))
        c2 = fmaf(w * m, m, c2);
        mx = fmaxf(mx, fmaxf(w, fabsf(2.0f * w * m)));
    }
    ls = warp_sum(ls);
    c2 = warp_sum(c2);
    mx = warp_max(mx);
    same = __all_sync(0xffffffffu, same);
    int ex = 0;
    if (mx > 0.f) frexpf(mx, &ex);                       // mx = f * 2^ex, f in [0.5, 1)
    const float scale = ldexpf(1.0f, 8 - ex);            // max |B'| * scale in [128, 256)
    for (int d = lane; d < D; d += 32) {
        const float r = 1.0f / (sr[d] + eps);
        const float w = r * r;
        const float v0 = w * scale, v1 = -2.0f * w * mr[d] * scale;
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        bh[(size_t)p * 2 * D + d] = h0;
        bl[(size_t)p * 2 * D + d] = __float2half_rn(v0 - __half2float(h0));
        bh[(size_t)p * 2 * D + D + d] = h1;
        bl[(size_t)p * 2 * D + D + d] = __float2half_rn(v1 - __half2float(h1));
    }
    if (lane == 0) {
        const float r0 = 1.0f / (s0 + eps);
        e0[p] = -ls - 0.5f * c2;
        e1[p] = -0.5f / (scale * X_SCALE);
        e2[p] = -0.5f * r0 * r0;                          // used only when every prototype is isotropic
        if (!same) atomicOr(noniso, 1);
    }
}

// Patch side: Ah/Al [N, 2D] fp16 = split of 256 * [ x^2 | x ] (the x^2 half only if some prototype is
// anisotropic), sn [N] = |x|^2.  Warp per row.
__global__ void tc_x_prep_kernel(const float* __restrict__ x, __half* __restrict__ ah, __half* __restrict__ al,
                                 float* __restrict__ sn, const int* __restrict__ noniso, int N, int D) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const bool gen = (*noniso != 0);
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)n * D);
    __half* hr = ah + (size_t)n * 2 * D;
    __half* lr = al + (size_t)n * 2 * D;
    float ss = 0.f;
    for (int d4 = lane; d4 < D / 4; d4 += 32) {
        const float4 v = __ldg(xr + d4);
        const float a[4] = {v.x, v.y, v.z, v.w};
        __align__(8) __half h[4], l[4], h2[4], l2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ss = fmaf(a[i], a[i], ss);
            const float s1 = a[i] * X_SCALE;
            h[i] = __float2half_rn(s1);
            l[i] = __float2half_rn(s1 - __half2float(h[i]));
            const float s2 = a[i] * a[i] * X_SCALE;
            h2[i] = __float2half_rn(s2);
            l2[i] = __float2half_rn(s2 - __half2float(h2[i]));
        }
        *reinterpret_cast<uint2*>(hr + D + d4 * 4) = *reinterpret_cast<uint2*>(h);
        *reinterpret_cast<uint2*>(lr + D + d4 * 4) = *reinterpret_cast<uint2*>(l);
        if (gen) {
            *reinterpret_cast<uint2*>(hr + d4 * 4) = *reinterpret_cast<uint2*>(h2);
            *reinterpret_cast<uint2*>(lr + d4 * 4) = *reinterpret_cast<uint2*>(l2);
        }
    }
    ss = warp_sum(ss);
    if (lane == 0) sn[n] = ss;
}

// ------------------------------------------------------------------------------------------ main
struct TcParams {
    const float* e0;
    const float* e1;
    const float* e2;
    const float* sn;
    const int* noniso;
    float* out;
    int N, HW, P, D;
    int n_ntiles, n_ptiles;
    int xbox;              // rows of the x tensor-map box
    int B, nti;            // image-aligned tiling of the [B,P,HW] TMA path: one x tile = one image, nti = round_up(HW, 32) columns
    int team;              // CTAs per team: the CTAs of a team work on the SAME x tile at the same time, on
                           // adjacent prototype tiles, so each output row receives team*512 contiguous bytes at once
    uint32_t smem_bytes;   // dynamic shared memory of the launch
    int x_no_sq;           // the staged patch operands lack the x^2 half (isotropic sigma asserted by the producer)
    int debug;   // ablation switches for profiling (MGP_TC_DEBUG): 1 no global stores, 2 no TMEM loads, 4 no MMAs,
                 // 8 no epilogue work, 16 no prototype TMA loads
};

constexpr int TC_THREADS = 384;   // warps: 0 proto TMA, 1 MMA, 2 TMEM alloc, 3 x-tile TMA, 4..11 epilogue

template <int LAYOUT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], const float* __restrict__ s_sn_c, float c0,
                                               float c1, float c2, int n0, int p, bool pok, const TcParams& prm,
                                               float* stg = nullptr, const CUtensorMap* map_out = nullptr, int img_b = 0,
                                               int img_hw0 = 0, bool img = false, float* run_v = nullptr,
                                               int* run_i = nullptr) {
    const int N = prm.N, P = prm.P, HW = prm.HW;
    float v[32];
    const float4* s4 = reinterpret_cast<const float4*>(s_sn_c);    // |x_n|^2 of the 32 columns (shared memory)
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 s = s4[j4];
        // two columns per FFMA2 (same rounding as two scalar fmaf: each half is an IEEE fused multiply-add)
        const float2 c00 = make_float2(c0, c0), c11 = make_float2(c1, c1), c22 = make_float2(c2, c2);
        const float2 v01 = ffma2(c11, make_float2(__uint_as_float(r[4 * j4 + 0]), __uint_as_float(r[4 * j4 + 1])),
                                 ffma2(c22, make_float2(s.x, s.y), c00));
        const float2 v23 = ffma2(c11, make_float2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])),
                                 ffma2(c22, make_float2(s.z, s.w), c00));
        v[4 * j4 + 0] = v01.x; v[4 * j4 + 1] = v01.y; v[4 * j4 + 2] = v23.x; v[4 * j4 + 3] = v23.y;
    }
    if (LAYOUT == LAYOUT_TOP1) {
        // only max_n log p[n, p] and its patch per image are wanted (labelled training step: the reference aliases
        // the other levels of wrong-class prototypes to level 0): nothing is stored, the log-likelihood matrix
        // never reaches HBM.  Image tiles keep a running (max, patch) across the warp's chunks (the caller issues
        // one 64-bit RED.MAX per tile); 128-patch tiles may cross image ends and reduce per image segment.
        if (img) {
            // chunk maximum by a tree of FMNMX (one instruction per column); the position is only looked up when the
            // chunk beats the running maximum (a few times per image): same result as the strict left-to-right scan
            // (first patch wins ties) at ~2 instead of ~5 instructions per column
            if (img_hw0 + 32 > HW) {                              // the image's last chunk: columns beyond HW do not exist
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (img_hw0 + j >= HW) v[j] = -INFINITY;
            }
            float m[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) m[j] = fmaxf(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                for (int j = 0; j < w; ++j) m[j] = fmaxf(m[j], m[j + w]);
            if (m[0] > *run_v) {
                int idx = 31;
#pragma unroll
                for (int j = 30; j >= 0; --j) idx = (v[j] == m[0]) ? j : idx;
                *run_v = m[0];
                *run_i = img_hw0 + idx;
            }
            return;
        }
        if (!pok) return;
        unsigned long long* best = reinterpret_cast<unsigned long long*>(prm.out);
        int b = n0 / HW, hw = n0 - b * HW;
        float mv = -INFINITY;
        int mi = -1;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (n0 + j < N) {
                if (mi < 0 || v[j] > mv) { mv = v[j]; mi = hw; }
                if (++hw == HW) {
                    atomicMax(best + (size_t)b * P + p, top1_pack(mv, mi));
                    ++b; hw = 0; mi = -1;
                }
            }
        }
        if (mi >= 0) atomicMax(best + (size_t)b * P + p, top1_pack(mv, mi));
        return;
    }
    if (LAYOUT == LAYOUT_NP_TMA) {
        // stage the [32 patches x 32 prototypes] block in shared memory (row = patch, 128 B) and hand it to
        // the TMA engine: the global writes are issued asynchronously as whole row segments, the warp only
        // pays 32 conflict-free STS.  Out-of-range rows / columns are clipped by the tensor map.
        const int lane = threadIdx.x & 31;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous block has been read
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[j * 32 + lane] = v[j];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !(prm.debug & 1)) {
            tma_store_2d(map_out, smem_u32(stg), p - lane, n0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        return;
    }
    if (LAYOUT == LAYOUT_BPHW_TMA || LAYOUT == LAYOUT_NEGP_TMA) {
        // [B,P,HW]: this thread's 32 values are 128 contiguous bytes of row (b, p).  Stage [32 prototypes x 128 B]
        // with the tensor map's 128B swizzle (8 x STS.128 per thread, conflict-free) and let the TMA engine write
        // it.  Tiles are image-aligned here, so a chunk lies in ONE image; columns past HW are clipped by the map.
        const int lane = threadIdx.x & 31;
        if (LAYOUT == LAYOUT_NEGP_TMA) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = -expf(v[j]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        uint8_t* rowp = reinterpret_cast<uint8_t*>(stg) + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(rowp + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && !(prm.debug & 1)) {
            tma_store_3d(map_out, smem_u32(stg), img_hw0, p - lane, img_b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        return;
    }
    if (!pok) return;
    if (prm.debug & 1) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc += v[j];
        if (acc == 123.456f) prm.out[0] = acc;
        return;
    }
    const bool full = (n0 + 32 <= N);
    if (LAYOUT == MGP_OUT_LOGP_NP) {
        float* dst = prm.out + (size_t)n0 * P + p;          // lanes = consecutive p: 128 B per warp store
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[(size_t)j * P] = v[j];
        } else {
            for (int j = 0; j < 32; ++j)
                if (n0 + j < N) dst[(size_t)j * P] = v[j];
        }
    } else {
        if (LAYOUT == MGP_OUT_NEGP_BPHW) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = -expf(v[j]);
        }
        int b = n0 / HW, hw = n0 - b * HW;
        if (full && (HW & 3) == 0) {                         // 4 consecutive patches share an image, 16 B aligned
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                *reinterpret_cast<float4*>(prm.out + ((size_t)b * P + p) * HW + hw) =
                    make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                hw += 4;
                if (hw >= HW) { hw -= HW; ++b; }
            }
        } else {
            for (int j = 0; j < 32; ++j) {
                if (n0 + j < N) prm.out[((size_t)b * P + p) * HW + hw] = v[j];
                if (++hw == HW) { hw = 0; ++b; }
            }
        }
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
logprob_tc_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                  const __grid_constant__ CUtensorMap map_ph, const __grid_constant__ CUtensorMap map_pl,
                  const __grid_constant__ CUtensorMap map_out, const TcParams prm) {
    constexpr bool TMA_ST = (LAYOUT == LAYOUT_NP_TMA || LAYOUT == LAYOUT_BPHW_TMA || LAYOUT == LAYOUT_NEGP_TMA);
    constexpr bool BPHW_TMA = (LAYOUT == LAYOUT_BPHW_TMA || LAYOUT == LAYOUT_NEGP_TMA);
    // the layout used by the non-TMA fallback of the same instantiation (anisotropic sigma: see `img` below)
    constexpr int STG_LAYOUT = (LAYOUT == LAYOUT_BPHW_TMA) ? MGP_OUT_LOGP_BPHW
                               : (LAYOUT == LAYOUT_NEGP_TMA) ? MGP_OUT_NEGP_BPHW : LAYOUT;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* base_ptr = smem_raw + (base - raw);

    const bool gen = (*prm.noniso != 0);
    if (gen && (prm.D > 128 || prm.x_no_sq)) __trap();            // isotropic sigma was promised (MGP_MATH_TC_ISO / staging): fail loudly
    const int nkb = (gen ? 2 * prm.D : prm.D) / KB;               // K blocks per tile
    const int kcol0 = gen ? 0 : prm.D;                            // isotropic: only the [x] / [-2 w mu] half
    // [B,P,HW] through TMA: one x tile = one image (nti = round_up(HW,32) columns, UMMA N = nti) so that no
    // 32-column chunk crosses an image end.  The wider tile only fits when sigma is isotropic (K = D);
    // otherwise this instantiation falls back to 128-patch tiles and register stores.
    const bool img = (BPHW_TMA || (LAYOUT == LAYOUT_TOP1 && prm.xbox == 32)) && !gen;
    const int NT = img ? prm.nti : 128;                           // patches per tile = UMMA N
    const int row_step = img ? prm.HW : 128;                      // first patch row of x tile nt = nt * row_step
    // image tiles: the x tile holds nti = round_up(HW, 32) rows (whole 32-row TMA boxes), the MMA only spans
    // round_up(HW, 16) of them (HW = 196: N = 208 instead of 224); the columns beyond are never read (masked / clipped)
    const uint32_t idesc = make_idesc(img ? ((prm.HW + 15) & ~15) : NT);
    const int n_ptiles = prm.n_ptiles, n_ntiles = img ? prm.B : prm.n_ntiles;
    const uint32_t xsub = (uint32_t)NT * KB * 2;                  // one [NT x 64] fp16 block of the x tile

    // carve-up: nbuf x tiles | S stages of (proto hi, proto lo) | [TMA-store staging] | barriers + sn tile
    const uint32_t x_bytes = (uint32_t)(2 * nkb) * xsub;          // hi blocks then lo blocks
    const uint32_t tile_budget = prm.smem_bytes - 1024u - 2048u - (TMA_ST ? (uint32_t)STAGING_BYTES : 0u);
    const int nbuf = (2 * x_bytes + 2 * 2 * SUB_BYTES <= tile_budget) ? 2 : 1;   // double-buffer x when it fits
    int S = (int)((tile_budget - nbuf * x_bytes) / (2 * SUB_BYTES));
    if (S > 6) S = 6;
    const uint32_t x_base = base;
    const uint32_t st_base = x_base + nbuf * x_bytes;
    const uint32_t stg_base = st_base + (uint32_t)S * 2 * SUB_BYTES;
    const uint32_t misc = stg_base + (TMA_ST ? (uint32_t)STAGING_BYTES : 0u);
    float* staging = reinterpret_cast<float*>(base_ptr + (stg_base - base));
    uint8_t* misc_ptr = base_ptr + (misc - base);
    uint64_t* bars = reinterpret_cast<uint64_t*>(misc_ptr);       // full[6] empty[6] xfull[2] xempty[2] tfull[2] tempty[2]
    const uint32_t bar0 = misc;
    auto FULL = [&](int i) { return bar0 + 8u * i; };
    auto EMPTY = [&](int i) { return bar0 + 8u * (6 + i); };
    auto XFULL = [&](int i) { return bar0 + 8u * (12 + i); };
    auto XEMPTY = [&](int i) { return bar0 + 8u * (14 + i); };
    auto TFULL = [&](int i) { return bar0 + 8u * (16 + i); };
    auto TEMPTY = [&](int i) { return bar0 + 8u * (18 + i); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* s_sn = reinterpret_cast<float*>(bars + 22);            // [256] |x|^2 of the current x tile, 16 B aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(XFULL(i), 1);
            mbar_init(XEMPTY(i), 1);
            mbar_init(TFULL(i), 1);
            mbar_init(TEMPTY(i), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // schedule: team t = blockIdx / team handles x tiles t, t + n_teams, ...; member k of the team takes the
    // prototype tiles k, k + team, ... of each of them
    const int TS = prm.team;
    const int n_teams = gridDim.x / TS, team = blockIdx.x / TS, k0 = blockIdx.x % TS;
    const bool has_work = (team < n_teams) && (k0 < n_ptiles);

    if (!has_work) {
        // nothing to do for this CTA (tiny problems)
    } else if (warp == 3 && lane == 0) {
        // =========================== x-tile TMA producer (next x tile prefetched when double-buffered) ===========
        int c = 0;
        for (int nt = team; nt < n_ntiles; nt += n_teams, ++c) {
            const int buf = c % nbuf;
            if (c >= nbuf) mbar_wait(XEMPTY(buf), (uint32_t)((c / nbuf - 1) & 1));
            mbar_expect_tx(XFULL(buf), x_bytes);
            const uint32_t xb = x_base + (uint32_t)buf * x_bytes;
            for (int kb = 0; kb < nkb; ++kb)
                for (int r = 0; r < NT; r += prm.xbox) {          // x-map box rows: 128 (128-patch tiles) or 32 (image tiles)
                    const uint32_t ro = (uint32_t)r * KB * 2;
                    tma_load_2d(xb + (uint32_t)kb * xsub + ro, &map_xh, kcol0 + kb * KB, nt * row_step + r, XFULL(buf));
                    tma_load_2d(xb + (uint32_t)(nkb + kb) * xsub + ro, &map_xl, kcol0 + kb * KB, nt * row_step + r, XFULL(buf));
                }
        }
    } else if (warp == 0 && lane == 0) {
        // =========================== prototype TMA producer ===========================
        int stage = 0;
        uint32_t phase = 0;
        for (int nt = team; nt < n_ntiles; nt += n_teams) {
            for (int pt = k0; pt < n_ptiles; pt += TS) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(EMPTY(stage), phase ^ 1u);
                    if (prm.debug & 16) {
                        mbar_arrive(FULL(stage));
                    } else {
                        mbar_expect_tx(FULL(stage), 2 * SUB_BYTES);
                        const uint32_t dst = st_base + (uint32_t)stage * 2 * SUB_BYTES;
                        tma_load_2d(dst, &map_ph, kcol0 + kb * KB, pt * PT, FULL(stage));
                        tma_load_2d(dst + SUB_BYTES, &map_pl, kcol0 + kb * KB, pt * PT, FULL(stage));
                    }
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // =========================== MMA issuer ===========================
        int stage = 0, acc = 0, c = 0;
        uint32_t phase = 0, acc_par = 0;
        for (int nt = team; nt < n_ntiles; nt += n_teams, ++c) {
            const int buf = c % nbuf;
            mbar_wait(XFULL(buf), (uint32_t)((c / nbuf) & 1));
            const uint32_t xb = x_base + (uint32_t)buf * x_bytes;
            for (int pt = k0; pt < n_ptiles; pt += TS) {
                mbar_wait(TEMPTY(acc), acc_par ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NT);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(FULL(stage), phase);
                    tc_fence_after();
                    const uint32_t ph = st_base + (uint32_t)stage * 2 * SUB_BYTES, pl = ph + SUB_BYTES;
                    const uint32_t xh = xb + (uint32_t)kb * xsub, xl = xb + (uint32_t)(nkb + kb) * xsub;
#pragma unroll
                    for (int k = 0; k < KB / 16; ++k) {
                        if (prm.debug & 4) continue;
                        const uint32_t off = (uint32_t)k * 32u;   // 16 fp16 = 32 B inside the 128 B swizzle row
                        const uint64_t a_h = umma_desc(ph + off), a_l = umma_desc(pl + off);
                        const uint64_t b_h = umma_desc(xh + off), b_l = umma_desc(xl + off);
                        tc_mma_f16(d_tmem, a_h, b_h, idesc, (kb | k) != 0);
                        tc_mma_f16(d_tmem, a_l, b_h, idesc, 1u);
                        tc_mma_f16(d_tmem, a_h, b_l, idesc, 1u);
                    }
                    tc_commit(EMPTY(stage));                      // frees the stage when these MMAs retire
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
                tc_commit(TFULL(acc));                            // accumulator ready for the epilogue
                acc ^= 1;
                if (acc == 0) acc_par ^= 1u;
            }
            tc_commit(XEMPTY(buf));                               // x buffer may be refilled
        }
    } else if (warp >= 4) {
        // =========================== epilogue ===========================
        // 8 warps drain each accumulator together: warp = (TMEM lane quarter q, column half h)
        const int e = warp - 4;
        const int q = e & 3, h = e >> 2;
        const int et = q * 32 + lane;                             // prototype row within the tile (TMEM lane)
        const int nch_all = NT / 32;                              // 4 chunks (NT = 128) or up to 8 (image tiles)
        const int ch0 = h ? (nch_all + 1) / 2 : 0, ch1 = h ? nch_all : (nch_all + 1) / 2;   // this warp's chunks
        float* stg = staging + e * 1024;                          // this warp's 4 KiB TMA-store block
        int acc = 0, c = 0;
        uint32_t acc_par = 0;
        const bool skip_epi = (prm.debug & 8) != 0;
        for (int nt = team; nt < n_ntiles; nt += n_teams, ++c) {
            const int row0 = nt * row_step;
            asm volatile("bar.sync 1, 256;" ::: "memory");        // readers of the previous x tile's norms are done
            for (int i = e * 32 + lane; i < NT; i += 256) {
                const int n = row0 + i;
                s_sn[i] = (n < prm.N) ? prm.sn[n] : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int pt = k0; pt < n_ptiles; pt += TS) {
                const int p = pt * PT + et;
                const bool pok = p < prm.P;
                const float c0 = pok ? __ldg(prm.e0 + p) : 0.f;
                const float c1 = pok ? __ldg(prm.e1 + p) : 0.f;
                const float c2 = (pok && !gen) ? __ldg(prm.e2 + p) : 0.f;
                mbar_wait(TFULL(acc), acc_par);
                tc_fence_after();
                if (!skip_epi) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT);
                    float run_v = -INFINITY;                      // LAYOUT_TOP1, image tiles: best of this warp's chunks
                    int run_i = -1;
#pragma unroll 1
                    for (int ch = ch0; ch < ch1; ch += 2) {
                        uint32_t r0[32], r1[32];
                        const bool two = ch + 1 < ch1;
                        if (!(prm.debug & 2)) {
                            tmem_ld32(taddr + (uint32_t)ch * 32u, r0);
                            if (two) tmem_ld32(taddr + (uint32_t)(ch + 1) * 32u, r1);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) { r0[j] = 0u; r1[j] = 0u; }
                        }
                        tmem_ld_wait();
                        if (ch + 2 >= ch1) {                      // accumulator slice is in registers: release it
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(TEMPTY(acc));
                        }
                        if (!BPHW_TMA || img) {
                            epilogue_chunk<LAYOUT>(r0, s_sn + ch * 32, c0, c1, c2, row0 + ch * 32, p, pok, prm, stg, &map_out,
                                                   nt, ch * 32, img, &run_v, &run_i);
                            if (two)
                                epilogue_chunk<LAYOUT>(r1, s_sn + (ch + 1) * 32, c0, c1, c2, row0 + (ch + 1) * 32, p, pok, prm,
                                                       stg, &map_out, nt, (ch + 1) * 32, img, &run_v, &run_i);
                        } else {
                            epilogue_chunk<STG_LAYOUT>(r0, s_sn + ch * 32, c0, c1, c2, row0 + ch * 32, p, pok, prm);
                            if (two)
                                epilogue_chunk<STG_LAYOUT>(r1, s_sn + (ch + 1) * 32, c0, c1, c2, row0 + (ch + 1) * 32, p, pok, prm);
                        }
                    }
                    if (LAYOUT == LAYOUT_TOP1 && img && pok && run_i >= 0 && !(prm.debug & 1))
                        atomicMax(reinterpret_cast<unsigned long long*>(prm.out) + (size_t)nt * prm.P + p,
                                  top1_pack(run_v, run_i));
                    if (ch0 >= ch1) {                             // (never for NT >= 64; keeps the barrier count right)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(TEMPTY(acc));
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(TEMPTY(acc));
                }
                acc ^= 1;
                if (acc == 0) acc_par ^= 1u;
            }
        }
        if (TMA_ST && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores landed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------ host
// [rows, cols] fp16 row-major, box = 64 cols x box_rows, 128 B swizzle; OOB rows read as zero (tc_ptx.cuh)
bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    return make_map_f16(m, ptr, rows, cols, box_rows);
}

// output [N, P] fp32 row-major, box = 32 prototypes x 32 patches, no swizzle (TMA-store epilogue)
bool make_out_map(CUtensorMap* m, const void* ptr, uint64_t N, uint64_t P) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {P, N};
    cuuint64_t strides[1] = {P * sizeof(float)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// output [B, P, HW] fp32, box = 32 patches x 32 prototypes x 1 image, 128B swizzle (inner box = 128 B)
bool make_out_map_bphw(CUtensorMap* m, const void* ptr, uint64_t B, uint64_t P, uint64_t HW) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {HW, P, B};
    cuuint64_t strides[2] = {HW * sizeof(float), P * HW * sizeof(float)};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct WsLayout {
    size_t bh, bl, e0, e1, e2, flag, ah, al, sn, total;
};
WsLayout ws_layout(long long N, int P, int D) {
    WsLayout w;
    size_t o = 0;
    w.bh = o; o = align256(o + (size_t)P * 2 * D * 2);
    w.bl = o; o = align256(o + (size_t)P * 2 * D * 2);
    w.e0 = o; o = align256(o + (size_t)P * 4);
    w.e1 = o; o = align256(o + (size_t)P * 4);
    w.e2 = o; o = align256(o + (size_t)P * 4);
    w.flag = o; o = align256(o + 4);
    w.ah = o; o = align256(o + (size_t)N * 2 * D * 2);
    w.al = o; o = align256(o + (size_t)N * 2 * D * 2);
    w.sn = o; o = align256(o + (size_t)((N + 255) / 256 * 256) * 4);
    w.total = o;
    return w;
}

}  // namespace

// logprob_tcz.cu: [N,P] output, isotropic sigma, D <= 128: patch tile resident in TMEM, x split fused
bool mgp_logprob_tcz_supported(int P, int D);
int mgp_logprob_tcz_launch(const float* xhat, const void* bh, const void* bl, const float* e0, const float* e1,
                           const float* e2, const int* noniso, float* out, long long N, int P, int D, cudaStream_t st);
int mgp_opt_tc_z();   // abi.cu

bool mgp_logprob_tc_supported(int layout, int B, int HW, int P, int D, int assume_iso) {
    (void)layout;
    // K blocks of 64; the x tile (128 patches x K x 4 B, K = 2D when some sigma is anisotropic) must fit in
    // shared memory: D <= 128 always, D = 256 only when the caller asserts isotropic sigma (MGP_MATH_TC_ISO)
    if (!(D == 64 || D == 128 || (D == 256 && assume_iso))) return false;
    if ((long long)B * HW < 1 || P < 1) return false;
    return get_encode() != nullptr;
}

size_t mgp_logprob_tc_ws_bytes(long long N, int P, int D) { return ws_layout(N, P, D).total; }

// the patch-side operand slots of the workspace, for a producer that writes them itself (mgp_normalize_fwd_stage)
bool mgp_logprob_tc_stage_ptrs(void* ws, size_t ws_bytes, long long N, int P, int D, __half** ah, __half** al, float** sn) {
    const WsLayout w = ws_layout(N, P, D);
    if (!ws || ws_bytes < w.total) return false;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
    *ah = reinterpret_cast<__half*>(wsb + w.ah);
    *al = reinterpret_cast<__half*>(wsb + w.al);
    *sn = reinterpret_cast<float*>(wsb + w.sn);
    return true;
}

int mgp_logprob_tc_launch(const float* xhat, const float* mu, const float* sigma, float eps, float eps_log, float* out,
                          int layout, int B, int HW, int P, int D, void* ws, size_t ws_bytes, int reuse_operands,
                          int assume_iso, int x_staged, cudaStream_t st) {
    const long long N = (long long)B * HW;
    const WsLayout w = ws_layout(N, P, D);
    if (ws_bytes < w.total) return MGP_ERR_WORKSPACE;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
    __half* bh = reinterpret_cast<__half*>(wsb + w.bh);
    __half* bl = reinterpret_cast<__half*>(wsb + w.bl);
    __half* ah = reinterpret_cast<__half*>(wsb + w.ah);
    __half* al = reinterpret_cast<__half*>(wsb + w.al);
    float* e0 = reinterpret_cast<float*>(wsb + w.e0);
    float* e1 = reinterpret_cast<float*>(wsb + w.e1);
    float* e2 = reinterpret_cast<float*>(wsb + w.e2);
    float* sn = reinterpret_cast<float*>(wsb + w.sn);
    int* flag = reinterpret_cast<int*>(wsb + w.flag);

    // [N,P] with isotropic sigma (asserted by the caller) and D <= 128: the TMEM-resident kernel reads fp32 x itself
    const bool use_z = (layout == MGP_OUT_LOGP_NP) && assume_iso && mgp_opt_tc_z() && mgp_logprob_tcz_supported(P, D);
    if (!reuse_operands) {
        MGP_CUDA(cudaMemsetAsync(flag, 0, 4, st));
        tc_proto_prep_kernel<<<(P + 7) / 8, 256, 0, st>>>(mu, sigma, eps, eps_log, bh, bl, e0, e1, e2, flag, P, D);
        MGP_CHECK_LAUNCH();
        if (!use_z && !(x_staged & 1)) {
            tc_x_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(xhat, ah, al, sn, flag, (int)N, D);
            MGP_CHECK_LAUNCH();
        }
    }
    if (use_z) return mgp_logprob_tcz_launch(xhat, bh, bl, e0, e1, e2, flag, out, N, P, D, st);

    CUtensorMap mxh, mxl, mph, mpl;
    CUtensorMap mout;
    const char* no_tma = getenv("MGP_TC_NO_TMA_STORE");
    const bool tma_ok = !(no_tma && atoi(no_tma));
    const bool tma_np = (layout == MGP_OUT_LOGP_NP) && (P % 4 == 0) && tma_ok;
    // [B,P,HW] through the 3-D map uses image-aligned x tiles (a chunk may not cross an image end)
    const bool tma_bphw = (layout != MGP_OUT_LOGP_NP) && (HW % 4 == 0) && HW >= 32 && HW <= 256 && tma_ok && D <= 128;
    const bool top1 = (layout == MGP_OUT_TOP1_BP);
    const bool tma_store = (tma_np || tma_bphw) && !top1;
    // image-aligned x tiles (box of 32 rows) for the [B,P,HW] TMA stores and for the top-1 epilogue
    const bool img_tiles = (tma_bphw && !top1) || (top1 && HW >= 32 && HW <= 256 && D <= 128);
    const uint32_t xbox = img_tiles ? 32u : 128u;
    if (top1) MGP_CUDA(cudaMemsetAsync(out, 0, (size_t)B * P * sizeof(unsigned long long), st));
    if (!make_map(&mxh, ah, (uint64_t)N, 2 * D, xbox) || !make_map(&mxl, al, (uint64_t)N, 2 * D, xbox) ||
        !make_map(&mph, bh, (uint64_t)P, 2 * D, 128) || !make_map(&mpl, bl, (uint64_t)P, 2 * D, 128))
        return MGP_ERR_UNSUPPORTED;
    if (tma_bphw && !top1) {
        if (!make_out_map_bphw(&mout, out, (uint64_t)B, (uint64_t)P, (uint64_t)HW)) return MGP_ERR_UNSUPPORTED;
    } else if (!make_out_map(&mout, tma_np ? out : (float*)ah, tma_np ? (uint64_t)N : 64, tma_np ? (uint64_t)P : 64)) {
        return MGP_ERR_UNSUPPORTED;
    }

    TcParams prm;
    prm.e0 = e0; prm.e1 = e1; prm.e2 = e2; prm.sn = sn; prm.noniso = flag; prm.out = out;
    prm.N = (int)N; prm.HW = HW; prm.P = P; prm.D = D;
    prm.x_no_sq = (x_staged == 1) ? 1 : 0;          // staged without the x^2 half: an anisotropic sigma must fault, not read stale data
    {
        const char* dbg = getenv("MGP_TC_DEBUG");
        prm.debug = dbg ? atoi(dbg) : 0;
    }
    prm.n_ptiles = (P + PT - 1) / PT;
    prm.n_ntiles = (int)((N + 127) / 128);
    prm.B = B;
    prm.xbox = (int)xbox;
    prm.nti = ((HW + 31) / 32) * 32;
    int dev = 0, sms = 148;
    MGP_CUDA(cudaGetDevice(&dev));
    MGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // teams of 4 CTAs (fewer when there are fewer prototype tiles) share an x tile and write adjacent tiles
    int team = 4;
    {
        const char* ts = getenv("MGP_TC_TEAM");
        if (ts && atoi(ts) > 0) team = atoi(ts);
    }
    if (team > prm.n_ptiles) team = prm.n_ptiles;
    if (team > sms) team = sms;
    int n_teams = sms / team;
    {
        const int nt_min = (img_tiles && B < prm.n_ntiles) ? B : prm.n_ntiles;
        if (n_teams > nt_min) n_teams = nt_min;
    }
    prm.team = team;
    const int grid = n_teams * team;
    // shared memory: 1 KiB alignment slack + x tile(s) + prototype stages [+ 32 KiB TMA-store staging] + 2 KiB misc
    const size_t x_max = (size_t)(assume_iso && D > 128 ? 512 : 1024) * D;   // general: 128 x 2D x 4 B (isotropic: half)
    if (1024 + x_max + (size_t)2 * 2 * SUB_BYTES + (tma_store ? STAGING_BYTES : 0) + 2048 > (size_t)227 * 1024)
        return MGP_ERR_UNSUPPORTED;
    const size_t smem = (size_t)227 * 1024;
    prm.smem_bytes = (uint32_t)smem;

#define MGP_TC_LAUNCH(L)                                                                                           \
    do {                                                                                                           \
        MGP_CUDA(cudaFuncSetAttribute(logprob_tc_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        logprob_tc_kernel<L><<<grid, TC_THREADS, smem, st>>>(mxh, mxl, mph, mpl, mout, prm);                               \
    } while (0)
    if (top1) MGP_TC_LAUNCH(LAYOUT_TOP1);
    else if (tma_np) MGP_TC_LAUNCH(LAYOUT_NP_TMA);
    else if (tma_bphw && layout == MGP_OUT_LOGP_BPHW) MGP_TC_LAUNCH(LAYOUT_BPHW_TMA);
    else if (tma_bphw) MGP_TC_LAUNCH(LAYOUT_NEGP_TMA);
    else if (layout == MGP_OUT_LOGP_NP) MGP_TC_LAUNCH(MGP_OUT_LOGP_NP);
    else if (layout == MGP_OUT_LOGP_BPHW) MGP_TC_LAUNCH(MGP_OUT_LOGP_BPHW);
    else MGP_TC_LAUNCH(MGP_OUT_NEGP_BPHW);
#undef MGP_TC_LAUNCH
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}
