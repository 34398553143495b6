// a10-a12 on the 5th-gen tensor cores: the whole update_GMM (ref model.py:277-301, :303-321, :367-401) of a single
// replica in ONE launch after em_plan, one CTA per class, for the shapes the shipped loop produces
// (K <= 16 components, D = 128 or 256, sigma constant over d inside every component).
//
// Per class and EM loop the two inner products are GEMMs over the class's bank rows X [cap x D]:
//     E-step     Q  [cap x K]  = X . A^T        A_k = -2 w_k mu_k          (q_nk = w_k (|x_n|^2 + |mu_k|^2) + Q_nk)
//     statistics S1^T [D x K]  = X^T . R        R_nk = smoothed responsibility
// They run as tcgen05.mma (kind::f16, fp32 accumulators in TMEM) on fp16 hi/lo splits -- hi*hi + lo*hi + hi*lo, 22
// mantissa bits, the scheme of logprob_tc.cu -- of
//   * the bank rows: a SHADOW of the fp32 bank kept in HBM as fp16 hi / lo of 256 x (written by the enqueue scatter,
//     csrc/bank.cu), so a 128-row tile is four TMA boxes into 128B-swizzled shared memory with no conversion pass.
//     The same tile serves both GEMMs: as the K-major A operand of the E-step (rows x d) and as the MN-major A operand
//     of the statistics GEMM (d x rows) -- one copy, two descriptors;
//   * the means operand A (rebuilt from the on-chip means after every Adam step) and the responsibilities R (written
//     by the E-step epilogue), both split in registers and stored in the UMMA K-major SWIZZLE_128B layout.
// Everything else (soft-max, S0, gradient, diversity term, Adam, pi momentum, the zero-gradient replay of the other
// classes' steps) is fp32 SIMT on the class's state, which stays in registers / shared memory for the whole timeline
// exactly as in em_fused_kernel (em.cu); thread d owns mean / moment elements (k, d) for all k.
//
// Per 128-row tile: TMA load (mbarrier) -> 2*D/16 E-step MMAs (hi rows 0-15 and lo rows 16-31 of the means share one
// B block: one N = 32 and one N = 16 MMA per k-step) -> epilogue (tcgen05.ld, soft-max, R) -> 2*128/16 statistics MMAs
// accumulating in TMEM across the tiles of the loop.  Two variants of the same kernel (template flag PIPE):
//   * serial (D = 256; D = 128 when more classes are active than there are SMs): the phases of a tile run one after
//     the other on one tile buffer, two CTAs per SM (D = 128) overlap each other's latencies;
//   * pipelined (D = 128, one CTA per SM, classes in the planner's order): three tile buffers, two E-step accumulators
//     and two R buffers; one thread issues TMA(t+1), the E-step MMAs of tile t and the statistics MMAs of tile t-1
//     while warps 0-3 run the soft-max of tile t; idle warps apply the replay of inactive classes.
// Both are enqueued and the planner's count of active classes decides on the device which one does the work.
// HBM/L2 traffic: num_em_loop x (4 D + 4) bytes per bank row -- the algorithmic bytes of SURVEY 8(d) K-D.
#include <cuda.h>
#include <cuda_fp16.h>

#include "mgp_common.cuh"
#include "em_common.cuh"
#include "tc_ptx.cuh"

namespace {
using namespace mgp_em;
using namespace mgp_tc;

constexpr float SX = 256.0f;      // shadow rows hold 256 x (fp16 hi + lo)
constexpr float SR = 1024.0f;     // responsibilities are stored as 1024 r
constexpr int TR = 128;           // bank rows per tile (UMMA M of the E-step, K extent of the statistics GEMM)
constexpr int NK = 16;            // UMMA N: components padded to 16

struct EmTcParams {
    const float* xx;              // [C*cap] |x|^2 of the bank rows (shadow)
    const float* bc;              // planner tables: [2 n] Adam bias corrections of steps step0+1.., then b1^i, b2^(i/2), b2^i for i <= n = L * n_active
    const int32_t* order;
    const int32_t* sched;
    const int32_t* clist;         // planner: classes in launch order, active first (stats scratch at 5 L C + 4)
    int pipe_max;                 // the pipelined kernel runs iff n_active <= pipe_max (0: never)
    float* mu;
    const float* sigma;
    float* weight;
    float* exp_avg;
    float* exp_avg_sq;
    int* status;                  // set to 1 if a class turned out to have anisotropic sigma (its update is skipped)
    long long* prof;              // profiling (mgp_debug_set_ptr("em_tc_prof")): clock64 stamps of class `prof_class`
    AdamCfg adam;
    float alpha, tau, omtau, lamda;
    int num_em_loop, C, K, cap, prof_class;
};
// stamp layout: prof[(tile_ctr * 8 + phase)]; phases: 0 TMA issued, 1 TMA landed, 2 E-step MMAs issued, 3 E-step done
// (seen by thread 0), 4 epilogue done, 5 statistics MMAs issued, 6 loop tail entered, 7 loop tail done
#define MGP_PROF(ctr, ph)                                                                          \
    do {                                                                                           \
        if (prm.prof && blockIdx.x == prm.prof_class && (ctr) < 64) prm.prof[(ctr) * 8 + (ph)] = clock64();  \
    } while (0)

// PIPE (D = 128): three row-tile buffers, two E-step accumulators and two R buffers, so that the TMA load of tile t+1,
// the E-step MMAs of tile t and the statistics MMAs of tile t-1 overlap the soft-max epilogue (one CTA per SM; the
// planner's class list puts the active classes first).  !PIPE (D = 256: one 128 KB tile buffer fits): serial per tile.
// Sum V = 32 R values per lane across the warp with V - R shuffles (a butterfly that halves the live set each round)
// instead of 5 V: afterwards a[i], i < R, holds the warp total of entry R * lane + i.
template <int V>
__device__ __forceinline__ void warp_multi_reduce(float (&a)[V], int lane) {
    static_assert(V % 32 == 0, "V must be a multiple of the warp size");
#pragma unroll
    for (int off = 16, n = V / 2; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? a[i] : a[i + n], keep = up ? a[i + n] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

template <int D, int KT, bool PIPE>
__global__ void __launch_bounds__(256, (D == 128 && !PIPE) ? 2 : 1)
em_tc_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_l, const EmTcParams prm) {
    constexpr int NCH = D / 64;                    // 64-element (128 B) chunks along d
    constexpr uint32_t CH_BYTES = TR * 128;        // one [128 rows x 64] fp16 block
    constexpr uint32_t X_BYTES = NCH * CH_BYTES;   // hi (lo follows)
    constexpr int DB = D / 128;                    // 128-wide d blocks (statistics accumulators)
    constexpr int OWN = D;                         // threads owning mean/moment elements: thread d <-> (k, d) for all k
    constexpr int TMEM_COLS = (D == 128 && !PIPE) ? 64 : 128;   // (PIPE: 2 x) 32 (E-step: hi.hi + lo.hi | hi.lo) + DB * 32 (statistics)
    constexpr int ISSUER = 128;                    // the TMA / MMA issuing thread: lane 0 of warp 4 (warps 0-3 run the E-step epilogue)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - raw);
    // carve-up (bytes from `base`)
    constexpr int NXBUF = PIPE ? 3 : 1, NRBUF = PIPE ? 2 : 1;
    const uint32_t o_xh = 0, o_xl = X_BYTES;                               // buffer b: + b * 2 * X_BYTES
    // means operand A and responsibilities R: per 64-wide K chunk one [32 rows x 128 B] block, rows 0-15 = hi,
    // rows 16-31 = lo, so ONE N = 32 MMA multiplies the row tile's hi half with both and an N = 16 MMA adds lo x hi
    const uint32_t o_a = NXBUF * 2 * X_BYTES;                             // [NCH][32][128 B]
    const uint32_t o_r = o_a + NCH * 4096;                                // NRBUF x [2 (64-row chunks)][32][128 B]
    const uint32_t o_misc = o_r + NRBUF * 8192;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bp + o_misc);            // !PIPE: tma, estep, stats | PIPE: xfull[3] xfree[3] efull[2] rfull[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* s_e = reinterpret_cast<float*>(bars + 12);                     // [KT][KT]
    float* s_red = s_e + KT * KT;                                         // [8] + [8][16]
    float* s_w = s_red + 136;                                             // [16] w_k
    float* s_ls = s_w + 16;                                               // [16] sum_d (log(sigma+eps) + log(2pi)/2)
    float* s_pi = s_ls + 16;                                              // [16]
    float* s_cst = s_pi + 16;                                             // [16]
    float* s_s0 = s_cst + 16;                                             // [16]
    float* s_misc = s_s0 + 16;                                            // [8]: replay sums T1..T3, min d; operand scale; moment decays
    constexpr int NPAIR = KT * (KT - 1) / 2;
    // per-warp partials of the loop top: |mu_i - mu_j|^2 for the NPAIR pairs, then |mu_k|^2 for the KT components
    constexpr bool MULTI = (NPAIR + KT) <= 64;                            // one transposing reduction (warp_multi_reduce)
    constexpr int PV = MULTI ? 32 * ((NPAIR + KT + 31) / 32) : NPAIR + KT;
    float* s_pair = s_misc + 8;                                           // [8 warps][PV]
    const uint32_t bar_tma = smem_u32(bars), bar_e = bar_tma + 8, bar_s = bar_tma + 16;

    const int n_active = prm.sched[0], step0 = prm.sched[1];
    // PIPE: one CTA per SM, so the planner's class list (active classes first, in order) decides who starts first
    // Launch regimes (both kernels are enqueued when D = 128; the planner's count decides on the device):
    //   n_active <= pipe_max (one CTA per SM covers every active class) -> the pipelined kernel, else the serial one.
    if (PIPE ? (n_active > prm.pipe_max) : (n_active <= prm.pipe_max)) return;
    // PIPE: CTA b < n_active runs active class clist[b]; its idle warps also take the inactive classes clist[n_active + b
    // (+ n_active)], so that those do not queue behind the 225 KB CTAs; inactive classes beyond 2 n_active get own CTAs
    constexpr int ABSORB = 2;
    if (PIPE && (int)blockIdx.x >= n_active && (int)blockIdx.x - n_active < ABSORB * n_active) return;
    const int c = PIPE ? prm.clist[blockIdx.x] : (int)blockIdx.x;
    const int ord = prm.order[c];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = prm.K, cap = prm.cap, L = prm.num_em_loop, P = prm.C * K, KD = K * D;
    const AdamCfg& adam = prm.adam;
    const bool own = tid < OWN;
    const float* sg_c = prm.sigma + (size_t)c * KD;

    if (tid == 0) MGP_PROF(63, 0);
    float p_[KT], m_[KT], v_[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        p_[k] = 0.f; m_[k] = 0.f; v_[k] = 1.f;
        if (own && k < K) {
            const size_t o = (size_t)c * KD + k * D + tid;
            p_[k] = prm.mu[o]; m_[k] = prm.exp_avg[o]; v_[k] = prm.exp_avg_sq[o];
        }
    }
    // `count` zero-gradient Adam steps first+1 .. first+count on the registers (em_common.cuh).  The per-step sum
    //   p <- p - m0 * sum_s c_s / (a d_s + eps),   a = sqrt(v0),
    // is evaluated through its expansion in eps / (a d_s) (< 1e-3 whenever a d_min > 1e3 eps, i.e. v0 > ~1e-10):
    //   sum_s c_s / (a d_s + eps) = T1 / a - eps T2 / a^2 + eps^2 T3 / a^3 - ...,   Tj = sum_s c_s / d_s^j,
    // three block-wide scalars per replay (truncation < 1e-9) instead of one reciprocal per (step, element); elements
    // with a tiny second moment (a fresh optimiser) take the explicit loop.
    // The step-dependent factors come from tables the planner wrote once for the whole launch (em.cu em_plan_kernel):
    //   bc[2i] = lr / (1 - b1^t), bc[2i+1] = sqrt(1 - b2^t), t = step0 + 1 + i;   B1[n] = b1^n, B2H[n] = b2^(n/2), B2[n] = b2^n
    // so that c_s = bc0[t] B1[s], d_s = B2H[s] / bc1[t]: no transcendental is evaluated per class.
    const int n_steps = L * n_active;
    const float* t_b1 = prm.bc + 2 * n_steps;
    const float* t_b2h = t_b1 + (n_steps + 1);
    const float* t_b2 = t_b2h + (n_steps + 1);
    // (a) one warp: the replay's block-wide scalars T1..T3 and min d_s
    auto replay_sums = [&](int first, int count, int ns, bool tail, float& T1, float& T2, float& T3, float& dmin_out) {
        const int i0 = first - step0;                       // table index of step first+1
        double t1 = 0.0, t2 = 0.0, t3 = 0.0;
        float dmin = INFINITY;
        for (int s = lane + 1; s <= ns + (tail ? 1 : 0); s += 32) {
            float cs, ds;
            if (s <= ns) {
                cs = __ldg(prm.bc + 2 * (i0 + s - 1)) * __ldg(t_b1 + s);
                ds = __ldg(t_b2h + s) / __ldg(prm.bc + 2 * (i0 + s - 1) + 1);
            } else {                                    // steps ns+1 .. count as one geometric term (em_common.cuh)
                const float geo = __ldg(t_b1 + ns + 1) * (1.0f - __ldg(t_b1 + (count - ns))) / (float)(1.0 - adam.beta1);
                const int ss = min(count, ns + 1 + (int)(adam.beta1 / (1.0 - adam.beta1)));
                cs = __ldg(prm.bc + 2 * (i0 + ns)) * geo;
                ds = __ldg(t_b2h + ss) / __ldg(prm.bc + 2 * (i0 + ss - 1) + 1);
            }
            const double c = (double)cs, r = 1.0 / (double)ds;
            t1 += c * r; t2 += c * r * r; t3 += c * r * r * r;
            dmin = fminf(dmin, ds);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t1 += __shfl_xor_sync(0xffffffffu, t1, o);
            t2 += __shfl_xor_sync(0xffffffffu, t2, o);
            t3 += __shfl_xor_sync(0xffffffffu, t3, o);
            dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        }
        T1 = (float)t1; T2 = (float)t2; T3 = (float)t3; dmin_out = dmin;
    };
    // (b) one element: p after the `count` zero-gradient steps (the moments decay separately)
    auto replay_elem = [&](float p, float m, float v, int first, int count, int ns, bool tail, float T1, float T2, float T3,
                           float dmin) -> float {
        const float a = sqrtf(v);
        if (a * dmin > 1000.0f * adam.epsf) {
            const float inv = 1.0f / a;
            const float e = adam.epsf * inv;
            return fmaf(-m * inv, fmaf(-e, fmaf(-e, T3, T2), T1), p);
        }
        if (m != 0.f) {                                     // tiny second moment (fresh optimiser): term by term
            const int i0 = first - step0;
            for (int s = 1; s <= ns; ++s) {
                const float cs = __ldg(prm.bc + 2 * (i0 + s - 1)) * __ldg(t_b1 + s);
                const float ds = __ldg(t_b2h + s) / __ldg(prm.bc + 2 * (i0 + s - 1) + 1);
                p = fmaf(-cs * m, 1.0f / fmaf(a, ds, adam.epsf), p);
            }
            if (tail) {
                const float geo = __ldg(t_b1 + ns + 1) * (1.0f - __ldg(t_b1 + (count - ns))) / (float)(1.0 - adam.beta1);
                const int ss = min(count, ns + 1 + (int)(adam.beta1 / (1.0 - adam.beta1)));
                const float cs = __ldg(prm.bc + 2 * (i0 + ns)) * geo;
                const float ds = __ldg(t_b2h + ss) / __ldg(prm.bc + 2 * (i0 + ss - 1) + 1);
                p = fmaf(-cs * m, 1.0f / fmaf(a, ds, adam.epsf), p);
            }
        }
        return p;
    };
    // the sums of a replay, by ONE warp, into sums[0..3] (shared memory; the caller's next block barrier publishes them)
    auto replay_prepare = [&](int first, int count, float* sums) {
        if (count <= 0) return;
        const int ns = replay_explicit_steps(count, first, (float)adam.beta1);
        float T1, T2, T3, dmin;
        replay_sums(first, count, ns, count > ns, T1, T2, T3, dmin);
        if (lane == 0) { sums[0] = T1; sums[1] = T2; sums[2] = T3; sums[3] = dmin; }
    };
    auto replay = [&](int first, int count, const float* sums) {
        if (count <= 0) return;
        const int ns = replay_explicit_steps(count, first, (float)adam.beta1);
        const bool tail = count > ns;
        if (own) {
            const float T1 = sums[0], T2 = sums[1], T3 = sums[2], dmin = sums[3];
#pragma unroll
            for (int k = 0; k < KT; ++k) p_[k] = replay_elem(p_[k], m_[k], v_[k], first, count, ns, tail, T1, T2, T3, dmin);
        }
        const float mdec = __ldg(t_b1 + count), vdec = __ldg(t_b2 + count);   // the moments decay by the full count
#pragma unroll
        for (int k = 0; k < KT; ++k) { m_[k] *= mdec; v_[k] *= vdec; }
    };
    // an inactive class only takes everybody's zero-gradient steps: warps [w0, w0 + nw) of this CTA, straight from / to
    // global memory (PIPE: done by the warps that idle during an active class's first tile loop)
    auto replay_inactive = [&](int ci, int w0, int nw) {
        const int count = L * n_active;
        if (count <= 0) return;
        const int ns = replay_explicit_steps(count, step0, (float)adam.beta1);
        const bool tail = count > ns;
        float T1, T2, T3, dmin;
        replay_sums(step0, count, ns, tail, T1, T2, T3, dmin);            // (every warp for itself: no block barrier here)
        const float mdec = __ldg(t_b1 + count), vdec = __ldg(t_b2 + count);
        for (int o = (warp - w0) * 32 + lane; o < KD; o += nw * 32) {
            const size_t g = (size_t)ci * KD + o;
            const float p = prm.mu[g], m = prm.exp_avg[g], v = prm.exp_avg_sq[g];
            prm.mu[g] = replay_elem(p, m, v, step0, count, ns, tail, T1, T2, T3, dmin);
            prm.exp_avg[g] = m * mdec;
            prm.exp_avg_sq[g] = v * vdec;
        }
    };
    auto write_back = [&]() {
        if (!own) return;
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                const size_t o = (size_t)c * KD + k * D + tid;
                prm.mu[o] = p_[k]; prm.exp_avg[o] = m_[k]; prm.exp_avg_sq[o] = v_[k];
            }
    };

    if (ord < 0) {                                   // inactive class: it only takes everybody's zero-gradient steps
        if (warp == 0) replay_prepare(step0, L * n_active, s_misc);
        __syncthreads();
        replay(step0, L * n_active, s_misc);
        write_back();
        return;
    }
    // the two replays' block-wide sums depend on the plan only: two otherwise idle warps evaluate them under the set-up
    if (warp == 7) replay_prepare(step0, L * ord, s_misc);
    if (warp == 6) replay_prepare(step0 + L * (ord + 1), L * (n_active - ord - 1), s_misc + 4);

    if (tid == 0) MGP_PROF(63, 1);
    // ---- set-up: barriers, TMEM, sigma-derived constants, zeroed operand tiles
    if (tid == 0) {
        if (PIPE) {
            for (int i = 0; i < 3; ++i) { mbar_init(bar_tma + 8u * i, 1); mbar_init(bar_tma + 8u * (3 + i), 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(bar_tma + 8u * (6 + i), 1); mbar_init(bar_tma + 8u * (8 + i), 4); }
        } else {
            mbar_init(bar_tma, 1);
            mbar_init(bar_e, 1);
            mbar_init(bar_s, 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    bool same = true;
    for (int i = tid; i < KD; i += 256) same = same && (sg_c[i] == sg_c[(i / D) * D]);
    for (uint32_t i = tid * 16u; i < NCH * 4096u + NRBUF * 8192u; i += 256u * 16u)     // A and R blocks: rows >= K stay zero
        *reinterpret_cast<uint4*>(bp + o_a + i) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < KT * KT; i += 256) s_e[i] = 0.f;
    if (tid < 16) {
        float w = 0.f, ls = 0.f, pi = 0.f;
        if (tid < K) {
            const float sg = sg_c[tid * D] + EM_EPS;                                     // ref :333-334
            w = 1.0f / (sg * sg);
            ls = (float)D * (logf(sg) + 0.5f * MGP_LOG_2PI);                             // D equal terms
            pi = prm.weight[(size_t)c * P + (size_t)c * K + tid];
        }
        s_w[tid] = w; s_ls[tid] = ls; s_pi[tid] = pi;
    }
    tc_fence_before();
    const bool iso = __syncthreads_and(same ? 1 : 0) != 0;
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!iso) {                                      // the host promised isotropic sigma: flag it, leave the class untouched
        if (tid == 0) atomicExch(prm.status, 1);
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
        return;
    }
    if (tid == 0) MGP_PROF(63, 2);
    replay(step0, L * ord, s_misc);
    if (tid == 0) MGP_PROF(63, 3);

    const int ntiles = (cap + TR - 1) / TR;
    const float n_rows = (float)cap;
    const float inv_den = 1.0f / (1.0f + (float)K * prm.alpha);
    const float div_scale = -4.0f * prm.lamda / ((float)K * (float)(K - 1));
    // E-step: A = X (K-major), B = [means hi ; means lo] (K-major): N = 32 with X hi, N = 16 (hi only) with X lo
    const uint32_t idesc_e32 = umma_idesc_f16(TR, 2 * NK, 0, 0), idesc_e16 = umma_idesc_f16(TR, NK, 0, 0);
    // statistics: A = X^T (MN-major), B = [R hi ; R lo] (K-major)
    const uint32_t idesc_s32 = umma_idesc_f16(128, 2 * NK, 1, 0), idesc_s16 = umma_idesc_f16(128, NK, 1, 0);
    const uint32_t d_e = tmem_base;                                  // [128 rows x (16 hi.hi + lo.hi | 16 hi.lo)]
    const uint32_t d_s = tmem_base + (PIPE ? 64 : 32);               // DB x [128 d x 32], same column split
    uint32_t tile_ctr = 0;                                           // tiles issued so far (mbarrier phases)

    auto load_tile = [&](int t) {                                    // issuer only: one 128-row tile, hi + lo, into the X buffer
        mbar_expect_tx(bar_tma, 2 * X_BYTES);
        const int row0 = c * cap + t * TR;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            tma_load_2d(base + o_xh + ch * CH_BYTES, &map_h, ch * 64, row0, bar_tma);
            tma_load_2d(base + o_xl + ch * CH_BYTES, &map_l, ch * 64, row0, bar_tma);
        }
    };
    for (int loop = 0; loop < L; ++loop) {
        // ---- means operand, |mu_k|^2, diversity kernel from the current means (all from the owners' registers)
        float amax = 0.f;
        if (own) {                                                   // (warp-uniform: OWN is a multiple of 32)
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K) amax = fmaxf(amax, fabsf(2.0f * s_w[k] * p_[k]));
            amax = warp_max(amax);
            // this thread's dimension of |mu_i - mu_j|^2 (ref utils/helpers.py:13-14; i < j) and of |mu_k|^2, summed over the warp
            if constexpr (MULTI) {
                float a[PV];
                int pi = 0;
#pragma unroll
                for (int i = 0; i < KT; ++i)
#pragma unroll
                    for (int j = i + 1; j < KT; ++j, ++pi) {
                        const float df = (j < K) ? p_[i] - p_[j] : 0.f;
                        a[pi] = df * df;
                    }
#pragma unroll
                for (int k = 0; k < KT; ++k) a[NPAIR + k] = (k < K) ? p_[k] * p_[k] : 0.f;
#pragma unroll
                for (int i = NPAIR + KT; i < PV; ++i) a[i] = 0.f;
                warp_multi_reduce<PV>(a, lane);                      // lane l now holds entries (PV/32) l + i
#pragma unroll
                for (int i = 0; i < PV / 32; ++i) s_pair[warp * PV + (PV / 32) * lane + i] = a[i];
            } else {
                int pi = 0;
#pragma unroll
                for (int i = 0; i < KT; ++i)
#pragma unroll
                    for (int j = i + 1; j < KT; ++j, ++pi) {
                        const float df = (j < K) ? p_[i] - p_[j] : 0.f;
                        const float t = warp_sum(df * df);
                        if (lane == 0) s_pair[warp * PV + pi] = t;
                    }
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    const float t = warp_sum((k < K) ? p_[k] * p_[k] : 0.f);
                    if (lane == 0) s_pair[warp * PV + NPAIR + k] = t;
                }
            }
        }
        if (lane == 0) s_red[warp] = amax;
        __syncthreads();
        float a_scale;
        {
            float mx = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) mx = fmaxf(mx, s_red[w8]);
            int ex = 0;
            if (mx > 0.f) frexpf(mx, &ex);
            a_scale = ldexpf(1.0f, 8 - ex);                          // max |a| * scale in [128, 256)
        }
        if (tid < K) {
            float mm = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < OWN / 32; ++w8) mm += s_pair[w8 * PV + NPAIR + tid];
            s_cst[tid] = -s_ls[tid] + logf(s_pi[tid] + EM_EPS) - 0.5f * s_w[tid] * mm;   // ref :316, :323-336
        }
        if (tid >= 32 && tid < 32 + NPAIR) {                         // exp(-|mu_i - mu_j|^2), ref model.py:390-392
            const int pr = tid - 32;
            int i = 0, rem = pr;
            while (rem >= KT - 1 - i) { rem -= KT - 1 - i; ++i; }
            const int j = i + 1 + rem;
            float t = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < OWN / 32; ++w8) t += s_pair[w8 * PV + pr];
            const float e = (j < K) ? expf(-t) : 0.f;
            s_e[i * KT + j] = e;
            s_e[j * KT + i] = e;
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (own && k < K) {
                const float a = -2.0f * s_w[k] * p_[k] * a_scale;
                const __half h = __float2half_rn(a);
                const uint32_t off = (uint32_t)(tid >> 6) * 4096u + (uint32_t)k * 128u +
                                     (uint32_t)(((((tid & 63) >> 3) ^ (k & 7)) & 7) << 4) + (uint32_t)(tid & 7) * 2u;
                *reinterpret_cast<__half*>(bp + o_a + off) = h;                                   // row k      (hi)
                *reinterpret_cast<__half*>(bp + o_a + off + 2048u) = __float2half_rn(a - __half2float(h));   // row 16 + k (lo): same swizzle phase
            }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // operand stores -> visible to the MMA (async proxy)
        __syncthreads();

        float s0[KT];
#pragma unroll
        for (int k = 0; k < KT; ++k) s0[k] = 0.f;
        const float inv_a = 1.0f / (a_scale * SX);

        if constexpr (PIPE) {
            // mbarriers: xfull[b] tile landed in buffer b | xfree[b] the statistics MMAs reading buffer b (and the R buffer
            // they used) have retired | efull[e] E-step accumulator e complete | rfull[e] soft-max wrote R buffer e.
            // Global tile counter g: row-tile buffer g % 3, accumulator / R buffer g & 1; every barrier completes exactly
            // once per tile that uses its buffer, so its phase parity is (g / 3) & 1 resp. (g / 2) & 1.
            auto XFULL = [&](uint32_t b) { return bar_tma + 8u * b; };
            auto XFREE = [&](uint32_t b) { return bar_tma + 8u * (3 + b); };
            auto EFULL = [&](uint32_t e) { return bar_tma + 8u * (6 + e); };
            auto RFULL = [&](uint32_t e) { return bar_tma + 8u * (8 + e); };
            const uint32_t g0 = tile_ctr;
            auto load_tile_p = [&](int t, uint32_t g) {              // issuer: tile t of this class -> buffer g % 3
                const uint32_t b = g % 3u;
                if (g >= 3) mbar_wait(XFREE(b), ((g - 3) / 3u) & 1u);   // the previous tile in this buffer has been consumed
                mbar_expect_tx(XFULL(b), 2 * X_BYTES);
                const int row0 = c * cap + t * TR;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    tma_load_2d(base + b * 2 * X_BYTES + o_xh + ch * CH_BYTES, &map_h, ch * 64, row0, XFULL(b));
                    tma_load_2d(base + b * 2 * X_BYTES + o_xl + ch * CH_BYTES, &map_l, ch * 64, row0, XFULL(b));
                }
            };
            auto stats_mma = [&](int t, uint32_t g) {                // issuer: D_s += X(g)^T . [R_hi ; R_lo](g)
                const uint32_t xb = base + (g % 3u) * 2 * X_BYTES, rb = base + o_r + (g & 1u) * 8192u;
                mbar_wait(RFULL(g & 1u), (g >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < TR / 16; ++ks) {
                    const uint32_t xo = (uint32_t)ks * 2048u;        // 16 rows x 128 B
                    const uint64_t bd = umma_desc(rb + (uint32_t)(ks >> 2) * 4096u + (uint32_t)(ks & 3) * 32u);
                    tc_mma_f16(d_s, umma_desc_mn(xb + o_xh + xo, CH_BYTES, 1024u), bd, idesc_s32, (t | ks) != 0);
                    tc_mma_f16(d_s, umma_desc_mn(xb + o_xl + xo, CH_BYTES, 1024u), bd, idesc_s16, 1u);
                }
                tc_commit(XFREE(g % 3u));
                MGP_PROF(g, 5);
            };
            if (warp == ISSUER / 32) {
                if (tid == ISSUER) {
                    if (loop == 0) load_tile_p(0, g0);               // (later loops: prefetched under the previous loop's tail)
                    for (int t = 0; t < ntiles; ++t) {
                        const uint32_t g = g0 + t;
                        MGP_PROF(g, 0);
                        if (t + 1 < ntiles) load_tile_p(t + 1, g + 1);
                        mbar_wait(XFULL(g % 3u), (g / 3u) & 1u);
                        MGP_PROF(g, 1);
                        tc_fence_after();
                        const uint32_t xb = base + (g % 3u) * 2 * X_BYTES, de = d_e + (g & 1u) * 32u;
#pragma unroll
                        for (int ks = 0; ks < D / 16; ++ks) {
                            const uint32_t xo = (uint32_t)(ks >> 2) * CH_BYTES + (uint32_t)(ks & 3) * 32u;
                            const uint64_t bd = umma_desc(base + o_a + (uint32_t)(ks >> 2) * 4096u + (uint32_t)(ks & 3) * 32u);
                            tc_mma_f16(de, umma_desc(xb + o_xh + xo), bd, idesc_e32, ks != 0);
                            tc_mma_f16(de, umma_desc(xb + o_xl + xo), bd, idesc_e16, 1u);
                        }
                        tc_commit(EFULL(g & 1u));
                        MGP_PROF(g, 2);
                        if (t >= 1) stats_mma(t - 1, g - 1);
                    }
                    stats_mma(ntiles - 1, g0 + ntiles - 1);
                    if (loop + 1 < L) load_tile_p(0, g0 + ntiles);   // the next loop's first tile, under this loop's tail
                }
                __syncwarp();
            }
            if (warp >= 5 && loop == 0) {                            // idle here: the absorbed inactive classes
                for (int j = (int)blockIdx.x; j < ABSORB * n_active && n_active + j < prm.C; j += n_active)
                    replay_inactive(prm.clist[n_active + j], 5, 3);
            }
            if (warp < 4) {
                for (int t = 0; t < ntiles; ++t) {
                    const uint32_t g = g0 + t;
                    const int row = t * TR + tid;
                    const bool valid = row < cap;
                    const float xxv = valid ? __ldg(prm.xx + (size_t)c * cap + row) : 0.f;
                    mbar_wait(EFULL(g & 1u), (g >> 1) & 1u);
                    if (tid == 0) MGP_PROF(g, 3);
                    tc_fence_after();
                    uint32_t q[16], q1[16];
                    const uint32_t de = d_e + (g & 1u) * 32u + ((uint32_t)(warp * 32) << 16);
                    tmem_ld16(de, q);
                    tmem_ld16(de + 16, q1);
                    tmem_ld_wait();
                    float wl[KT], mx = -INFINITY;
#pragma unroll
                    for (int k = 0; k < KT; ++k) {
                        const float acc = __uint_as_float(q[k]) + __uint_as_float(q1[k]);
                        const float qq = fmaf(s_w[k], xxv, acc * inv_a);
                        wl[k] = (k < K) ? s_cst[k] - 0.5f * qq : -INFINITY;             // lp + log(pi + eps)  (ref :316)
                        mx = fmaxf(mx, wl[k]);
                    }
                    float se = 0.f;
#pragma unroll
                    for (int k = 0; k < KT; ++k) {
                        wl[k] = (k < K) ? expf(wl[k] - mx) : 0.f;
                        se += wl[k];
                    }
                    const float inv_se = 1.0f / se;
                    if (t >= 2) mbar_wait(XFREE((g - 2) % 3u), ((g - 2) / 3u) & 1u);   // the MMAs that read this R buffer have retired
                    uint8_t* rbp = bp + o_r + (g & 1u) * 8192u;
                    const uint32_t rbase = (uint32_t)(tid >> 6) * 4096u + (uint32_t)(tid & 7) * 2u;
                    const int c16 = (tid & 63) >> 3;
#pragma unroll
                    for (int k = 0; k < KT; ++k)
                        if (k < K) {
                            const float r = valid ? fmaf(wl[k], inv_se, prm.alpha) * inv_den : 0.f;   // ref :380-383
                            s0[k] += r;
                            const float rs = r * SR;
                            const __half h = __float2half_rn(rs);
                            const uint32_t off = rbase + (uint32_t)k * 128u + (uint32_t)(((c16 ^ (k & 7)) & 7) << 4);
                            *reinterpret_cast<__half*>(rbp + off) = h;                                         // row k
                            *reinterpret_cast<__half*>(rbp + off + 2048u) = __float2half_rn(rs - __half2float(h));   // row 16 + k
                        }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (tid == 0) MGP_PROF(g, 4);
                    if (lane == 0) mbar_arrive(RFULL(g & 1u));
                }
            }
            tile_ctr += (uint32_t)ntiles;
            if (tid == 0) MGP_PROF(tile_ctr - 1, 6);
            // ---- S0 over the class; the owners wait for the last statistics MMAs of this loop
            if (warp < 4) {
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    const float v = warp_sum(s0[k]);
                    if (lane == 0) s_red[warp * 16 + k] = v;
                }
                const uint32_t gl = tile_ctr - 1;
                mbar_wait(XFREE(gl % 3u), (gl / 3u) & 1u);
                tc_fence_after();
            }
            __syncthreads();
        } else {
        for (int t = 0; t < ntiles; ++t, ++tile_ctr) {
            const uint32_t par = tile_ctr & 1u;
            if (tid == ISSUER) {
                MGP_PROF(tile_ctr, 0);
                if (!(t == 0 && loop > 0)) {                                 // (a loop's first tile was prefetched by the previous loop)
                    if (tile_ctr > 0) mbar_wait(bar_s, (tile_ctr - 1) & 1u); // previous statistics MMAs have read X and R
                    load_tile(t);
                }
                mbar_wait(bar_tma, par);
                MGP_PROF(tile_ctr, 1);
                tc_fence_after();
                // E-step: D_e[row, 0:32] = X_hi . [A_hi ; A_lo]^T (N = 32);  D_e[row, 0:16] += X_lo . A_hi^T (N = 16)
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t xo = (uint32_t)(ks >> 2) * CH_BYTES + (uint32_t)(ks & 3) * 32u;
                    const uint64_t bd = umma_desc(base + o_a + (uint32_t)(ks >> 2) * 4096u + (uint32_t)(ks & 3) * 32u);
                    tc_mma_f16(d_e, umma_desc(base + o_xh + xo), bd, idesc_e32, ks != 0);
                    tc_mma_f16(d_e, umma_desc(base + o_xl + xo), bd, idesc_e16, 1u);      // lo.hi adds onto the hi.hi columns
                }
                tc_commit(bar_e);
                MGP_PROF(tile_ctr, 2);
            }
            if (warp < 4) {
                // ---- E-step epilogue: thread = bank row
                const int row = t * TR + tid;
                const bool valid = row < cap;
                const float xxv = valid ? __ldg(prm.xx + (size_t)c * cap + row) : 0.f;
                mbar_wait(bar_e, par);
                if (tid == 0) MGP_PROF(tile_ctr, 3);
                tc_fence_after();
                uint32_t q[16], q1[16];
                tmem_ld16(d_e + ((uint32_t)(warp * 32) << 16), q);
                tmem_ld16(d_e + 16 + ((uint32_t)(warp * 32) << 16), q1);
                tmem_ld_wait();
                float wl[KT], mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    const float acc = __uint_as_float(q[k]) + __uint_as_float(q1[k]);
                    const float qq = fmaf(s_w[k], xxv, acc * inv_a);
                    wl[k] = (k < K) ? s_cst[k] - 0.5f * qq : -INFINITY;                 // lp + log(pi + eps)  (ref :316)
                    mx = fmaxf(mx, wl[k]);
                }
                float se = 0.f;
#pragma unroll
                for (int k = 0; k < KT; ++k) {
                    wl[k] = (k < K) ? expf(wl[k] - mx) : 0.f;
                    se += wl[k];
                }
                const float inv_se = 1.0f / se;
                const uint32_t rbase = (uint32_t)(tid >> 6) * 4096u + (uint32_t)(tid & 7) * 2u;
                const int c16 = (tid & 63) >> 3;
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if (k < K) {
                        const float r = valid ? fmaf(wl[k], inv_se, prm.alpha) * inv_den : 0.f;   // ref :380-383
                        s0[k] += r;
                        const float rs = r * SR;
                        const __half h = __float2half_rn(rs);
                        const uint32_t off = rbase + (uint32_t)k * 128u + (uint32_t)(((c16 ^ (k & 7)) & 7) << 4);
                        *reinterpret_cast<__half*>(bp + o_r + off) = h;                                           // row k
                        *reinterpret_cast<__half*>(bp + o_r + off + 2048u) = __float2half_rn(rs - __half2float(h));   // row 16 + k
                    }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                if (tid == 0) MGP_PROF(tile_ctr, 4);
            }
            __syncthreads();
            if (tid == ISSUER) {
                tc_fence_after();
                // statistics: D_s[d, 0:32] += X_hi^T . [R_hi ; R_lo] (N = 32);  D_s[d, 0:16] += X_lo^T . R_hi (N = 16)
#pragma unroll
                for (int db = 0; db < DB; ++db) {
#pragma unroll
                    for (int ks = 0; ks < TR / 16; ++ks) {
                        const uint32_t xo = (uint32_t)db * 2u * CH_BYTES + (uint32_t)ks * 2048u;   // 16 rows x 128 B
                        const uint64_t bd = umma_desc(base + o_r + (uint32_t)(ks >> 2) * 4096u + (uint32_t)(ks & 3) * 32u);
                        tc_mma_f16(d_s + db * 32, umma_desc_mn(base + o_xh + xo, CH_BYTES, 1024u), bd, idesc_s32, (t | ks) != 0);
                        tc_mma_f16(d_s + db * 32, umma_desc_mn(base + o_xl + xo, CH_BYTES, 1024u), bd, idesc_s16, 1u);
                    }
                }
                tc_commit(bar_s);
                MGP_PROF(tile_ctr, 5);
            }
        }
        if (tid == 0) MGP_PROF(tile_ctr - 1, 6);
        if (tid == ISSUER && loop + 1 < L) {          // the next loop starts on the same rows: fetch its first tile under the tail
            mbar_wait(bar_s, (tile_ctr - 1) & 1u);
            load_tile(0);
        }
        // ---- S0 over the class, S1 from TMEM
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            float v = s0[k];
            v = warp_sum(v);
            if (lane == 0 && warp < 4) s_red[warp * 16 + k] = v;
        }
        mbar_wait(bar_s, (tile_ctr - 1) & 1u);                        // all statistics MMAs of this loop have retired
        tc_fence_after();
        __syncthreads();
        }   // !PIPE
        if (tid < K) s_s0[tid] = (s_red[tid] + s_red[16 + tid]) + (s_red[32 + tid] + s_red[48 + tid]);
        uint32_t sacc[16], sacc1[16];
        if (own) {
            const uint32_t ta = d_s + (uint32_t)(tid >> 7) * 32u + ((uint32_t)((warp & 3) * 32) << 16);
            tmem_ld16(ta, sacc);
            tmem_ld16(ta + 16, sacc1);
            tmem_ld_wait();
        }
        tc_fence_before();
        __syncthreads();
        // ---- gradient + diversity + Adam on the owned elements (ref model.py:385-397)
        if (own) {
            const int sidx = L * ord + loop;                                  // this Adam step's bias corrections
            const float step_size = __ldg(prm.bc + 2 * sidx), bc2_sqrt = __ldg(prm.bc + 2 * sidx + 1);
            float newp[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                newp[k] = p_[k];
                if (k < K) {
                    const float muv = p_[k];
                    const float s1 = (__uint_as_float(sacc[k]) + __uint_as_float(sacc1[k])) * (1.0f / (SX * SR));
                    float g = -(s1 - muv * s_s0[k]) * s_w[k] / n_rows;
                    float esum = 0.f, emu = 0.f;
#pragma unroll
                    for (int j = 0; j < KT; ++j)
                        if (j < K) {
                            const float e = s_e[k * KT + j];                  // broadcast
                            esum += e;
                            emu = fmaf(e, p_[j], emu);                        // mu_j[d] is this thread's own register
                        }
                    g += div_scale * (esum * muv - emu);
                    const float mm = m_[k] + (g - m_[k]) * adam.omb1;             // torch.optim.Adam (_single_tensor_adam)
                    const float vv = v_[k] * adam.b2f + adam.omb2 * g * g;
                    const float denom = sqrtf(vv) / bc2_sqrt + adam.epsf;
                    newp[k] = muv - step_size * (mm / denom);
                    m_[k] = mm; v_[k] = vv;
                }
            }
#pragma unroll
            for (int k = 0; k < KT; ++k) p_[k] = newp[k];
        }
        __syncthreads();                                              // every reader of s_mu / s_s0 is done
        if (tid < K) s_pi[tid] = prm.tau * s_pi[tid] + prm.omtau * ((s_s0[tid] + EM_EPS) / n_rows);   // ref :385, :399, :297
        if (tid == 0) MGP_PROF(tile_ctr - 1, 7);
    }
    if (tid == 0) MGP_PROF(63, 4);
    replay(step0 + L * (ord + 1), L * (n_active - ord - 1), s_misc + 4);
    if (tid == 0) MGP_PROF(63, 5);
    write_back();
    if (tid < K) prm.weight[(size_t)c * P + (size_t)c * K + tid] = s_pi[tid];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
    if (tid == 0) MGP_PROF(63, 6);
}

template <int D>
size_t em_tc_smem(int kt, bool pipe) {
    return 1024 + (pipe ? 3 : 1) * 2 * (size_t)(D / 64) * TR * 128 + (size_t)(D / 64) * 4096 + (pipe ? 2 : 1) * 8192 +
           ((size_t)kt * kt + 136 + 6 * 16 + 8 + 8 * (size_t)(32 * ((kt * (kt - 1) / 2 + kt + 31) / 32))) * 4 + 128;
}

}  // namespace

int mgp_opt_em_pipe();   // abi.cu
static long long* g_em_tc_prof = nullptr;
static int g_em_tc_prof_class = 0;
void mgp_em_tc_set_prof(void* p, int cls) { g_em_tc_prof = reinterpret_cast<long long*>(p); g_em_tc_prof_class = cls; }

bool mgp_em_tc_supported(int K, int D, int cap) {
    return K >= 2 && K <= 16 && (D == 128 || D == 256) && cap >= 1 && get_encode() != nullptr;
}

int mgp_em_tc_launch(const void* shadow_h, const void* shadow_l, const float* shadow_xx, const float* bias_corr, const int32_t* order,
                     const int32_t* sched, float* mu, const float* sigma, float* weight, float* exp_avg, float* exp_avg_sq,
                     int* status, int num_em_loop, float alpha, double lr, double beta1, double beta2, double adam_eps,
                     double tau, float lamda, int C, int K, int D, int cap, cudaStream_t st) {
    CUtensorMap mh, ml;
    const uint64_t rows = (uint64_t)C * cap;
    if (!make_map_f16(&mh, shadow_h, rows, D, TR) || !make_map_f16(&ml, shadow_l, rows, D, TR)) return MGP_ERR_UNSUPPORTED;
    EmTcParams prm;
    prm.xx = shadow_xx; prm.bc = bias_corr; prm.order = order; prm.sched = sched; prm.mu = mu; prm.sigma = sigma; prm.weight = weight;
    prm.exp_avg = exp_avg; prm.exp_avg_sq = exp_avg_sq; prm.status = status;
    prm.adam = make_adam(lr, beta1, beta2, adam_eps);
    prm.alpha = alpha; prm.tau = (float)tau; prm.omtau = (float)(1.0 - tau); prm.lamda = lamda;
    prm.num_em_loop = num_em_loop; prm.C = C; prm.K = K; prm.cap = cap;
    prm.prof = g_em_tc_prof; prm.prof_class = g_em_tc_prof_class;
    prm.clist = reinterpret_cast<const int32_t*>(bias_corr + (size_t)5 * num_em_loop * C + 4);
    const bool pipe = mgp_opt_em_pipe() != 0 && D == 128;
    static int n_sm = 0;
    if (n_sm == 0) {
        int devi = 0;
        MGP_CUDA(cudaGetDevice(&devi));
        MGP_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, devi));
    }
    prm.pipe_max = pipe ? n_sm : 0;
#define MGP_EMTC(DD, KK, PP)                                                                                        \
    do {                                                                                                            \
        const size_t smem = em_tc_smem<DD>(KK, PP);                                                                 \
        MGP_CUDA(cudaFuncSetAttribute(em_tc_kernel<DD, KK, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        em_tc_kernel<DD, KK, PP><<<C, 256, smem, st>>>(mh, ml, prm);                                                \
    } while (0)
#define MGP_EMTC_K(DD, PP)                                                                                          \
    do {                                                                                                            \
        if (K <= 5) MGP_EMTC(DD, 5, PP);                                                                            \
        else if (K <= 10) MGP_EMTC(DD, 10, PP);                                                                     \
        else MGP_EMTC(DD, 16, PP);                                                                                  \
    } while (0)
    if (D == 128) {
        if (pipe) MGP_EMTC_K(128, true);
        MGP_CHECK_LAUNCH();
        MGP_EMTC_K(128, false);                      // (returns at once when the pipelined kernel took the call)
    } else {
        MGP_EMTC_K(256, false);
    }
#undef MGP_EMTC_K
#undef MGP_EMTC
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}
