// API-parity kernel for MGProto._m_step_diversified (ref model.py:367-401) on EXPLICIT rows and log-responsibilities:
// the reference's update_GMM calls it per (class, EM loop); the product's update_GMM runs the fused kernels
// (em.cu / em_tc.cu) instead, so this entry point only serves callers that drive the reference's private methods
// themselves.  Given x [n,D], log_resp [n,K], the class's means / sigmas [K,D] and pi_old is NOT needed for the
// gradient (log(pi_old + eps) is constant in mu):
//     r      = (exp(log_resp) + alpha) / sum_k(exp(log_resp) + alpha)                    (:380-383)
//     pi_new = (sum_n r + eps) / n                                                        (:385, :399)
//     grad   = d/d mu [ -mean_n sum_k r (lp(mu) + log(pi_old + eps)) + lamda * diversity ]  (:387-396)
//            = -(S1 - mu S0) / (sigma + eps)^2 / n  +  lamda * (-4 / (K (K-1))) * (sum_j e_kj (mu_k - mu_j)),  e = exp(-|mu_k - mu_j|^2)
#include "mgp_common.cuh"
#include "em_common.cuh"

namespace {
using namespace mgp_em;

// r [n,K] smoothed responsibilities; warp per row
__global__ void resp_smooth_kernel(const float* __restrict__ log_resp, float alpha, float* __restrict__ r, int n, int K) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float den = 0.f;
    for (int k = lane; k < K; k += 32) den += expf(log_resp[(size_t)row * K + k]) + alpha;
    den = warp_sum(den);
    for (int k = lane; k < K; k += 32) r[(size_t)row * K + k] = (expf(log_resp[(size_t)row * K + k]) + alpha) / den;
}

// one CTA; thread per (k, d) element (strided), serial over rows (API parity, not a hot path)
__global__ void __launch_bounds__(256)
mstep_div_kernel(const float* __restrict__ x, const float* __restrict__ r, const float* __restrict__ mu,
                 const float* __restrict__ sigma, float lamda, float* __restrict__ pi_out, float* __restrict__ grad, int n,
                 int K, int D) {
    extern __shared__ float sm[];
    float* s_e = sm;            // [K][K]
    float* s_s0 = sm + K * K;   // [K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int pr = warp; pr < K * K; pr += 8) {                   // ref utils/helpers.py:13-14, model.py:390-392
        const int i = pr / K, j = pr - i * K;
        float t = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float df = mu[i * D + d] - mu[j * D + d];
            t = fmaf(df, df, t);
        }
        t = warp_sum(t);
        if (lane == 0) s_e[pr] = (i == j) ? 0.f : expf(-t);
    }
    for (int k = warp; k < K; k += 8) {
        float s = 0.f;
        for (int row = lane; row < n; row += 32) s += r[(size_t)row * K + k];
        s = warp_sum(s);
        if (lane == 0) {
            s_s0[k] = s;
            pi_out[k] = (s + EM_EPS) / (float)n;
        }
    }
    __syncthreads();
    const float div_scale = (K > 1) ? -4.0f * lamda / ((float)K * (float)(K - 1)) : 0.f;
    for (int o = tid; o < K * D; o += 256) {
        const int k = o / D, d = o - k * D;
        float s1 = 0.f;
        for (int row = 0; row < n; ++row) s1 = fmaf(r[(size_t)row * K + k], x[(size_t)row * D + d], s1);
        const float sg = sigma[o] + EM_EPS;
        const float muv = mu[o];
        float g = -(s1 - muv * s_s0[k]) / (sg * sg) / (float)n;
        float esum = 0.f, emu = 0.f;
        for (int j = 0; j < K; ++j) {
            const float e = s_e[k * K + j];
            esum += e;
            emu = fmaf(e, mu[j * D + d], emu);
        }
        grad[o] = g + div_scale * (esum * muv - emu);
    }
}

}  // namespace

extern "C" int mgp_em_mstep_div(const float* x, const float* log_resp, const float* mu, const float* sigma, float alpha,
                                float lamda, float* ws_nk, float* pi_out, float* grad_out, int n, int K, int D,
                                void* stream) {
    if (!x || !log_resp || !mu || !sigma || !ws_nk || !pi_out || !grad_out || n <= 0 || K <= 0 || D <= 0) return MGP_ERR_INVALID;
    if (K > 64) return MGP_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    resp_smooth_kernel<<<(n + 7) / 8, 256, 0, st>>>(log_resp, alpha, ws_nk, n, K);
    MGP_CHECK_LAUNCH();
    mstep_div_kernel<<<1, 256, (size_t)(K * K + K) * sizeof(float), st>>>(x, ws_nk, mu, sigma, lamda, pi_out, grad_out, n, K, D);
    MGP_CHECK_LAUNCH();
    return MGP_OK;
}
