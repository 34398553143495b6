// Store-path microbenchmark: how fast can W warps/SM write a [N, P] fp32 matrix with the
// epilogue's access pattern (each warp instruction = 128 B of one row, rows 4*P bytes apart)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

// tile = 256 rows x 128 cols; each warp of a CTA writes 32 cols x (256/ (warps/4)) rows
template <int VEC>
__global__ void store_pattern(float* out, int N, int P, int tiles_n, int tiles_p, int order) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    const int q = warp & 3, g = warp >> 2, groups = nw >> 2;
    const long long total = (long long)tiles_n * tiles_p;
    // order 0: tile t = blockIdx + k*gridDim (neighbouring CTAs write neighbouring column tiles of the same rows)
    // order 1: each CTA owns a contiguous range of tiles (row tile major) -- the GEMM kernel's schedule
    const long long t_begin = order ? total * blockIdx.x / gridDim.x : blockIdx.x;
    const long long t_end = order ? total * (blockIdx.x + 1) / gridDim.x : total;
    const long long t_step = order ? 1 : gridDim.x;
    for (long long t = t_begin; t < t_end; t += t_step) {
        const int tn = (int)(t / tiles_p), tp = (int)(t % tiles_p);
        const int p = tp * 128 + q * 32 + lane;
        if (VEC == 1) {
            for (int r = g; r < 256; r += groups) {
                const int n = tn * 256 + r;
                if (n < N && p < P) out[(size_t)n * P + p] = (float)r;
            }
        } else {
            // 4 rows per lane-quad: lanes 0..7 write 32 cols (float4) of row r, 8..15 row r+1, ...
            const int sub = lane >> 3, c4 = (lane & 7) * 4;
            for (int r = g * 4 + sub; r < 256; r += groups * 4) {
                const int n = tn * 256 + r;
                const int pp = tp * 128 + q * 32 + c4;
                if (n < N && pp + 3 < P) *reinterpret_cast<float4*>(out + (size_t)n * P + pp) = make_float4(1, 2, 3, 4);
            }
        }
    }
}

int main(int argc, char** argv) {
    const int N = 50176, P = argc > 1 ? atoi(argv[1]) : 2000;
    float* out;
    cudaMalloc(&out, (size_t)N * P * 4);
    const int tiles_n = (N + 255) / 256, tiles_p = (P + 127) / 128;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int order = 0; order <= 1; ++order)
    for (int vec = 1; vec <= 4; vec += 3)
        for (int warps = 8; warps <= 32; warps *= 2) {
            for (int it = 0; it < 3; ++it) {
                if (vec == 1) store_pattern<1><<<148, warps * 32>>>(out, N, P, tiles_n, tiles_p, order);
                else store_pattern<4><<<148, warps * 32>>>(out, N, P, tiles_n, tiles_p, order);
            }
            cudaEventRecord(e0);
            const int reps = 10;
            for (int it = 0; it < reps; ++it) {
                if (vec == 1) store_pattern<1><<<148, warps * 32>>>(out, N, P, tiles_n, tiles_p, order);
                else store_pattern<4><<<148, warps * 32>>>(out, N, P, tiles_n, tiles_p, order);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("P=%d order=%d vec=%d warps/SM=%2d : %.1f us  %.2f TB/s  (%s)\n", P, order, vec, warps, ms / reps * 1e3,
                   (double)N * P * 4 / (ms / reps * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
