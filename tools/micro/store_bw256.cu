// Store-path microbenchmark 2: thread = output row, 32 B (st.global.v8.f32) per lane, a warp instruction
// writes 32 rows x 32 B; consecutive instructions walk along the row (the "patch on TMEM lane" epilogue).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void st256(float* q, float a) {
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(q), "f"(a) : "memory");
}

// tile = 128 rows x 256 cols; warp (q = lane quarter, h = column part) writes rows q*32+lane, cols of its part
__global__ void store_rows(float* out, int N, int P, int tiles_n, int tiles_p, int vec) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = blockDim.x >> 5;
    const int q = warp & 3, h = warp >> 2, parts = nw >> 2;
    const int cols_per = 256 / parts;
    const long long total = (long long)tiles_n * tiles_p;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int tn = (int)(t / tiles_p), tp = (int)(t % tiles_p);
        const int n = tn * 128 + q * 32 + lane;
        if (n >= N) continue;
        float* row = out + (size_t)n * P + tp * 256 + h * cols_per;
        const int pmax = P - (tp * 256 + h * cols_per);
        if (vec == 8) {
#pragma unroll 4
            for (int c = 0; c < cols_per; c += 8)
                if (c + 8 <= pmax) st256(row + c, (float)c);
        } else {
#pragma unroll 4
            for (int c = 0; c < cols_per; c += 4)
                if (c + 4 <= pmax) *reinterpret_cast<float4*>(row + c) = make_float4(1, 2, 3, 4);
        }
    }
}

int main(int argc, char** argv) {
    const int N = 50176, P = argc > 1 ? atoi(argv[1]) : 2000;
    float* out;
    cudaMalloc(&out, (size_t)N * P * 4);
    const int tiles_n = (N + 127) / 128, tiles_p = (P + 255) / 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int vec = 4; vec <= 8; vec += 4)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int it = 0; it < 3; ++it) store_rows<<<148, warps * 32>>>(out, N, P, tiles_n, tiles_p, vec);
            cudaEventRecord(e0);
            const int reps = 10;
            for (int it = 0; it < reps; ++it) store_rows<<<148, warps * 32>>>(out, N, P, tiles_n, tiles_p, vec);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("P=%d rows-on-lanes vec=%d warps/SM=%2d : %.1f us  %.2f TB/s  (%s)\n", P, vec, warps,
                   ms / reps * 1e3, (double)N * P * 4 / (ms / reps * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
