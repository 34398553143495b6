// TMA-store microbenchmark for the [N,P] epilogue of logprob_tc_kernel: W warps per CTA (one CTA per SM) stage
// [ROWS x 32] fp32 blocks in shared memory and hand them to cp.async.bulk.tensor.2d stores, with DEPTH staging
// buffers per warp (the kernel ships DEPTH = 1) and the kernel's team schedule (TEAM CTAs on adjacent prototype
// tiles of the same patch tile).  No MMAs, no TMEM: the ceiling of the store stream alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_store_bw tma_store_bw.cu && ./tma_store_bw [P]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int DEPTH, int ROWS>
__global__ void __launch_bounds__(256, 1)
tma_store_kernel(const __grid_constant__ CUtensorMap map, int n_tiles_n, int n_tiles_p, int team) {
    extern __shared__ __align__(1024) float stage[];          // [8 warps][DEPTH][ROWS][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, h = warp >> 2;                    // prototype quarter of the 128-wide tile, patch half
    const int n_teams = gridDim.x / team, tm = blockIdx.x / team, k0 = blockIdx.x % team;
    float* my = stage + warp * DEPTH * ROWS * 32;
    int buf = 0;
    for (int nt = tm; nt < n_tiles_n; nt += n_teams) {        // 128 patches per tile
        for (int pt = k0; pt < n_tiles_p; pt += team) {       // 128 prototypes per tile
            for (int ch = h * (64 / ROWS); ch < (h + 1) * (64 / ROWS); ++ch) {      // this warp's row chunks
                float* s = my + buf * ROWS * 32;
                if (lane == 0) {
                    if (DEPTH == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < ROWS; ++j) s[j * 32 + lane] = (float)(j + lane);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    const int c0 = pt * 128 + q * 32, c1 = nt * 128 + ch * ROWS;
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map),
                                 "r"(smem_u32(s)), "r"(c0), "r"(c1)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                buf = (buf + 1) % DEPTH;
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int DEPTH, int ROWS>
static float run(const CUtensorMap& map, int tn, int tp, int team, int reps) {
    const size_t smem = (size_t)8 * DEPTH * ROWS * 32 * 4;
    cudaFuncSetAttribute(tma_store_kernel<DEPTH, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = (148 / team) * team;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) tma_store_kernel<DEPTH, ROWS><<<grid, 256, smem>>>(map, tn, tp, team);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) tma_store_kernel<DEPTH, ROWS><<<grid, 256, smem>>>(map, tn, tp, team);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int N = 50176, P = argc > 1 ? atoi(argv[1]) : 2000;
    float* out;
    cudaMalloc(&out, (size_t)N * P * 4);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) != cudaSuccess || !sym) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
    const int tn = (N + 127) / 128, tp = (P + 127) / 128;
    for (int rows = 32; rows <= 64; rows *= 2) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)P * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)rows};
        cuuint32_t es[2] = {1, 1};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("encode failed\n");
            return 1;
        }
        for (int team = 1; team <= 4; team *= 2) {
            float a, b, c;
            if (rows == 32) { a = run<1, 32>(map, tn, tp, team, 10); b = run<2, 32>(map, tn, tp, team, 10); c = run<4, 32>(map, tn, tp, team, 10); }
            else            { a = run<1, 64>(map, tn, tp, team, 10); b = run<2, 64>(map, tn, tp, team, 10); c = run<3, 64>(map, tn, tp, team, 10); }   // 4 x 64-row buffers would not fit in 227 KB
            const double gb = (double)N * P * 4 / 1e9;
            printf("box %2d x 32  team %d : depth1 %.1f us %.2f TB/s | depth2 %.1f us %.2f TB/s | depth%d %.1f us %.2f TB/s\n", rows,
                   team, a * 1e3, gb / a, b * 1e3, gb / b, rows == 32 ? 4 : 3, c * 1e3, gb / c);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
