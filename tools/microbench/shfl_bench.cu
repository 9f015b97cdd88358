// Microbenchmark: per-SM throughput of LDS.32, SHFL and their mix (do shuffles share the shared-memory crossbar?).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o shfl_bench shfl_bench.cu && ./shfl_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
    const float *p = sm + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        const float *q = p + (it & 7) * 256;
        if (MODE == 0 || MODE == 2) {   // 4 (mode 2) or 8 (mode 0) LDS.32, conflict-free
            float v0, v1, v2, v3;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"((unsigned)__cvta_generic_to_shared(q)));
            asm volatile("ld.shared.f32 %0, [%1+1024];" : "=f"(v1) : "r"((unsigned)__cvta_generic_to_shared(q)));
            asm volatile("ld.shared.f32 %0, [%1+2048];" : "=f"(v2) : "r"((unsigned)__cvta_generic_to_shared(q)));
            asm volatile("ld.shared.f32 %0, [%1+3072];" : "=f"(v3) : "r"((unsigned)__cvta_generic_to_shared(q)));
            a0 += v0; a1 += v1; a2 += v2; a3 += v3;
            if (MODE == 0) {
                float w0, w1, w2, w3;
                asm volatile("ld.shared.f32 %0, [%1+4096];" : "=f"(w0) : "r"((unsigned)__cvta_generic_to_shared(q)));
                asm volatile("ld.shared.f32 %0, [%1+5120];" : "=f"(w1) : "r"((unsigned)__cvta_generic_to_shared(q)));
                asm volatile("ld.shared.f32 %0, [%1+6144];" : "=f"(w2) : "r"((unsigned)__cvta_generic_to_shared(q)));
                asm volatile("ld.shared.f32 %0, [%1+7168];" : "=f"(w3) : "r"((unsigned)__cvta_generic_to_shared(q)));
                a4 += w0; a5 += w1; a6 += w2; a7 += w3;
            }
        }
        if (MODE == 1 || MODE == 2) {   // 4 (mode 2) or 8 (mode 1) shuffles
            a4 += __shfl_down_sync(0xffffffffu, a0, 1);
            a5 += __shfl_up_sync(0xffffffffu, a1, 1);
            a6 += __shfl_down_sync(0xffffffffu, a2, 1);
            a7 += __shfl_up_sync(0xffffffffu, a3, 1);
            if (MODE == 1) {
                a0 += __shfl_down_sync(0xffffffffu, a4, 1);
                a1 += __shfl_up_sync(0xffffffffu, a5, 1);
                a2 += __shfl_down_sync(0xffffffffu, a6, 1);
                a3 += __shfl_up_sync(0xffffffffu, a7, 1);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int MODE>
void run(const char *name, float *out, int sms, double ghz) {
    const int iters = 20000, blocks = sms * 4;
    k<MODE><<<blocks, 256>>>(out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    // warp-level memory-pipe instructions per SM: 4 blocks x 8 warps x iters x 8
    const double inst = 4.0 * 8 * iters * 8, clk = ms * 1e-3 * ghz * 1e9;
    printf("%-28s %8.3f ms  %6.2f clk per warp-instruction per SM\n", name, ms, clk / inst);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 4 * 256 * 4);
    const double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    run<0>("8 x LDS.32", out, p.multiProcessorCount, ghz);
    run<1>("8 x SHFL", out, p.multiProcessorCount, ghz);
    run<2>("4 x LDS.32 + 4 x SHFL", out, p.multiProcessorCount, ghz);
    return 0;
}
