// Microbenchmark: fp32 add throughput on sm_100a -- scalar FADD vs packed FADD2 (add.f32x2),
// alone and mixed with LDS.128 at the ratio the fused stereo kernel needs.  Gives the measured
// CUDA-core roofline denominator (lane-adds / clk / SM) for kernel B.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void k_fadd(float *out, const float *in, int iters, long long *cyc) {
    float a[NACC], v[4];
    for (int j = 0; j < 4; j++) v[j] = in[j + (threadIdx.x & 7)];
    for (int j = 0; j < NACC; j++) a[j] = in[4 + j] + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < NACC; j++) a[j] = __fadd_rn(a[j], v[(j + r) & 3]);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < NACC; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NACC>  // NACC float2 accumulators
__global__ void k_fadd2(float *out, const float *in, int iters, long long *cyc) {
    float2 a[NACC], v[4];
    for (int j = 0; j < 4; j++) v[j] = make_float2(in[j + (threadIdx.x & 7)], in[j + 1 + (threadIdx.x & 3)]);
    for (int j = 0; j < NACC; j++) a[j] = make_float2(in[4 + j] + threadIdx.x, in[5 + j]);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < NACC; j++) a[j] = __fadd2_rn(a[j], v[(j + r) & 3]);
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < NACC; j++) s += a[j].x + a[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mix: per LDS.128 (4 floats = 2 float2 addends) do ADDS_PER_LDS packed adds (or 2x scalar)
template <int NACC, int ADDS_PER_LDS, bool PACKED>
__global__ void k_mix(float *out, const float *in, int iters, long long *cyc) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float4(in[i & 15], in[(i + 1) & 15], in[(i + 2) & 15], in[(i + 3) & 15]);
    __syncthreads();
    float2 a[NACC];
    for (int j = 0; j < NACC; j++) a[j] = make_float2(in[4 + j] + threadIdx.x, in[5 + j]);
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float4 q = sm[(idx + r * 32) & 2047];
            float2 v0 = make_float2(q.x, q.y), v1 = make_float2(q.z, q.w);
#pragma unroll
            for (int j = 0; j < ADDS_PER_LDS; j++) {
                float2 v = (j & 1) ? v1 : v0;
                if (PACKED) a[j % NACC] = __fadd2_rn(a[j % NACC], v);
                else { a[j % NACC].x = __fadd_rn(a[j % NACC].x, v.x); a[j % NACC].y = __fadd_rn(a[j % NACC].y, v.y); }
            }
        }
        idx += 256;
    }
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < NACC; j++) s += a[j].x + a[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
int run(const char *name, F launch, int blocks, int threads, double lane_ops_per_thread, float *out, long long *cyc) {
    launch();  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < blocks; i++) if (h[i] > mx) mx = h[i];
    delete[] h;
    int sms = 148; double blocks_per_sm = (double)blocks / sms;
    double per_sm_clk = lane_ops_per_thread * threads * blocks_per_sm / mx;
    double total = lane_ops_per_thread * threads * blocks;
    printf("%-44s blocks=%4d thr=%4d  %8.3f ms  %7.2f T lane-adds/s  %6.1f lane-adds/clk/SM (max blk cycles %.0f, eff clk %.0f MHz)\n",
           name, blocks, threads, ms, total / ms * 1e-9, per_sm_clk, mx, mx / ms * 1e-3);
    return 0;
}

int main() {
    float *in, *out; long long *cyc;
    CK(cudaMalloc(&in, 4096 * 4)); CK(cudaMalloc(&out, 148 * 8 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8 * 8));
    float hin[4096]; for (int i = 0; i < 4096; i++) hin[i] = 1.0f + i * 1e-3f;
    CK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
    const int iters = 4000;
    cudaFuncSetAttribute(k_mix<16, 14, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(k_mix<16, 14, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(k_mix<16, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaFuncSetAttribute(k_mix<16, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int wpsm : {4, 8, 16, 32}) {
        int threads = 256, blocks = 148 * (wpsm * 32 / threads > 0 ? wpsm * 32 / threads : 1);
        if (wpsm == 4) { threads = 128; blocks = 148; }
        printf("--- %d warps/SM\n", wpsm);
        run("scalar FADD, 16 chains/thread", [&] { k_fadd<16><<<blocks, threads>>>(out, in, iters, cyc); }, blocks, threads, 16.0 * 8 * iters, out, cyc);
        run("scalar FADD, 32 chains/thread", [&] { k_fadd<32><<<blocks, threads>>>(out, in, iters, cyc); }, blocks, threads, 32.0 * 8 * iters, out, cyc);
        run("packed FADD2, 8 float2 chains/thread", [&] { k_fadd2<8><<<blocks, threads>>>(out, in, iters, cyc); }, blocks, threads, 16.0 * 8 * iters, out, cyc);
        run("packed FADD2, 16 float2 chains/thread", [&] { k_fadd2<16><<<blocks, threads>>>(out, in, iters, cyc); }, blocks, threads, 32.0 * 8 * iters, out, cyc);
        run("FADD2 + LDS.128 (14 add2 per LDS)", [&] { k_mix<16, 14, true><<<blocks, threads, 32768>>>(out, in, iters, cyc); }, blocks, threads, 2.0 * 14 * 8 * iters, out, cyc);
        run("scalar FADD + LDS.128 (28 add per LDS)", [&] { k_mix<16, 14, false><<<blocks, threads, 32768>>>(out, in, iters, cyc); }, blocks, threads, 2.0 * 14 * 8 * iters, out, cyc);
        run("FADD2 + LDS.128 (8 add2 per LDS)", [&] { k_mix<16, 8, true><<<blocks, threads, 32768>>>(out, in, iters, cyc); }, blocks, threads, 2.0 * 8 * 8 * iters, out, cyc);
        run("scalar FADD + LDS.128 (16 add per LDS)", [&] { k_mix<16, 8, false><<<blocks, threads, 32768>>>(out, in, iters, cyc); }, blocks, threads, 2.0 * 8 * 8 * iters, out, cyc);
    }
    return 0;
}
