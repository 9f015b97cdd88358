// Consumers of the disparity map that the reference runs in Python after the hot path:
//   * accuracy metrics D1 / Threshold_N / MAE over the masked ground truth
//     (src/python/pipeline/depth_estimation_pipeline_metrics.py:18-56, mask from depth_estimation_pipeline_runner.py:85)
//   * disparity -> depth -> point list (x = column, y = row, z = baseline*focal/disparity), row-major order,
//     invalid disparities skipped (depth_estimation_pipeline_hooks.py:84-92, helpers/point_cloud_helpers.py:5-13 --
//     a Python double loop in the reference).
// Both are single-pass HBM-bound kernels: one read of the disparity map, a few bytes out.
#include "common.cuh"

namespace sd {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// out[0] = masked pixel count, out[1] = D1 outliers, out[2] = |E| > threshold, out[3] = sum |E|   (doubles)
__global__ void __launch_bounds__(256) metrics_kernel(const float *__restrict__ est, const float *__restrict__ gt, size_t n,
                                                      float max_disp, float threshold, double *__restrict__ out) {
    double cnt = 0, d1 = 0, th = 0, sum = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g = __ldg(gt + i);
        if (g <= max_disp && g > 0.0f) {
            const float e = fabsf(__fsub_rn(__ldg(est + i), g));
            cnt += 1.0;
            if (e > 3.0f && __fdiv_rn(e, fabsf(g)) > 0.05f) d1 += 1.0;
            if (e > threshold) th += 1.0;
            sum += (double)e;
        }
    }
    __shared__ double sh[4][8];
    cnt = warp_sum(cnt); d1 = warp_sum(d1); th = warp_sum(th); sum = warp_sum(sum);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = cnt; sh[1][w] = d1; sh[2][w] = th; sh[3][w] = sum; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int k = 0; k < 8; k++) s += sh[threadIdx.x][k];
        atomicAdd(out + threadIdx.x, s);
    }
}

constexpr int PC_BLOCK = 1024;  // pixels per block in the ordered compaction

__global__ void __launch_bounds__(256) pc_count_kernel(const float *__restrict__ disp, size_t n, float invalid,
                                                       int *__restrict__ counts) {
    const size_t base = (size_t)blockIdx.x * PC_BLOCK;
    int c = 0;
    for (int k = threadIdx.x; k < PC_BLOCK; k += 256) {
        const size_t i = base + k;
        if (i < n && __ldg(disp + i) != invalid) c++;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < 8; k++) s += sh[k];
        counts[blockIdx.x] = s;
    }
}

// exclusive scan of the per-block counts by ONE block (n_blocks <= a few thousand); total -> counts[n_blocks]
__global__ void __launch_bounds__(1024) pc_scan_kernel(int *__restrict__ counts, int n_blocks) {
    __shared__ int sh[1024];
    int carry = 0;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? counts[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n_blocks) counts[i] = carry + sh[threadIdx.x] - v;
        carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[n_blocks] = carry;
}

__global__ void __launch_bounds__(256) pc_write_kernel(const float *__restrict__ disp, size_t n, int W, float invalid,
                                                       float fb, const int *__restrict__ offsets, float *__restrict__ xyz) {
    const size_t base = (size_t)blockIdx.x * PC_BLOCK;
    __shared__ int warp_tot[8];
    __shared__ int running;
    if (threadIdx.x == 0) running = offsets[blockIdx.x];
    __syncthreads();
    for (int k0 = 0; k0 < PC_BLOCK; k0 += 256) {       // 256 consecutive pixels per round keeps row-major order
        const size_t i = base + k0 + threadIdx.x;
        const float d = i < n ? __ldg(disp + i) : invalid;
        const bool ok = (i < n) && (d != invalid);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) warp_tot[w] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int k = 0; k < w; k++) before += warp_tot[k];
        if (ok) {
            const size_t o = (size_t)(running + before + __popc(m & ((1u << l) - 1)));
            xyz[3 * o + 0] = (float)(i % W);                 // [y, x, depth] in the reference's naming: column first
            xyz[3 * o + 1] = (float)(i / W);
            xyz[3 * o + 2] = __fdiv_rn(fb, d);               // (baseline * focal_length) / disparity
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 8; k++) t += warp_tot[k];
            running += t;
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace sd

using namespace sd;

extern "C" {

// metrics_out: 4 doubles in DEVICE memory, zeroed by this call: {count, D1 outliers, threshold outliers, sum |E|}
int sd_metrics(const float *disparity, const float *gt_disparity, long long n, float max_disparity, float threshold,
               double *metrics_out, void *stream) {
    if (!disparity || !gt_disparity || !metrics_out || n <= 0) return SD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(metrics_out, 0, 4 * sizeof(double), st) != cudaSuccess) return SD_ERR_CUDA;
    const int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
    metrics_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(disparity, gt_disparity, (size_t)n, max_disparity, threshold,
                                                                  metrics_out);
    return cudaGetLastError() == cudaSuccess ? SD_OK : SD_ERR_CUDA;
}

// xyz: [H*W,3] floats (device), scratch: ceil(H*W/1024)+1 ints (device); the number of points ends up in
// scratch[ceil(H*W/1024)].  Points are in row-major pixel order, like the reference's double loop.
int sd_point_cloud(const float *disparity, int H, int W, float focal_times_baseline, float invalid_disparity, float *xyz,
                   int *scratch, void *stream) {
    if (!disparity || !xyz || !scratch || H <= 0 || W <= 0) return SD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    const int nb = (int)((n + PC_BLOCK - 1) / PC_BLOCK);
    pc_count_kernel<<<nb, 256, 0, st>>>(disparity, n, invalid_disparity, scratch);
    pc_scan_kernel<<<1, 1024, 0, st>>>(scratch, nb);
    pc_write_kernel<<<nb, 256, 0, st>>>(disparity, n, W, invalid_disparity, focal_times_baseline, scratch, xyz);
    return cudaGetLastError() == cudaSuccess ? SD_OK : SD_ERR_CUDA;
}

}  // extern "C"
