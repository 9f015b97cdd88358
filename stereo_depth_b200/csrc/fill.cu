// Kernel D+E: disparity upscale + vertical bilateral fill + horizontal bilateral fill, fused.
//
// Reference: upscale_disparity_vertical_fill.cu:20-51 then horizontal_disparity_fill.cu:19-40
// (two kernels, the second in place).  The horizontal pass only ever reads columns that are
// multiples of K, and those hold exactly what the vertical pass wrote, so every output pixel is a
// pure function of the refined disparity and the left gray image:
//   vf(r, c)  for c % K == 0:  x = r / K, i = r % K, p = K*disp[x][c/K]
//        i == 0            -> p
//        n = K*disp[x-1][c/K];  |p-n| <= thr -> p + (i*(n-p))/K
//        else  cur=GL[r][c]:  |cur-GL[Kx][c]| <= |cur-GL[(K+1)x][c]| ? p : n      (sic: (K+1)*x)
//   out(r, c): m = c % K, nk = c - m, p = vf(r,nk), n = vf(r,nk+K)
//        |p-n| <= thr -> p + (m*(n-p))/K   else colour pick between GL[r][nk], GL[r][nk+K]
// SAFE definitions where the reference is undefined (oracle/stereo_oracle.c so_vfill/so_hfill):
//   rows 1..K-1 of the image (never written by the reference) replicate p;
//   GL[(K+1)x] wraps modulo H;  column nk+K == W reads the next row's column 0 like the
//   reference's flat index does, and p on the last row / when W % K != 0.
// HBM-bound: reads disp [Hd,Wd] and (rarely) gray, writes [H,W]; 4 pixels per thread, STG.128.
#include "common.cuh"

namespace sd {
namespace {

// x / K for the compile-time factors the reference is used with (exact: K is a power of two), IEEE division otherwise.
template <int KT>
__device__ __forceinline__ float div_k(float x, float fk) {
    if (KT == 1) return x;
    if (KT == 2) return __fmul_rn(x, 0.5f);
    if (KT == 4) return __fmul_rn(x, 0.25f);
    return __fdiv_rn(x, fk);
}

// Row `row` of the GLOBAL left gray image (only the rare colour-pick branch gets here).
__device__ __forceinline__ const float *global_gray_row(const GrayView &gv, const float *gl, int row, int W) {
    if (gv.n == 0) return (gv.flat ? gv.flat : gl) + (size_t)row * W;
    // compile-time indices only: the view stays in the kernel's parameter space (no local copy)
    const float *p = gv.band[0];
    int first = gv.row0[0];
#pragma unroll
    for (int q = 1; q < 8; q++)
        if (q < gv.n && row >= gv.row0[q]) {
            p = gv.band[q];
            first = gv.row0[q];
        }
    return p + (size_t)(row - first) * W;
}

// gl = left gray of this handle's (local) image; glg = where the GLOBAL image lives (== gl outside band mode).
template <int KT>
__device__ __forceinline__ float vfill_value(const Geom &g, const float *__restrict__ gl, const GrayView &glg,
                                             const float *__restrict__ disp, int r, int c) {
    const int K = KT > 0 ? KT : g.K, x = r / K, i = r - x * K, yd = c / K;
    const float fk = (float)K;
    const float p = __fmul_rn(fk, __ldg(disp + (size_t)x * g.Wd + yd));
    const int xg = wrapm(x + g.band_x_off, g.Hd_glob);  // global pooled row (the reference's x)
    if (i == 0 || xg == 0 || x == 0) return p;
    const float n = __fmul_rn(fk, __ldg(disp + (size_t)(x - 1) * g.Wd + yd));
    if (fabsf(__fsub_rn(p, n)) <= g.threshold)
        return __fadd_rn(p, div_k<KT>(__fmul_rn((float)i, __fsub_rn(n, p)), fk));
    const float prev_color = __ldg(gl + (size_t)(K * x) * g.W + c);
    const float next_color = __ldg(global_gray_row(glg, gl, wrapm((K + 1) * xg, g.H_glob), g.W) + c);
    const float cur = __ldg(gl + (size_t)r * g.W + c);
    return (fabsf(__fsub_rn(cur, prev_color)) <= fabsf(__fsub_rn(cur, next_color))) ? p : n;
}

template <int KT>
__device__ __forceinline__ float hfill_value(const Geom &g, const float *__restrict__ gl, int r, int c, int nk,
                                             float p, float n) {
    const int m = c - nk, K = KT > 0 ? KT : g.K;
    if (fabsf(__fsub_rn(p, n)) <= g.threshold)
        return __fadd_rn(p, div_k<KT>(__fmul_rn((float)m, __fsub_rn(n, p)), (float)K));
    const size_t o = (size_t)r * g.W;
    const float pc = __ldg(gl + o + nk);
    const size_t cf = o + nk + K;
    const float nc = (cf < (size_t)g.H * g.W) ? __ldg(gl + cf) : pc;
    const float cur = __ldg(gl + o + c);
    return (fabsf(__fsub_rn(cur, pc)) <= fabsf(__fsub_rn(cur, nc))) ? p : n;
}

// value of the "next" mod-K sample to the right of nk on row r
template <int KT>
__device__ __forceinline__ float next_sample(const Geom &g, const float *__restrict__ gl, const GrayView &glg,
                                             const float *__restrict__ disp, int r, int nk, float p) {
    const int K = KT > 0 ? KT : g.K;
    if (nk + K < g.W) return vfill_value<KT>(g, gl, glg, disp, r, nk + K);
    // the reference's flat index lands on column 0 of the next row -- unless this is the GLOBAL last row
    const int rg = wrapm(r + g.band_x_off * K, g.H_glob);
    if (g.W % K == 0 && rg + 1 < g.H_glob && r + 1 < g.H) return vfill_value<KT>(g, gl, glg, disp, r + 1, 0);
    return p;
}

template <int KT>
__global__ void __launch_bounds__(256) fill_kernel(Geom g, const float *__restrict__ gray, const GrayView glg,
                                                   const float *__restrict__ refined, float *__restrict__ out, bool vec_ok) {
    const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int frame = blockIdx.z;
    if (r >= g.H || c4 >= g.W) return;
    const size_t plane = (size_t)g.H * g.W;
    const float *gl = gray + (size_t)frame * 2 * plane;
    const float *disp = refined + (size_t)frame * g.Hd * g.Wd;
    float *o = out + (size_t)frame * plane + (size_t)r * g.W + c4;
    float v[4];
    int nk_cached = -1;
    float p = 0.0f, n = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = c4 + k;
        if (c >= g.W) break;
        const int K = KT > 0 ? KT : g.K;
        const int nk = c - c % K;
        if (nk != nk_cached) {
            p = (nk_cached >= 0 && nk == nk_cached + K) ? n : vfill_value<KT>(g, gl, glg, disp, r, nk);
            n = next_sample<KT>(g, gl, glg, disp, r, nk, p);
            nk_cached = nk;
        }
        v[k] = hfill_value<KT>(g, gl, r, c, nk, p, n);
    }
    if (vec_ok && c4 + 3 < g.W) {
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        for (int k = 0; k < 4 && c4 + k < g.W; k++) o[k] = v[k];
    }
}

// K = 2 fast path: a thread produces a 2-row x 4-column output block (rows 2x, 2x+1) from 3 + 3 refined
// disparities; the rare colour-pick branches fall back to the generic helpers above, so semantics are identical.
__global__ void __launch_bounds__(256) fill_k2_kernel(Geom g, const float *__restrict__ gray, const GrayView glg,
                                                      const float *__restrict__ refined, float *__restrict__ out) {
    const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int x = blockIdx.y * blockDim.y + threadIdx.y;
    const int frame = blockIdx.z;
    if (x >= g.Hd || c4 >= g.W) return;
    const size_t plane = (size_t)g.H * g.W;
    const float *gl = gray + (size_t)frame * 2 * plane;
    const float *disp = refined + (size_t)frame * g.Hd * g.Wd;
    const int y0 = c4 >> 1;
    const bool has3 = c4 + 4 < g.W;
    const float *da = disp + (size_t)x * g.Wd + y0;
    const float a0 = __ldg(da), a1 = __ldg(da + 1), a2 = has3 ? __ldg(da + 2) : 0.0f;
    const int xg = wrapm(x + g.band_x_off, g.Hd_glob);
    const bool first = (xg == 0) || (x == 0);  // the reference returns before the vertical fill for x == 0
    float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
    if (!first) {
        b0 = __ldg(da - g.Wd);
        b1 = __ldg(da - g.Wd + 1);
        b2 = has3 ? __ldg(da - g.Wd + 2) : 0.0f;
    }
    const float thr = g.threshold;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int r = 2 * x + i;
        if (r >= g.H) break;
        auto vf = [&](float d, float dprev, int c) {
            const float p = __fmul_rn(2.0f, d);
            if (i == 0 || first) return p;
            const float n = __fmul_rn(2.0f, dprev);
            if (fabsf(__fsub_rn(p, n)) <= thr) return __fadd_rn(p, __fmul_rn(__fsub_rn(n, p), 0.5f));
            return vfill_value<2>(g, gl, glg, disp, r, c);
        };
        auto hf = [&](float p, float n, int c) {
            if (fabsf(__fsub_rn(p, n)) <= thr) return __fadd_rn(p, __fmul_rn(__fsub_rn(n, p), 0.5f));
            return hfill_value<2>(g, gl, r, c, c - 1, p, n);
        };
        const float v0 = vf(a0, b0, c4), v1 = vf(a1, b1, c4 + 2);
        const float v2 = has3 ? vf(a2, b2, c4 + 4) : next_sample<2>(g, gl, glg, disp, r, c4 + 2, v1);
        *reinterpret_cast<float4 *>(out + (size_t)frame * plane + (size_t)r * g.W + c4) =
            make_float4(v0, hf(v0, v1, c4 + 1), v1, hf(v1, v2, c4 + 3));
    }
}

}  // namespace

cudaError_t launch_fill(const Geom &g, int frames, const Scratch &s, const GrayView &gl_glob, float *out, cudaStream_t st) {
    dim3 block(32, 8), grid(((g.W + 3) / 4 + 31) / 32, (g.H + 7) / 8, frames);
    const bool vec_ok = (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    if (g.K == 2 && vec_ok) {
        dim3 grid2((g.W / 4 + 31) / 32, (g.Hd + 7) / 8, frames);
        fill_k2_kernel<<<grid2, block, 0, st>>>(g, s.gray, gl_glob, s.refined, out);
        return cudaGetLastError();
    }
    switch (g.K) {
        case 1: fill_kernel<1><<<grid, block, 0, st>>>(g, s.gray, gl_glob, s.refined, out, vec_ok); break;
        case 2: fill_kernel<2><<<grid, block, 0, st>>>(g, s.gray, gl_glob, s.refined, out, vec_ok); break;
        case 4: fill_kernel<4><<<grid, block, 0, st>>>(g, s.gray, gl_glob, s.refined, out, vec_ok); break;
        default: fill_kernel<0><<<grid, block, 0, st>>>(g, s.gray, gl_glob, s.refined, out, vec_ok); break;
    }
    return cudaGetLastError();
}

}  // namespace sd
