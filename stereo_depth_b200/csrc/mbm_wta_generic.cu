// Kernel B, generic variant: fused similarity cost + multi-block aggregation + winner-take-all for
// ANY radii / image size.  The specialised variant (mbm_wta_fast.cu) covers the reference's default
// radii; this one is the always-correct path and the parity anchor for it.
//
// Never materialises the [Hd,Wd,L] volumes the reference round-trips through HBM
// (buffer/device_buffer.cc:9-10): per disparity level the block builds the cost plane of its tile
// (+ large_radius halo) in shared memory, aggregates it, and folds it into the running arg-max.
//
// Reference semantics (all fp32, order-faithful; see oracle/stereo_oracle.c):
//   cost(x,y,d) = chain from 0.0f over i,j in [-r,r] (rows outer) of 255 - |PL[x+i][y+j] - PR[x+i][y+j-disp]|
//                 device_functions.cuh:53-73, ncc_matching_cost_volume_construction.cu:67-76
//   Hs/Vs/Cs    = chains from 0.0f over (2rs+1)x(2rl+1), (2rl+1)x(2rs+1), (2rm+1)^2 windows, rows outer
//   agg         = (Hs*Vs)*Cs              multi_block_matching_cost_aggregation.cu:56-87
//   d*          = first arg-max over ascending d with strict >, best initialised to FLT_MIN
//                 wta_disparity_selection.cu:22-30
// Out-of-range indices use true modulo (SAFE definition of pad_index, common.cuh).
#include "common.cuh"

namespace sd {
namespace {

constexpr int TH = 16, TW = 32;  // pixels per block (one thread each)

__device__ __forceinline__ float cost_cell(const float *__restrict__ pl, const float *__restrict__ pr, int Hd, int Wd,
                                           int x, int y, int disp, int r) {
    float c = 0.0f;
    for (int i = -r; i <= r; i++) {
        const int xi = wrapm(x + i, Hd);
        const float *lrow = pl + (size_t)xi * Wd, *rrow = pr + (size_t)xi * Wd;
        for (int j = -r; j <= r; j++) {
            const float l = __ldg(lrow + wrapm(y + j, Wd));
            const float rr = __ldg(rrow + wrapm(y + j - disp, Wd));
            c = __fadd_rn(c, __fsub_rn(255.0f, fabsf(__fsub_rn(l, rr))));
        }
    }
    return c;
}

__global__ void __launch_bounds__(TH *TW) mbm_wta_generic_kernel(Geom g, const float *__restrict__ pool,
                                                                  float4 *__restrict__ wta4, float2 *__restrict__ edge2,
                                                                  float *__restrict__ dbg_cost, float *__restrict__ dbg_agg,
                                                                  float *__restrict__ agg_planes) {
    extern __shared__ float plane[];
    const int PR = TH + 2 * g.rl, PC = TW + 2 * g.rl;
    const int frame = blockIdx.z;
    const int r0 = blockIdx.y * TH, c0 = blockIdx.x * TW;
    const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
    const int x = r0 + ty, y = c0 + tx;
    const bool valid = (x < g.Hd) && (y < g.Wd);
    const size_t np = (size_t)g.Hd * g.Wd;
    const float *pl = pool + (size_t)frame * 2 * np, *pr = pl + np;

    float best = kFltMin, prev = 0.0f, am1 = 0.0f, ap1 = 0.0f, a0 = 0.0f;
    int bd = 0;
    for (int d = 0; d < g.L; d++) {
        const int disp = g.min_ds + d;
        for (int idx = threadIdx.x; idx < PR * PC; idx += TH * TW) {
            const int u = idx / PC, v = idx - u * PC;
            // cost cell the aggregation window sees at virtual (r0-rl+u, c0-rl+v)
            plane[idx] = cost_cell(pl, pr, g.Hd, g.Wd, wrapm(r0 - g.rl + u, g.Hd), wrapm(c0 - g.rl + v, g.Wd), disp, g.r_cost);
        }
        __syncthreads();
        if (valid) {
            const float *ctr = plane + (ty + g.rl) * PC + tx + g.rl;
            float hs = 0.0f, vs = 0.0f, cs = 0.0f;
            for (int i = -g.rs; i <= g.rs; i++)
                for (int j = -g.rl; j <= g.rl; j++) hs = __fadd_rn(hs, ctr[i * PC + j]);
            for (int i = -g.rl; i <= g.rl; i++)
                for (int j = -g.rs; j <= g.rs; j++) vs = __fadd_rn(vs, ctr[i * PC + j]);
            for (int i = -g.rm; i <= g.rm; i++)
                for (int j = -g.rm; j <= g.rm; j++) cs = __fadd_rn(cs, ctr[i * PC + j]);
            const float agg = __fmul_rn(__fmul_rn(hs, vs), cs);
            if (frame == 0) {  // debug volumes, reference layout [Hd][Wd][L]
                const size_t o = ((size_t)x * g.Wd + y) * g.L + d;
                if (dbg_cost) dbg_cost[o] = ctr[0];
                if (dbg_agg) dbg_agg[o] = agg;
            }
            if (agg_planes) agg_planes[((size_t)frame * g.L + d) * np + (size_t)x * g.Wd + y] = agg;  // compat mode
            if (d == 0) a0 = agg;
            if (d == bd + 1) ap1 = agg;  // before bd moves: value right after the current best
            if (agg > best) {
                best = agg;
                bd = d;
                am1 = prev;
            }
            prev = agg;
        }
        __syncthreads();
    }
    if (valid) {
        const size_t o = (size_t)frame * np + (size_t)x * g.Wd + y;
        wta4[o] = make_float4((float)bd, am1, best, ap1);
        edge2[o] = make_float2(a0, prev);
    }
}

}  // namespace

cudaError_t launch_mbm_wta_generic(const Geom &g, int frames, const Scratch &s, float *dbg_cost, float *dbg_agg,
                                   cudaStream_t st) {
    const size_t smem = (size_t)(TH + 2 * g.rl) * (TW + 2 * g.rl) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(mbm_wta_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((g.Wd + TW - 1) / TW, (g.Hd + TH - 1) / TH, frames);
    mbm_wta_generic_kernel<<<grid, TH * TW, smem, st>>>(g, s.pool, s.wta4, s.edge2, dbg_cost, dbg_agg, s.agg_vol);
    return cudaGetLastError();
}

}  // namespace sd
