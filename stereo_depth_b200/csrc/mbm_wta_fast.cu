// Kernel B, specialised variant: fused similarity cost + multi-block aggregation + winner-take-all
// for the reference's default radii (cost 3x3, windows 3x21 / 21x3 / 9x9).  CUDA-core (fp32 add)
// bound: 237 lane-ops per (pixel, disparity) cell, ~0.4 B of HBM traffic per cell -- no tensor cores.
//
// Bit-exactness rules out box filters / sliding windows: every window sum must be the reference's
// sequential fp32 chain (rows outer, columns inner, from 0.0f).  The kernel therefore spends its time
// on 204 dependent-order adds per cell and is organised to issue them at full rate:
//   * two disparity levels per pass, interleaved as float2 -> every add is one packed FADD2
//     (add.rn.f32x2, IEEE-exact per lane) and LDS.128 fetches two cells x two levels;
//   * a thread owns 4x4 pixels x 2 levels: one LDS.128 feeds ~14 packed adds (each cost cell is loaded
//     once per window per thread and reused by every pixel of the 4x4 block whose window covers it);
//   * the cost plane lives in shared memory with its 16-byte chunks split by parity (even chunks first,
//     odd chunks second) so the 32-byte lane stride of the 4-pixel ownership is bank-conflict free;
//   * the pooled left/right row bands (with the reference's circular padding already applied) are staged
//     in shared memory once per tile and reused by all L/2 passes;
//   * the 3x3 cost is computed by row-streaming 4-column strips: every |L-R| tap is evaluated once and
//     reused by the 3 cost rows and up to 3 cost columns that contain it;
//   * WTA state (best, d*, A[d*-1], A[d*+1], previous level) stays in registers; the volume never exists.
//
// Semantics: identical to mbm_wta_generic.cu / oracle so_cost + so_aggregate + so_wta (SAFE padding).
// References: device_functions.cuh:53-73, ncc_matching_cost_volume_construction.cu:67-76,
// multi_block_matching_cost_aggregation.cu:56-87, wta_disparity_selection.cu:22-30.
#include "common.cuh"

namespace sd {
namespace {

constexpr int BW = 64;            // tile width in pixels
constexpr int NCHUNK = 42;        // 16-byte chunks per cost-plane row: (BW + 20) cells x float2 / 16 B
constexpr int HALF = 21;          // even chunks [0,21), odd chunks [21,42)
constexpr int NSTRIP = 21;        // 4-column strips per cost-plane row
constexpr int SEG = 9;            // cost-plane rows per cost-phase work item
constexpr int LW = 88;            // left band row pitch (floats): 64 + 22 rounded up to 4

template <int BH>
struct Cfg {
    static constexpr int NT = (BH / 4) * 16;
    static constexpr int PRW = BH + 20;                        // cost-plane rows
    static constexpr int NSEG = (PRW + SEG - 1) / SEG;
    static constexpr int BR = SEG * NSEG + 2;                  // band rows (incl. rows only dead items touch)
    static constexpr int ITEMS = NSTRIP * NSEG;
};

__host__ __device__ inline int right_band_pitch(int L) {
    const int Lp = (L + 1) & ~1;
    return (Lp + 86 + 3) & ~3;
}

template <int BH>
__host__ __device__ inline size_t smem_bytes(int L) {
    return (size_t)Cfg<BH>::PRW * NCHUNK * 16 + (size_t)Cfg<BH>::BR * (LW + right_band_pitch(L)) * 4;
}

__device__ __forceinline__ float2 lo2(const float4 &q) { return make_float2(q.x, q.y); }
__device__ __forceinline__ float2 hi2(const float4 &q) { return make_float2(q.z, q.w); }
__device__ __forceinline__ float tap(float l, float r) { return __fsub_rn(255.0f, fabsf(__fsub_rn(l, r))); }

// Shared-memory position (in 16 B chunks) of logical chunk q within a cost-plane row.
__device__ __forceinline__ constexpr int chunk_pos(int q) { return (q >> 1) + (q & 1) * HALF; }

template <int BH, bool DBG>
__global__ void __launch_bounds__(Cfg<BH>::NT, (BH <= 32) ? 2 : 1)
mbm_wta_fast_kernel(Geom g, const float *__restrict__ pool, float4 *__restrict__ wta4, float2 *__restrict__ edge2,
                    int RW, float *__restrict__ dbg_cost, float *__restrict__ dbg_agg) {
    using C = Cfg<BH>;
    extern __shared__ float4 smem4[];
    float4 *plane = smem4;                                              // [PRW][42] chunks of (cell,level) pairs
    float *bandL = reinterpret_cast<float *>(plane + C::PRW * NCHUNK);  // [BR][LW]
    float *bandR = bandL + C::BR * LW;                                  // [BR][RW]

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int frame = blockIdx.z, r0 = blockIdx.y * BH, c0 = blockIdx.x * BW;
    const int Hd = g.Hd, Wd = g.Wd, L = g.L;
    const int Lp = (L + 1) & ~1, M = Lp >> 1;
    const size_t np = (size_t)Hd * Wd;
    const float *pl = pool + (size_t)frame * 2 * np, *pr = pl + np;

    // ---- stage the pooled row bands once per tile (circular padding applied here) ----------------
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = C::NT / 32;
        const int originR = c0 - 10 - g.min_ds - Lp;  // virtual column of bandR[.][0]
        for (int rr = warp; rr < C::BR; rr += nwarps) {
            const size_t ro = (size_t)wrapm(r0 - 11 + rr, Hd) * Wd;
            for (int cc = lane; cc < LW; cc += 32) bandL[rr * LW + cc] = __ldg(pl + ro + wrapm(c0 - 11 + cc, Wd));
            for (int cc = lane; cc < RW; cc += 32) bandR[rr * RW + cc] = __ldg(pr + ro + wrapm(originR + cc, Wd));
        }
    }

    // ---- per-pixel WTA state (16 pixels, index k = a*4 + b) ---------------------------------------
    float best[16], am1[16], ap1[16], prev[16];
    int bd[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        best[k] = kFltMin;
        am1[k] = ap1[k] = prev[k] = 0.0f;
        bd[k] = 0;
    }
    const int px0 = r0 + 4 * ty, py0 = c0 + 4 * tx;  // first owned pixel

    // cost-phase work item of this thread
    const bool has_item = tid < C::ITEMS;
    const int strip = tid % NSTRIP, seg = tid / NSTRIP;

    __syncthreads();

    for (int m = 0; m < M; m++) {
        const int d0 = 2 * m;
        // ================= cost phase: plane[R][s] = (cost(d0), cost(d0+1)) ==========================
        if (has_item) {
            const int R0 = seg * SEG;
            const float *bl = bandL + R0 * LW + strip * 4;
            const float *br = bandR + R0 * RW + strip * 4 + (Lp - 2 - d0);
            float2 T[3][6];
            auto taps = [&](int brow, float2(&t)[6]) {
                const float4 l4 = *reinterpret_cast<const float4 *>(bl + brow * LW);
                const float2 l2 = *reinterpret_cast<const float2 *>(bl + brow * LW + 4);
                const float2 ra = *reinterpret_cast<const float2 *>(br + brow * RW);
                const float2 rb = *reinterpret_cast<const float2 *>(br + brow * RW + 2);
                const float2 rc = *reinterpret_cast<const float2 *>(br + brow * RW + 4);
                const float2 rd = *reinterpret_cast<const float2 *>(br + brow * RW + 6);
                const float lv[6] = {l4.x, l4.y, l4.z, l4.w, l2.x, l2.y};
                const float rv[8] = {ra.x, ra.y, rb.x, rb.y, rc.x, rc.y, rd.x, rd.y};
#pragma unroll
                for (int j = 0; j < 6; j++) t[j] = make_float2(tap(lv[j], rv[j + 1]), tap(lv[j], rv[j]));
            };
            taps(0, T[0]);
            taps(1, T[1]);
#pragma unroll
            for (int rr = 0; rr < SEG; rr++) {
                float2(&top)[6] = T[rr % 3];
                float2(&mid)[6] = T[(rr + 1) % 3];
                float2(&bot)[6] = T[(rr + 2) % 3];
                taps(rr + 2, bot);
                float2 c[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 s = top[i];  // 0.0f + x == x: the chain's first add is exact
                    s = __fadd2_rn(s, top[i + 1]);
                    s = __fadd2_rn(s, top[i + 2]);
                    s = __fadd2_rn(s, mid[i]);
                    s = __fadd2_rn(s, mid[i + 1]);
                    s = __fadd2_rn(s, mid[i + 2]);
                    s = __fadd2_rn(s, bot[i]);
                    s = __fadd2_rn(s, bot[i + 1]);
                    s = __fadd2_rn(s, bot[i + 2]);
                    c[i] = s;
                }
                const int R = R0 + rr;
                if (R < C::PRW) {
                    plane[R * NCHUNK + strip] = make_float4(c[0].x, c[0].y, c[1].x, c[1].y);
                    plane[R * NCHUNK + HALF + strip] = make_float4(c[2].x, c[2].y, c[3].x, c[3].y);
                }
            }
        }
        __syncthreads();

        // ================= aggregation phase ==========================================================
        float2 hv[16], acc[16];
        // ---- H: 3 rows x 21 cols.  plane rows 4ty+9 .. 4ty+14, cells 4tx .. 4tx+23 -----------------------
        {
            const float4 *hp = plane + (4 * ty + 9) * NCHUNK + tx;
#pragma unroll
            for (int t = 0; t < 6; t++) {
                float2 v[24];
#pragma unroll
                for (int j = 0; j < 12; j++) {
                    const float4 q = hp[t * NCHUNK + chunk_pos(j)];
                    v[2 * j] = lo2(q);
                    v[2 * j + 1] = hi2(q);
                }
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const int rel = t - 1 - a;  // window row offset of plane row t for pixel row a
                    if (rel < -1 || rel > 1) continue;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
#pragma unroll
                        for (int j = 0; j < 21; j++) {
                            if (rel == -1 && j == 0) acc[a * 4 + b] = v[b];
                            else acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], v[b + j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++) hv[k] = acc[k];
        }
        // ---- V: 21 rows x 3 cols.  plane rows 4ty .. 4ty+23, cells 4tx+8 .. 4tx+15 (uses +9..+14) -----
        {
            const float4 *vp = plane + (4 * ty) * NCHUNK + tx + 2;
            auto vrow = [&](const float4 *p, int amin, int amax, int first_a) {
                // first_a: pixel row whose window starts at this plane row (-1: none)
                const float4 q0 = p[0], q1 = p[HALF], q2 = p[1], q3 = p[HALF + 1];
                const float2 w[8] = {lo2(q0), hi2(q0), lo2(q1), hi2(q1), lo2(q2), hi2(q2), lo2(q3), hi2(q3)};
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (a < amin || a > amax) continue;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        if (a == first_a) acc[a * 4 + b] = w[b + 1];
                        else acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], w[b + 1]);
                        acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], w[b + 2]);
                        acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], w[b + 3]);
                    }
                }
            };
            vrow(vp + 0 * NCHUNK, 0, 0, 0);
            vrow(vp + 1 * NCHUNK, 0, 1, 1);
            vrow(vp + 2 * NCHUNK, 0, 2, 2);
            vrow(vp + 3 * NCHUNK, 0, 3, 3);
            const float4 *p = vp + 4 * NCHUNK;
#pragma unroll 2
            for (int t = 4; t <= 20; t++, p += NCHUNK) vrow(p, 0, 3, -1);
            vrow(vp + 21 * NCHUNK, 1, 3, -1);
            vrow(vp + 22 * NCHUNK, 2, 3, -1);
            vrow(vp + 23 * NCHUNK, 3, 3, -1);
#pragma unroll
            for (int k = 0; k < 16; k++) hv[k] = __fmul2_rn(hv[k], acc[k]);
        }
        // ---- C: 9 rows x 9 cols.  plane rows 4ty+6 .. 4ty+17, cells 4tx+6 .. 4tx+17 ----------------------
        {
            const float4 *cp = plane + (4 * ty + 6) * NCHUNK + tx;
            auto crow = [&](const float4 *p, int amin, int amax, int first_a) {
                // logical chunks 2tx+3 .. 2tx+8
                const float4 q0 = p[HALF + 1], q1 = p[2], q2 = p[HALF + 2], q3 = p[3], q4 = p[HALF + 3], q5 = p[4];
                const float2 u[12] = {lo2(q0), hi2(q0), lo2(q1), hi2(q1), lo2(q2), hi2(q2),
                                      lo2(q3), hi2(q3), lo2(q4), hi2(q4), lo2(q5), hi2(q5)};
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (a < amin || a > amax) continue;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
#pragma unroll
                        for (int j = 0; j < 9; j++) {
                            if (a == first_a && j == 0) acc[a * 4 + b] = u[b];
                            else acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], u[b + j]);
                        }
                    }
                }
            };
            crow(cp + 0 * NCHUNK, 0, 0, 0);
            crow(cp + 1 * NCHUNK, 0, 1, 1);
            crow(cp + 2 * NCHUNK, 0, 2, 2);
            crow(cp + 3 * NCHUNK, 0, 3, 3);
            const float4 *p = cp + 4 * NCHUNK;
#pragma unroll 1
            for (int t = 4; t <= 8; t++, p += NCHUNK) crow(p, 0, 3, -1);
            crow(cp + 9 * NCHUNK, 1, 3, -1);
            crow(cp + 10 * NCHUNK, 2, 3, -1);
            crow(cp + 11 * NCHUNK, 3, 3, -1);
#pragma unroll
            for (int k = 0; k < 16; k++) hv[k] = __fmul2_rn(hv[k], acc[k]);  // (H*V)*C
        }

        // ================= winner-take-all update (ascending d, strict >) =============================
        if (DBG && frame == 0) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int x = px0 + (k >> 2), y = py0 + (k & 3);
                if (x < Hd && y < Wd) {
                    const int s = 4 * tx + (k & 3) + 10, R = 4 * ty + (k >> 2) + 10;
                    const float4 q = plane[R * NCHUNK + chunk_pos(s >> 1)];
                    const float2 cc = (s & 1) ? hi2(q) : lo2(q);
                    const size_t o = ((size_t)x * Wd + y) * L + d0;
                    if (dbg_cost) dbg_cost[o] = cc.x;
                    if (dbg_agg) dbg_agg[o] = hv[k].x;
                    if (d0 + 1 < L) {
                        if (dbg_cost) dbg_cost[o + 1] = cc.y;
                        if (dbg_agg) dbg_agg[o + 1] = hv[k].y;
                    }
                }
            }
        }
        if (m == 0) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int x = px0 + (k >> 2), y = py0 + (k & 3);
                if (x < Hd && y < Wd) edge2[(size_t)frame * np + (size_t)x * Wd + y].x = hv[k].x;  // A[0]
            }
        }
        const bool has2 = (d0 + 1 < L);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float a0 = hv[k].x, a1 = hv[k].y;
            if (bd[k] == d0 - 1) ap1[k] = a0;
            if (a0 > best[k]) {
                best[k] = a0;
                bd[k] = d0;
                am1[k] = prev[k];
            }
            if (bd[k] == d0) ap1[k] = a1;
            if (has2 && a1 > best[k]) {
                best[k] = a1;
                bd[k] = d0 + 1;
                am1[k] = a0;
            }
            prev[k] = has2 ? a1 : a0;
        }
        __syncthreads();  // everyone is done reading the plane before the next pass overwrites it
    }

#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int x = px0 + (k >> 2), y = py0 + (k & 3);
        if (x < Hd && y < Wd) {
            const size_t o = (size_t)frame * np + (size_t)x * Wd + y;
            wta4[o] = make_float4((float)bd[k], am1[k], best[k], ap1[k]);
            edge2[o].y = prev[k];  // A[L-1]
        }
    }
}

template <int BH, bool DBG>
cudaError_t launch_t(const Geom &g, int frames, const Scratch &s, float *dbg_cost, float *dbg_agg, cudaStream_t st) {
    const size_t smem = smem_bytes<BH>(g.L);
    // per-device attribute: set on every launch (cheap) so multi-GPU processes stay correct
    cudaError_t e = cudaFuncSetAttribute(mbm_wta_fast_kernel<BH, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((g.Wd + BW - 1) / BW, (g.Hd + BH - 1) / BH, frames);
    mbm_wta_fast_kernel<BH, DBG><<<grid, Cfg<BH>::NT, smem, st>>>(g, s.pool, s.wta4, s.edge2, right_band_pitch(g.L),
                                                                   dbg_cost, dbg_agg);
    return cudaGetLastError();
}

}  // namespace

bool mbm_wta_fast_supported(const Geom &g) {
    return g.r_cost == 1 && g.rs == 1 && g.rm == 4 && g.rl == 10 && g.L >= 1 && smem_bytes<32>(g.L) <= 227 * 1024;
}

cudaError_t launch_mbm_wta_fast(const Geom &g, int frames, const Scratch &s, float *dbg_cost, float *dbg_agg,
                                cudaStream_t st) {
    if (!mbm_wta_fast_supported(g)) return cudaErrorNotSupported;
    if (dbg_cost || dbg_agg) return launch_t<32, true>(g, frames, s, dbg_cost, dbg_agg, st);
    return launch_t<32, false>(g, frames, s, dbg_cost, dbg_agg, st);
}

}  // namespace sd
