// Kernel B, warp-specialised variant (experimental, sd_set_variant(h, 3)): same arithmetic as mbm_wta_fast.cu, different
// schedule.  One block per SM owns a 64x64 tile: 8 CONSUMER warps run the aggregation + WTA of pass m while 4 PRODUCER
// warps build the cost plane of pass m+1 into the other half of a double-buffered plane; full/empty mbarriers replace
// the two block-wide barriers per pass, `setmaxnreg` moves registers from the producers (96) to the consumers (200; 256*200 + 128*96 fits the 384*168 launch allocation), and
// the 64-row tile cuts the halo recomputation of the cost phase from 2.13x to 1.72x.
// The code of the two roles is the cost phase and the aggregation phase of mbm_wta_fast.cu, unchanged.
#include <type_traits>

#include "common.cuh"
#include "mbm_helpers.cuh"

namespace sd {
namespace {

using namespace mbm;

constexpr int WBH = 64;                 // tile height
constexpr int WPRW = WBH + 20;          // cost-plane rows
constexpr int WSEG = 14;                // cost-plane rows per producer work item
constexpr int WNSEG = WPRW / WSEG;      // 6 (exact)
constexpr int WBR = WSEG * WNSEG + 2;   // band rows
constexpr int WITEMS = NSTRIP * WNSEG;  // 126 work items for 128 producer threads
constexpr int WNCONS = 256, WNPROD = 128, WNT = WNCONS + WNPROD;
constexpr int WPLANE = WPRW * NCHUNK;   // float4 chunks per plane buffer
static_assert(WPRW % WSEG == 0, "segments must tile the plane");

__host__ __device__ inline size_t ws_smem_bytes(int L, int min_ds) {
    return (size_t)2 * WPLANE * 16 + (size_t)WBR * (LW + make_pad_geom(64, 64, L, min_ds).rw) * 4;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1 << 26)) __trap();  // a lost arrival must not hang the GPU
}
// Producers run far ahead of the consumers: back off between polls so the waiting does not eat issue slots.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, unsigned parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(400);
        if (++spins > (1 << 22)) __trap();
    }
}

template <bool STORE>
__global__ void __launch_bounds__(WNT, 1)
mbm_wta_ws_kernel(Geom g, PadGeom pg, const float *__restrict__ padl, const float *__restrict__ padr,
                  float4 *__restrict__ wta4, float2 *__restrict__ edge2, float *__restrict__ agg_planes) {
    extern __shared__ float4 smem4[];
    float4 *planes = smem4;                                          // [2][WPRW][42]
    float *bandL = reinterpret_cast<float *>(planes + 2 * WPLANE);   // [WBR][LW]
    float *bandR = bandL + WBR * LW;                                 // [WBR][RW]
    __shared__ __align__(8) uint64_t band_bar, full_bar[2], empty_bar[2];

    const int tid = threadIdx.x;
    const int frame = blockIdx.z, r0 = blockIdx.y * WBH, c0 = blockIdx.x * BW;
    const int Hd = g.Hd, Wd = g.Wd, L = g.L;
    const int Lp = (L + 1) & ~1, M = Lp >> 1;
    const size_t np = (size_t)Hd * Wd;
    const int RW = pg.rw;

    if (tid == 0) {
        mbar_init(&band_bar, 1);
        mbar_init(&full_bar[0], WNPROD);
        mbar_init(&full_bar[1], WNPROD);
        mbar_init(&empty_bar[0], WNCONS);
        mbar_init(&empty_bar[1], WNCONS);
    }
    __syncthreads();

    if (tid >= WNCONS) {
        // =========================== PRODUCERS: cost planes ===========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        const int ptid = tid - WNCONS;
        if (ptid < 32) {
            if (ptid == 0) mbar_expect_tx(&band_bar, (unsigned)(WBR * (LW + RW) * 4));
            __syncwarp();
            const float *sl = padl + ((size_t)frame * pg.rows + r0) * pg.pwl + c0;
            const float *sr = padr + ((size_t)frame * pg.rows + r0) * pg.pwr + c0;
            for (int rr = ptid; rr < WBR; rr += 32) {
                tma_bulk_g2s(bandL + rr * LW, sl + (size_t)rr * pg.pwl, LW * 4, &band_bar);
                tma_bulk_g2s(bandR + rr * RW, sr + (size_t)rr * pg.pwr, (unsigned)(RW * 4), &band_bar);
            }
        }
        mbar_wait(&band_bar, 0);
        const bool has_item = ptid < WITEMS;
        const int strip = ptid % NSTRIP, seg = ptid / NSTRIP;
        for (int m = 0; m < M; m++) {
            const int d0 = 2 * m, b = m & 1;
            float4 *pl = planes + b * WPLANE;
            mbar_wait_relaxed(&empty_bar[b], ((m >> 1) & 1) ^ 1);  // consumers are done with the previous contents
            auto cost_phase = [&](auto aligned_tag) {
                constexpr bool ALIGNED = decltype(aligned_tag)::value;
            const int R0 = seg * WSEG;
            const float *bl = bandL + R0 * LW + strip * 4 + 4;
            const float *br = bandR + R0 * RW + strip * 4 + (Lp - 2 - d0);
            float2 T[3][6];
            // Band loads are issued one full row ahead of their use through volatile asm (kept in program
            // order by the compiler): the LDS latency then overlaps the previous row's taps and chains.
            struct Raw {
                float4 l4;
                float2 l2;
                float r[8];
            };
            auto load_raw = [&](int brow) {
                Raw w;
                w.l4 = lds128(bl + brow * LW);
                w.l2 = lds64(bl + brow * LW + 4);
                if (ALIGNED) {
                    const float4 q0 = lds128(br + brow * RW), q1 = lds128(br + brow * RW + 4);
                    w.r[0] = q0.x; w.r[1] = q0.y; w.r[2] = q0.z; w.r[3] = q0.w;
                    w.r[4] = q1.x; w.r[5] = q1.y; w.r[6] = q1.z; w.r[7] = q1.w;
                } else {
                    const float2 ra = lds64(br + brow * RW);
                    const float4 q = lds128(br + brow * RW + 2);
                    const float2 rd = lds64(br + brow * RW + 6);
                    w.r[0] = ra.x; w.r[1] = ra.y; w.r[2] = q.x; w.r[3] = q.y;
                    w.r[4] = q.z; w.r[5] = q.w; w.r[6] = rd.x; w.r[7] = rd.y;
                }
                return w;
            };
            auto taps = [&](const Raw &w, float2(&t)[6]) {
                const float lv[6] = {w.l4.x, w.l4.y, w.l4.z, w.l4.w, w.l2.x, w.l2.y};
#pragma unroll
                for (int j = 0; j < 6; j++) t[j] = make_float2(tap(lv[j], w.r[j + 1]), tap(lv[j], w.r[j]));
            };
            Raw raw[2];
            raw[0] = load_raw(0);
            raw[1] = load_raw(1);
            taps(raw[0], T[0]);
            raw[0] = load_raw(2);
            taps(raw[1], T[1]);
#pragma unroll
            for (int rr = 0; rr < WSEG; rr++) {
                float2(&top)[6] = T[rr % 3];
                float2(&mid)[6] = T[(rr + 1) % 3];
                float2(&bot)[6] = T[(rr + 2) % 3];
                if (rr + 1 < WSEG) raw[(rr + 1) & 1] = load_raw(rr + 3);  // next row's band values
                taps(raw[rr & 1], bot);                                  // band row rr + 2
                // four independent 9-tap chains, interleaved; (0.0f + x) + y == x + y exactly
                float2 c[4];
#pragma unroll
                for (int i = 0; i < 4; i++) c[i] = __fadd2_rn(top[i], top[i + 1]);
#pragma unroll
                for (int i = 0; i < 4; i++) c[i] = __fadd2_rn(c[i], top[i + 2]);
#pragma unroll
                for (int q = 0; q < 3; q++)
#pragma unroll
                    for (int i = 0; i < 4; i++) c[i] = __fadd2_rn(c[i], mid[i + q]);
#pragma unroll
                for (int q = 0; q < 3; q++)
#pragma unroll
                    for (int i = 0; i < 4; i++) c[i] = __fadd2_rn(c[i], bot[i + q]);
                const int R = R0 + rr;
                if (R < WPRW) {
                    pl[R * NCHUNK + strip] = make_float4(c[0].x, c[0].y, c[1].x, c[1].y);
                    pl[R * NCHUNK + HALF + strip] = make_float4(c[2].x, c[2].y, c[3].x, c[3].y);
                }
            }
            };
            if (has_item) {
                if (((Lp - 2 - d0) & 3) == 0) cost_phase(std::true_type{});
                else cost_phase(std::false_type{});
            }
            mbar_arrive(&full_bar[b]);  // release: this thread's plane rows are written
        }
    } else {
        // =========================== CONSUMERS: aggregation + WTA =====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int tx = tid & 15, ty = tid >> 4;
        const int px0 = r0 + 4 * ty, py0 = c0 + 4 * tx;  // first owned pixel
        const size_t o00 = (size_t)frame * np + (size_t)px0 * Wd + py0;
        float best[16], prev[16];
        unsigned pend = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const bool valid = (px0 + (k >> 2) < Hd) && (py0 + (k & 3) < Wd);
            best[k] = valid ? kFltMin : __int_as_float(0x7f800000);
            prev[k] = 0.0f;
        }
        for (int m = 0; m < M; m++) {
            const int d0 = 2 * m, b = m & 1;
            const float4 *pl = planes + b * WPLANE;
            mbar_wait(&full_bar[b], (m >> 1) & 1);
            float2 hv[16], acc[16];
            // ---- H: 3 rows x 21 cols.  plane rows 4ty+9 .. 4ty+14, cells 4tx .. 4tx+23 -----------------------
            {
                const float4 *hp = pl + (4 * ty + 9) * NCHUNK + tx;
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
                    for (int j = 0; j < 21; j++) {
#pragma unroll
                        for (int a = 0; a < 4; a++) {
                            const int rel = t - 1 - a;  // window row offset of plane row t for pixel row a
                            if (rel < -1 || rel > 1) continue;
                            if (rel == -1 && j == 0) continue;  // folded into j == 1
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                if (rel == -1 && j == 1) hv[a * 4 + b] = __fadd2_rn(v[b], v[b + 1]);
                                else hv[a * 4 + b] = __fadd2_rn(hv[a * 4 + b], v[b + j]);
                            }
                        }
                    }
                }
            }
            // ---- V: 21 rows x 3 cols.  plane rows 4ty .. 4ty+23, cells 4tx+8 .. 4tx+15 (uses +9..+14) -----
            {
                const float4 *vp = pl + (4 * ty) * NCHUNK + tx + 2;
                auto vrow = [&](const float4 *p, int amin, int amax, int first_a) {
                    // first_a: pixel row whose window starts at this plane row (-1: none)
                    const float4 q0 = p[0], q1 = p[HALF], q2 = p[1], q3 = p[HALF + 1];
                    const float2 w[8] = {lo2(q0), hi2(q0), lo2(q1), hi2(q1), lo2(q2), hi2(q2), lo2(q3), hi2(q3)};
#pragma unroll
                    for (int c3 = 0; c3 < 3; c3++) {
#pragma unroll
                        for (int a = 0; a < 4; a++) {
                            if (a < amin || a > amax) continue;
                            if (a == first_a && c3 == 0) continue;  // folded into c3 == 1
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                if (a == first_a && c3 == 1) acc[a * 4 + b] = __fadd2_rn(w[b + 1], w[b + 2]);
                                else acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], w[b + 1 + c3]);
                            }
                        }
                    }
                };
                vrow(vp + 0 * NCHUNK, 0, 0, 0);
                vrow(vp + 1 * NCHUNK, 0, 1, 1);
                vrow(vp + 2 * NCHUNK, 0, 2, 2);
                vrow(vp + 3 * NCHUNK, 0, 3, 3);
                if (false) {
#pragma unroll
                    for (int t = 4; t <= 20; t++) vrow(vp + t * NCHUNK, 0, 3, -1);
                } else if (true) {
                    // manual software pipeline: the next row's loads are issued before this row's adds
                    auto vload = [&](const float4 *p, float4(&q)[4]) { q[0] = p[0]; q[1] = p[HALF]; q[2] = p[1]; q[3] = p[HALF + 1]; };
                    auto vcomp = [&](const float4(&q)[4]) {
                        const float2 w[8] = {lo2(q[0]), hi2(q[0]), lo2(q[1]), hi2(q[1]), lo2(q[2]), hi2(q[2]), lo2(q[3]), hi2(q[3])};
#pragma unroll
                        for (int c3 = 0; c3 < 3; c3++)
#pragma unroll
                            for (int k = 0; k < 16; k++) acc[k] = __fadd2_rn(acc[k], w[(k & 3) + 1 + c3]);
                    };
                    float4 qa[4], qb[4];
                    const float4 *p = vp + 4 * NCHUNK;
                    vload(p, qa);
#pragma unroll 1
                    for (int t = 4; t < 20; t += 2, p += 2 * NCHUNK) {
                        vload(p + NCHUNK, qb);
                        vcomp(qa);
                        vload(p + 2 * NCHUNK, qa);
                        vcomp(qb);
                    }
                    vcomp(qa);  // row 20
                } else {
                    const float4 *p = vp + 4 * NCHUNK;
#pragma unroll 2
                    for (int t = 4; t <= 20; t++, p += NCHUNK) vrow(p, 0, 3, -1);
                }
                vrow(vp + 21 * NCHUNK, 1, 3, -1);
                vrow(vp + 22 * NCHUNK, 2, 3, -1);
                vrow(vp + 23 * NCHUNK, 3, 3, -1);
#pragma unroll
                for (int k = 0; k < 16; k++) hv[k] = __fmul2_rn(hv[k], acc[k]);
            }
            // ---- C: 9 rows x 9 cols.  plane rows 4ty+6 .. 4ty+17, cells 4tx+6 .. 4tx+17 ----------------------
            {
                const float4 *cp = pl + (4 * ty + 6) * NCHUNK + tx;
                auto crow = [&](const float4 *p, int amin, int amax, int first_a) {
                    // logical chunks 2tx+3 .. 2tx+8
                    const float4 q0 = p[HALF + 1], q1 = p[2], q2 = p[HALF + 2], q3 = p[3], q4 = p[HALF + 3], q5 = p[4];
                    const float2 u[12] = {lo2(q0), hi2(q0), lo2(q1), hi2(q1), lo2(q2), hi2(q2),
                                          lo2(q3), hi2(q3), lo2(q4), hi2(q4), lo2(q5), hi2(q5)};
#pragma unroll
                    for (int j = 0; j < 9; j++) {
#pragma unroll
                        for (int a = 0; a < 4; a++) {
                            if (a < amin || a > amax) continue;
                            if (a == first_a && j == 0) continue;  // folded into j == 1
#pragma unroll
                            for (int b = 0; b < 4; b++) {
                                if (a == first_a && j == 1) acc[a * 4 + b] = __fadd2_rn(u[b], u[b + 1]);
                                else acc[a * 4 + b] = __fadd2_rn(acc[a * 4 + b], u[b + j]);
                            }
                        }
                    }
                };
                crow(cp + 0 * NCHUNK, 0, 0, 0);
                crow(cp + 1 * NCHUNK, 0, 1, 1);
                crow(cp + 2 * NCHUNK, 0, 2, 2);
                crow(cp + 3 * NCHUNK, 0, 3, 3);
                if (0) {
#pragma unroll
                    for (int t = 4; t <= 8; t++) crow(cp + t * NCHUNK, 0, 3, -1);
                } else {
                    const float4 *p = cp + 4 * NCHUNK;
#pragma unroll 1
                    for (int t = 4; t <= 8; t++, p += NCHUNK) crow(p, 0, 3, -1);
                }
                crow(cp + 9 * NCHUNK, 1, 3, -1);
                crow(cp + 10 * NCHUNK, 2, 3, -1);
                crow(cp + 11 * NCHUNK, 3, 3, -1);
#pragma unroll
                for (int k = 0; k < 16; k++) hv[k] = __fmul2_rn(hv[k], acc[k]);  // (H*V)*C
            }


            mbar_arrive(&empty_bar[b]);  // plane buffer b may be refilled
            if (STORE && agg_planes) {
                // reference-compat mode: materialise the aggregated volume, plane-major [F][L][Hd*Wd] so that the 4
                // pixels of a thread row are one coalesced 16-byte store per level
                float *pl0 = agg_planes + ((size_t)frame * L + d0) * np + (size_t)px0 * Wd + py0;
                const bool vec = ((Wd & 3) == 0) && (py0 + 3 < Wd);
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    if (px0 + a >= Hd) break;
                    float *q0 = pl0 + (size_t)a * Wd;
                    if (vec) {
                        *reinterpret_cast<float4 *>(q0) = make_float4(hv[a * 4].x, hv[a * 4 + 1].x, hv[a * 4 + 2].x, hv[a * 4 + 3].x);
                        if (d0 + 1 < L)
                            *reinterpret_cast<float4 *>(q0 + np) =
                                make_float4(hv[a * 4].y, hv[a * 4 + 1].y, hv[a * 4 + 2].y, hv[a * 4 + 3].y);
                    } else {
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (py0 + b < Wd) {
                                q0[b] = hv[a * 4 + b].x;
                                if (d0 + 1 < L) q0[np + b] = hv[a * 4 + b].y;
                            }
                    }
                }
            }
            if (m == 0) {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int x = px0 + (k >> 2), y = py0 + (k & 3);
                    if (x < Hd && y < Wd) {
                        const size_t o = o00 + (size_t)(k >> 2) * Wd + (k & 3);
                        edge2[o].x = hv[k].x;                                    // A[0]
                        wta4[o] = make_float4(0.0f, 0.0f, hv[k].x, hv[k].y);      // record if nothing ever beats FLT_MIN
                    }
                }
            }
            const bool has2 = (d0 + 1 < L);
            const float fd0 = (float)d0, fd1 = (float)(d0 + 1);
            unsigned npend = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float a0 = hv[k].x, a1 = hv[k].y;
                float4 *rec = wta4 + o00 + (size_t)(k >> 2) * Wd + (k & 3);
                if (pend & (1u << k)) rec->w = a0;  // A[d*+1] for a maximum found at d0-1
                if (a0 > best[k]) {
                    best[k] = a0;
                    *rec = make_float4(fd0, prev[k], a0, a1);
                }
                if (has2 && a1 > best[k]) {
                    best[k] = a1;
                    *rec = make_float4(fd1, a0, a1, 0.0f);
                    npend |= 1u << k;
                }
                prev[k] = has2 ? a1 : a0;
            }
            pend = npend;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int x = px0 + (k >> 2), y = py0 + (k & 3);
            if (x < Hd && y < Wd) edge2[o00 + (size_t)(k >> 2) * Wd + (k & 3)].y = prev[k];  // A[L-1]
        }
    }
}

}  // namespace

bool mbm_wta_ws_supported(const Geom &g) {
    return g.r_cost == 1 && g.rs == 1 && g.rm == 4 && g.rl == 10 && g.L >= 1 && ws_smem_bytes(g.L, g.min_ds) + 64 <= 227 * 1024;
}

cudaError_t launch_mbm_wta_ws(const Geom &g, int frames, const Scratch &s, cudaStream_t st) {
    if (!mbm_wta_ws_supported(g) || !s.padl || !s.padr) return cudaErrorNotSupported;
    cudaError_t e = launch_pad_pooled(g, frames, s, st);
    if (e != cudaSuccess) return e;
    const size_t smem = ws_smem_bytes(g.L, g.min_ds);
    const PadGeom pg = make_pad_geom(g.Hd, g.Wd, g.L, g.min_ds);
    dim3 grid((g.Wd + BW - 1) / BW, (g.Hd + WBH - 1) / WBH, frames);
    if (s.agg_vol) {
        e = cudaFuncSetAttribute(mbm_wta_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mbm_wta_ws_kernel<true><<<grid, WNT, smem, st>>>(g, pg, s.padl, s.padr, s.wta4, s.edge2, s.agg_vol);
    } else {
        e = cudaFuncSetAttribute(mbm_wta_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        mbm_wta_ws_kernel<false><<<grid, WNT, smem, st>>>(g, pg, s.padl, s.padr, s.wta4, s.edge2, s.agg_vol);
    }
    return cudaGetLastError();
}

}  // namespace sd
