// Kernel S: certified level screen for the specialised fused kernel (mbm_wta_fast.cu).
//
// The fused kernel must evaluate every window sum as the reference's sequential fp32 chain (204 adds per
// (pixel, level) cell) because the arg-max is sensitive to the summation order.  But only the levels that can
// still BE the arg-max need that treatment.  This kernel computes every aggregated cost approximately -- the same
// fp32 differences l-r as the reference, |l-r| summed separably (per tap column: 3-row sum -> nested vertical 3/9/21-row
// sums; then sliding horizontal 21/9/3-column sums with the 3x3 cost's horizontal 3-tap folded in; ~45 lane-ops per cell
// instead of 237) -- and keeps, per pixel, the set
// of levels whose approximate cost is within a RIGOROUS error bound of the pixel's approximate maximum.  The union over
// a 32x64 tile of those levels and their two neighbours (the secondary matching reads A[d*-1], A[d*+1]; circular,
// secondary_matching.cu:28-31), as level pairs, is the tile's pass mask: mbm_wta_fast_kernel then runs its exact
// passes only for those pairs.  Results are bit-identical to evaluating all levels (tests/test_gpu_parity.py runs both
// ways).  The bound:
//
//   * The reference's tap is t = fl(255 - |a|), a = fl(l - r) (device_functions.cuh:66-70).  The screen forms the same
//     a and sums |a| (dissimilarities); N*255 - sum|a| differs from the real sum of the reference's taps by at most
//     N*255*u (u = 2^-24).  For pooled values in [0, 255] all terms are >= 0, so every fp32 summation order has a
//     relative error below (#additions on the longest path) * u: the reference's chains stay within 89 u S_max of the
//     real sum, the screen (10 nested vertical adds, at most 39 sliding-window operations + 2 for the 3-tap, one final
//     subtraction) within 52 u S_max, with S_max = 63*9*255 resp. 81*9*255: together below 1.6 in absolute terms; the
//     analysis uses E = 4.
//   * A pixel whose approximate maximum A'max is >= T = 2^15 * 144600 * 185910 has all three sums of that level
//     >= F = 2^15, hence A_ref[max] >= A'max (1 - 3.7e-4).  A level with A' < (1 - 2e-3) A'max has
//     A_ref <= (H'+E)(V'+E)(C'+E)(1+2u) < A'max (1 - 3.7e-4)  (all sums >= F/2: factor (1 + E/(F/2))^3 = 1.00073;
//     else the product is < 4.41e14 < T/2) -- it cannot be the reference's arg-max and is dropped.
//   * The per-pixel set is maintained on the fly: a level enters when A' >= kKeep * running max; the set is cleared
//     when a new value exceeds kClear * running max (then everything seen so far is below (1-eps) of the final
//     maximum, kClear - 1 >= eps / (1 - eps)).  It is therefore always a superset of {A' >= kKeep * final max}.
//   * Pixels with A'max < T (mean tap < 58 of 255: not seen on image data) flag every pair of their tile;
//     out-of-range float inputs (pooled value outside [0, 255] or NaN) are detected by pad_pooled_kernel and
//     make the fused kernel ignore the masks altogether.
//
// Layout: one block of 128 threads per 32x64 tile (the fused kernel's tile), two blocks per SM (110 KB each: the
// TMA-staged row bands + three row-sum buffers).  The block sees every level pair of its tile in ascending order, so
// its running maxima settle early and the candidate bookkeeping goes quiet.  For L > 64 the right band is staged one
// window of 32 level pairs at a time.  Phase A: thread = one of the 86 TAP columns (each |l-r| is formed exactly once),
// walks the 54 band rows as a software pipeline with all running sums in registers (all-positive nested sums: 3-row sum
// -> Y3 -> Z9 -> W21).  Phase B: thread = (row, 16 columns): sliding sums along the row with the cost's horizontal 3-tap
// folded in (box filters commute), similarities, product, candidate bookkeeping.  At the end thread 0
// writes the tile's mask, its cost class (heaviest-first schedule of the fused kernel) and the statistics that feed the
// adaptive policy in api.cu.
#include "common.cuh"
#include "mbm_helpers.cuh"

namespace sd {
namespace {

using namespace mbm;

constexpr int SXW = 86;         // TAP columns per tile (64 + 2*10 cost columns + 1 tap column on either side)
constexpr int SBR = 54;         // band rows used (32 + 2*10 + 2*1)
// Row pitches of the three row-sum buffers in float2 cells.  Each is 16*odd bytes modulo 128, so the LDS.128 of eight
// consecutive rows (phase B: lane = row) hit eight different 16-byte bank groups.  The buffers hold VERTICAL sums of
// single tap columns (the horizontal 3-tap of the 3x3 cost is applied in phase B, where horizontal sums are cheap):
// Y3 for all 86 tap columns, Z9 for tap columns 6..79 and W21 for tap columns 9..76 (all that phase B reads).
constexpr int SPY = 86, SPZ = 74, SPW = 70;
constexpr int SZ0 = 6, SZ1 = 80, SW0 = 9, SW1 = 77;
constexpr int SBUF = 32 * (SPY + SPZ + SPW);   // cells per group
constexpr int SGT = 128;                       // threads per block
constexpr float kKeep = 0.998f;             // candidate:  A' >= kKeep * running max        (eps  = 2e-3)
constexpr float kClear = 1.005f;            // new max > kClear * old max clears the set    (eps' = 5e-3 >= eps/(1-eps))
constexpr float kMinMax = 32768.0f * 144600.0f * 185910.0f;   // T

constexpr float kHVmax = 63.0f * 9.0f * 255.0f, kCmax = 81.0f * 9.0f * 255.0f;   // window sums of 255 per tap

__host__ __device__ inline size_t screen_smem_bytes(int L, int min_ds, int groups) {
    return (size_t)groups * SBUF * sizeof(float2) + (size_t)SBR * (LW + make_pad_geom(64, 64, L, min_ds).rw) * 4;
}

__host__ __device__ inline size_t screen_smem_bytes_windowed() {
    return (size_t)SBUF * sizeof(float2) + (size_t)SBR * (LW + 152) * 4;
}

__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }

// ---- phase A: TAP column t of level pair (d0, d0+1); Y3 / Z9 / W21 = 3 / 9 / 21-row sums of the column's 3-row sums ----
// Band row rr holds image row r0-11+rr; plane row R = rr-2 (image row r0-10+R) is complete at step rr.
// Tap column t is cost column t-1 (cost column s uses tap columns s, s+1, s+2): left band column t+4, right band column
// t-1+Lp-d for level d (same indexing as the cost phase of mbm_wta_fast.cu).  Lane .x = level d0, .y = level d0+1.
// The 3x3 cost of the reference is a box sum of the tap field 255-|l-r|, and every window is a box sum of costs, so the
// horizontal 3-tap commutes with everything done here: each |l-r| is formed ONCE (by the thread of its column) and summed
// vertically; phase B applies the horizontal 3-tap together with its sliding window sums.
__device__ __forceinline__ void screen_phase_a(const float *__restrict__ pl, const float *__restrict__ pr, int RW,
                                               float2 *__restrict__ bY, float2 *__restrict__ bZ, float2 *__restrict__ bW,
                                               bool keep_z, bool keep_w) {
    // Software pipeline with one dependent operation per stage: iteration `it` runs stage E_k on band row it-k-1, and
    // the stages are listed last-first, so every instruction of an iteration only reads results of EARLIER
    // iterations (the loop is fully unrolled: all ring indices are compile-time, rings are just names).  Pair sums
    // (p, q, zb, wb) are formed as soon as their operands exist, which leaves one dependent add per stage:
    //   E0 loads | E1 a = l-r | E2 p, X | E3 q, Y3 | E4 zb, Z9, wb | E5 W21
    // The screen sums DISSIMILARITIES |l-r| (the abs is a free operand modifier) and phase B turns the window sums
    // into similarities: N*255 - sum|a| differs from the sum of the reference's taps fl(255-|a|) by at most N*255*u.
    //   X(R)   = |a|(R) + |a|(R+1) + |a|(R+2)           3-row sum of this tap column at plane row R (band rows R..R+2)
    //   Y3(c)  = X(c-1) + X(c) + X(c+1)      c in [1,50]
    //   Z9(c)  = Y3(c-3) + Y3(c) + Y3(c+3)   c in [4,47]
    //   W21(c) = Z9(c-6) + Y3(c) + Z9(c+6)   c in [10,41]   (the 32 output rows of the tile are plane rows 10..41)
    constexpr int PF = 3;    // the band loads of a row are issued PF iterations before their first use
    float rw[8][3];          // raw band values: l | r(level d0+1) r(level d0)
    float2 a[4];             // l - r per level
    float2 p[4], X[4], q[4];
    float2 Y[16], Z[16], zb[16], wb[16];
#pragma unroll
    for (int it = 0; it < SBR + PF + 5; it++) {
        {   // E5: row i delivered Z9(i-6) in the previous iteration -> W21(i-12)
            const int i = it - PF - 4, cz = i - 6, cw = cz - 6;
            if (i >= 0 && i < SBR && cw >= 10 && cw <= 41 && keep_w) bW[(cw - 10) * SPW] = add2(wb[cw % 16], Z[cz % 16]);
        }
        {   // E4: row i delivered Y3(i-3) in the previous iteration
            const int i = it - PF - 3, cy = i - 3;
            if (i >= 0 && i < SBR && cy >= 1) {
                if (cy >= 7) {
                    const int cz = cy - 3;
                    Z[cz % 16] = add2(zb[cz % 16], Y[cy % 16]);
                    if (cz >= 10 && cz <= 41 && keep_z) bZ[(cz - 10) * SPZ] = Z[cz % 16];
                }
                if (cy >= 4) zb[cy % 16] = add2(Y[(cy - 3) % 16], Y[cy % 16]);
                if (cy >= 10 && cy <= 41) wb[cy % 16] = add2(Z[(cy - 6) % 16], Y[cy % 16]);   // Z9(cy-6) is 3 iterations old
            }
        }
        {   // E3: row i delivered X(i-2) in the previous iteration
            const int i = it - PF - 2, R = i - 2;
            if (i >= 0 && i < SBR && R >= 0) {
                if (R >= 2) {
                    const int cy = R - 1;
                    Y[cy % 16] = add2(q[(R - 1) % 4], X[R % 4]);
                    if (cy >= 10 && cy <= 41) bY[(cy - 10) * SPY] = Y[cy % 16];
                }
                if (R >= 1) q[R % 4] = add2(X[(R - 1) % 4], X[R % 4]);
            }
        }
        {   // E2: a(i) is one iteration old; scalar adds, |.| is an operand modifier
            const int i = it - PF - 1;
            if (i >= 0 && i < SBR) {
                const float2 ai = a[i % 4];
                if (i >= 2) X[(i - 2) % 4] = make_float2(__fadd_rn(p[(i - 1) % 4].x, fabsf(ai.x)), __fadd_rn(p[(i - 1) % 4].y, fabsf(ai.y)));
                if (i >= 1) {
                    const float2 am = a[(i - 1) % 4];
                    p[i % 4] = make_float2(__fadd_rn(fabsf(am.x), fabsf(ai.x)), __fadd_rn(fabsf(am.y), fabsf(ai.y)));
                }
            }
        }
        {   // E1: the loads were issued PF iterations ago
            const int i = it - PF;
            if (i >= 0 && i < SBR) {
                const float *w = rw[i % 8];
                a[i % 4] = make_float2(__fsub_rn(w[0], w[2]), __fsub_rn(w[0], w[1]));
            }
        }
        if (it < SBR) {   // E0 (volatile asm keeps the loads where they are written, PF iterations ahead of their use)
            float *w = rw[it % 8];
            const unsigned la = smem_u32(pl + it * LW), ra = smem_u32(pr);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[0]) : "r"(la));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[1]) : "r"(ra));
            asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(w[2]) : "r"(ra));
            pr += RW;
        }
    }
}

// ---- candidate-set words ----------------------------------------------------------------------------------------------
struct Mask128 {
    unsigned long long lo, hi;
};
template <typename T> __device__ __forceinline__ T m_zero() { return (T)0; }
template <> __device__ __forceinline__ Mask128 m_zero<Mask128>() { return Mask128{0ull, 0ull}; }
template <typename T> __device__ __forceinline__ T m_bit(int idx) { return (T)1 << idx; }
template <> __device__ __forceinline__ Mask128 m_bit<Mask128>(int idx) {
    return Mask128{idx < 64 ? 1ull << idx : 0ull, idx >= 64 ? 1ull << (idx - 64) : 0ull};
}
template <typename T> __device__ __forceinline__ void m_or(T &a, const T &b) { a |= b; }
__device__ __forceinline__ void m_or(Mask128 &a, const Mask128 &b) { a.lo |= b.lo; a.hi |= b.hi; }
template <typename T> __device__ __forceinline__ bool m_any(const T &a) { return a != 0; }
__device__ __forceinline__ bool m_any(const Mask128 &a) { return (a.lo | a.hi) != 0ull; }
template <typename T> __device__ __forceinline__ bool m_test(const T &a, int idx) { return ((a >> idx) & 1) != 0; }
__device__ __forceinline__ bool m_test(const Mask128 &a, int idx) { return (((idx < 64 ? a.lo : a.hi) >> (idx & 63)) & 1ull) != 0ull; }
__device__ __forceinline__ unsigned warp_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned long long warp_or(unsigned long long v) {
    return ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(v >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (unsigned)v);
}
__device__ __forceinline__ Mask128 warp_or(const Mask128 &v) { return Mask128{warp_or(v.lo), warp_or(v.hi)}; }
__device__ __forceinline__ void shared_or(unsigned *p, unsigned v) { atomicOr(p, v); }
__device__ __forceinline__ void shared_or(unsigned long long *p, unsigned long long v) { atomicOr(p, v); }
__device__ __forceinline__ void shared_or(Mask128 *p, const Mask128 &v) {
    if (v.lo) atomicOr(&p->lo, v.lo);
    if (v.hi) atomicOr(&p->hi, v.hi);
}

constexpr int SWIN = 152;       // right-band window (floats) of the windowed schedule: 32 level pairs + 84 cost columns + taps
constexpr int SWPASS = 32;      // level pairs per window

// MaskT: candidate set of one pixel, two bits per level pair (bit 2m = level 2m, bit 2m+1 = level 2m+1): 64 bits for
// L <= 64, 128 bits up to L = 128.
// WIN: the right row band is staged one window of SWPASS level pairs at a time (SWIN columns) instead of whole, so the
// shared-memory footprint does not grow with L and two blocks per SM fit up to L = 128.
// (An earlier version ran several groups of 128 threads per block on disjoint level pairs and merged their sets at the
// end; one group per block and two blocks per SM was faster: the running maxima see every pair and settle early.)
// DBG: additionally store the approximate aggregated costs A' of frame 0 into dbg_vol ([Hd][Wd][L], d innermost) -- the
// parity tests check the kernel's own sums against the bound the header claims (sd_set_debug_screen).
template <typename MaskT, bool WIN, bool DBG>
__global__ void __launch_bounds__(SGT, 2)
mbm_screen_kernel(Geom g, PadGeom pg, const float *__restrict__ padl, const float *__restrict__ padr,
                  unsigned *__restrict__ pass_mask, unsigned long long *__restrict__ stats, int *__restrict__ tile_order,
                  int *__restrict__ bucket_count, unsigned long long *__restrict__ host_word, int epoch,
                  float *__restrict__ dbg_vol) {
    extern __shared__ float4 smem4[];
    float2 *bufs = reinterpret_cast<float2 *>(smem4);                   // [Y3 | Z9 | W21]
    float *bandL = reinterpret_cast<float *>(bufs + SBUF);              // [SBR][LW]
    float *bandR = bandL + SBR * LW;                                    // [SBR][RW]
    __shared__ __align__(8) uint64_t band_bar;
    __shared__ MaskT s_mask;
    __shared__ int s_all;

    const int tid = threadIdx.x, gt = tid;
    const int frame = blockIdx.z, r0 = blockIdx.y * kTileH, c0 = blockIdx.x * BW;
    const int Hd = g.Hd, Wd = g.Wd, L = g.L;
    const int Lp = (L + 1) & ~1, M = Lp >> 1;
    const int RW = WIN ? SWIN : pg.rw;   // pitch of the staged right band

    if (tid == 0) {
        mbar_init(&band_bar, 1);
        s_all = 0;
        s_mask = m_zero<MaskT>();
    }
    __syncthreads();
    const float *sl = padl + ((size_t)frame * pg.rows + r0) * pg.pwl + c0;
    const float *sr = padr + ((size_t)frame * pg.rows + r0) * pg.pwr + c0;
    // Window of level pairs [m0, m1): right-band columns Lp-2(m1-1)-2 .. Lp-2 m0+84, from a 16-byte aligned start that
    // keeps the whole window inside the padded row (pg.rw is a multiple of 4 and >= Lp + 86).
    auto window_start = [&](int m1) {
        const int c_lo = (Lp - 2 * (m1 - 1) - 2) & ~3;
        return c_lo < pg.rw - SWIN ? c_lo : pg.rw - SWIN;
    };
    int win0 = WIN ? window_start(M < SWPASS ? M : SWPASS) : 0;
    if (WIN && win0 < 0) win0 = 0;
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(&band_bar, (unsigned)(SBR * (LW + RW) * 4));
        __syncwarp();
        for (int rr = tid; rr < SBR; rr += 32) {
            tma_bulk_g2s(bandL + rr * LW, sl + (size_t)rr * pg.pwl, LW * 4, &band_bar);
            tma_bulk_g2s(bandR + rr * RW, sr + (size_t)rr * pg.pwr + win0, (unsigned)(RW * 4), &band_bar);
        }
    }

    float2 *bY = bufs, *bZ = bY + 32 * SPY, *bW = bZ + 32 * SPZ;
    // phase-B ownership: pooled row `row` of the tile, columns 16*seg .. 16*seg+15 (a warp = one segment: its 32
    // lanes read 32 different rows, conflict-free with the pitches above)
    const int row = gt & 31, seg = gt >> 5;
    float rmax[16], lo[16];
    MaskT cand[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const bool valid = (r0 + row < Hd) && (c0 + 16 * seg + k < Wd);
        rmax[k] = 0.0f;
        lo[k] = valid ? 0.0f : __int_as_float(0x7f800000);   // pixels outside the image never become candidates
        cand[k] = m_zero<MaskT>();
    }

    {
        int spins = 0;
        while (!mbar_try_wait(&band_bar, 0))
            if (++spins > (1 << 24)) __trap();  // a lost transaction must not hang the GPU
    }

    for (int m = 0; m < M; m++) {
        const int d0 = 2 * m;
        if (WIN && m > 0 && (m % SWPASS) == 0) {
            // next window: everyone has left phase A of pair m-1 (the barrier that ended it), so the right band
            // may be overwritten; the mbarrier's phase flips with every use
            const int m1 = (m + SWPASS < M) ? m + SWPASS : M;
            win0 = window_start(m1);
            if (win0 < 0) win0 = 0;
            if (tid < 32) {
                if (tid == 0) mbar_expect_tx(&band_bar, (unsigned)(SBR * RW * 4));
                __syncwarp();
                for (int rr = tid; rr < SBR; rr += 32)
                    tma_bulk_g2s(bandR + rr * RW, sr + (size_t)rr * pg.pwr + win0, (unsigned)(RW * 4), &band_bar);
            }
            int spins = 0;
            while (!mbar_try_wait(&band_bar, (unsigned)((m / SWPASS) & 1)))
                if (++spins > (1 << 24)) __trap();
        }
        if (gt < SXW)
            screen_phase_a(bandL + gt + 4, bandR + gt + (Lp - d0 - 2 - win0), RW, bY + gt, bZ + (gt - SZ0), bW + (gt - SW0),
                           gt >= SZ0 && gt < SZ1, gt >= SW0 && gt < SW1);
        __syncthreads();

        // ---- phase B ---------------------------------------------------------------------------------------
        // Sliding sums along the row: the first window is a tree sum, every further position is one dependent add of
        // a precomputed difference (entering - leaving cell).  Absolute error of a window sum <= (5 + 2*17 + 2) u S_max
        // (see the header).
        const float2 neg1 = make_float2(-1.0f, -1.0f);
        // Window sums over TAP columns: a cost-column window [c, c+n) is the tap-column window [c, c+n+2) with the
        // horizontal 3-tap folded in, i.e. the sum of three n-wide box sums at consecutive positions.  Each n-wide sum
        // slides (one dependent add of a precomputed difference per position); A' = (H' V') C' with similarities
        // N*255 - dissimilarity sum.
        float2 A[16];
        {   // H: cost columns y .. y+20  ->  S21(p) = sum of Y3 over tap columns p .. p+20, H = S21(y) + S21(y+1) + S21(y+2)
            const float4 *pp = reinterpret_cast<const float4 *>(bY + row * SPY + 16 * seg);
            float2 y[38];
#pragma unroll
            for (int j = 0; j < 19; j++) {
                const float4 v4 = pp[j];
                y[2 * j] = lo2(v4);
                y[2 * j + 1] = hi2(v4);
            }
            float2 s1[10], s2[5], S[18];
#pragma unroll
            for (int j = 0; j < 10; j++) s1[j] = add2(y[2 * j], y[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 5; j++) s2[j] = add2(s1[2 * j], s1[2 * j + 1]);
            S[0] = add2(add2(add2(s2[0], s2[1]), add2(s2[2], s2[3])), add2(s2[4], y[20]));
#pragma unroll
            for (int k = 1; k < 18; k++) S[k] = add2(S[k - 1], __ffma2_rn(y[k - 1], neg1, y[k + 20]));   // + (y[k+20] - y[k-1])
            const float2 hmax = make_float2(kHVmax, kHVmax);
#pragma unroll
            for (int k = 0; k < 16; k++) A[k] = __ffma2_rn(add2(add2(S[k], S[k + 1]), S[k + 2]), neg1, hmax);
        }
        {   // V: cost columns y+9 .. y+11 -> tap columns y+9 .. y+13 with weights 1 2 3 2 1 (buffer column = tap column - 9)
            const float4 *pp = reinterpret_cast<const float4 *>(bW + row * SPW + 16 * seg);
            const float2 vmax = make_float2(kHVmax, kHVmax);
            float2 w[20], T[18];
#pragma unroll
            for (int j = 0; j < 10; j++) {
                const float4 v4 = pp[j];
                w[2 * j] = lo2(v4);
                w[2 * j + 1] = hi2(v4);
            }
#pragma unroll
            for (int k = 0; k < 18; k++) T[k] = add2(add2(w[k], w[k + 1]), w[k + 2]);
#pragma unroll
            for (int k = 0; k < 16; k++) A[k] = __fmul2_rn(A[k], __ffma2_rn(add2(add2(T[k], T[k + 1]), T[k + 2]), neg1, vmax));
        }
        {   // C: cost columns y+6 .. y+14 -> S9(p) = sum of Z9 over tap columns p+6 .. p+14 (buffer column = tap column - 6)
            const float4 *pp = reinterpret_cast<const float4 *>(bZ + row * SPZ + 16 * seg);
            float2 z[26], S[18];
#pragma unroll
            for (int j = 0; j < 13; j++) {
                const float4 v4 = pp[j];
                z[2 * j] = lo2(v4);
                z[2 * j + 1] = hi2(v4);
            }
            S[0] = add2(add2(add2(z[0], z[1]), add2(z[2], z[3])), add2(add2(z[4], z[5]), add2(add2(z[6], z[7]), z[8])));
#pragma unroll
            for (int k = 1; k < 18; k++) S[k] = add2(S[k - 1], __ffma2_rn(z[k - 1], neg1, z[k + 8]));
            const float2 cmax = make_float2(kCmax, kCmax);
#pragma unroll
            for (int k = 0; k < 16; k++) A[k] = __fmul2_rn(A[k], __ffma2_rn(add2(add2(S[k], S[k + 1]), S[k + 2]), neg1, cmax));
        }
        if (DBG && frame == 0 && r0 + row < Hd) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int yy = c0 + 16 * seg + k;
                if (yy < Wd) {
                    float *q = dbg_vol + ((size_t)(r0 + row) * Wd + yy) * L + d0;
                    q[0] = A[k].x;
                    if (d0 + 1 < L) q[1] = A[k].y;
                }
            }
        }
        // ---- candidate bookkeeping: the set always contains every LEVEL within kKeep of the final maximum -----------
        const bool has2 = (d0 + 1 < L);
        const MaskT bitx = m_bit<MaskT>(2 * m), bity = m_bit<MaskT>(2 * m + 1);
        float v[16];
        bool hit = false;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            v[k] = has2 ? fmaxf(A[k].x, A[k].y) : A[k].x;
            hit |= (v[k] >= lo[k]);
        }
        if (hit) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (v[k] >= lo[k]) {
                    if (v[k] > rmax[k] * kClear) cand[k] = m_zero<MaskT>();   // everything seen so far is below (1-eps) of the new max
                    if (v[k] > rmax[k]) {
                        rmax[k] = v[k];
                        lo[k] = v[k] * kKeep;
                    }
                    // both levels against the UPDATED threshold (<= kKeep * final maximum: still a superset)
                    if (A[k].x >= lo[k]) m_or(cand[k], bitx);
                    if (has2 && A[k].y >= lo[k]) m_or(cand[k], bity);
                }
            }
        }
        __syncthreads();   // the buffers are free for the next pass
    }

    // ---- the tile's set: union over its pixels; pixels the bound does not cover flag everything -----------------------
    MaskT mine = m_zero<MaskT>();
    bool weak = false;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const bool valid = (r0 + row < Hd) && (c0 + 16 * seg + k < Wd);
        if (!valid) continue;
        m_or(mine, cand[k]);
        if (!(rmax[k] >= kMinMax)) weak = true;   // below the bound's floor, or NaN
    }
    mine = warp_or(mine);
    weak = __any_sync(0xffffffffu, weak);
    if ((tid & 31) == 0) {
        if (m_any(mine)) shared_or(&s_mask, mine);
        if (weak) atomicOr(&s_all, 1);
    }
    __syncthreads();
    if (tid == 0) {
        unsigned w[4] = {0u, 0u, 0u, 0u};
        if (s_all) {
            for (int m = 0; m < M; m++) w[m >> 5] |= 1u << (m & 31);
        } else {
            // a candidate level d needs d-1, d, d+1 (circular in L, secondary_matching.cu:28-31) evaluated exactly
            for (int d = 0; d < L; d++) {
                const int m = d >> 1;
                if (!m_test(s_mask, d)) continue;
                const int a = ((d + L - 1) % L) >> 1, b = ((d + 1) % L) >> 1;
                w[m >> 5] |= 1u << (m & 31);
                w[a >> 5] |= 1u << (a & 31);
                w[b >> 5] |= 1u << (b & 31);
            }
        }
        const int tile = (frame * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        unsigned *dst = pass_mask + (size_t)tile * 4;
        dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
        const int pc = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
        // cost class of the tile for the fused kernel's heaviest-first schedule
        const int b = (pc * kScreenBuckets) / (M + 1);
        tile_order[b * (gridDim.x * gridDim.y * gridDim.z) + atomicAdd(&bucket_count[b], 1)] = tile;
        if (stats) {
            atomicAdd(&stats[0], (unsigned long long)pc);
            atomicAdd(&stats[1], (unsigned long long)M);
        }
        if (host_word) {
            // Adaptive policy feed (api.cu): the last block of the launch posts {chunk tag, pairs screened, pairs
            // flagged} of THIS chunk as one 64-bit store to mapped host memory; the host polls it, never waits.
            unsigned long long *acc = reinterpret_cast<unsigned long long *>(bucket_count + kScreenBuckets);
            int *done = bucket_count + kScreenBuckets + 2;
            atomicAdd(acc, ((unsigned long long)M << 24) | (unsigned long long)pc);
            __threadfence();
            const int total = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(done, 1) == total - 1) {
                __threadfence();
                const unsigned long long v = atomicAdd(acc, 0ull);
                *reinterpret_cast<volatile unsigned long long *>(host_word) = ((unsigned long long)(epoch & 0xffff) << 48) | (v & 0xffffffffffffull);
            }
        }
    }
}

template <typename MaskT, bool WIN, bool DBG>
cudaError_t launch_screen_t(const Geom &g, int frames, const Scratch &s, cudaStream_t st) {
    const size_t smem = WIN ? screen_smem_bytes_windowed() : screen_smem_bytes(g.L, g.min_ds, 1);
    const PadGeom pg = make_pad_geom(g.Hd, g.Wd, g.L, g.min_ds);
    cudaError_t e = cudaFuncSetAttribute(mbm_screen_kernel<MaskT, WIN, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(pg.tiles_x, pg.tiles_y, frames);
    e = cudaMemsetAsync(s.bucket_count, 0, kScreenCtrlInts * sizeof(int), st);
    if (e != cudaSuccess) return e;
    mbm_screen_kernel<MaskT, WIN, DBG><<<grid, SGT, smem, st>>>(g, pg, s.padl, s.padr, s.pass_mask, s.screen_stats, s.tile_order,
                                                            s.bucket_count, s.screen_host_word, s.range_epoch, s.dbg_screen);
    return cudaGetLastError();
}

}  // namespace

bool mbm_screen_supported(const Geom &g) {
    const int Lp = (g.L + 1) & ~1;
    return mbm_wta_fast_supported(g) && Lp / 2 <= 64 && Lp / 2 >= 2;
}

cudaError_t launch_mbm_screen(const Geom &g, int frames, const Scratch &s, cudaStream_t st) {
    if (!mbm_screen_supported(g) || !s.padl || !s.padr || !s.pass_mask || !s.tile_order || !s.bucket_count) return cudaErrorNotSupported;
    // One group of 128 threads per block, two blocks per SM.  L <= 64: the whole right band fits (64-bit candidate sets);
    // larger L: the right band is staged window by window (128-bit sets).
    const int M = ((g.L + 1) & ~1) / 2;
    if (M <= 32 && 2 * (screen_smem_bytes(g.L, g.min_ds, 1) + 1024) <= 227 * 1024)
        return s.dbg_screen ? launch_screen_t<unsigned long long, false, true>(g, frames, s, st)
                            : launch_screen_t<unsigned long long, false, false>(g, frames, s, st);
    return s.dbg_screen ? launch_screen_t<Mask128, true, true>(g, frames, s, st) : launch_screen_t<Mask128, true, false>(g, frames, s, st);
}

}  // namespace sd
