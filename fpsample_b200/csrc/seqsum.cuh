// seqsum.cuh -- the kd build's split value is a STRICTLY SEQUENTIAL binary32 sum (reference src/_ext/KDTreeBase.h:151-158:
// std::accumulate with a float accumulator): s <- RN(s + x_i), one dependent FADD per element.  This header evaluates a
// tile of it in parallel, bit for bit:
//
//   while every partial sum stays inside the binade of the incoming sum s (|s| in [2^e, 2^(e+1)), ulp u = 2^(e-23)),
//   s is an integer multiple k of u with 2^23 <= |k| < 2^24, and RN(s + x) = (k + RN_int(x / u)) * u: an INTEGER addition.
//   The only data-dependent rounding is an exact tie (x / u = m + 1/2), which rounds to the even neighbour of k + m + 1/2,
//   i.e. depends on the parity of the running k -- and leaves the running k EVEN.  So the parity entering any element is
//   the parity of the increments since the last tie (or since the tile's start), and a tile is
//     (1) per lane, EPL consecutive elements: increments under "even on entry", their sum / min / max prefix, and how
//         the lane's first tie would differ under "odd on entry";
//     (2) two ballots give every lane its entry parity (last lane below with a tie, flips in between);
//     (3) one warp prefix sum of the lane sums, min / max of all prefixes;
//     (4) accepted iff no element is too large for the integer model and 2^23 < |k| < 2^24 holds for every prefix
//         (with a slack of one unit for the hypothesis a lane did not track).  Otherwise the caller runs the plain chain.
//
// scripts/sim_blocksum.py is the CPU model (bit-equal to the sequential sum on uniform, LiDAR-like, tie-heavy and
// zero-mean columns); scripts/micro/seqsum_test.cu checks this device code against the chain on the GPU.
#pragma once
#include "common.cuh"

namespace fps {

__device__ __forceinline__ float4 sq_lds128(u32 a) {
    float4 f;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(a));
    return f;
}

// the plain chain over `cnt` floats (a multiple of 16) at the 16-byte aligned shared address `a`, every lane the same
__device__ __forceinline__ float sq_chain16(u32 a, u32 cnt, float sum) {
    const u32 nblk = cnt >> 4;
    float4 a0 = sq_lds128(a), a1 = sq_lds128(a + 16u), a2 = sq_lds128(a + 32u), a3 = sq_lds128(a + 48u);
    for (u32 b = 1; b <= nblk; ++b) {
        const u32 nb = a + 64u * (b < nblk ? b : b - 1);   // the last round reloads its own block: no branch in the chain
        const float4 n0 = sq_lds128(nb), n1 = sq_lds128(nb + 16u), n2 = sq_lds128(nb + 32u), n3 = sq_lds128(nb + 48u);
        sum = __fadd_rn(sum, a0.x), sum = __fadd_rn(sum, a0.y), sum = __fadd_rn(sum, a0.z), sum = __fadd_rn(sum, a0.w);
        sum = __fadd_rn(sum, a1.x), sum = __fadd_rn(sum, a1.y), sum = __fadd_rn(sum, a1.z), sum = __fadd_rn(sum, a1.w);
        sum = __fadd_rn(sum, a2.x), sum = __fadd_rn(sum, a2.y), sum = __fadd_rn(sum, a2.z), sum = __fadd_rn(sum, a2.w);
        sum = __fadd_rn(sum, a3.x), sum = __fadd_rn(sum, a3.y), sum = __fadd_rn(sum, a3.z), sum = __fadd_rn(sum, a3.w);
        a0 = n0, a1 = n1, a2 = n2, a3 = n3;
    }
    return sum;
}

// One tile of 32 * EPL floats at the 16-byte aligned shared address `a`, summed onto `s` (warp-uniform) exactly as the
// sequential chain would.  Returns false (s untouched) when the tile does not satisfy the integer model.
// hint: carried by the caller from tile to tile of one column (start at 0): non-zero after a tile that met an exact tie.
template <int EPL>
__device__ __forceinline__ bool seq_sum_tile(u32 a, float &s, u32 &hint) {
    static_assert(EPL % 4 == 0 && EPL >= 4 && EPL <= 32, "a lane reads whole float4s");
    const u32 lane = lane_id();
    const u32 ef = (__float_as_uint(s) >> 23) & 0xffu;
    if (ef < 24u || ef == 255u) return false;                        // zero, tiny, inf or nan incoming sum
    const float scale = __uint_as_float((277u - ef) << 23);          // 1 / u = 2^(150 - ef)
    const float u = __uint_as_float((ef - 23u) << 23);               // ulp of the binade, 2^(ef - 150)
    const int k_in = __float2int_rn(__fmul_rn(s, scale));            // exact, 2^23 <= |k_in| < 2^24

    int sm = 0, mn = 0x7fffffff, mx = (int)0x80000000, dlt = 0;
    u32 q = 0;                                                       // parity of the running k, "even on entry"
    bool seen = false, bad = false;
    const u32 base = a + lane * (u32)(EPL * 4);
    bool with_ties = hint != 0;   // the previous tile of this column had an exact tie: expect more, skip the tie-free pass
    if (!with_ties) {
        // tie-free pass: everything on the FP32 pipe, integers held in floats (|v| < 2^19 and EPL <= 32 keep every
        // partial sum below 2^24, so the float additions are exact)
        float smf = 0.0f, mnf = __int_as_float(0x7f800000), mxf = __int_as_float(0xff800000), vmax = 0.0f;
        bool anytie = false;
#pragma unroll
        for (int t = 0; t < EPL / 4; ++t) {
            const float4 f = sq_lds128(base + 16u * t);
            const float x[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float v = __fmul_rn(x[c], scale);              // exact (power of two) unless it overflows: caught below
                vmax = fmaxf(vmax, fabsf(v));
                // round to the nearest even integer without F2I / I2F: adding 1.5 * 2^23 leaves the integer in the low
                // mantissa bits (valid for |v| < 2^22)
                const float rf = __fsub_rn(__fadd_rn(v, 12582912.0f), 12582912.0f);
                anytie = anytie || fabsf(__fsub_rn(v, rf)) == 0.5f;
                smf = __fadd_rn(smf, rf);
                mnf = fminf(mnf, smf);
                mxf = fmaxf(mxf, smf);
            }
        }
        // fmaxf drops a nan operand, but a nan (or inf - inf) reaches smf; anything too large for the model trips vmax
        bad = !(vmax < 524288.0f) || !(fabsf(smf) < 16777216.0f);
        if (__any_sync(FULL, bad)) return false;
        sm = __float2int_rn(smf), mn = __float2int_rn(mnf), mx = __float2int_rn(mxf);
        q = (u32)sm & 1u;
        with_ties = __any_sync(FULL, anytie);
    }
    if (with_ties) {   // the full rule, every lane (a tie's rounding depends on the parity of the running sum)
        sm = 0, mn = 0x7fffffff, mx = (int)0x80000000, q = 0;
#pragma unroll
        for (int t = 0; t < EPL / 4; ++t) {
            const float4 f = sq_lds128(base + 16u * t);
            const float x[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float v = __fmul_rn(x[c], scale);
                bad = bad || !(fabsf(v) < 524288.0f);                // also catches nan
                const float tm = __fadd_rn(v, 12582912.0f);
                const int r = (__float_as_int(tm) & 0x7fffff) - 0x400000;
                const float d = __fsub_rn(v, __fsub_rn(tm, 12582912.0f));   // exact
                const bool tie = fabsf(d) == 0.5f;
                const int alt = r + (d > 0.0f ? 1 : -1);             // the other neighbour of a tie
                const int inc = (tie && q) ? alt : r;                // r is the even neighbour: right when k is even
                if (tie && !seen) {
                    dlt = (q ? r : alt) - inc;                       // what "odd on entry" adds here instead
                    seen = true;
                }
                q = tie ? 0u : (q ^ (u32)(inc & 1));
                sm += inc;
                mn = min(mn, sm);
                mx = max(mx, sm);
            }
        }
    }
    if (__any_sync(FULL, bad)) return false;

    // parity on entry of every lane: the last lane below with a tie fixes it, lanes without one flip it by their sum
    const u32 C = __ballot_sync(FULL, seen), V = __ballot_sync(FULL, q & 1u);
    hint = C;
    const u32 below = (1u << lane) - 1u, cm = C & below;
    u32 p, span;
    if (cm) {
        const u32 c = 31u - (u32)__clz(cm);
        p = (V >> c) & 1u;
        span = below & ~((2u << c) - 1u);
    } else {
        p = (u32)k_in & 1u;
        span = below;
    }
    p ^= (u32)__popc(V & span) & 1u;
    const int sl = sm + (p ? dlt : 0);

    int x = sl;   // inclusive prefix sum over lanes
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const int y = __shfl_up_sync(FULL, x, dd);
        if ((int)lane >= dd) x += y;
    }
    const int off = x - sl, total = __shfl_sync(FULL, x, 31);
    const int LO = __reduce_min_sync(FULL, off + mn - 1), HI = __reduce_max_sync(FULL, off + mx + 1);   // slack: the other hypothesis
    const int klo = k_in + LO, khi = k_in + HI;
    const bool ok = k_in > 0 ? (klo > (1 << 23) && khi < (1 << 24)) : (khi < -(1 << 23) && klo > -(1 << 24));
    if (!ok) return false;
    s = __fmul_rn(__int2float_rn(k_in + total), u);
    return true;
}


// ---- the same tile, WITHOUT knowing the incoming sum: a record for the two-phase sum of long columns (kdbuild.cu) ---------
// Under the hypothesis that the running sum enters the tile inside binade `ef` (biased exponent) and never leaves it, the
// tile adds an integer number of ulps that depends on the incoming sum only through its PARITY (the first exact tie):
// tot0 for an even, tot1 for an odd incoming k.  lo / hi bound every partial sum of the tile relative to the incoming k
// (with the one-unit slack of seq_sum_tile), so the consumer -- which walks the tiles in order with the true running sum --
// accepts the record iff its sum really is in binade ef and k + lo, k + hi stay inside (2^23, 2^24); otherwise it runs
// the plain chain over the tile.  Either way the result is the sequential sum, bit for bit; the hypothesis only decides
// how fast.  Returns false when the tile does not fit the integer model at all (an element too large for the binade).
struct SeqTileRec {
    u32 ef;
    int tot0, tot1, lo, hi;
    u32 pad[3];
};
template <int EPL>
__device__ __forceinline__ bool seq_sum_tile_record(u32 a, u32 ef, int &tot0, int &tot1, int &LO, int &HI) {
    static_assert(EPL % 4 == 0 && EPL >= 4 && EPL <= 32, "a lane reads whole float4s");
    const u32 lane = lane_id();
    if (ef < 24u || ef == 255u) return false;
    const float scale = __uint_as_float((277u - ef) << 23);          // 1 / u = 2^(150 - ef)
    int sm = 0, mn = 0x7fffffff, mx = (int)0x80000000, dlt = 0;
    u32 q = 0;                                                       // parity of the running k, "even on entry" of the lane
    bool seen = false, bad = false;
    const u32 base = a + lane * (u32)(EPL * 4);
#pragma unroll
    for (int t = 0; t < EPL / 4; ++t) {
        const float4 f = sq_lds128(base + 16u * t);
        const float x[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v = __fmul_rn(x[c], scale);
            bad = bad || !(fabsf(v) < 524288.0f);                    // also catches nan
            const float tm = __fadd_rn(v, 12582912.0f);
            const int r = (__float_as_int(tm) & 0x7fffff) - 0x400000;
            const float d = __fsub_rn(v, __fsub_rn(tm, 12582912.0f));   // exact
            const bool tie = fabsf(d) == 0.5f;
            const int alt = r + (d > 0.0f ? 1 : -1);                 // the other neighbour of a tie
            const int inc = (tie && q) ? alt : r;                    // r is the even neighbour: right when k is even
            if (tie && !seen) {
                dlt = (q ? r : alt) - inc;                           // what "odd on entry" adds here instead
                seen = true;
            }
            q = tie ? 0u : (q ^ (u32)(inc & 1));
            sm += inc;
            mn = min(mn, sm);
            mx = max(mx, sm);
        }
    }
    if (__any_sync(FULL, bad)) return false;
    // parity on entry of every lane, for an even (p0) and an odd (p1) incoming k: the last lane below with a tie fixes it
    // for both, lanes without one flip it by their sum
    const u32 C = __ballot_sync(FULL, seen), V = __ballot_sync(FULL, q & 1u);
    const u32 below = (1u << lane) - 1u, cm = C & below;
    u32 p0, p1, span;
    if (cm) {
        const u32 c = 31u - (u32)__clz(cm);
        p0 = p1 = (V >> c) & 1u;
        span = below & ~((2u << c) - 1u);
    } else {
        p0 = 0u, p1 = 1u;
        span = below;
    }
    const u32 flip = (u32)__popc(V & span) & 1u;
    p0 ^= flip, p1 ^= flip;
    const int sl0 = sm + (p0 ? dlt : 0), sl1 = sm + (p1 ? dlt : 0);
    int x0 = sl0, x1 = sl1;   // inclusive prefix sums over lanes
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const int y0 = __shfl_up_sync(FULL, x0, dd), y1 = __shfl_up_sync(FULL, x1, dd);
        if ((int)lane >= dd) x0 += y0, x1 += y1;
    }
    const int off0 = x0 - sl0, off1 = x1 - sl1;
    tot0 = __shfl_sync(FULL, x0, 31), tot1 = __shfl_sync(FULL, x1, 31);
    LO = min(__reduce_min_sync(FULL, off0 + mn - 1), __reduce_min_sync(FULL, off1 + mn - 1));   // slack: the hypothesis a lane did not track
    HI = max(__reduce_max_sync(FULL, off0 + mx + 1), __reduce_max_sync(FULL, off1 + mx + 1));
    return true;
}

// the consumer side: apply one record to the running sum (warp-uniform); false = run the chain over the tile instead
__device__ __forceinline__ bool seq_sum_apply_record(float &s, u32 ef, int tot0, int tot1, int lo, int hi) {
    const u32 es = (__float_as_uint(s) >> 23) & 0xffu;
    if (ef == 0u || es != ef) return false;
    const float scale = __uint_as_float((277u - ef) << 23);
    const float u = __uint_as_float((ef - 23u) << 23);
    const int k_in = __float2int_rn(__fmul_rn(s, scale));            // exact, 2^23 <= |k_in| < 2^24
    const int klo = k_in + lo, khi = k_in + hi;
    const bool ok = k_in > 0 ? (klo > (1 << 23) && khi < (1 << 24)) : (khi < -(1 << 23) && klo > -(1 << 24));
    if (!ok) return false;
    s = __fmul_rn(__int2float_rn(k_in + ((k_in & 1) ? tot1 : tot0)), u);
    return true;
}

// the whole sequential sum of `count` floats at shared address `a` (4-byte aligned), tiles of 32 * EPL where they apply
template <int EPL>
__device__ __forceinline__ float seq_sum_shared(u32 a, u32 count) {
    constexpr u32 TILE = 32u * EPL;
    float sum = 0.0f;
    u32 i = 0, hint = 0;
    while (i < count && ((a + 4u * i) & 15u)) {   // up to 3 values
        float f;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(a + 4u * i));
        sum = __fadd_rn(sum, f), ++i;
    }
    for (; i + TILE <= count; i += TILE)
        if (!seq_sum_tile<EPL>(a + 4u * i, sum, hint)) sum = sq_chain16(a + 4u * i, TILE, sum);
    const u32 rest16 = (count - i) & ~15u;
    if (rest16) sum = sq_chain16(a + 4u * i, rest16, sum), i += rest16;
    for (; i < count; ++i) {
        float f;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f) : "r"(a + 4u * i));
        sum = __fadd_rn(sum, f);
    }
    return sum;
}

}  // namespace fps
