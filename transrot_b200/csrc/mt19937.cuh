// MT19937 as the reference uses it: org.apache.commons.math3.random.MersenneTwister +
// BitsStreamGenerator (commons-math3 3.6.1, vendored as bytecode in the reference's
// src/resources; bit recipe in SURVEY.md Appendix B).  Call sites in the reference:
// Main.java:15,24,34-37,199,200,208; Space.java:31,77-79,677,694-696,721,736;
// Molecule.java:13,67,68.
//
// Two forms:
//   * mt_std_*  : one thread owns a whole 624-word state (seeding, the propagate kernel).
//   * MtStream  : one WARP owns a state that lives in global memory; lane i caches the
//                 tempered word (base+i) of a 32-word window in a register, a draw is one
//                 shuffle, and the 624-word block is regenerated incrementally, 32 words
//                 at a time, in place, when a window is first touched.
#pragma once
#include <stdint.h>

namespace tr {

constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int MT_WORDS = 625;   // exported form: 624 state words + mti

__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// ---- single-thread form ------------------------------------------------------------------

// MersenneTwister.setSeed(int)
__host__ __device__ inline void mt_std_seed_int(uint32_t *mt, uint32_t seed)
{
    mt[0] = seed;
    for (int i = 1; i < MT_N; i++) {
        uint32_t p = mt[i - 1];
        mt[i] = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
    }
}

// MersenneTwister.setSeed(long) = setSeed(int[]{hi, lo}) = init_by_array; leaves mti = 624.
__host__ __device__ inline void mt_std_seed_long(uint32_t *mt, int64_t seed)
{
    const uint32_t key[2] = { (uint32_t)((uint64_t)seed >> 32), (uint32_t)((uint64_t)seed & 0xffffffffu) };
    mt_std_seed_int(mt, 19650218u);
    int i = 1, j = 0;
    for (int k = MT_N; k != 0; k--) {
        uint32_t p = mt[i - 1];
        mt[i] = (mt[i] ^ ((p ^ (p >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
        if (j >= 2) j = 0;
    }
    for (int k = MT_N - 1; k != 0; k--) {
        uint32_t p = mt[i - 1];
        mt[i] = (mt[i] ^ ((p ^ (p >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
}

__host__ __device__ inline void mt_std_regen(uint32_t *mt)
{
    for (int k = 0; k < MT_N; k++) {
        uint32_t nxt = mt[k == MT_N - 1 ? 0 : k + 1];
        uint32_t y = (mt[k] & 0x80000000u) | (nxt & 0x7fffffffu);
        uint32_t far = mt[k < MT_N - MT_M ? k + MT_M : k + MT_M - MT_N];
        mt[k] = far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
}

__host__ __device__ inline uint32_t mt_std_next32(uint32_t *mt, int &mti)
{
    if (mti >= MT_N) { mt_std_regen(mt); mti = 0; }
    return mt_temper(mt[mti++]);
}

__host__ __device__ inline double mt_std_next_double(uint32_t *mt, int &mti)
{
    uint64_t hi = (uint64_t)(mt_std_next32(mt, mti) >> 6) << 26;
    uint64_t lo = mt_std_next32(mt, mti) >> 6;
    return (double)(int64_t)(hi | lo) * 0x1.0p-52;
}

__host__ __device__ inline int mt_std_next_int(uint32_t *mt, int &mti, int n)
{
    if ((n & -n) == n) return (int)(((int64_t)n * (int64_t)(mt_std_next32(mt, mti) >> 1)) >> 31);
    int bits, val;
    do {
        bits = (int)(mt_std_next32(mt, mti) >> 1);
        val = bits % n;
    } while ((int)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
    return val;
}

__host__ __device__ inline int64_t mt_std_next_long(uint32_t *mt, int &mti)
{
    uint64_t hi = (uint64_t)mt_std_next32(mt, mti) << 32;
    uint64_t lo = mt_std_next32(mt, mti);
    return (int64_t)(hi | lo);
}

// ---- warp-windowed form --------------------------------------------------------------------

#ifdef __CUDACC__

// State of one stream between draws.  pos = words consumed from the current 624-block (the
// reference's mti); gen = words of the current block already regenerated in place
// (a state produced by the single-thread form has gen = 624).  Both are warp-uniform.
struct MtStream {
    uint32_t win;   // lane i: tempered word (pos & ~31) + i
    int pos;
    int gen;
};

// Make the window [base, base+32) current.  In-place incremental regeneration is exact:
// word k needs old[k], old[k+1] and old[k+397] (k < 227) or new[k-227] (k >= 227), and
// new[0] for k = 623; windows are visited in increasing order, so every "new" operand was
// written by an earlier window and every "old" operand has not been touched yet.
__device__ __forceinline__ void mt_fetch_window(MtStream &s, uint32_t *mt, int base, int lane)
{
    const int k = base + lane;
    if (base < s.gen) {
        s.win = (k < MT_N) ? mt_temper(mt[k]) : 0u;
        return;
    }
    uint32_t cur = 0, nxt = 0, far = 0, v = 0;
    if (k < MT_N) {
        cur = mt[k];
        nxt = mt[k == MT_N - 1 ? 0 : k + 1];
        far = mt[k < MT_N - MT_M ? k + MT_M : k + MT_M - MT_N];
    }
    __syncwarp();
    if (k < MT_N) {
        uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
        v = far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mt[k] = v;
    }
    __syncwarp();
    s.gen = (base + 32 < MT_N) ? base + 32 : MT_N;
    s.win = mt_temper(v);
}

__device__ __forceinline__ void mt_stream_open(MtStream &s, uint32_t *mt, int pos, int gen, int lane)
{
    s.pos = pos;
    s.gen = gen;
    s.win = 0u;
    if (pos < MT_N && (pos & 31) != 0) mt_fetch_window(s, mt, pos & ~31, lane);
}

__device__ __forceinline__ uint32_t mt_next32(MtStream &s, uint32_t *mt, int lane)
{
    if (s.pos >= MT_N) { s.pos = 0; s.gen = 0; }
    if ((s.pos & 31) == 0) mt_fetch_window(s, mt, s.pos, lane);
    uint32_t w = __shfl_sync(0xffffffffu, s.win, s.pos & 31);
    s.pos++;
    return w;
}

// BitsStreamGenerator.nextDouble: (((long)next(26) << 26) | next(26)) * 0x1.0p-52
__device__ __forceinline__ double mt_next_double(MtStream &s, uint32_t *mt, int lane)
{
    uint64_t hi = (uint64_t)(mt_next32(s, mt, lane) >> 6) << 26;
    uint64_t lo = mt_next32(s, mt, lane) >> 6;
    return (double)(int64_t)(hi | lo) * 0x1.0p-52;
}

// BitsStreamGenerator.nextInt(n): power-of-two fast path, else the overflow-rejection loop
__device__ __forceinline__ int mt_next_int(MtStream &s, uint32_t *mt, int lane, int n)
{
    if ((n & -n) == n) return (int)(((int64_t)n * (int64_t)(mt_next32(s, mt, lane) >> 1)) >> 31);
    int bits, val;
    do {
        bits = (int)(mt_next32(s, mt, lane) >> 1);
        val = bits % n;
    } while ((int)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
    return val;
}

// Finish the in-place regeneration of the current block so the state reads like the
// reference's (all 624 words of the block valid, mti = pos).
__device__ __forceinline__ void mt_stream_canonicalize(MtStream &s, uint32_t *mt, int lane)
{
    if (s.pos >= MT_N) return;   // exhausted block: already canonical (mti = 624)
    uint32_t keep = s.win;
    for (int base = s.gen; base < MT_N; base += 32) mt_fetch_window(s, mt, base, lane);
    s.win = keep;
}

#endif  // __CUDACC__

}  // namespace tr
