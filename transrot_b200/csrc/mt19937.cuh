// MT19937 as the reference uses it: org.apache.commons.math3.random.MersenneTwister +
// BitsStreamGenerator (commons-math3 3.6.1, vendored as bytecode in the reference's
// src/resources; bit recipe in SURVEY.md Appendix B).  Call sites in the reference:
// Main.java:15,24,34-37,199,200,208; Space.java:31,77-79,677,694-696,721,736;
// Molecule.java:13,67,68.
//
// Two forms:
//   * mt_std_*  : one thread owns a whole 624-word state (seeding, the propagate kernel).
//   * MtRing    : one WARP owns a stream.  The 624-word block lives in global memory
//                 (L2-resident) in two ping-pong buffers; the warp regenerates it 32 words at a
//                 time (one word per lane, coalesced) and appends the TEMPERED words to a 64-word
//                 ring in shared memory.  A draw is one broadcast LDS; the "is there enough in
//                 the ring" test is hoisted to once per move.
#pragma once
#include <stdint.h>

namespace tr {

constexpr int MT_N = 624;
constexpr int MT_M = 397;
constexpr int MT_WORDS = 625;   // exported form: 624 state words + mti
constexpr int MT_RING = 64;     // tempered words per stream kept in shared memory

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

// ---- warp form -------------------------------------------------------------------------------

// What persists of one stream between launches (inside ChainCtrl).  The reference's state
// (624 words + mti) maps to { par = 0, gen = 624, fed = mti } with the words in buffer 0.
struct MtCursor {
    int par;    // which of the two global buffers holds the block being consumed / generated
    int gen;    // words of that block already generated (624 for an imported state)
    int fed;    // words of that block already handed to the ring (= consumed, when the ring is empty)
};

#ifdef __CUDACC__

// A stream as seen by its owning TEAM of L lanes (L = 16 or 32, all inside one warp).
struct MtRing {
    uint32_t *mt;        // global: [2][624] ping-pong buffers of this stream
    uint32_t *ring;      // shared: [MT_RING] tempered words
    MtCursor *cur;       // shared: persistent cursor (inside the chain's ChainCtrl copy)
    int rd;              // next ring slot to read          (team-uniform)
    int avail;           // unread words in the ring        (team-uniform)
};

__device__ __forceinline__ void mt_ring_open(MtRing &g, uint32_t *mt, uint32_t *ring, MtCursor *cur)
{
    g.mt = mt; g.ring = ring; g.cur = cur; g.rd = 0; g.avail = 0;
}

// Append the next <= 32 words of the stream to the ring (all L lanes of the owning team call this;
// tl = lane index inside the team, mask = the team's lanes).  Block b+1 is generated from block b
// held in the other buffer, so the regeneration is exact and a block stays readable until the one
// after next is started:
//   new[k] = (k < 227 ? old[k+397] : new[k-227]) ^ twist(old[k], k == 623 ? new[0] : old[k+1]).
// Words of one 32-word window never depend on each other (227 > 32), so L < 32 lanes may take
// them in any order.
template <int L>
__device__ __noinline__ int mt_ring_refill_words(uint32_t *mt, uint32_t *ring, MtCursor *cur, int wr, int tl, unsigned mask)
{
    int par = cur->par, gen = cur->gen, fed = cur->fed;
    __syncwarp(mask);
    if (fed >= MT_N) { par ^= 1; gen = 0; fed = 0; }
    uint32_t *B = mt + par * MT_N;
    const uint32_t *O = mt + (par ^ 1) * MT_N;
    const bool load = fed < gen;
    const int limit = load ? gen : MT_N;             // load mode stops at the generated prefix
    const int count = (limit - fed < 32) ? limit - fed : 32;
#pragma unroll
    for (int i = tl; i < 32; i += L) {
        if (i < count) {
            const int k = fed + i;
            uint32_t v;
            if (load) v = B[k];
            else {
                const uint32_t a = O[k];
                const uint32_t b = (k == MT_N - 1) ? B[0] : O[k + 1];
                const uint32_t far = (k < MT_N - MT_M) ? O[k + MT_M] : B[k + MT_M - MT_N];
                const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
                v = far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
                B[k] = v;
            }
            ring[(wr + i) & (MT_RING - 1)] = mt_temper(v);
        }
    }
    if (!load) gen = fed + count;
    fed += count;
    if (tl == 0) { cur->par = par; cur->gen = gen; cur->fed = fed; }
    // pull the operands of the NEXT window towards L1 now: its refill then finds them there instead of
    // waiting ~700 cycles on L2 with few other warps to cover the gap
    if (tl < 4 && fed >= gen) {
        const bool wrap = fed >= MT_N;                       // next window opens the next block
        const uint32_t *o2 = wrap ? B : O;                   // its "old" block
        const uint32_t *b2 = wrap ? O : B;                   // its "new" block (only [k-227] is read from it)
        const int k2 = wrap ? 0 : fed;
        const uint32_t *src = (tl < 2) ? o2 + k2 : ((k2 < MT_N - MT_M) ? o2 + k2 + MT_M : b2 + k2 + MT_M - MT_N);
        src += (tl & 1) * 32;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
    }
    __syncwarp(mask);
    return count;
}

// The same window, split so that the global loads can be in flight while the warp does other work
// (crew.cuh): mt_window_begin issues the loads of the next <= 32 words, mt_window_end twists, stores,
// tempers and appends them to the ring.  Whole warp (L = 32); `s` < 0 marks "nothing to do".
struct MtWindow {
    int s;               // stream, or -1
    int par, gen, fed;   // cursor before the window (after the block switch, if one was due)
    int count;
    int load;            // words were generated earlier (imported state / rewound cursor): plain reload
    uint32_t a, b, far;  // this lane's operands
};

__device__ __forceinline__ MtWindow mt_window_begin(uint32_t *mt, const MtCursor *cur, int s, int lane)
{
    MtWindow w;
    w.s = s; w.par = 0; w.gen = 0; w.fed = 0; w.count = 0; w.load = 0; w.a = w.b = w.far = 0u;
    if (s < 0) return w;
    int par = cur->par, gen = cur->gen, fed = cur->fed;
    if (fed >= MT_N) { par ^= 1; gen = 0; fed = 0; }
    const uint32_t *B = mt + par * MT_N, *O = mt + (par ^ 1) * MT_N;
    w.load = fed < gen;
    const int limit = w.load ? gen : MT_N;
    w.count = (limit - fed < 32) ? limit - fed : 32;
    w.par = par; w.gen = gen; w.fed = fed;
    if (lane < w.count) {
        const int k = fed + lane;
        if (w.load) w.a = B[k];
        else {
            w.a = O[k];
            w.b = (k == MT_N - 1) ? B[0] : O[k + 1];
            w.far = (k < MT_N - MT_M) ? O[k + MT_M] : B[k + MT_M - MT_N];
        }
    }
    return w;
}

// Returns the number of words appended at ring slot `wr`.
__device__ __forceinline__ int mt_window_end(const MtWindow &w, uint32_t *mt, uint32_t *ring, MtCursor *cur, int wr, int lane)
{
    if (w.s < 0) return 0;
    if (lane < w.count) {
        uint32_t v = w.a;
        if (!w.load) {
            const uint32_t y = (w.a & 0x80000000u) | (w.b & 0x7fffffffu);
            v = w.far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            mt[w.par * MT_N + w.fed + lane] = v;
        }
        ring[(wr + lane) & (MT_RING - 1)] = mt_temper(v);
    }
    if (lane == 0) { cur->par = w.par; cur->gen = w.load ? w.gen : w.fed + w.count; cur->fed = w.fed + w.count; }
    __syncwarp();
    return w.count;
}

template <int L>
__device__ __forceinline__ void mt_ring_refill(MtRing &g, int tl, unsigned mask)
{
    g.avail += mt_ring_refill_words<L>(g.mt, g.ring, g.cur, g.rd + g.avail, tl, mask);
}

// Hoisted availability test: call once per move with the most words the move can draw.
template <int L>
__device__ __forceinline__ void mt_ring_reserve(MtRing &g, int need, int tl, unsigned mask)
{
    while (g.avail < need) mt_ring_refill<L>(g, tl, mask);
}

__device__ __forceinline__ uint32_t mt_ring_next32(MtRing &g)
{
    const uint32_t w = g.ring[g.rd];
    g.rd = (g.rd + 1) & (MT_RING - 1);
    g.avail--;
    return w;
}

// The 52-bit integer of BitsStreamGenerator.nextDouble: ((long)next(26) << 26) | next(26)
__device__ __forceinline__ unsigned long long mt_ring_next52(MtRing &g)
{
    const uint32_t hi = mt_ring_next32(g) >> 6;
    const uint32_t lo = mt_ring_next32(g) >> 6;
    return ((unsigned long long)hi << 26) | lo;
}

// BitsStreamGenerator.nextDouble: that integer * 0x1.0p-52
__device__ __forceinline__ double mt_ring_next_double(MtRing &g)
{
    return (double)(long long)mt_ring_next52(g) * 0x1.0p-52;
}

// u >= p for u = nextDouble() without forming u: U * 2^-52 >= p  <=>  U >= ceil(p * 2^52) (both exact).
__host__ __device__ inline unsigned long long double_threshold52(double p)
{
    if (!(p > 0.0)) return 0ull;                      // u >= p always (also for p = -0.0)
    if (p > 1.0) return ~0ull;                        // never
    return (unsigned long long)ceil(p * 4503599627370496.0);
}

// Exact bits % n for 0 <= bits < 2^31 by multiply-high: magic = ceil(2^(31+L) / n), L = ceil(log2 n)
// (Granlund-Montgomery; the error term magic*n - 2^(31+L) < 2^L keeps the quotient exact below 2^31).
struct FastMod {
    uint32_t n, magic, shift;
};

__host__ __device__ inline FastMod fastmod_make(uint32_t n)
{
    FastMod f;
    f.n = n;
    uint32_t lg = 0;
    while ((1u << lg) < n) lg++;
    f.shift = 31 + lg;
    const uint64_t p = (uint64_t)1 << f.shift;
    f.magic = (uint32_t)((p + n - 1) / n);
    return f;
}

__host__ __device__ __forceinline__ uint32_t fastmod(const FastMod &f, uint32_t bits)
{
    const uint32_t q = (uint32_t)(((uint64_t)bits * f.magic) >> f.shift);
    return bits - q * f.n;
}

// BitsStreamGenerator.nextInt(n): power-of-two fast path, else the overflow-rejection loop.
// The caller reserved one word; a rejection (probability < n / 2^31) re-reserves.
template <int L>
__device__ __forceinline__ int mt_ring_next_int(MtRing &g, const FastMod &f, int tl, unsigned mask)
{
    const int n = (int)f.n;
    if ((n & -n) == n) return (int)(((int64_t)n * (int64_t)(mt_ring_next32(g) >> 1)) >> 31);
    for (;;) {
        const int bits = (int)(mt_ring_next32(g) >> 1);
        const int val = (int)fastmod(f, (uint32_t)bits);
        if ((int)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) >= 0) return val;
        mt_ring_reserve<L>(g, 1, tl, mask);
    }
}

// Give `avail` unread ring words back to the cursor so nothing but (par, gen, fed) has to persist.
__device__ __forceinline__ void mt_cursor_rewind(MtCursor &c, int avail)
{
    if (avail <= c.fed) c.fed -= avail;
    else { c.par ^= 1; c.fed = MT_N - (avail - c.fed); c.gen = MT_N; }   // back into the previous block
}

// Give the unread ring words back to the cursor so nothing but (par, gen, fed) has to persist.
__device__ __forceinline__ void mt_ring_close(MtRing &g, int tl, unsigned mask)
{
    __syncwarp(mask);
    if (tl == 0) {
        MtCursor c = *g.cur;
        if (g.avail <= c.fed) c.fed -= g.avail;
        else { c.par ^= 1; c.fed = MT_N - (g.avail - c.fed); c.gen = MT_N; }   // back into the previous block
        *g.cur = c;
    }
    g.avail = 0;
    __syncwarp(mask);
}

#endif  // __CUDACC__

}  // namespace tr
