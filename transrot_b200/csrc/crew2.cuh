// The annealing hot path, second warp-specialised kernel ("crew2"): particle-major energy teams, a per-chain cache of
// particle-pair energies, and the control flow split over three control warps.  Same reference code as crew.cuh:
// Main.sawtoothAnneal (Main.java:177-263), Space.move / Space.rotate (Space.java:660-734), Molecule.rotateTemp
// (Molecule.java:64-97), Atom.calcEnergy (Atom.java:91-119), Space.calcEnergy (Space.java:648-658).
//
// What changed against crew.cuh, and why (profiles/r2_*: the crew kernel issued 472 warp instructions per chain-move, 41 %
// of them FP64, and a round took 5 900 cycles against a 2 600-cycle FP64 floor - it was bound by instruction count and
// by the latency of its serial phases, not by the FP64 pipe):
//   * PAIR-ENERGY CACHE.  The reference recomputes eStart = sum_n m.calcEnergy(n) for every move (Space.java:711-718)
//     although nothing moved since the values were last formed.  Each chain keeps E(m, n) for every particle pair in
//     shared memory (N (N-1) / 2 doubles, packed lower triangle); a move evaluates only the TRIAL positions, takes
//     eStart from the cache, and an accepted move writes its per-neighbour energies back.  Half the FP64 work, half the
//     staged data.  The entries are overwritten, never accumulated, so nothing drifts; they are refreshed from scratch by
//     every Space.calcEnergy round (once per point) and persist across launches (KParams::pair_cache), which keeps
//     resumed runs bit-identical.  A rejected rotation still applies Molecule.rotateTemp's (a.x - x) + x round trip to the
//     committed coordinates (Molecule.java:70-72,90-92): cached entries of that particle then describe coordinates that
//     differ by at most one ulp, once (the round trip is idempotent) - far below the rounding of the sums themselves.
//   * PARTICLE-MAJOR TEAMS.  A team is L = 8 lanes; lane l owns the particles l, l + L, l + 2L ... (M per lane, SS site
//     slots each), so the per-neighbour energies the cache needs are plain per-lane sums, dE is reduced from
//     per-neighbour differences e_new(n) - E(m, n) (less cancellation than differencing two large sums), and one energy
//     warp serves FOUR chains: the per-round overhead (plan, commit, geometry, staging, reduction, barrier) is paid once
//     per four moves.  The staged sites of the moved particle sit in registers for the whole evaluation.
//   * THREE CONTROL WARPS, lane c <-> chain c.  The single control lane of crew.cuh needed 2 200 - 4 000 cycles per round;
//     with the teams faster it would be the critical path.  C1 keeps the S stream and the bookkeeping (settle, accumulators,
//     picks, translation vectors, modes).  C2 owns everything transcendental: the rotation draw (R stream, sincos) and the
//     Metropolis threshold -kT log(rand); it needs nothing C1 computes in the same round (which move is in flight follows
//     from the plan C1 posted a round earlier and from the sign of the previous dE).  C3 is the random-number generator:
//     it keeps the three MT19937 streams' 128-word rings of tempered words topped up behind the consumers' cursors, four
//     words at a time - the energy warps no longer spend a sixth of their instructions on ring windows.
//   * TWO GROUPS OF ENERGY WARPS half a round apart (TR_C2_LAG): the two energy warps of an SM sub-partition no longer reach
//     their FP64-heavy evaluation phase together (+ 8 %).
//   * Clusters of 33 and more particles have variants with whole-warp teams (L = 32) and the pair cache in global memory; they
//     are correct and tested but lose against the team kernels (a round lasts as long as its slowest energy warp, and a 5-site
//     move next to 1-site moves leaves most of them waiting), so tr_set_kernel(TR_KERNEL_CREW2) is the only way to get them.
// Plans and results are double-buffered by round parity; see "Round synchronisation" below for the barriers.
#pragma once
#include "crew.cuh"

namespace tr {

#ifdef TR_CREW_TIMING
#define C2TM(...) __VA_ARGS__
// ordered against the barrier (asm volatile with a memory clobber), unlike clock64()
__device__ __forceinline__ long long c2_clock() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#else
#define C2TM(...)
#endif

// Round synchronisation.  The energy warps and the control warps meet through two pairs of named barriers instead of one
// block-wide barrier per round:
//   plan barrier p (ids 2, 3; parity of the round the plan is for): control warps ARRIVE after writing the plan, energy warps WAIT.
//   result barrier p (ids 4, 5; parity of the round): energy warps ARRIVE after posting results, control warps WAIT.
// So a control warp never blocks on posting a plan and an energy warp never blocks on posting its results.  With one plan
// barrier per GROUP of energy warps (the whole-warp variants: too many warps for the 16 barrier ids) the warps of a group still
// advance round by round together - a round lasts as long as its slowest energy warp, which is what makes those variants lose on
// mixed clusters ((NH4Cl)32: a 5-site move next to 1-site moves; 33 % of all samples wait at the plan barrier).  With one plan
// barrier per energy WARP (TR_C2_PER_WARP, up to 10 energy warps) a warp waits for the control warps only; it can run up to a
// round ahead of the slowest one, because plan r + 1 is built from the results of round r - 1.
__device__ __forceinline__ void c2_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void c2_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
constexpr int C2_BAR_PLAN = 2, C2_BAR_RES = 4, C2_BAR_PLAN_B = 6;
constexpr int C2_CTRL_WARPS_ = 3;

// Two groups of energy warps (TR_C2_LAG > 0).  Every SM sub-partition hosts two energy warps (ids w and w + 4); started
// together they reach their FP64-heavy evaluation phase together and share the pipe, then leave it idle together.  Group B
// (`c2_group(warp) == 1`) has its own plan barrier (ids 6, 7) and starts every round TR_C2_LAG cycles after group A did (a
// time stamp group A leaves in shared memory), so one group evaluates while the other runs its serial phase.  The control
// warps arrive at both plan barriers and still wait for the results of all energy warps: plan r + 1 is built from the results
// of round r - 1, a whole round back, which is what leaves room for the lag.
// Measured on (H2O)20 x 4096 chains (profiles/r2_experiments.md): lag 0 (one group): 1.51 G moves/s; 600: 1.53; 1000: 1.56;
// 1400: 1.61; 1800: 1.63; 2200: 1.62; 2600: 1.54 (group A then waits for the plan: lag + group B's round + the control
// warps' iteration have to fit in two rounds).  The sample cluster (5 site slots per lane, shorter rounds): 1000: 2.02, 1800: 1.97.
#ifndef TR_C2_LAG
#define TR_C2_LAG 1800
#endif
// which of the two energy warps of a sub-partition leads: 1 = warps 4..7 are group A, warps 3, 8, 9 lag (+ 0.9 % over 0)
#ifndef TR_C2_GROUP_SWAP
#define TR_C2_GROUP_SWAP 1
#endif
__host__ __device__ constexpr int c2_group(int warp) { return TR_C2_LAG > 0 ? (((warp >> 2) & 1) ^ TR_C2_GROUP_SWAP) : 0; }
constexpr int C2_STAMP_WARP = TR_C2_GROUP_SWAP ? 4 : 3;          // the first energy warp of group A
// threads that meet at a group's plan barrier: the three control warps + the group's energy warps (warps 3 .. NW-1 by role map 0)
template <int NW>
__host__ __device__ constexpr int c2_plan_threads(int group)
{
    int n = 3;
    for (int w = 3; w < NW; w++) n += c2_group(w) == group ? 1 : 0;
    return 32 * n;
}
// TR_C2_PER_WARP (experiment, OFF): one named plan barrier per energy warp (ids 6 ..; the three control warps arrive, that warp
// waits) instead of one per group, so that an energy warp waits for the control warps only (+ 1.8 % on (H2O)20).  NOT SAFE:
// there are not enough barrier ids for two per warp (one per round parity), and with one per warp nothing but the length of a
// control iteration keeps the control warps' next arrival behind the warp's wait for the previous one.  In the short rounds at
// the end of a launch the arrival can overtake the wait, the barrier then counts arrivals of two phases into one, and the warp
// waits forever: about one launch in 140 deadlocked (tools/stress_hang.py).  The group barriers alternate by round parity and
// are safe by construction (an arrival for plan r + 2 follows the results of round r, hence the wait for plan r).  A safe
// per-warp hand-off with two mbarrier objects per warp ran 4 000 launches clean but measured no gain over the group barriers
// (1.672 vs 1.67 G), so it was not kept.
#ifndef TR_C2_PER_WARP
#define TR_C2_PER_WARP 0
#endif
constexpr int C2_BAR_PLAN_W = 6;
template <int NW>
__device__ __forceinline__ void c2_plan_posted(int parity)
{
    if (TR_C2_PER_WARP && NW - C2_CTRL_WARPS_ <= 10) {
#pragma unroll
        for (int w = 0; w < NW - C2_CTRL_WARPS_; w++) c2_bar_arrive(C2_BAR_PLAN_W + w, 32 * (C2_CTRL_WARPS_ + 1));
        return;
    }
    c2_bar_arrive(C2_BAR_PLAN + parity, c2_plan_threads<NW>(0));
    if (TR_C2_LAG > 0) c2_bar_arrive(C2_BAR_PLAN_B + parity, c2_plan_threads<NW>(1));
}

constexpr int C2_RING_NEED_M = 4;      // M words a move consumes, always (Main.java:199-208)
constexpr int C2_RING_NEED_R = 8;      // R words guaranteed before a move is posted: angle double + up to 6 nextInt(3) tries
constexpr int C2_MAX_BUBBLES = 256;    // rounds a chain may wait for random words before it is abandoned (CHAIN_ABORT_RNG)

struct __align__(16) C2Cand {          // one drawn move, the S-stream part (Space.randMolecule + the translation vector)
    double va, vb, vc;
    int m;
    int info;                          // site count | rot << 8
};

struct __align__(16) C2Plan {
    // ---- written by control warp C1
    C2Cand cand[2];                    // PLAN_AUTO: successor if no Metropolis double was drawn / if one was; PLAN_MOVE: cand[0]
    int mode, flags, snap_slot, movie_slot;
    int endS0, endS1;                  // S-ring position after each candidate's own draws (where its Metropolis double sits)
    double kBt;
    // ---- written by control warp C2
    double rnd, thr;                   // Metropolis double of the move evaluated in the PREVIOUS round, -kT log(rnd)
    double grd, sn;                    // its guard band; sin of this round's rotation angle
    double cs;                         // cos
    int axis, nR;                      // rotation axis; R words the rotation draw consumes (< 0: ran out of guaranteed words)
};

struct __align__(16) C2Stage {         // one site of the moved particle at its trial position
    double nx, ny, nz, kq;
    int row, lj;
    int pad_[2];
};

// Control region of a chain (what the control lanes touch); the team region (stage, com, coordinates, pair cache) is
// separate so that each can have the stride that keeps its accesses off each other's banks.
struct __align__(16) C2Chain {
    ChainCtrl ctrl;
    tr_schedule sch;
    uint32_t *mt;                      // global: [3][2][624] ping-pong blocks of this chain
    __align__(16) uint32_t ring[3][CREW_RING];     // (C3 appends four words with one 16-byte store)
    C2Plan plan[2];
    CrewResult result[2];
    double etot;
    double bound;
    int rd[3];                         // ring words consumed so far (C1 -> C3)
    int pad0_;
    int wr[3];                         // ring words appended so far (C3 -> C1)
    int pad1_;
    __align__(16) uint32_t pre[4][12]; // C3: operands of the next chunk of S, S + 1, M, R, copied in asynchronously (cp.async)
};

// Chain-independent tables at compile-time offsets; the chain regions follow at runtime strides.
template <int L, int M, int SS>
struct C2Tables {
    static constexpr int SPL = M * SS;                 // site slots per lane
    static constexpr int NPOS = L * SPL;
    static constexpr int NMOL = L * M;                 // particle slots
    // Whole-warp teams (L = 32: 33 or more particles) keep the pair cache in global memory (N (N-1) / 2 doubles per chain no
    // longer fit beside enough chains to fill an SM: 16 KB at N = 64, 255 KB at N = 256) and evaluate particle slot by
    // particle slot; small teams keep it in shared memory and evaluate all their site slots at once.
    static constexpr bool GLOBAL_CACHE = L >= 32;
    static constexpr bool WIDE = !GLOBAL_CACHE && M * SS <= 10;
    static constexpr int OFF_STAMP = 16;                                       // [0..15]: block-wide scalars; [16..31]: round stamps
    static constexpr int OFF_SLOT_NS = 32;                                     // then [M] ints
    static constexpr int OFF_MOL_NS = OFF_SLOT_NS + align16(M * 4);
    static constexpr int OFF_MOVEABLE = OFF_MOL_NS + align16(NMOL * 4);
    static constexpr int OFF_POS_INFO = OFF_MOVEABLE + align16(NMOL * 4);
    static constexpr int OFF_QPOS = OFF_POS_INFO + align16(NPOS * 4);
    static constexpr int OFF_CD = OFF_QPOS + NPOS * 8;

    __host__ __device__ static int table_bytes(int n_types, bool has_exp) { return (OFF_CD + n_types * (n_types + 1) * 16 * (has_exp ? 2 : 1) + 127) & ~127; }
    __host__ __device__ static int tri_count(int n_mol) { return (n_mol * (n_mol - 1) / 2 + 1) & ~1; }     // global copy (KParams::pair_cache)
    // team region, every offset a compile-time constant (sized for the NMOL particle slots, whatever n_mol is): one base
    // register per chain serves every access
    static constexpr int T_STAGE = 0;
    static constexpr int T_COM = T_STAGE + SS * (int)sizeof(C2Stage);
    // Coordinates: position p sits at element xi(p) = p + p / L - every row of L positions (one site slot of every lane) is
    // followed by one element of padding, so that the sites of ONE particle (positions L apart: the trial geometry, the commit)
    // fall into different banks while a team's row stays contiguous.
    static constexpr int XROW = L + 1;
    static constexpr int NXY = (NPOS / L) * XROW;
    static constexpr int T_XY = T_COM + align16(NMOL * 24);     // [NXY] double2 {x, y}: one 16-byte load per site slot
    static constexpr int T_ZS = T_XY + NXY * 16;                // [NXY] double
    static constexpr int T_PEND = T_ZS + align16(NXY * 8);
    static constexpr int T_CACHE = T_PEND + NMOL * 8;
    static constexpr int T_END = T_CACHE + (GLOBAL_CACHE ? 0 : ((NMOL * (NMOL - 1) / 2 + 1) & ~1) * 8);
    // control region: 16 mod 128 bytes, so that the lanes of a control warp (one chain each) spread over the banks
    __host__ __device__ static constexpr int ctrl_bytes() { return ((align16((int)sizeof(C2Chain)) + 127) & ~127) + 16; }
    // team region: 64 mod 128 bytes, so that the two teams a 64-bit access serves in one pass (8 lanes x 8 bytes each) fall
    // into different banks
    __host__ __device__ static constexpr int team_bytes() { return ((T_END + 127) & ~127) + 64; }

    __device__ static __forceinline__ int &alive(unsigned char *smem, int round) { return ((int *)smem)[round & 1]; }
    // (round << 32 | clock) of the moment group A's first energy warp started the round
    __device__ static __forceinline__ volatile unsigned long long *stamp(unsigned char *smem, int round)
    {
        return (volatile unsigned long long *)(smem + OFF_STAMP) + (round & 1);
    }
    // the most sites any lane's particle has in particle slot k (warp-uniform loop bounds for the slot-by-slot evaluation)
    __device__ static __forceinline__ int slot_ns(const unsigned char *smem, int k) { return *(const int *)(smem + OFF_SLOT_NS + 4 * k); }
    __device__ static __forceinline__ int mol_ns(const unsigned char *smem, int m) { return *(const int *)(smem + OFF_MOL_NS + 4 * m); }
    __device__ static __forceinline__ int moveable(const unsigned char *smem, int i) { return *(const int *)(smem + OFF_MOVEABLE + 4 * i); }
    __device__ static __forceinline__ int pos_info(const unsigned char *smem, int p) { return *(const int *)(smem + OFF_POS_INFO + 4 * p); }
    __device__ static __forceinline__ double qpos(const unsigned char *smem, int p) { return *(const double *)(smem + OFF_QPOS + 8 * p); }
    __device__ static __forceinline__ const unsigned char *cd(const unsigned char *smem) { return smem + OFF_CD; }
    __device__ static __forceinline__ const unsigned char *ab(const unsigned char *smem, const DevSystem &sys)
    {
        return smem + OFF_CD + (sys.has_exp ? sys.n_types * sys.row_stride : 0);
    }
    // position of site a of particle m: particle slot k = m / L of lane m % L
    __device__ static __forceinline__ int pos_of(int m, int a) { return ((m / L) * SS + a) * L + (m % L); }
    __host__ __device__ static constexpr int xi(int p) { return p + p / L; }

    __device__ static void load(unsigned char *smem, const DevSystem &sys)
    {
        const int T = sys.n_types, tab = T * (T + 1) * 16;
        double2 *cdp = (double2 *)(smem + OFF_CD), *abp = (double2 *)(smem + OFF_CD + tab);
        for (int i = threadIdx.x; i < T * (T + 1); i += blockDim.x) {
            const int r = i / (T + 1), c = i % (T + 1);
            cdp[i] = (c < T) ? sys.pair_cd[r * T + c] : make_double2(0.0, 0.0);
            if (sys.has_exp) abp[i] = (c < T) ? sys.pair_ab[r * T + c] : make_double2(0.0, 0.0);
        }
        int *ns = (int *)(smem + OFF_MOL_NS), *mv = (int *)(smem + OFF_MOVEABLE), *pi = (int *)(smem + OFF_POS_INFO);
        double *qp = (double *)(smem + OFF_QPOS);
        for (int i = threadIdx.x; i < NMOL; i += blockDim.x) ns[i] = i < sys.n_mol ? sys.mol_off[i + 1] - sys.mol_off[i] : 0;
        if (threadIdx.x < M) {
            int most = 0;
            for (int n = threadIdx.x * L; n < (threadIdx.x + 1) * L && n < sys.n_mol; n++) most = max(most, sys.mol_off[n + 1] - sys.mol_off[n]);
            *(int *)(smem + OFF_SLOT_NS + 4 * threadIdx.x) = most;
        }
        for (int i = threadIdx.x; i < sys.n_moveable; i += blockDim.x) mv[i] = sys.moveable[i];
        for (int i = threadIdx.x; i < NPOS; i += blockDim.x) {
            pi[i] = sys.pos_info[i];
            const int s = sys.pos_site[i];
            qp[i] = s >= 0 ? sys.site_q[s] : 0.0;
        }
    }
};

struct C2View {
    C2Chain *cs;
    C2Stage *stage;
    double2 *xy;
    double *com, *zs, *cache, *pend;
};

// Block layout: tables | C control regions | C team regions
template <int L, int M, int SS>
__device__ __forceinline__ C2View c2_view(unsigned char *chains, int c, int C, double *global_cache)
{
    using TL = C2Tables<L, M, SS>;
    C2View v;
    v.cs = (C2Chain *)(chains + c * TL::ctrl_bytes());
    unsigned char *b = chains + C * TL::ctrl_bytes() + c * TL::team_bytes();
    v.stage = (C2Stage *)(b + TL::T_STAGE);
    v.com = (double *)(b + TL::T_COM);
    v.xy = (double2 *)(b + TL::T_XY); v.zs = (double *)(b + TL::T_ZS);
    v.pend = (double *)(b + TL::T_PEND);           // [M][L]: per-neighbour energies of the move staged last
    v.cache = TL::GLOBAL_CACHE ? global_cache : (double *)(b + TL::T_CACHE);
    return v;
}

__device__ __forceinline__ int tri_index(int lo, int hi) { return ((hi * (hi - 1)) >> 1) + lo; }

// ---- C3: the random-number generator ------------------------------------------------------------------------------------------
// One lane keeps one chain's three streams.  The 624-word blocks live in global memory (L2 / L1) in two ping-pong buffers
// (mt19937.cuh): new[k] = (k < 227 ? old[k+397] : new[k-227]) ^ twist(old[k], k == 623 ? new[0] : old[k+1]).  Tempered words are
// appended to the stream's 128-word ring in shared memory four at a time; the ring's write position is always a multiple of 4.

struct C2Gen {                       // one stream as its generator lane sees it
    uint32_t *mt;                    // global: [2][624]
    uint32_t *ring;                  // shared: [CREW_RING]
    int par, gen, fed;               // MtCursor
    int wr;                          // words appended so far
};

// One word, whatever the state (used until the stream position is a multiple of 4: imported / rewound cursors).
__device__ __forceinline__ void c2_gen_word(C2Gen &g)
{
    if (g.fed >= MT_N) { g.par ^= 1; g.gen = 0; g.fed = 0; }
    uint32_t *B = g.mt + g.par * MT_N;
    const uint32_t *O = g.mt + (g.par ^ 1) * MT_N;
    const int k = g.fed;
    uint32_t v;
    if (k < g.gen) v = B[k];
    else {
        const uint32_t a = O[k];
        const uint32_t b = (k == MT_N - 1) ? B[0] : O[k + 1];
        const uint32_t far = (k < MT_N - MT_M) ? O[k + MT_M] : B[k + MT_M - MT_N];
        const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        v = far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        B[k] = v;
        g.gen = k + 1;
    }
    g.ring[g.wr & (CREW_RING - 1)] = mt_temper(v);
    g.fed = k + 1;
    g.wr += 1;
}

// An aligned chunk of four words in two steps a round apart, so that C3 never waits for global memory: c2_chunk_issue starts
// asynchronous copies (cp.async) of the chunk's operands into the chain's staging area and moves the ISSUE cursor;
// c2_chunk_finish (next round or later, when the ring has room) twists, stores the new state words, tempers and appends.
// A chunk's operands never depend on the chunks still in flight ahead of it: it reads old[k .. k+4], old[k+397 ..] or
// new[k-227 ..] and (the last one of a block) new[0] - all at least 56 chunks old.
struct C2Issue {                     // where the next chunk to be ISSUED sits (runs ahead of C2Gen by the chunks in flight)
    int par, gen, fed;
};

__device__ __forceinline__ void c2_cp_async4(uint32_t *dst_smem, const uint32_t *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void c2_cp_async16(uint32_t *dst_smem, const uint32_t *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}

// slot layout: [0..3] old[k..k+3] (or the words themselves: reload), [4] old[k+4] / new[0], [5..8] far words, [9] = 1: reload
__device__ __forceinline__ void c2_chunk_issue(uint32_t *mt, C2Issue &it, uint32_t *slot)
{
    if (it.fed >= MT_N) { it.par ^= 1; it.gen = 0; it.fed = 0; }
    uint32_t *B = mt + it.par * MT_N;
    const uint32_t *O = mt + (it.par ^ 1) * MT_N;
    const int k = it.fed;
    if (k < it.gen) {                                          // (gen is a multiple of 4 whenever fed is)
        c2_cp_async16(slot, B + k);
        slot[9] = 1u;
    } else {
        c2_cp_async16(slot, O + k);
        c2_cp_async4(slot + 4, (k + 4 == MT_N) ? B : O + k + 4);
#pragma unroll
        for (int i = 0; i < 4; i++) c2_cp_async4(slot + 5 + i, (k + i < MT_N - MT_M) ? O + k + i + MT_M : B + k + i + MT_M - MT_N);
        slot[9] = 0u;
        it.gen = k + 4;
    }
    it.fed = k + 4;
}

__device__ __forceinline__ void c2_chunk_finish(C2Gen &g, const uint32_t *slot)
{
    if (g.fed >= MT_N) { g.par ^= 1; g.gen = 0; g.fed = 0; }
    const int k = g.fed;
    const uint4 a = *(const uint4 *)slot;
    uint4 v = a;
    if (slot[9] == 0u) {
        const uint32_t a4 = slot[4];
        const uint4 f = make_uint4(slot[5], slot[6], slot[7], slot[8]);
        uint32_t y;
        y = (a.x & 0x80000000u) | (a.y & 0x7fffffffu); v.x = f.x ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        y = (a.y & 0x80000000u) | (a.z & 0x7fffffffu); v.y = f.y ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        y = (a.z & 0x80000000u) | (a.w & 0x7fffffffu); v.z = f.z ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        y = (a.w & 0x80000000u) | (a4 & 0x7fffffffu);  v.w = f.w ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        *(uint4 *)(g.mt + g.par * MT_N + k) = v;
        g.gen = k + 4;
    }
    g.ring[g.wr & (CREW_RING - 1)] = mt_temper(v.x);
    g.ring[(g.wr + 1) & (CREW_RING - 1)] = mt_temper(v.y);
    g.ring[(g.wr + 2) & (CREW_RING - 1)] = mt_temper(v.z);
    g.ring[(g.wr + 3) & (CREW_RING - 1)] = mt_temper(v.w);
    g.fed = k + 4;
    g.wr += 4;
}

// ---- energy -------------------------------------------------------------------------------------------------------------

// What a team keeps in registers between rounds to commit the move it staged last (setTemps, Molecule.java:128-135); the
// per-neighbour energies at the trial position wait in the team's shared-memory region (C2View::pend).
struct C2Pending {
    double tx, ty, tz;         // lane a < s_m: trial coordinates of site a
    double cx, cy, cz;         // lane 0: the particle's centre of mass if the move is committed
    int pp;                    // position of this lane's staged site, or -1
    int pm;                    // lane 0: particle whose centre of mass moves, or -1 (rotation)
    int m;                     // the moved particle
};

// Trial geometry of a posted move (Space.java:660-667 / 692-710 after the draws): lane a < s_m builds site a of particle m at
// its trial position and stages it; returns this lane's box-test result (0.75 * spaceLen, Space.java:663,706).
template <int L, int M, int SS>
__device__ __forceinline__ bool c2_geometry(const unsigned char *smem, const C2View &v, const DevSystem &sys, const C2Cand &d, const C2Plan *pl, int tl,
                                            C2Pending &pend, int &lj_out)
{
    using TL = C2Tables<L, M, SS>;
    const int m = d.m, sm = d.info & 0xff, rot = (d.info >> 8) & 1;
    bool oob = false;
    pend.pp = -1; pend.m = m;
    lj_out = 0;
    if (tl < sm) {
        const int p = TL::pos_of(m, tl);
        const int inf = TL::pos_info(smem, p);
        const int px = TL::xi(p);
        const double2 oxy = v.xy[px];
        double ox = oxy.x, oy = oxy.y, oz = v.zs[px], tx, ty, tz;
        if (rot) {
            // Molecule.java:70-95, operation for operation, no contraction
            const double cx = v.com[3 * m], cy = v.com[3 * m + 1], cz = v.com[3 * m + 2];
            const double va = pl->sn, vb = pl->cs, nsn = __dmul_rn(-1.0, va);
            const int axis = pl->axis;
            const double ax = __dsub_rn(ox, cx), ay = __dsub_rn(oy, cy), az = __dsub_rn(oz, cz);
            // The Givens rotation mixes two coordinates (pu, pw) and leaves the third alone.  Axis 0: (y, z), axis 1: (x, z),
            // axis 2: (x, y); first result cos*pu + sin*pw (axis 1: cos*pu - sin*pw), second -sin*pu + cos*pw (axis 1:
            // sin*pu + cos*pw) - selected, not branched, so that teams with different axes do not serialise.
            const double pu = axis == 0 ? ay : ax, pw = axis == 2 ? ay : az;
            const double s1 = axis == 1 ? nsn : va, s2 = axis == 1 ? va : nsn;
            const double ru = __dadd_rn(__dmul_rn(vb, pu), __dmul_rn(s1, pw));
            const double rw = __dadd_rn(__dmul_rn(s2, pu), __dmul_rn(vb, pw));
            tx = axis == 0 ? ax : ru;
            ty = axis == 0 ? ru : (axis == 1 ? ay : rw);
            tz = axis == 2 ? az : rw;
            // a.x += x: the committed coordinate keeps the (a.x - x) + x round trip even if the move is rejected
            v.xy[px] = make_double2(__dadd_rn(ax, cx), __dadd_rn(ay, cy)); v.zs[px] = __dadd_rn(az, cz);
            tx = __dadd_rn(tx, cx); ty = __dadd_rn(ty, cy); tz = __dadd_rn(tz, cz);
        } else {
            tx = __dadd_rn(ox, d.va); ty = __dadd_rn(oy, d.vb); tz = __dadd_rn(oz, d.vc);
        }
        C2Stage st;
        st.nx = tx; st.ny = ty; st.nz = tz;
        st.kq = __dmul_rn(K_VALUE, TL::qpos(smem, p));
        st.row = info_type(inf) * sys.row_stride; st.lj = info_lj(inf);
        st.pad_[0] = st.pad_[1] = 0;
        v.stage[tl] = st;
        lj_out = st.lj;
        const double b = v.cs->bound;                // t > L*0.75 || t < L*-0.75; the lower bound is the exact negative
        oob = (fabs(tx) > b) | (fabs(ty) > b) | (fabs(tz) > b);
        pend.tx = tx; pend.ty = ty; pend.tz = tz; pend.pp = px;       // (element index, padding included)
    }
    if (tl == 0) {
        pend.pm = rot ? -1 : m;
        if (!rot) {
            pend.cx = __dadd_rn(v.com[3 * m], d.va);
            pend.cy = __dadd_rn(v.com[3 * m + 1], d.vb);
            pend.cz = __dadd_rn(v.com[3 * m + 2], d.vc);
        }
    }
    return oob;
}

// Per-neighbour energies of the staged particle at its trial position against this lane's particles, and this lane's part of
// eEnd - eStart (cached energies subtracted per neighbour).  `sm_w` / `ljw` are warp-uniform: the most staged sites any team of
// the warp has this round, and which staged sites carry Lennard-Jones parameters in any team - a team with fewer staged sites
// sees neutral entries (zero charge, far away), so the loop bounds and the arithmetic form never diverge inside a warp.
// The slot that holds the moved particle itself (and empty slots) is evaluated like any other and discarded afterwards.
// The loop over the particle slots is a real loop: the whole kernel's hot code has to stay inside the 32 KB instruction cache.
template <int L, int M, int SS, bool HAS_EXP, bool TRACE>
__device__ __forceinline__ void c2_evaluate(const unsigned char *smem, const C2View &v, const DevSystem &sys, int m, int sm_w, unsigned ljw,
                                            unsigned ljb, int tl, double &dl, double &sl, double &el)
{
    using TL = C2Tables<L, M, SS>;
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    const int nmol = sys.n_mol;
    // the staged sites, in registers for the whole evaluation
    double snx[SS], sny[SS], snz[SS], skq[SS];
    int srow[SS];
#pragma unroll
    for (int a = 0; a < SS; a++) {
        if (a < sm_w) {
            const C2Stage st = v.stage[a];
            snx[a] = st.nx; sny[a] = st.ny; snz[a] = st.nz; skq[a] = st.kq; srow[a] = st.row;
        } else { snx[a] = sny[a] = snz[a] = skq[a] = 0.0; srow[a] = 0; }
    }
    dl = 0.0; sl = 0.0; el = 0.0;
    const double2 *pxy = v.xy + tl;
    const double *pz = v.zs + tl;
    const double *pq = (const double *)(smem + TL::OFF_QPOS) + tl;
    const int *pi = (const int *)(smem + TL::OFF_POS_INFO) + tl;
    double *pe = v.pend + tl;
#pragma unroll 1
    for (int k = 0; k < M; k++) {
        double x[SS], y[SS], z[SS], q[SS];
        int info[SS];
#pragma unroll
        for (int b = 0; b < SS; b++) {
            const double2 xy = pxy[b * TL::XROW];
            x[b] = xy.x; y[b] = xy.y; z[b] = pz[b * TL::XROW]; q[b] = pq[b * L];
            info[b] = (HAS_EXP || ljw) ? pi[b * L] : 0;
        }
        double e = 0.0;
#pragma unroll
        for (int a = 0; a < SS; a++)
            if (a < sm_w) e = trial_site_energy<SS, HAS_EXP>(snx[a], sny[a], snz[a], skq[a], srow[a], ((ljw >> a) & 1u) != 0u, tab_cd, tab_ab, x, y, z, q, info, ljb, e);
        const int n = tl + L * k;
        const bool on = (n != m) & (n < nmol);
        const int lo = n < m ? n : m, hi = n < m ? m : n;
        const double eold = v.cache[on ? tri_index(lo, hi) : 0];
        *pe = e;                                              // waits for the commit
        dl += on ? e - eold : 0.0;                            // eEnd - eStart from per-neighbour differences
        if (TRACE) { sl += on ? eold : 0.0; el += on ? e : 0.0; }
        pxy += SS * TL::XROW; pz += SS * TL::XROW; pq += SS * L; pi += SS * L; pe += L;
    }
}

// The same for small lanes (M * SS site slots <= C2_WIDE_MAX): ALL of this lane's site slots against one staged site at a time
// - M * SS independent evaluations in flight, which hides the FP64 latency from a single warp - in a real loop over the staged
// sites.  Also folds in the pair-cache update of the move committed at the top of this round (`commit_prev`, particle
// `m_prev`): entry (m_prev, n) is written before entry (m, n) is read, by the lane that owns n.
constexpr int C2_WIDE_MAX = 10;
// this lane's charges: in registers for the whole launch, or re-read from shared memory per staged site (frees 2 M SS registers:
// with the two-group schedule, TR_C2_LAG > 0, the 9-wide evaluation of (H2O)20 otherwise spills)
#ifndef TR_C2_Q_SMEM
#define TR_C2_Q_SMEM (TR_C2_LAG > 0)
#endif
constexpr bool C2_Q_IN_REGISTERS = !(TR_C2_Q_SMEM);

template <int L, int M, int SS, bool HAS_EXP, bool TRACE>
__device__ __forceinline__ void c2_evaluate_wide(const unsigned char *smem, const C2View &v, const DevSystem &sys, int m, int sm_w, unsigned ljw,
                                                 unsigned ljb, int tl, const double (&x)[M * SS], const double (&y)[M * SS], const double (&z)[M * SS],
                                                 const double (&q)[M * SS], double &dl, double &sl, double &el)
{
    using TL = C2Tables<L, M, SS>;
    constexpr int NE = M * SS;
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    const int nmol = sys.n_mol;
    // the cached energies this move is measured against (the commit of the previous move is in the cache: the team synchronised)
    double e[M], eold[M];
    bool on[M];
    const int row_m = (m * (m - 1)) >> 1;                     // packed lower triangle: (lo, hi) sits at hi (hi - 1) / 2 + lo
#pragma unroll
    for (int k = 0; k < M; k++) {
        const int n = tl + L * k;
        on[k] = (n != m) & (n < nmol);
        const int idx = n < m ? row_m + n : ((n * (n - 1)) >> 1) + m;      // (n's own row offset is loop-invariant)
        eold[k] = v.cache[on[k] ? idx : 0];
        e[k] = 0.0;
    }
#pragma unroll 1
    for (int a = 0; a < sm_w; a++) {
        const C2Stage st = v.stage[a];
        double r2[NE], y0[NE], t[NE], u[NE], p[NE], ri[NE];
#pragma unroll
        for (int j = 0; j < NE; j++) { const double d = x[j] - st.nx; r2[j] = d * d; }
#pragma unroll
        for (int j = 0; j < NE; j++) { const double d = y[j] - st.ny; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
        for (int j = 0; j < NE; j++) { const double d = z[j] - st.nz; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
        for (int j = 0; j < NE; j++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[j]) : "d"(r2[j]));
#pragma unroll
        for (int j = 0; j < NE; j++) t[j] = r2[j] * y0[j];
#pragma unroll
        for (int j = 0; j < NE; j++) u[j] = fma(-t[j], y0[j], 1.0);
#pragma unroll
        for (int j = 0; j < NE; j++) p[j] = fma(u[j], 0.375, 0.5);
#pragma unroll
        for (int j = 0; j < NE; j++) u[j] = y0[j] * u[j];
#pragma unroll
        for (int j = 0; j < NE; j++) ri[j] = fma(p[j], u[j], y0[j]);
#pragma unroll
        for (int j = 0; j < NE; j++) e[j / SS] = fma(st.kq * (C2_Q_IN_REGISTERS ? q[j] : TL::qpos(smem, j * L + tl)), ri[j], e[j / SS]);
        if (HAS_EXP || ((ljw >> a) & 1u)) {
#pragma unroll
            for (int j = 0; j < NE; j++) {
                if (HAS_EXP || ((ljb >> (j % SS)) & 1u)) {
                    const int cj = info_col(TL::pos_info(smem, j * L + tl));
                    const double2 cd = *(const double2 *)(tab_cd + st.row + cj);
                    const double i2 = ri[j] * ri[j];
                    const double i6 = i2 * i2 * i2;
                    e[j / SS] = fma(fma(cd.y, i6, -cd.x), i6, e[j / SS]);                  // + D/r^12 - C/r^6
                    if (HAS_EXP) {
                        const double2 ab = *(const double2 *)(tab_ab + st.row + cj);
                        e[j / SS] = fma(ab.x, exp(-ab.y * (r2[j] * ri[j])), e[j / SS]);
                    }
                }
            }
        }
    }
    dl = 0.0; sl = 0.0; el = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        v.pend[k * L + tl] = e[k];                            // waits for the commit
        dl += on[k] ? e[k] - eold[k] : 0.0;                   // eEnd - eStart from per-neighbour differences
        if (TRACE) { sl += on[k] ? eold[k] : 0.0; el += on[k] ? e[k] : 0.0; }
    }
}

// Whole-warp teams (L = 32): particle slot by particle slot, each against the staged sites in a real loop, with as many site
// slots in flight as the fullest particle of that slot has sites (warp-uniform: C2Tables::slot_ns) - a slot of single-site ions
// costs one evaluation per staged site, not SS.  The cached energies come from global memory (L2): all M loads are issued
// before the first evaluation and are consumed after the last.
template <int NS, bool HAS_EXP>
__device__ __forceinline__ double c2_slot_loop(const C2Stage *stage, int sm_w, unsigned ljw, unsigned ljb, const unsigned char *tab_cd, const unsigned char *tab_ab,
                                               const double *x, const double *y, const double *z, const double *q, const int *info)
{
    double e = 0.0;
#pragma unroll 1
    for (int a = 0; a < sm_w; a++) {
        const C2Stage st = stage[a];
        e = trial_site_energy<NS, HAS_EXP>(st.nx, st.ny, st.nz, st.kq, st.row, ((ljw >> a) & 1u) != 0u, tab_cd, tab_ab, x, y, z, q, info, ljb, e);
    }
    return e;
}

template <int L, int M, int SS, bool HAS_EXP, bool TRACE>
__device__ __forceinline__ void c2_evaluate_slots(const unsigned char *smem, const C2View &v, const DevSystem &sys, int m, int sm_w, unsigned ljw,
                                                  unsigned ljb, int tl, double &dl, double &sl, double &el)
{
    using TL = C2Tables<L, M, SS>;
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    const int nmol = sys.n_mol;
    const int row_m = (m * (m - 1)) >> 1;
    double eold[M];
#pragma unroll
    for (int k = 0; k < M; k++) {
        const int n = tl + L * k;
        const bool on = (n != m) & (n < nmol);
        eold[k] = v.cache[on ? (n < m ? row_m + n : ((n * (n - 1)) >> 1) + m) : 0];
    }
    double enew = 0.0;
#pragma unroll 1
    for (int k = 0; k < M; k++) {
        const int ns = TL::slot_ns(smem, k);
        double x[SS], y[SS], z[SS], q[SS];
        int info[SS];
#pragma unroll
        for (int b = 0; b < SS; b++) {
            const int p = (k * SS + b) * L + tl;
            const double2 xy = v.xy[TL::xi(p)];
            x[b] = xy.x; y[b] = xy.y; z[b] = v.zs[TL::xi(p)]; q[b] = TL::qpos(smem, p);
            info[b] = (HAS_EXP || ljw) ? TL::pos_info(smem, p) : 0;
        }
        double e;
        if (SS > 3 && ns > 3) e = c2_slot_loop<SS, HAS_EXP>(v.stage, sm_w, ljw, ljb, tab_cd, tab_ab, x, y, z, q, info);
        else if (SS > 1 && ns > 1) e = c2_slot_loop<(SS < 3 ? SS : 3), HAS_EXP>(v.stage, sm_w, ljw, ljb, tab_cd, tab_ab, x, y, z, q, info);
        else e = c2_slot_loop<1, HAS_EXP>(v.stage, sm_w, ljw, ljb, tab_cd, tab_ab, x, y, z, q, info);
        const int n = tl + L * k;
        v.pend[k * L + tl] = e;                               // waits for the commit
        enew += ((n != m) & (n < nmol)) ? e : 0.0;
    }
    // this lane's part of eStart from the cache (Space.java:711-713), of eEnd, and their difference
    double sold = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        const int n = tl + L * k;
        sold += ((n != m) & (n < nmol)) ? eold[k] : 0.0;
    }
    dl = enew - sold; sl = sold; el = enew;
}

// The pair-cache half of a commit (the energies the accepted move left in C2View::pend), when no evaluation follows it.
template <int L, int M>
__device__ __noinline__ void c2_commit_cache(const C2View &v, int n_mol, int m_prev, int tl)
{
#pragma unroll 1
    for (int k = 0; k < M; k++) {
        const int n = tl + L * k;
        if ((n != m_prev) & (n < n_mol)) v.cache[tri_index(n < m_prev ? n : m_prev, n < m_prev ? m_prev : n)] = v.pend[k * L + tl];
    }
}

// Space.calcEnergy (Space.java:648-658) of the committed structure, one team; every particle-pair energy is stored into
// the cache on the way (this = the lower-indexed particle, as Space.calcEnergy has it).
template <int L, int M, int SS, bool HAS_EXP>
__device__ __forceinline__ double c2_total_energy(const unsigned char *smem, const C2View &v, const DevSystem &sys, int tl, unsigned mask)
{
    using TL = C2Tables<L, M, SS>;
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    const int nmol = sys.n_mol;
    double tot = 0.0;
    for (int xm = 0; xm < nmol; xm++) {
        const int sx = TL::mol_ns(smem, xm);
        double e[M];
#pragma unroll
        for (int k = 0; k < M; k++) e[k] = 0.0;
        for (int a = 0; a < sx; a++) {
            const int p = TL::pos_of(xm, a);
            const int infi = TL::pos_info(smem, p);
            const int row = info_type(infi) * sys.row_stride;
            const double xi = v.xy[TL::xi(p)].x, yi = v.xy[TL::xi(p)].y, zi = v.zs[TL::xi(p)];
            const double kqi = K_VALUE * TL::qpos(smem, p);
#pragma unroll
            for (int k = 0; k < M; k++) {
                const int n = tl + L * k;
                if (n > xm && n < nmol) {
#pragma unroll
                    for (int b = 0; b < SS; b++) {
                        const int pj = (k * SS + b) * L + tl;
                        const int cj = info_col(TL::pos_info(smem, pj));
                        const double2 cd = *(const double2 *)(tab_cd + row + cj);
                        double2 ab = make_double2(0.0, 0.0);
                        if (HAS_EXP) ab = *(const double2 *)(tab_ab + row + cj);
                        // an empty site slot of a shorter particle sits far away with a zero charge and the all-zero column
                        e[k] = pair_accumulate<HAS_EXP ? 2 : 1>(e[k], v.xy[TL::xi(pj)].x - xi, v.xy[TL::xi(pj)].y - yi, v.zs[TL::xi(pj)] - zi, 0.0, kqi * TL::qpos(smem, pj), cd, ab);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < M; k++) {
            const int n = tl + L * k;
            if (n > xm && n < nmol) { v.cache[tri_index(xm, n)] = e[k]; tot += e[k]; }
        }
    }
    __syncwarp(mask);
    return crew_team_sum<L>(tot, mask);
}

template <int L, int M, int SS>
__device__ __forceinline__ void c2_sites_to_global(const C2View &v, double *dst, const DevSystem &sys, int tl)
{
    using TL = C2Tables<L, M, SS>;
#pragma unroll
    for (int j = 0; j < M * SS; j++) {
        const int p = j * L + tl, i = sys.pos_site[p];
        if (i >= 0) { dst[3 * i] = v.xy[TL::xi(p)].x; dst[3 * i + 1] = v.xy[TL::xi(p)].y; dst[3 * i + 2] = v.zs[TL::xi(p)]; }
    }
}

// A Space.calcEnergy round (once per point): the energy for C1, the write(n) / writeMovie structure dumps.  Out of line: rare.
template <int L, int M, int SS, bool HAS_EXP>
__device__ __noinline__ void c2_total_round(const unsigned char *smem, const C2View &v, const KParams &P, long long chain, int snap_slot, int movie_slot,
                                            int tl, unsigned mask)
{
    const DevSystem &sys = P.sys;
    const double e = c2_total_energy<L, M, SS, HAS_EXP>(smem, v, sys, tl, mask);
    if (tl == 0) v.cs->etot = e;
    if (snap_slot >= 0)
        c2_sites_to_global<L, M, SS>(v, P.snap_sites + ((size_t)chain * P.max_snaps + snap_slot) * (size_t)sys.n_sites * 3, sys, tl);
    if (movie_slot >= 0)
        c2_sites_to_global<L, M, SS>(v, P.movie_sites + ((size_t)chain * P.max_points + movie_slot) * (size_t)sys.n_sites * 3, sys, tl);
}

template <int L, int M, int SS>
__device__ __forceinline__ void c2_sites_from_global(const C2View &v, const double *src, const DevSystem &sys, int tl)
{
    using TL = C2Tables<L, M, SS>;
#pragma unroll
    for (int j = 0; j < M * SS; j++) {
        const int p = j * L + tl, i = sys.pos_site[p];
        v.xy[TL::xi(p)] = i >= 0 ? make_double2(src[3 * i], src[3 * i + 1]) : make_double2(FAR_AWAY, FAR_AWAY);
        v.zs[TL::xi(p)] = i >= 0 ? src[3 * i + 2] : FAR_AWAY;
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------

struct C2Pick {                      // the S-stream part of a move starting at a given ring position
    double va, vb, vc;
    int m, sm, rot, end;             // end: ring position after its draws
    bool ok;
};

constexpr int C2_CTRL_WARPS = 3;

// Warp w runs on SM sub-partition w % 4: warps 0, 1, 2 = C1, C2, C3 (one control warp on each of three sub-partitions), the
// rest are energy warps.  (Measured alternatives, (H2O)20: C1 and C3 together on sub-partition 0: -17 %; all three control warps
// on sub-partition 0: -13 %; the same with six energy warps, two per remaining sub-partition: -18 %.)
enum { C2_ROLE_C1 = 0, C2_ROLE_C2 = 1, C2_ROLE_C3 = 2, C2_ROLE_E = 3 };
__device__ __forceinline__ int c2_role(int warp) { return warp < C2_CTRL_WARPS ? warp : C2_ROLE_E; }
__device__ __forceinline__ int c2_energy_index(int warp) { return warp - C2_CTRL_WARPS; }

#ifndef TR_CREW2_MINB
#define TR_CREW2_MINB 1
#endif
template <int L, int M, int SS, bool HAS_EXP, bool TRACE, int EW>
__global__ void __launch_bounds__(32 * (C2_CTRL_WARPS + EW), TR_CREW2_MINB)
crew2_kernel(const __grid_constant__ KParams P, const int C)
{
    using TL = C2Tables<L, M, SS>;
    constexpr int TPW = 32 / L;                    // chains an energy warp evaluates side by side
    constexpr int NW = C2_CTRL_WARPS + EW;
    constexpr int NT = 32 * NW;
    static_assert(EW * TPW <= 32, "one control lane per chain");
    static_assert(TR_C2_LAG == 0 || C2_STAMP_WARP < NW, "the lagging group follows the time stamps of an energy warp that exists");
    static_assert(SS <= L, "one lane per site of the moved particle");
    extern __shared__ __align__(16) unsigned char smem[];
    const DevSystem &sys = P.sys;
    TL::load(smem, sys);
    unsigned char *chains = smem + TL::table_bytes(sys.n_types, sys.has_exp);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long chain0 = (long long)blockIdx.x * C;
    const int nloc = (int)((P.n_chains - chain0 < C) ? P.n_chains - chain0 : C);
    const int tl = lane % L;
    const int team = lane / L;
    const unsigned mask = team_mask<L, 1>();
    const int ntri = TL::tri_count(sys.n_mol);

    // chain state -> shared memory (every team takes whole chains)
    for (int c0 = warp * TPW; c0 < nloc; c0 += NW * TPW) {
        const int c = c0 + team;
        if (c < nloc) {
            const long long ch = chain0 + c;
            const C2View v = c2_view<L, M, SS>(chains, c, C, P.pair_cache + (size_t)ch * ntri);
            const int *src = (const int *)(P.ctrl + ch);
            int *dst = (int *)&v.cs->ctrl;
            for (int i = tl; i < (int)(sizeof(ChainCtrl) / 4); i += L) dst[i] = src[i];
            const double *cg = P.com + (size_t)ch * sys.n_mol * 3;
            for (int i = tl; i < sys.n_mol * 3; i += L) v.com[i] = cg[i];
            if constexpr (!TL::GLOBAL_CACHE) {
                const double *pc = P.pair_cache + (size_t)ch * ntri;
                for (int i = tl; i < ntri; i += L) v.cache[i] = pc[i];
            }
            c2_sites_from_global<L, M, SS>(v, P.sites + (size_t)ch * sys.n_sites * 3, sys, tl);
            __syncwarp(mask);
            if (tl == 0) {
                v.cs->sch = P.sched[v.cs->ctrl.sched];
                v.cs->mt = P.mt + (size_t)ch * 3 * 2 * MT_N;
                v.cs->plan[0].mode = PLAN_IDLE; v.cs->plan[1].mode = PLAN_IDLE;
                v.cs->plan[0].nR = 0; v.cs->plan[1].nR = 0;
                v.cs->bound = v.cs->ctrl.space_len * 0.75;
            }
            if (tl < 3) { v.cs->wr[tl] = 0; v.cs->rd[tl] = 0; }
        }
    }
    __syncthreads();

    const int role = c2_role(warp);
    if (role == C2_ROLE_C3) {
        // ================================================================ control warp C3: the random-number generator
        const bool mine = lane < nloc;
        const C2View v = c2_view<L, M, SS>(chains, mine ? lane : 0, C, nullptr);
        C2Chain *cs = v.cs;
        C2Gen gM, gS, gR;
        auto open_stream = [&](C2Gen &g, int s) {
            const MtCursor cur = cs->ctrl.rng[s];
            g.mt = cs->mt + s * 2 * MT_N; g.ring = cs->ring[s];
            g.par = cur.par; g.gen = cur.gen; g.fed = cur.fed; g.wr = 0;
        };
        open_stream(gM, STREAM_M); open_stream(gS, STREAM_S); open_stream(gR, STREAM_R);
        const bool work = mine && !cs->ctrl.done;
        // chunks in flight: slot 0, 1 = the next two chunks of S (a queue: slot qS is the older one), slot 2 = M, slot 3 = R
        C2Issue iM, iS, iR;
        int qS = 0;
        if (work) {
            // prologue: align every stream to a multiple of 4 words, fill every ring but its last 8 words, start the first copies
            while (gM.fed & 3) c2_gen_word(gM);
            while (gS.fed & 3) c2_gen_word(gS);
            while (gR.fed & 3) c2_gen_word(gR);
            iM.par = gM.par; iM.gen = gM.gen; iM.fed = gM.fed;
            iS.par = gS.par; iS.gen = gS.gen; iS.fed = gS.fed;
            iR.par = gR.par; iR.gen = gR.gen; iR.fed = gR.fed;
#pragma unroll 1
            for (int i = 0; i < CREW_RING / 4 - 2; i++) {
                c2_chunk_issue(gM.mt, iM, cs->pre[2]); c2_chunk_issue(gS.mt, iS, cs->pre[0]); c2_chunk_issue(gR.mt, iR, cs->pre[3]);
                asm volatile("cp.async.wait_all;" ::: "memory");
                c2_chunk_finish(gM, cs->pre[2]); c2_chunk_finish(gS, cs->pre[0]); c2_chunk_finish(gR, cs->pre[3]);
            }
            c2_chunk_issue(gS.mt, iS, cs->pre[0]); c2_chunk_issue(gS.mt, iS, cs->pre[1]);
            c2_chunk_issue(gM.mt, iM, cs->pre[2]); c2_chunk_issue(gR.mt, iR, cs->pre[3]);
            cs->wr[STREAM_M] = gM.wr; cs->wr[STREAM_S] = gS.wr; cs->wr[STREAM_R] = gR.wr;
        }
        __syncwarp();
        crew_bar(NT);                                       // "rings filled"
        C2TM(long long tm_work = 0, tm_rounds = 0;)
        for (int r = -1;; r++) {
            C2TM(const long long tm0 = c2_clock();)
            if (work) {
                // words [rd, wr) are unread (rd as C1 published it last round: never ahead of the truth).  Per round at most two
                // chunks of S (a move draws 5 words on average, 11 at most), one of M (always 4) and one of R (3 per rotation).
                // Every chunk finished here was issued a round ago or earlier; its successor is issued right away.
                asm volatile("cp.async.wait_all;" ::: "memory");
                int roomS = (CREW_RING - (gS.wr - cs->rd[STREAM_S])) >> 2;
#pragma unroll 1
                for (int i = 0; i < 2 && roomS > 0; i++, roomS--) {
                    c2_chunk_finish(gS, cs->pre[qS]);
                    c2_chunk_issue(gS.mt, iS, cs->pre[qS]);
                    qS ^= 1;
                }
                if (CREW_RING - (gM.wr - cs->rd[STREAM_M]) >= 4) { c2_chunk_finish(gM, cs->pre[2]); c2_chunk_issue(gM.mt, iM, cs->pre[2]); }
                if (CREW_RING - (gR.wr - cs->rd[STREAM_R]) >= 4) { c2_chunk_finish(gR, cs->pre[3]); c2_chunk_issue(gR.mt, iR, cs->pre[3]); }
                __threadfence_block();                      // ring words before the cursors that announce them
                cs->wr[STREAM_M] = gM.wr; cs->wr[STREAM_S] = gS.wr; cs->wr[STREAM_R] = gR.wr;
            }
            C2TM(tm_work += c2_clock() - tm0; tm_rounds++;)
            if (r < 0) crew_bar(NT);                        // plan[0] is written
            else c2_plan_posted<NW>((r + 1) & 1);
            if (r >= 0) c2_bar_sync(C2_BAR_RES + (r & 1), NT);      // the teams have posted round r
            if (TL::alive(smem, r + 1) == 0) break;
        }
        if (work) asm volatile("cp.async.wait_all;" ::: "memory");       // (the copies in flight are simply dropped: nothing was consumed from them)
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("C3: rounds %lld work %lld\n", tm_rounds, tm_work / tm_rounds);)
        // park the streams: unread ring words go back to the cursors (C1 published its final read positions)
        if (mine) {
            auto park = [&](const C2Gen &g, int s) {
                MtCursor cur;
                cur.par = g.par; cur.gen = g.gen; cur.fed = g.fed;
                if (work) mt_cursor_rewind(cur, g.wr - cs->rd[s]);
                cs->ctrl.rng[s] = cur;
            };
            park(gM, STREAM_M); park(gS, STREAM_S); park(gR, STREAM_R);
        }
    } else if (role == C2_ROLE_C1) {
        // ================================================================ control warp C1: lane c <-> chain c
        // S stream, bookkeeping, modes.  Iteration r runs while the teams execute round r: it settles what they did in round
        // r - 1, adopts the move of round r, and writes its part of plan[(r + 1) & 1].
        const bool mine = lane < nloc;
        const C2View v = c2_view<L, M, SS>(chains, mine ? lane : 0, C, nullptr);
        C2Chain *cs = v.cs;
        ChainCtrl *c = &cs->ctrl;
        const tr_schedule &sch = cs->sch;
        const long long chain = chain0 + lane;
        const int S = sys.n_sites;

        double t = 0.0, kBt = 0.0, running = 0.0, totalE = 0.0, totalSqE = 0.0, maxD = 0.0, maxTmag = 0.0;
        unsigned long long thr_rot = 0, thr_trans = 0;
        int z = 0, mpp = 1;
        int seg_moves = 0, seg_acc = 0, seg_eval = 0;
        unsigned seg_sm = 0;                                       // executed pair evaluations = S * sum s_m - sum s_m^2 (trial positions only)
        unsigned seg_sm2 = 0;
        int rdM = 0, rdS = 0, rdR = 0, wrM = 0, wrS = 0, wrR = 0;
        long long budget = P.max_moves > 0 ? P.max_moves : 0x7fffffffffffffffLL;
        bool live = mine && !c->done;
        bool hot = false;
        int mode_prev = PLAN_IDLE, mode_cur = PLAN_IDLE;
        bool total_for_init = false;
        int starved = 0;                                         // consecutive rounds without the random words for the next move
        int commit_pending = 0;
        int cur_info = 0, prev_info = 0;                         // m | sm << 16 | rot << 21 | magwalk << 22 | (axis + 1) << 23
        bool metro_ok = false;
        double rnd = 0.0, thr = 0.0, grd = 0.0;
        int kindbit = 0, magw_rot = 0, magw_trans = 0;           // M-stream draws of the move being planned
        int adopt_magw_rot = 0, adopt_magw_trans = 0;            // ... of the move posted last (adopted one iteration later)
        C2Pick cA, cB;
        cA.va = cA.vb = cA.vc = 0.0; cA.m = cA.sm = cA.rot = cA.end = 0; cA.ok = false;
        cB = cA;

        auto load_hot = [&]() {
            t = c->t; kBt = __dmul_rn(BOLTZMANN_CONSTANT, c->t);
            running = c->running; totalE = c->totalE; totalSqE = c->totalSqE;
            maxD = c->maxD; maxTmag = __dmul_rn(c->maxD, sch.magwalk_factor_trans);
            thr_rot = double_threshold52(c->probRot); thr_trans = double_threshold52(c->probTrans);
            z = c->z; mpp = sch.moves_per_point; hot = true;
        };
        auto flush_hot = [&]() {
            c->z = z; c->running = running; c->totalE = totalE; c->totalSqE = totalSqE;
            c->total += seg_moves;
            c->accepted += seg_acc; c->accepted_all += seg_acc;
            c->moves += seg_moves; c->evaluated += seg_eval;
            c->pair_evals += (long long)S * seg_sm - seg_sm2;
            seg_moves = seg_acc = seg_eval = 0; seg_sm = seg_sm2 = 0;
        };
        auto ringM = [&](int pos) { return cs->ring[STREAM_M][pos & (CREW_RING - 1)]; };
        auto ringS = [&](int pos) { return cs->ring[STREAM_S][pos & (CREW_RING - 1)]; };
        // Main.java:199-213 at the current M position: r.nextDouble() >= 0.5 is the top bit of the first word; the second
        // double is compared as a 52-bit integer against ceil(prob * 2^52)
        auto draw_kind = [&]() {
            const uint32_t w0 = ringM(rdM);
            const unsigned long long u2 = ((unsigned long long)(ringM(rdM + 2) >> 6) << 26) | (ringM(rdM + 3) >> 6);
            kindbit = (int)(w0 >> 31);
            magw_rot = (u2 >= thr_rot) ? 0 : 1;
            magw_trans = (u2 >= thr_trans) ? 0 : 1;
        };
        auto draw_pick = [&](C2Pick &k, int pos) {
            k.ok = false;
            int p = pos, pick;                         // Space.java:736: r.nextInt(moveableMolecules.size())
            const int n = (int)sys.pick_mod.n;
            if ((n & -n) == n) {
                if (p >= wrS) return;
                pick = (int)(((long long)n * (long long)(ringS(p) >> 1)) >> 31);
                p++;
            } else for (;;) {
                if (p >= wrS) return;
                const int bits = (int)(ringS(p) >> 1);
                p++;
                pick = (int)fastmod(sys.pick_mod, (uint32_t)bits);
                if ((int)((uint32_t)bits - (uint32_t)pick + (uint32_t)(n - 1)) >= 0) break;
            }
            k.m = TL::moveable(smem, pick);
            k.sm = TL::mol_ns(smem, k.m);
            k.rot = (kindbit && k.sm > 1) ? 1 : 0;
            k.va = k.vb = k.vc = 0.0;
            if (!k.rot) {
                if (p + 6 > wrS) return;
                const double maxT = magw_trans ? maxTmag : maxD;
                const double ux = crew_next_double(ringS(p), ringS(p + 1));                          // Space.java:694-696
                const double uy = crew_next_double(ringS(p + 2), ringS(p + 3));
                const double uz = crew_next_double(ringS(p + 4), ringS(p + 5));
                k.va = __dsub_rn(__dmul_rn(__dmul_rn(ux, 2.0), maxT), maxT);
                k.vb = __dsub_rn(__dmul_rn(__dmul_rn(uy, 2.0), maxT), maxT);
                k.vc = __dsub_rn(__dmul_rn(__dmul_rn(uz, 2.0), maxT), maxT);
                p += 6;
            }
            if (p + CREW_SETTLE_SLACK > wrS) return;      // the move's own Metropolis double must already be in the ring
            k.end = p;
            k.ok = true;
        };
        auto write_cand = [&](const C2Pick &k, C2Cand *dst) {
            C2Cand d;
            d.va = k.va; d.vb = k.vb; d.vc = k.vc; d.m = k.m; d.info = k.sm | (k.rot << 8);
            *dst = d;
        };
        // The move evaluated in round r becomes the current one: cursors move past its draws.  Its rotation draw (axis, R
        // words) was posted by C2 one iteration ago.
        auto adopt = [&](const C2Pick &k, const C2Plan *plc) {
            const int magwalk = k.rot ? adopt_magw_rot : adopt_magw_trans;
            const int axis = k.rot ? plc->axis : -1;
            cur_info = k.m | (k.sm << 16) | (k.rot << 21) | (magwalk << 22) | ((axis + 1) << 23);
            rdS = k.end; rdM += C2_RING_NEED_M;
            if (k.rot) {
                const int nR = plc->nR;
                if (nR < 0) c->done = CHAIN_ABORT_RNG; else rdR += nR;
            }
        };
        if (live) load_hot();
        C2TM(long long tm_work = 0, tm_rounds = 0;)
        crew_bar(NT);                                       // "rings filled"
        C2TM(const long long tm_begin = c2_clock();)

        for (int r = -1;; r++) {
            C2TM(const long long tm0 = c2_clock();)
            C2Plan *pl = &cs->plan[(r + 1) & 1];
            const C2Plan *plc = &cs->plan[r & 1];
            int mode = PLAN_IDLE;
            bool bubble = false;                                  // nothing is posted for this chain this round, but it is not finished
            if (live) {
                wrM = cs->wr[STREAM_M]; wrS = cs->wr[STREAM_S]; wrR = cs->wr[STREAM_R];
                // ---------------------------------------------- settle round r - 1
                if (mode_prev == PLAN_AUTO || mode_prev == PLAN_MOVE) {
                    const CrewResult *res = &cs->result[(r - 1) & 1];
                    const int oob = res->oob;
                    const double dE = oob ? 0.0 : res->dE;
                    int acc, drew;
                    if (metro_ok) { rnd = plc->rnd; thr = plc->thr; grd = plc->grd; }           // C2 posted them during round r - 1
                    else if (!oob && dE > 0.0) rnd = crew_next_double(ringS(rdS), ringS(rdS + 1));   // in the ring: CREW_SETTLE_SLACK
                    crew_metropolis(oob, dE, t == 0.0, metro_ok, rnd, thr, grd, kBt, acc, drew);
                    if (acc) commit_pending = 1;
                    rdS += 2 * drew;
                    running = __dadd_rn(running, acc ? dE : 0.0);                                 // Main.java:215-223
                    if (!sch.static_temp || z > sch.eq_configs) {
                        totalE = __dadd_rn(totalE, running);
                        totalSqE = __dadd_rn(totalSqE, __dmul_rn(running, running));
                    }
                    seg_acc += acc;
                    if (!oob) {
                        const int sm = (prev_info >> 16) & 31;
                        seg_eval++; seg_sm += (unsigned)sm; seg_sm2 += (unsigned)(sm * sm);
                    }
                    if (TRACE) {
                        const long long idx = c->moves + seg_moves;
                        if (idx < P.trace_moves) {
                            tr_trace_rec tr;
                            tr.mol = prev_info & 0xffff; tr.kind = (int8_t)((prev_info >> 21) & 1); tr.magwalk = (int8_t)((prev_info >> 22) & 1);
                            tr.bounds = (int8_t)oob; tr.accepted = (int8_t)acc; tr.axis = ((prev_info >> 23) & 7) - 1; tr.drew_rand = drew;
                            tr.e_start = oob ? 0.0 : res->eS; tr.e_end = oob ? 0.0 : res->eE; tr.dE = dE;
                            tr.rand = drew ? rnd : __longlong_as_double(0x7ff8000000000000LL);
                            tr.temp = t; tr.running = running;
                            P.trace[chain * P.trace_moves + idx] = tr;
                        }
                    }
                    seg_moves++; z++;
                    if (seg_moves >= (1 << 20)) flush_hot();
                    if (mode_cur == PLAN_AUTO) {
                        // the team is already evaluating the successor it chose by the same test
                        adopt(drew ? cB : cA, plc);
                        commit_pending = 0;
                    }
                } else if (mode_prev == PLAN_TOTAL) {
                    const double etot = cs->etot;
                    if (total_for_init) {
                        // Main.java:64,71: write(0) and startingEnergy = calcEnergy()
                        const int k = c->n_snaps;
                        if (k < P.max_snaps) { P.snap_energy[chain * P.max_snaps + k] = etot; P.snap_tooth[chain * P.max_snaps + k] = 0; }
                        c->n_snaps = k + 1;
                        if (etot < c->min_energy) { c->min_energy = etot; c->min_tooth = 0; }
                        c->start_energy = etot; c->running = etot; c->started = 1;
                        c->delT = c->t / (double)(c->ptsPerTooth - 1);      // Main.java:188 for tooth 0
                        total_for_init = false;
                    } else {
                        crew_finish_point(P, c, sch, chain, etot, true);
                    }
                    load_hot();
                    bubble = true;                                 // t / maxRot changed under C2's eyes: post the next move a round later
                }
                if (mode_cur == PLAN_MOVE) adopt(cA, plc);
                metro_ok = false;

                // ---------------------------------------------- plan round r + 1
                int flags = 0, snap_slot = -1, movie_slot = -1;
                if (mode_cur == PLAN_AUTO || mode_cur == PLAN_MOVE) {
                    // a move is being evaluated now; its Metropolis double sits at rdS (C2 forms the threshold from it) and both
                    // possible successors are drawn here; if all of that exists, the team may go on by itself
                    bool chain_ok = false;
                    metro_ok = !c->done;                           // (the double is in the ring: CREW_SETTLE_SLACK at posting time)
                    if (metro_ok && z + 1 < mpp && budget > 0 && wrM - rdM >= C2_RING_NEED_M && wrR - rdR >= C2_RING_NEED_R) {
                        draw_kind();
                        draw_pick(cA, rdS); draw_pick(cB, rdS + 2);
                        chain_ok = cA.ok && cB.ok;
                    }
                    if (chain_ok) {
                        mode = PLAN_AUTO;
                        write_cand(cA, &pl->cand[0]); write_cand(cB, &pl->cand[1]);
                        pl->endS0 = cA.end; pl->endS1 = cB.end;
                        adopt_magw_rot = magw_rot; adopt_magw_trans = magw_trans;
                        flags = (t == 0.0) ? PF_TZERO : 0;
                        budget--;
                    }
                    prev_info = cur_info;
                } else if (mode_cur == PLAN_IDLE && !bubble) {
                    // nothing is running: everything that is not "next move of the same point" is decided here
                    if (!c->done && !c->started) {
                        mode = PLAN_TOTAL; total_for_init = true;
                        snap_slot = (P.snap_sites && c->n_snaps < P.max_snaps) ? c->n_snaps : -1;
                    } else if (!c->done && z >= mpp) {
                        flush_hot();
                        if (crew_point_needs_total(P, c, sch, snap_slot, movie_slot)) mode = PLAN_TOTAL;
                        else {
                            crew_finish_point(P, c, sch, chain, 0.0, false);
                            load_hot();
                            bubble = true;
                        }
                    }
                    if (!c->done && mode == PLAN_IDLE && !bubble && budget > 0) {
                        bool ok = wrM - rdM >= C2_RING_NEED_M && wrR - rdR >= C2_RING_NEED_R;
                        if (ok) { draw_kind(); draw_pick(cA, rdS); ok = cA.ok; }
                        if (ok) {
                            write_cand(cA, &pl->cand[0]);
                            pl->endS0 = cA.end; pl->endS1 = cA.end;
                            adopt_magw_rot = magw_rot; adopt_magw_trans = magw_trans;
                            mode = PLAN_MOVE; budget--; starved = 0;
                        } else if (++starved < C2_MAX_BUBBLES) bubble = true;      // C3 is still refilling: try again next round
                        else c->done = CHAIN_ABORT_RNG;
                    }
                    if (mode == PLAN_IDLE && commit_pending) mode = PLAN_SYNC;   // (parked with an accepted move not yet committed)
                    if (mode != PLAN_IDLE) {
                        flags = commit_pending ? PF_COMMIT : 0;
                        commit_pending = 0;
                    }
                    if (c->done && mode == PLAN_IDLE && !bubble) live = false;
                }
                // (PLAN_TOTAL / PLAN_SYNC running now, or a bubble: wait)
                pl->mode = mode; pl->flags = flags; pl->snap_slot = snap_slot; pl->movie_slot = movie_slot;
                pl->kBt = kBt;
                cs->rd[STREAM_M] = rdM; cs->rd[STREAM_S] = rdS; cs->rd[STREAM_R] = rdR;      // for C3: everything before these is consumed
            } else if (mine) pl->mode = PLAN_IDLE;
            const unsigned any = __ballot_sync(0xffffffffu, mode != PLAN_IDLE || mode_cur != PLAN_IDLE || (live && bubble));
            mode_prev = mode_cur; mode_cur = mode;
            if (lane == 0) TL::alive(smem, r + 1) = (int)any;
            C2TM(tm_work += c2_clock() - tm0; tm_rounds++;)
            if (r < 0) crew_bar(NT);                        // plan[0] is written
            else c2_plan_posted<NW>((r + 1) & 1);
            if (r >= 0) c2_bar_sync(C2_BAR_RES + (r & 1), NT);      // the teams have posted round r
            if (!any) break;
        }
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("C1: rounds %lld work %lld total/round %lld\n", tm_rounds, tm_work / tm_rounds, (c2_clock() - tm_begin) / tm_rounds);)
        // park the chain: counters into c (C3 hands the unread ring words back to the stream cursors)
        if (mine && hot) flush_hot();
    } else if (role == C2_ROLE_C2) {
        // ================================================================ control warp C2: lane c <-> chain c
        // Everything transcendental.  Iteration r: (a) which move is in flight in round r follows from the plan C1 posted
        // for it and from the sign of the dE of round r - 1; (b) its Metropolis double and the threshold -kT log(rnd) on dE;
        // (c) the rotation draw of the NEXT move (Molecule.java:67-68) with sincos.  Both go into plan[(r + 1) & 1].
        const bool mine = lane < nloc;
        const C2View v = c2_view<L, M, SS>(chains, mine ? lane : 0, C, nullptr);
        C2Chain *cs = v.cs;
        const ChainCtrl *c = &cs->ctrl;
        const FastMod three = fastmod_make(3);
        int rdM = 0, rdR = 0, nR_next = 0;                       // nR_next: R words the rotation draw posted last consumes
        auto ringM = [&](int pos) { return cs->ring[STREAM_M][pos & (CREW_RING - 1)]; };
        auto ringS = [&](int pos) { return cs->ring[STREAM_S][pos & (CREW_RING - 1)]; };
        auto ringR = [&](int pos) { return cs->ring[STREAM_R][pos & (CREW_RING - 1)]; };
        C2TM(long long tm_work = 0, tm_rounds = 0;)
        crew_bar(NT);                                       // "rings filled"
        for (int r = -1;; r++) {
            C2TM(const long long tm0 = c2_clock();)
            if (mine) {
                C2Plan *pl = &cs->plan[(r + 1) & 1];
                const C2Plan *plc = &cs->plan[r & 1];
                const int mode_cur = r < 0 ? PLAN_IDLE : plc->mode;
                double rnd = 0.0, thr = 0.0, grd = 0.0;
                if (mode_cur == PLAN_AUTO || mode_cur == PLAN_MOVE) {
                    int which = 0;
                    if (mode_cur == PLAN_AUTO) {
                        const CrewResult *res = &cs->result[(r - 1) & 1];
                        which = (!res->oob && res->dE > 0.0) ? 1 : 0;       // a Metropolis double was drawn for the move of round r - 1
                    }
                    const int info = plc->cand[which].info;
                    const int endS = which ? plc->endS1 : plc->endS0;
                    rdM += C2_RING_NEED_M;
                    if ((info >> 8) & 1) rdR += nR_next > 0 ? nR_next : 0;
                    // Space.java:676-687: rand > exp(-dE / kT)  <=>  dE > -kT log(rand)
                    const double t = c->t;
                    const double kBt = __dmul_rn(BOLTZMANN_CONSTANT, t);
                    rnd = crew_next_double(ringS(endS), ringS(endS + 1));
                    if (t != 0.0) {
                        thr = __dmul_rn(-kBt, log(rnd));
                        grd = (kBt + fabs(thr)) * 1e-11;
                    }
                }
                // the next move's M and R draws (Main.java:199-205, Molecule.java:67-68); C1 only posts a move when the words
                // read here are in the rings
                double sn = 0.0, csn = 1.0;
                int axis = -1, nR = 0;
                {
                    const uint32_t w0 = ringM(rdM);
                    if (w0 >> 31) {
                        const unsigned long long u2 = ((unsigned long long)(ringM(rdM + 2) >> 6) << 26) | (ringM(rdM + 3) >> 6);
                        const double maxR = (u2 >= double_threshold52(c->probRot)) ? c->maxRot : JAVA_TWO_PI;
                        const double u = crew_next_double(ringR(rdR), ringR(rdR + 1));                       // Molecule.java:67
                        const double rotAmount = __dmul_rn(__dmul_rn(__dsub_rn(u, 0.5), 2.0), maxR);
                        int n = 2;
                        nR = -1;
                        for (; n < C2_RING_NEED_R;) {                                                        // Molecule.java:68: nextInt(3)
                            const int bits = (int)(ringR(rdR + n) >> 1);
                            n++;
                            axis = (int)fastmod(three, (uint32_t)bits);
                            if ((int)((uint32_t)bits - (uint32_t)axis + 2u) >= 0) { nR = n; break; }
                        }
                        sincos(rotAmount, &sn, &csn);
                    }
                }
                nR_next = nR;
                pl->rnd = rnd; pl->thr = thr; pl->grd = grd; pl->sn = sn; pl->cs = csn; pl->axis = axis; pl->nR = nR;
            }
            C2TM(tm_work += c2_clock() - tm0; tm_rounds++;)
            if (r < 0) crew_bar(NT);                        // plan[0] is written
            else c2_plan_posted<NW>((r + 1) & 1);
            if (r >= 0) c2_bar_sync(C2_BAR_RES + (r & 1), NT);      // the teams have posted round r
            if (TL::alive(smem, r + 1) == 0) break;
        }
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("C2: rounds %lld work %lld\n", tm_rounds, tm_work / tm_rounds);)
    } else {
        // ================================================================ energy warps: one team <-> one chain
        const int c = c2_energy_index(warp) * TPW + team;
        const bool has = c < nloc;
        const C2View v = c2_view<L, M, SS>(chains, has ? c : 0, C, P.pair_cache + (size_t)(chain0 + (has ? c : 0)) * ntri);
        C2Chain *cs = v.cs;
        unsigned ljb = 0;                                   // site slots b that hold a Lennard-Jones site in ANY particle (block-uniform)
        for (int p = 0; p < TL::NPOS; p++)
            if (info_valid(TL::pos_info(smem, p)) && info_lj(TL::pos_info(smem, p))) ljb |= 1u << ((p / L) % SS);
        C2Pending pend;
        pend.tx = pend.ty = pend.tz = pend.cx = pend.cy = pend.cz = 0.0;
        pend.pp = -1; pend.pm = -1; pend.m = -1;
        // this lane's charges: chain-independent, in registers for the whole launch (wide evaluation only)
        constexpr int WN = TL::WIDE ? M * SS : 1;
        double wq[WN];
#pragma unroll
        for (int j = 0; j < WN; j++) wq[j] = C2_Q_IN_REGISTERS ? TL::qpos(smem, j * L + tl) : 0.0;
        double last_dE = 0.0;
        int last_oob = 0;
        const int grp = c2_group(warp);
        crew_bar(NT);                                       // "rings filled"
        crew_bar(NT);                                       // plan[0] is written
        C2TM(long long tm_work = 0, tm_rounds = 0, tm_a = 0, tm_b = 0;)
        // Start the two energy warps of an SM sub-partition (warp ids w and w + 4) half an evaluation apart: nothing re-aligns them
        // (they only wait for the control warps), and their FP64 bursts overlap less (measured: + 3 %).
        if ((warp >> 2) & 1) { const long long t0 = clock64(); while (clock64() - t0 < 1000) { } }
        for (int r = 0;; r++) {
            if (r > 0) {                                    // plan r is posted
                if (TR_C2_PER_WARP && EW <= 10) c2_bar_sync(C2_BAR_PLAN_W + c2_energy_index(warp), 32 * (C2_CTRL_WARPS + 1));
                else c2_bar_sync(grp ? C2_BAR_PLAN_B + (r & 1) : C2_BAR_PLAN + (r & 1), c2_plan_threads<NW>(grp));
            }
            if (TL::alive(smem, r) == 0) break;
            if (TR_C2_LAG > 0) {
                if (warp == C2_STAMP_WARP) {
                    if (lane == 0) *TL::stamp(smem, r) = ((unsigned long long)(unsigned)r << 32) | (unsigned)clock();
                } else if (grp) {
                    unsigned long long st;
                    do st = *TL::stamp(smem, r); while ((unsigned)(st >> 32) != (unsigned)r);
                    constexpr unsigned LAG = (M * SS <= 5) ? TR_C2_LAG * 5 / 9 : TR_C2_LAG;      // shorter rounds, shorter lag
                    while ((unsigned)clock() - (unsigned)st < LAG) { }
                }
            }
            C2TM(const long long tm0 = c2_clock();)
            const C2Plan *pl = &cs->plan[r & 1];
            const int4 hdr = *(const int4 *)&pl->mode;       // mode, flags, snap_slot, movie_slot
            const int mode = has ? hdr.x : PLAN_IDLE;
            const bool moving = mode == PLAN_AUTO || mode == PLAN_MOVE;
            // Which successor was posted for "a Metropolis double was drawn for the move of last round" is known without the plan
            // (Space.java:720: drawn iff dE > 0): its descriptor is fetched while the Metropolis test is still being decided.
            const int which = (mode == PLAN_AUTO && !last_oob && last_dE > 0.0) ? 1 : 0;
            const C2Cand d = pl->cand[which];
            int m = -1, sm = 0, lj = 0, commit = 0;
            const int m_prev = pend.m;
            double wx[WN], wy[WN], wz[WN];
#pragma unroll
            for (int j = 0; j < WN; j++) wx[j] = wy[j] = wz[j] = 0.0;     // (dead stores: every path that reads them loads them first)
            bool oob = false;
            if (mode != PLAN_IDLE) {
                const int flags = hdr.y;
                commit = flags & PF_COMMIT;
                if (mode == PLAN_AUTO) {
                    int acc, drew;
                    crew_metropolis(last_oob, last_dE, (flags & PF_TZERO) != 0, true, pl->rnd, pl->thr, pl->grd, pl->kBt, acc, drew);
                    commit = acc;
                }
                if (commit) {
                    // setTemps (Molecule.java:128-135) of the move staged last, and its per-neighbour energies into the pair cache
                    if (pend.pp >= 0) { v.xy[pend.pp] = make_double2(pend.tx, pend.ty); v.zs[pend.pp] = pend.tz; }
                    if (tl == 0 && pend.pm >= 0) { v.com[3 * pend.pm] = pend.cx; v.com[3 * pend.pm + 1] = pend.cy; v.com[3 * pend.pm + 2] = pend.cz; }
                    if constexpr (TL::WIDE) {
                        const int row_p = (m_prev * (m_prev - 1)) >> 1;
#pragma unroll
                        for (int k = 0; k < M; k++) {
                            const int n = tl + L * k;
                            if ((n != m_prev) & (n < sys.n_mol)) v.cache[n < m_prev ? row_p + n : ((n * (n - 1)) >> 1) + m_prev] = v.pend[k * L + tl];
                        }
                    } else c2_commit_cache<L, M>(v, sys.n_mol, m_prev, tl);
                }
            }
            __syncwarp();                                   // the commit is visible to the whole team (all lanes of the warp meet here)
            // this lane's own coordinates: the loads fly while the trial geometry is built (the slot of the moved particle
            // itself is discarded, so the rotateTemp round trip below does not matter here)
            if constexpr (TL::WIDE) {
#pragma unroll
                for (int j = 0; j < WN; j++) {
                    const double2 xy = v.xy[j * TL::XROW + tl];
                    wx[j] = xy.x; wy[j] = xy.y; wz[j] = v.zs[j * TL::XROW + tl];
                }
            }
            if (moving) {
                m = d.m; sm = d.info & 0xff;
                oob = c2_geometry<L, M, SS>(smem, v, sys, d, pl, tl, pend, lj);
                // a team with fewer staged sites than another team of the warp: neutral entries up to the common bound
                if (tl >= sm && tl < SS) {
                    C2Stage st;
                    st.nx = st.ny = st.nz = -FAR_AWAY;       // (not +FAR_AWAY: empty site slots sit there, and r = 0 would give NaN)
                    st.kq = 0.0; st.row = 0; st.lj = 0; st.pad_[0] = st.pad_[1] = 0;
                    v.stage[tl] = st;
                }
            } else if (mode == PLAN_TOTAL) c2_total_round<L, M, SS, HAS_EXP>(smem, v, P, chain0 + c, hdr.z, hdr.w, tl, mask);
            C2TM(const long long tma = c2_clock(); tm_a += tma - tm0;)
            // warp-wide: who is out of bounds (per team), the most staged sites of any team, which staged sites are Lennard-Jones
            const unsigned oob_w = __ballot_sync(0xffffffffu, oob);
            const unsigned lj_w = __ballot_sync(0xffffffffu, lj != 0);
            __syncwarp();                                   // staging stores before the reads below
            const bool team_oob = (oob_w & mask) != 0u;
            const bool eval = moving && !team_oob;
            const int sm_w = __reduce_max_sync(0xffffffffu, eval ? sm : 0);
            unsigned ljw = 0;                               // bit a: staged site a is a Lennard-Jones site in some team
#pragma unroll
            for (int t4 = 0; t4 < TPW; t4++) ljw |= (lj_w >> (t4 * L)) & ((1u << SS) - 1u);
            if (sm_w > 0) {
                double dl = 0.0, sl = 0.0, el = 0.0;
                if (eval) {
                    if constexpr (TL::WIDE) c2_evaluate_wide<L, M, SS, HAS_EXP, TRACE>(smem, v, sys, m, sm_w, ljw, ljb, tl, wx, wy, wz, wq, dl, sl, el);
                    else if constexpr (TL::GLOBAL_CACHE) c2_evaluate_slots<L, M, SS, HAS_EXP, TRACE>(smem, v, sys, m, sm_w, ljw, ljb, tl, dl, sl, el);
                    else c2_evaluate<L, M, SS, HAS_EXP, TRACE>(smem, v, sys, m, sm_w, ljw, ljb, tl, dl, sl, el);
                }
                const double dE = crew_team_sum<L>(dl, mask);
                if (TRACE) { sl = crew_team_sum<L>(sl, mask); el = crew_team_sum<L>(el, mask); }
                if (eval) {
                    if (tl == 0) {
                        CrewResult *res = &cs->result[r & 1];
                        res->oob = 0; res->dE = dE;
                        if (TRACE) { res->eS = sl; res->eE = el; }
                    }
                    last_dE = dE; last_oob = 0;
                }
            }
            if (moving && team_oob) {
                if (tl == 0) cs->result[r & 1].oob = 1;
                last_dE = 0.0; last_oob = 1;
            }
            C2TM(tm_b += c2_clock() - tma; tm_work += c2_clock() - tm0; tm_rounds++;)
            c2_bar_arrive(C2_BAR_RES + (r & 1), NT);        // results of round r are posted
        }
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("E%d: rounds %lld work %lld (plan+commit+geometry %lld, evaluate %lld)\n", warp, tm_rounds, tm_work / tm_rounds, tm_a / tm_rounds, tm_b / tm_rounds);)
    }
    __syncthreads();

    // chain state -> global memory
    for (int c0 = warp * TPW; c0 < nloc; c0 += NW * TPW) {
        const int c = c0 + team;
        if (c >= nloc) continue;
        const long long ch = chain0 + c;
        const C2View v = c2_view<L, M, SS>(chains, c, C, P.pair_cache + (size_t)ch * ntri);
        int *dst = (int *)(P.ctrl + ch);
        const int *src = (const int *)&v.cs->ctrl;
        for (int i = tl; i < (int)(sizeof(ChainCtrl) / 4); i += L) dst[i] = src[i];
        double *cg = P.com + (size_t)ch * sys.n_mol * 3;
        for (int i = tl; i < sys.n_mol * 3; i += L) cg[i] = v.com[i];
        if constexpr (!TL::GLOBAL_CACHE) {
            double *pc = P.pair_cache + (size_t)ch * ntri;
            for (int i = tl; i < ntri; i += L) pc[i] = v.cache[i];
        }
        c2_sites_to_global<L, M, SS>(v, P.sites + (size_t)ch * sys.n_sites * 3, sys, tl);
    }
}

}  // namespace tr
