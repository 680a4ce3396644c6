// The annealing hot path on the device: Main.sawtoothAnneal (reference Main.java:177-263),
// Space.move / Space.rotate (Space.java:660-734), Molecule.rotateTemp (Molecule.java:64-97),
// the Atom pair potential (Atom.java:91-119) and Space.calcEnergy (Space.java:648-658).
//
// Mapping.  A chain (one independent annealing run = one reference JVM process) is owned by
// a TEAM of W warps (W = 1: warp per chain, several chains per block; W > 1: block per
// chain).  The sites of `space` live in REGISTERS for the whole launch: position p of a
// host-chosen permutation sits in team thread (p mod 32W), slot (p div 32W).  The permutation
// puts the sites that have Lennard-Jones parameters first, so whole slots are Coulomb-only
// and skip the r^-6 / r^-12 arithmetic (warp-uniform decision per staged site x slot).
// Per move the moved particle's old and trial site coordinates are staged in shared memory
// (<= TR_MAX_MOL_SITES entries, read back as broadcasts), every thread evaluates the site
// pairs between its own sites and the staged ones in FP64, and eEnd - eStart is reduced with
// shuffles (plus one shared-memory hop when W > 1).  The three MT19937 streams stay in global
// memory (L2-resident) behind 64-word rings of tempered words in shared memory (mt19937.cuh).
// The schedule cursor lives in shared memory, so a launch can stop after any move and the
// next one resumes bit-identically.
#pragma once
#include <math.h>
#include "device_types.cuh"
#include "mt19937.cuh"

namespace tr {

struct __align__(16) StageSite {
    double ox, oy, oz;   // committed coordinates (after the rotateTemp round trip)
    double nx, ny, nz;   // trial coordinates (Atom.tempx/y/z)
    double kq;           // Atom.kTimesQ = K_VALUE * q
    int type;
    int lj;              // type has Lennard-Jones parameters with some type
};

struct MoveDesc {        // W > 1 only: what warp 0 decided, broadcast through shared memory
    double a, b, c;      // translation vector, or (sin, cos, unused)
    int m;
    int rot;
    int axis;
    int magwalk;
};

template <int SPL>
struct SiteRegs {
    double x[SPL], y[SPL], z[SPL], q[SPL];
    int info[SPL];       // mol | type << 16 | index-in-particle << 24 | has-LJ << 29, or INFO_UNUSED
};

constexpr int INFO_UNUSED = 0xffff;   // mol field of a slot that holds no site (type 0, so table lookups stay in range)
__device__ __forceinline__ int info_mol(int info) { return info & 0xffff; }
__device__ __forceinline__ bool info_valid(int info) { return (info & 0xffff) != INFO_UNUSED; }
__device__ __forceinline__ int info_type(int info) { return (info >> 16) & 0xff; }
__device__ __forceinline__ int info_atom(int info) { return (info >> 24) & 0x1f; }
__device__ __forceinline__ int info_lj(int info) { return (info >> 29) & 1; }

// ---- the pair potential --------------------------------------------------------------------
// Atom.java:102-103: (kTimesQ*q/r) + A*exp(-1*B*r) - (C/pow(r,6)) + (D/pow(r,12)).  Evaluated
// through 1/r so no divide or sqrt is issued: rsqrt seed (MUFU.RSQ64H, ~22 bits) + one cubic
// Newton step (error ~e^3, below half an ulp), then r^-6 by multiplication.  Every operation
// is FP64; the result differs from the reference's pow/sqrt/divide form by a few ulp per term
// (bar: 1e-12 relative).  MODE 0: Coulomb only (C = D = A = 0 for the whole slot), 12 FP64
// instructions; MODE 1: + Lennard-Jones, 17; MODE 2: + Born-Mayer exp.
__device__ __forceinline__ double rsqrt_f64(double r2)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    double t = r2 * y;
    double e = fma(-t, y, 1.0);            // 1 - r2*y^2
    double p = fma(e, 0.375, 0.5);         // 1/2 + 3/8 e
    double ye = y * e;
    return fma(p, ye, y);                  // y (1 + e/2 + 3e^2/8)
}

template <int MODE>
__device__ __forceinline__ double pair_accumulate(double acc, double dx, double dy, double dz, double kqq, double2 cd, double2 ab)
{
    const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    const double ri = rsqrt_f64(r2);
    double e = fma(kqq, ri, acc);
    if (MODE >= 1) {
        const double ri2 = ri * ri;
        const double ri6 = ri2 * ri2 * ri2;
        e = fma(fma(cd.y, ri6, -cd.x), ri6, e);        // + D/r^12 - C/r^6
    }
    if (MODE >= 2) e = fma(ab.x, exp(-ab.y * (r2 * ri)), e);
    return e;
}

// ---- team primitives -------------------------------------------------------------------------

template <int W>
__device__ __forceinline__ void team_sync()
{
    if (W == 1) __syncwarp(); else __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the team; every thread gets bitwise the same result.  `part` holds W doubles and
// may be reused only after another team barrier.
template <int W>
__device__ __forceinline__ double team_sum(double a, double *part, int warp, int lane)
{
    a = warp_sum(a);
    if (W > 1) {
        if (lane == 0) part[warp] = a;
        __syncthreads();
        double s = 0;
#pragma unroll
        for (int w = 0; w < W; w++) s += part[w];
        a = s;
    }
    return a;
}

// ---- energies ----------------------------------------------------------------------------------

// Per-thread partial sums of eStart / eEnd (Space.java:711-718) for the particle staged in
// `stage` (sm sites, index m).  Pairs with the thread's unused slots or with sites of m itself
// are computed and discarded (no divergence; a self pair gives NaN before the select).
template <int SPL, bool HAS_EXP>
__device__ __forceinline__ void move_energy(const SiteRegs<SPL> &R, const StageSite *stage, int sm, int m,
                                            const double2 *s_cd, const double2 *s_ab, int T, int lj_slots,
                                            double &eS, double &eE)
{
    eS = 0.0; eE = 0.0;
    bool on[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) on[k] = info_valid(R.info[k]) && info_mol(R.info[k]) != m;
    for (int a = 0; a < sm; a++) {
        const StageSite st = stage[a];
        const double2 *row_cd = s_cd + st.type * T;
#pragma unroll
        for (int k = 0; k < SPL; k++) {
            const double kqq = st.kq * R.q[k];               // kTimesQ * atom.q
            const double ox = R.x[k] - st.ox, oy = R.y[k] - st.oy, oz = R.z[k] - st.oz;
            const double nx = R.x[k] - st.nx, ny = R.y[k] - st.ny, nz = R.z[k] - st.nz;
            double s, e;
            if (HAS_EXP) {
                const int tj = info_type(R.info[k]);
                const double2 cd = row_cd[tj], ab = s_ab[st.type * T + tj];
                s = pair_accumulate<2>(eS, ox, oy, oz, kqq, cd, ab);
                e = pair_accumulate<2>(eE, nx, ny, nz, kqq, cd, ab);
            } else if (k < lj_slots && st.lj) {
                const double2 cd = row_cd[info_type(R.info[k])], z2 = make_double2(0.0, 0.0);
                s = pair_accumulate<1>(eS, ox, oy, oz, kqq, cd, z2);
                e = pair_accumulate<1>(eE, nx, ny, nz, kqq, cd, z2);
            } else {
                const double2 z2 = make_double2(0.0, 0.0);
                s = pair_accumulate<0>(eS, ox, oy, oz, kqq, z2, z2);
                e = pair_accumulate<0>(eE, nx, ny, nz, kqq, z2, z2);
            }
            eS = on[k] ? s : eS;
            eE = on[k] ? e : eE;
        }
    }
}

// Space.calcEnergy (Space.java:648-658): sum over particle pairs x < y of Molecule.calcEnergy(x -> y);
// the site of the lower-indexed particle is `this` (its kTimesQ multiplies the other's q).
// `scratch` must hold all S site coordinates in site order (written by the caller, team-synced).
template <int SPL, int W, bool HAS_EXP>
__device__ __forceinline__ double team_total_energy(const SiteRegs<SPL> &R, const double *scratch, const DevSystem &sys,
                                                    const double2 *s_cd, const double2 *s_ab, double *part,
                                                    int warp, int lane)
{
    double acc = 0.0;
    const int S = sys.n_sites, T = sys.n_types;
    for (int i = 0; i < S; i++) {
        const int infi = sys.site_info[i];
        const int mi = info_mol(infi), ti = info_type(infi);
        const double kqi = K_VALUE * sys.site_q[i];
        const double xi = scratch[3 * i], yi = scratch[3 * i + 1], zi = scratch[3 * i + 2];
#pragma unroll
        for (int k = 0; k < SPL; k++) {
            const int inf = R.info[k];
            if (info_valid(inf) && mi < info_mol(inf)) {
                const int tj = info_type(inf);
                const double2 cd = s_cd[ti * T + tj];
                double2 ab = make_double2(0.0, 0.0);
                if (HAS_EXP) ab = s_ab[ti * T + tj];
                acc = pair_accumulate<HAS_EXP ? 2 : 1>(acc, R.x[k] - xi, R.y[k] - yi, R.z[k] - zi, kqi * R.q[k], cd, ab);
            }
        }
    }
    return team_sum<W>(acc, part, warp, lane);
}

template <int SPL, int W>
__device__ __forceinline__ void sites_to_scratch(const SiteRegs<SPL> &R, double *scratch, const int *pos_site, int tid)
{
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        if (info_valid(R.info[k])) {
            const int i = pos_site[tid + k * 32 * W];
            scratch[3 * i] = R.x[k]; scratch[3 * i + 1] = R.y[k]; scratch[3 * i + 2] = R.z[k];
        }
    }
    team_sync<W>();
}

template <int SPL, int W>
__device__ __forceinline__ void sites_to_global(const SiteRegs<SPL> &R, double *dst, const int *pos_site, int tid)
{
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        if (info_valid(R.info[k])) {
            const int i = pos_site[tid + k * 32 * W];
            dst[3 * i] = R.x[k]; dst[3 * i + 1] = R.y[k]; dst[3 * i + 2] = R.z[k];
        }
    }
}

template <int SPL, int W>
__device__ __forceinline__ void sites_from_global(SiteRegs<SPL> &R, const double *src, const DevSystem &sys, int tid)
{
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        const int i = sys.pos_site[tid + k * 32 * W];
        if (i >= 0) {
            R.x[k] = src[3 * i]; R.y[k] = src[3 * i + 1]; R.z[k] = src[3 * i + 2];
            R.q[k] = sys.site_q[i];
            R.info[k] = sys.site_info[i];
        } else {
            R.x[k] = R.y[k] = R.z[k] = 0.0; R.q[k] = 0.0; R.info[k] = INFO_UNUSED;
        }
    }
}

// ---- shared-memory carving -------------------------------------------------------------------

struct SmemTables {
    int *mol_off;
    int *moveable;
    double2 *cd;
    double2 *ab;
};

struct SmemChain {
    StageSite *stage;     // [TR_MAX_MOL_SITES]
    ChainCtrl *ctrl;
    uint32_t *ring;       // [3][MT_RING]
    double *com;          // [N*3]
    double *scratch;      // [S*3]
    double *part;         // [W]
    MoveDesc *desc;
    int *flag;
};

__host__ __device__ inline int align16(int v) { return (v + 15) & ~15; }

__host__ __device__ inline int smem_tables_bytes(int n_mol, int n_moveable, int n_types, bool has_exp)
{
    int b = align16((n_mol + 1) * 4) + align16(n_moveable * 4) + n_types * n_types * 16;
    if (has_exp) b += n_types * n_types * 16;
    return align16(b);
}

__host__ __device__ inline int smem_chain_bytes(int n_mol, int n_sites, int W)
{
    return align16(TR_MAX_MOL_SITES * (int)sizeof(StageSite)) + align16((int)sizeof(ChainCtrl)) + 3 * MT_RING * 4 +
           align16(n_mol * 24) + align16(n_sites * 24) + align16(W * 8) + align16((int)sizeof(MoveDesc)) + 16;
}

__device__ __forceinline__ SmemTables carve_tables(unsigned char *base, const DevSystem &sys)
{
    SmemTables t;
    unsigned char *p = base;
    t.mol_off = (int *)p;  p += align16((sys.n_mol + 1) * 4);
    t.moveable = (int *)p; p += align16(sys.n_moveable * 4);
    t.cd = (double2 *)p;   p += sys.n_types * sys.n_types * 16;
    t.ab = (double2 *)p;
    return t;
}

__device__ __forceinline__ SmemChain carve_chain(unsigned char *p, const DevSystem &sys, int W)
{
    SmemChain c;
    c.stage = (StageSite *)p;  p += align16(TR_MAX_MOL_SITES * (int)sizeof(StageSite));
    c.ctrl = (ChainCtrl *)p;   p += align16((int)sizeof(ChainCtrl));
    c.ring = (uint32_t *)p;    p += 3 * MT_RING * 4;
    c.com = (double *)p;       p += align16(sys.n_mol * 24);
    c.scratch = (double *)p;   p += align16(sys.n_sites * 24);
    c.part = (double *)p;      p += align16(W * 8);
    c.desc = (MoveDesc *)p;    p += align16((int)sizeof(MoveDesc));
    c.flag = (int *)p;
    return c;
}

__device__ __forceinline__ void load_tables(const SmemTables &t, const DevSystem &sys, bool has_exp)
{
    for (int i = threadIdx.x; i <= sys.n_mol; i += blockDim.x) t.mol_off[i] = sys.mol_off[i];
    for (int i = threadIdx.x; i < sys.n_moveable; i += blockDim.x) t.moveable[i] = sys.moveable[i];
    const int TT = sys.n_types * sys.n_types;
    for (int i = threadIdx.x; i < TT; i += blockDim.x) {
        t.cd[i] = sys.pair_cd[i];
        if (has_exp) t.ab[i] = sys.pair_ab[i];
    }
}

// ---- schedule bookkeeping -----------------------------------------------------------------------

// write(n) (Space.java:550-623 as far as the device is concerned): calcEnergy, keep the
// structure, track the minimum (Space.java:614-617).
template <int SPL, int W, bool HAS_EXP>
__device__ __noinline__ double snapshot(const KParams &P, long long chain, const SiteRegs<SPL> R, const SmemChain sc,
                                        const SmemTables tb, int n, int tid, int warp, int lane)
{
    sites_to_scratch<SPL, W>(R, sc.scratch, P.sys.pos_site, tid);
    const double e = team_total_energy<SPL, W, HAS_EXP>(R, sc.scratch, P.sys, tb.cd, tb.ab, sc.part, warp, lane);
    ChainCtrl *c = sc.ctrl;
    const int k = c->n_snaps;
    if (k < P.max_snaps) {
        if (tid == 0) {
            P.snap_energy[chain * P.max_snaps + k] = e;
            P.snap_tooth[chain * P.max_snaps + k] = n;
        }
        if (P.snap_sites)
            sites_to_global<SPL, W>(R, P.snap_sites + ((size_t)chain * P.max_snaps + k) * (size_t)P.sys.n_sites * 3,
                                    P.sys.pos_site, tid);
    }
    team_sync<W>();
    if (tid == 0) {
        c->n_snaps = k + 1;
        if (e < c->min_energy) { c->min_energy = e; c->min_tooth = n; }
    }
    team_sync<W>();
    return e;
}

// End of a point (Main.java:225-240) and, when it was the last one, end of the tooth (:242-257).
// Returns true when the chain has finished.  Rare (once per movePerPt moves): kept out of line.
template <int SPL, int W, bool HAS_EXP>
__device__ __noinline__ bool end_of_point(const KParams &P, long long chain, const SiteRegs<SPL> R, const SmemChain sc,
                                          const SmemTables tb, const tr_schedule &sch, int numTeeth, int tid, int warp, int lane)
{
    ChainCtrl *c = sc.ctrl;
    const double t = c->t;
    {
        const int pi = c->n_points;
        bool movie = true;
        if (sch.write_conf_heat_capacities && t < 0.005 && t > -0.005) movie = false;   // the `continue` at :228
        double t_next = t;
        if (movie) {
            if (!sch.static_temp) t_next = __dsub_rn(t, c->delT);
            if (t_next < 0.0) t_next = 0.0;
        }
        double me = __longlong_as_double(0x7ff8000000000000LL);
        if (movie && P.points) {
            sites_to_scratch<SPL, W>(R, sc.scratch, P.sys.pos_site, tid);
            me = team_total_energy<SPL, W, HAS_EXP>(R, sc.scratch, P.sys, tb.cd, tb.ab, sc.part, warp, lane);
        }
        if (tid == 0) {
            if (P.points && pi < P.max_points) {
                tr_point_rec r;
                r.tooth = c->x; r.point = c->y; r.accepted = c->accepted; r.total = c->total;
                r.movie_written = movie ? 1 : 0; r.finale = c->isDoubleCycle;
                r.temp = t; r.total_energy = c->totalE; r.total_squared_energy = c->totalSqE; r.movie_energy = me;
                P.points[chain * P.max_points + pi] = r;
            }
            c->n_points = pi + 1;
            c->t = t_next;
            c->y += 1; c->z = 0;
            c->totalE = 0.0; c->totalSqE = 0.0;
        }
        team_sync<W>();
    }
    if (c->y < c->ptsPerTooth) return false;

    const int x = c->x;
    const bool start_finale = (x == numTeeth - 1) && !c->isDoubleCycle && sch.finale;
    if (tid == 0) {
        c->saveT = __dmul_rn(c->saveT, sch.temp_decrease_per_tooth);
        c->t = c->saveT;
        c->ptsPerTooth += sch.points_added_per_tooth;
        if (start_finale) {
            c->t = 0.0;
            c->isDoubleCycle = 1;
            c->maxD = __ddiv_rn(c->maxD, 100.0);
            c->maxRot = __ddiv_rn(c->maxRot, 100.0);
            c->probRot = 0.0; c->probTrans = 0.0;
            // x-- then x++: the finale repeats the same tooth index
        } else {
            c->x = x + 1;
        }
        c->y = 0; c->accepted = 0; c->total = 0;
        c->delT = c->t / (double)(c->ptsPerTooth - 1);
    }
    team_sync<W>();
    if (!start_finale) {
        snapshot<SPL, W, HAS_EXP>(P, chain, R, sc, tb, x + 1, tid, warp, lane);   // write(x+1)
        if (c->x >= numTeeth) {
            if (tid == 0) c->done = 1;
            team_sync<W>();
            return true;
        }
    }
    return false;
}

// ---- the kernel ------------------------------------------------------------------------------

template <int W> struct LaunchCfg { static constexpr int CPB = (W == 1) ? 4 : 1; };

template <int SPL, int W, bool HAS_EXP, bool TRACE, int MINB>
__global__ void __launch_bounds__(32 * W * LaunchCfg<W>::CPB, MINB) anneal_kernel(const __grid_constant__ KParams P)
{
    constexpr int CPB = LaunchCfg<W>::CPB;
    constexpr int TPC = 32 * W;
    extern __shared__ __align__(16) unsigned char smem[];
    const DevSystem &sys = P.sys;
    const SmemTables tb = carve_tables(smem, sys);
    load_tables(tb, sys, HAS_EXP);

    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int team = (W == 1) ? warp_in_block : 0;
    const int warp = (W == 1) ? 0 : warp_in_block;          // warp index inside the team
    const int tid = (W == 1) ? lane : (int)threadIdx.x;     // thread index inside the team
    const long long chain = (long long)blockIdx.x * CPB + team;
    const SmemChain sc = carve_chain(smem + P.smem_tables + team * P.smem_per_chain, sys, W);
    const bool active = chain < P.n_chains;
    const long long ch = active ? chain : 0;                 // idle teams shadow chain 0 read-only

    // chain state -> shared memory / registers
    ChainCtrl *c = sc.ctrl;
    {
        const int *src = (const int *)(P.ctrl + ch);
        int *dst = (int *)c;
        for (int i = tid; i < (int)(sizeof(ChainCtrl) / 4); i += TPC) dst[i] = src[i];
        const double *cs = P.com + (size_t)ch * sys.n_mol * 3;
        for (int i = tid; i < sys.n_mol * 3; i += TPC) sc.com[i] = cs[i];
    }
    SiteRegs<SPL> R;
    sites_from_global<SPL, W>(R, P.sites + (size_t)ch * sys.n_sites * 3, sys, tid);
    __syncthreads();
    if (!active || c->done) return;     // whole teams leave together (W > 1: the whole block)

    const tr_schedule &sch = P.sched[c->sched];   // global (L1-resident, uniform): fields are re-read on use
    const int S = sys.n_sites, T = sys.n_types;
    const int lj_slots = (sys.n_lj_pos + TPC - 1) / TPC;
    const FastMod three = fastmod_make(3);
    MtRing gM, gS, gR;
    {
        uint32_t *mt = P.mt + (size_t)chain * 3 * 2 * MT_N;
        mt_ring_open(gM, mt + STREAM_M * 2 * MT_N, sc.ring + STREAM_M * MT_RING, &c->rng[STREAM_M]);
        mt_ring_open(gS, mt + STREAM_S * 2 * MT_N, sc.ring + STREAM_S * MT_RING, &c->rng[STREAM_S]);
        mt_ring_open(gR, mt + STREAM_R * 2 * MT_N, sc.ring + STREAM_R * MT_RING, &c->rng[STREAM_R]);
    }

    const int numTeeth = sch.static_temp ? 1 : sch.num_teeth;   // Main.java:74-76
    const double bound_hi = c->space_len * 0.75;                // Space.java:663,706
    const double bound_lo = c->space_len * -0.75;

    // Main.java:64,71: write(0) and startingEnergy = calcEnergy()
    if (!c->started) {
        const double e0 = snapshot<SPL, W, HAS_EXP>(P, chain, R, sc, tb, 0, tid, warp, lane);
        if (tid == 0) {
            c->start_energy = e0;
            c->running = e0;
            c->started = 1;
            c->delT = c->t / (double)(c->ptsPerTooth - 1);      // Main.java:188 for tooth 0
        }
        team_sync<W>();
    }

    long long budget = P.max_moves > 0 ? P.max_moves : 0x7fffffffffffffffLL;

    while (budget > 0) {
        // ---------------------------------------------------------------- z-loop (Main.java:195)
        int z = c->z;
        int nz = sch.moves_per_point - z;
        if ((long long)nz > budget) nz = (int)budget;
        const long long moves_before = c->moves;
        const double t = c->t;
        // loop-carried statistics live in registers of every thread (uniform values)
        double running = c->running, totalE = c->totalE, totalSqE = c->totalSqE;
        int accepted = 0;
        int evaluated = 0;
        long long pair_evals = 0;

        for (int it = 0; it < nz; it++, z++) {
            // -- choose the move: Main.java:197-213 (warp 0 of the team owns the RNG)
            int m = 0, rot = 0, axis = -1, magwalk = 0;
            double va = 0.0, vb = 0.0, vc = 0.0;       // translation vector or (sin, cos)
            if (warp == 0) {
                mt_ring_reserve(gS, 9, lane);          // pick + 3 doubles + Metropolis double
                mt_ring_reserve(gM, 4, lane);          // two doubles, always (Main.java:199-208)
                mt_ring_reserve(gR, 3, lane);          // angle double + axis
                const int pick = mt_ring_next_int(gS, sys.pick_mod, lane);       // Space.java:736
                m = tb.moveable[pick];
                const int sm0 = tb.mol_off[m + 1] - tb.mol_off[m];
                const double u1 = mt_ring_next_double(gM);                       // Main.java:199
                const double u2 = mt_ring_next_double(gM);                       // :200 or :208
                rot = (u1 >= 0.5 && sm0 > 1) ? 1 : 0;
                if (rot) {
                    magwalk = (u2 >= c->probRot) ? 0 : 1;
                    const double maxR = magwalk ? JAVA_TWO_PI : c->maxRot;
                    const double u = mt_ring_next_double(gR);                    // Molecule.java:67
                    const double rotAmount = __dmul_rn(__dmul_rn(__dsub_rn(u, 0.5), 2.0), maxR);
                    axis = mt_ring_next_int(gR, three, lane);                    // Molecule.java:68
                    sincos(rotAmount, &va, &vb);
                } else {
                    magwalk = (u2 >= c->probTrans) ? 0 : 1;
                    const double maxD = c->maxD;
                    const double maxT = magwalk ? __dmul_rn(maxD, sch.magwalk_factor_trans) : maxD;
                    const double ux = mt_ring_next_double(gS);                   // Space.java:694-696
                    const double uy = mt_ring_next_double(gS);
                    const double uz = mt_ring_next_double(gS);
                    va = __dsub_rn(__dmul_rn(__dmul_rn(ux, 2.0), maxT), maxT);
                    vb = __dsub_rn(__dmul_rn(__dmul_rn(uy, 2.0), maxT), maxT);
                    vc = __dsub_rn(__dmul_rn(__dmul_rn(uz, 2.0), maxT), maxT);
                }
                if (W > 1 && lane == 0) {
                    MoveDesc d; d.a = va; d.b = vb; d.c = vc; d.m = m; d.rot = rot; d.axis = axis; d.magwalk = magwalk;
                    *sc.desc = d;
                }
            }
            if (W > 1) {
                __syncthreads();
                const MoveDesc d = *sc.desc;
                va = d.a; vb = d.b; vc = d.c; m = d.m; rot = d.rot; axis = d.axis; magwalk = d.magwalk;
            }
            const int sm = tb.mol_off[m + 1] - tb.mol_off[m];

            // -- trial geometry by the owners of m's sites; stage old + trial coordinates
            bool oob = false;
            if (rot) {
                const double cx = sc.com[3 * m], cy = sc.com[3 * m + 1], cz = sc.com[3 * m + 2];
                const double sn = va, cs = vb, nsn = __dmul_rn(-1.0, va);
#pragma unroll
                for (int k = 0; k < SPL; k++) {
                    if (info_mol(R.info[k]) == m) {
                        // Molecule.java:70-95, operation for operation, no contraction
                        const double ax = __dsub_rn(R.x[k], cx), ay = __dsub_rn(R.y[k], cy), az = __dsub_rn(R.z[k], cz);
                        double tx, ty, tz;
                        if (axis == 0) {
                            tx = ax;
                            ty = __dadd_rn(__dmul_rn(cs, ay), __dmul_rn(sn, az));
                            tz = __dadd_rn(__dmul_rn(nsn, ay), __dmul_rn(cs, az));
                        } else if (axis == 1) {
                            tx = __dsub_rn(__dmul_rn(cs, ax), __dmul_rn(sn, az));
                            ty = ay;
                            tz = __dadd_rn(__dmul_rn(sn, ax), __dmul_rn(cs, az));
                        } else {
                            tx = __dadd_rn(__dmul_rn(cs, ax), __dmul_rn(sn, ay));
                            ty = __dadd_rn(__dmul_rn(nsn, ax), __dmul_rn(cs, ay));
                            tz = az;
                        }
                        R.x[k] = __dadd_rn(ax, cx); R.y[k] = __dadd_rn(ay, cy); R.z[k] = __dadd_rn(az, cz);
                        tx = __dadd_rn(tx, cx); ty = __dadd_rn(ty, cy); tz = __dadd_rn(tz, cz);
                        StageSite st;
                        st.ox = R.x[k]; st.oy = R.y[k]; st.oz = R.z[k];
                        st.nx = tx; st.ny = ty; st.nz = tz;
                        st.kq = __dmul_rn(K_VALUE, R.q[k]);
                        st.type = info_type(R.info[k]); st.lj = info_lj(R.info[k]);
                        sc.stage[info_atom(R.info[k])] = st;
                        oob |= (tx > bound_hi) | (tx < bound_lo) | (ty > bound_hi) | (ty < bound_lo) | (tz > bound_hi) | (tz < bound_lo);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < SPL; k++) {
                    if (info_mol(R.info[k]) == m) {
                        const double tx = __dadd_rn(R.x[k], va), ty = __dadd_rn(R.y[k], vb), tz = __dadd_rn(R.z[k], vc);
                        StageSite st;
                        st.ox = R.x[k]; st.oy = R.y[k]; st.oz = R.z[k];
                        st.nx = tx; st.ny = ty; st.nz = tz;
                        st.kq = __dmul_rn(K_VALUE, R.q[k]);
                        st.type = info_type(R.info[k]); st.lj = info_lj(R.info[k]);
                        sc.stage[info_atom(R.info[k])] = st;
                        oob |= (tx > bound_hi) | (tx < bound_lo) | (ty > bound_hi) | (ty < bound_lo) | (tz > bound_hi) | (tz < bound_lo);
                    }
                }
            }
            int any_oob;
            if (W == 1) { any_oob = __any_sync(0xffffffffu, oob); __syncwarp(); }
            else any_oob = __syncthreads_or(oob);

            // -- energy, Metropolis, commit: Space.java:668-689 / :711-733
            int acc = 0, drew = 0;
            double eS = 0.0, eE = 0.0, dE = 0.0, rnd = 0.0;
            if (!any_oob) {
                move_energy<SPL, HAS_EXP>(R, sc.stage, sm, m, tb.cd, tb.ab, T, lj_slots, eS, eE);
                // eEnd - eStart, differenced per thread before the reduction (less cancellation than
                // reducing the two ~100x larger sums; the reference's own value carries the same rounding noise)
                dE = team_sum<W>(eE - eS, sc.part, warp, lane);
                if (TRACE) {
                    team_sync<W>();
                    eS = team_sum<W>(eS, sc.part, warp, lane);
                    team_sync<W>();
                    eE = team_sum<W>(eE, sc.part, warp, lane);
                }
                acc = 1;
                if (dE > 0.0) {
                    if (warp == 0) {
                        rnd = mt_ring_next_double(gS);                           // drawn before the T == 0 test
                        if (t == 0.0) acc = 0;
                        else {
                            const double test = exp(__ddiv_rn(__dmul_rn(-1.0, dE), __dmul_rn(BOLTZMANN_CONSTANT, t)));
                            if (rnd > test) acc = 0;
                        }
                        if (W > 1 && lane == 0) { sc.flag[0] = acc; }
                    }
                    drew = 1;
                    if (W > 1) { __syncthreads(); acc = sc.flag[0]; }
                }
                if (acc) {
#pragma unroll
                    for (int k = 0; k < SPL; k++) {
                        if (info_mol(R.info[k]) == m) {
                            const StageSite *st = &sc.stage[info_atom(R.info[k])];
                            R.x[k] = st->nx; R.y[k] = st->ny; R.z[k] = st->nz;
                        }
                    }
                    if (!rot && tid == 0) {   // m.x = m.tempx (Molecule.java:129-131)
                        sc.com[3 * m] = __dadd_rn(sc.com[3 * m], va);
                        sc.com[3 * m + 1] = __dadd_rn(sc.com[3 * m + 1], vb);
                        sc.com[3 * m + 2] = __dadd_rn(sc.com[3 * m + 2], vc);
                    }
                }
                evaluated++;
                pair_evals += 2LL * sm * (S - sm);
            }
            team_sync<W>();   // stage / com / flag are free for the next move

            // -- Main.java:215-223
            running = __dadd_rn(running, acc ? dE : 0.0);
            if (!sch.static_temp || z > sch.eq_configs) {
                totalE = __dadd_rn(totalE, running);
                totalSqE = __dadd_rn(totalSqE, __dmul_rn(running, running));
            }
            accepted += acc;
            if (TRACE) {
                if (tid == 0 && moves_before + it < P.trace_moves) {
                    tr_trace_rec r;
                    r.mol = m; r.kind = (int8_t)rot; r.magwalk = (int8_t)magwalk; r.bounds = (int8_t)(any_oob ? 1 : 0);
                    r.accepted = (int8_t)acc; r.axis = axis; r.drew_rand = drew;
                    r.e_start = eS; r.e_end = eE; r.dE = dE; r.rand = drew ? rnd : __longlong_as_double(0x7ff8000000000000LL);
                    r.temp = t; r.running = running;
                    P.trace[chain * P.trace_moves + moves_before + it] = r;
                }
            }
        }
        budget -= nz;
        if (tid == 0) {
            c->z = z;
            c->running = running; c->totalE = totalE; c->totalSqE = totalSqE;
            c->total += nz;                       // total++ per move (Main.java:196)
            c->accepted += accepted;
            c->accepted_all += accepted;
            c->moves = moves_before + nz;
            c->evaluated += evaluated;
            c->pair_evals += pair_evals;
        }
        team_sync<W>();
        if (z < sch.moves_per_point) break;       // budget exhausted inside a point
        if (end_of_point<SPL, W, HAS_EXP>(P, chain, R, sc, tb, sch, numTeeth, tid, warp, lane)) break;
    }

    // chain state -> global memory
    if (warp == 0) {
        mt_ring_close(gM, lane);
        mt_ring_close(gS, lane);
        mt_ring_close(gR, lane);
    }
    team_sync<W>();
    {
        int *dst = (int *)(P.ctrl + chain);
        const int *src = (const int *)c;
        for (int i = tid; i < (int)(sizeof(ChainCtrl) / 4); i += TPC) dst[i] = src[i];
        double *cd = P.com + (size_t)chain * sys.n_mol * 3;
        for (int i = tid; i < sys.n_mol * 3; i += TPC) cd[i] = sc.com[i];
    }
    sites_to_global<SPL, W>(R, P.sites + (size_t)chain * S * 3, sys.pos_site, tid);
}

}  // namespace tr
