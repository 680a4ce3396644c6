// The annealing hot path, second warp-specialised kernel ("crew2"): molecule-major energy teams, a per-chain cache of
// particle-pair energies, and the control flow split over two control warps.  Same reference code as crew.cuh:
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
//   * MOLECULE-MAJOR TEAMS.  A team is L = 8 (or 16) lanes; lane l owns the particles l, l + L, l + 2L ... (M per lane,
//     SS site slots each), so the per-neighbour energies the cache needs are plain per-lane sums, dE is reduced from
//     per-neighbour differences e_new(n) - E(m, n) (less cancellation than differencing two large sums), and with L = 8
//     one energy warp serves FOUR chains: the per-round overhead (plan, commit, geometry, staging, ring window,
//     reduction, barrier) is paid once per four moves.
//   * TWO CONTROL WARPS.  The control lane of crew.cuh needed 2 200 - 4 000 cycles per round; with the teams twice as
//     fast it would be the critical path.  C1 keeps the S stream and the bookkeeping (settle, accumulators, picks,
//     translation vectors, modes); C2 owns everything transcendental: the rotation draw (R stream, sincos) and the
//     Metropolis threshold -kT log(rand).  C2 needs nothing C1 computes in the same round: it learns which move is in flight
//     from the plan C1 posted a round earlier and from the sign of the previous dE.
// One named barrier per round over all warps; plans and results are double-buffered by round parity.
#pragma once
#include "crew.cuh"

namespace tr {

#ifdef TR_CREW_TIMING
#define C2TM(...) __VA_ARGS__
#else
#define C2TM(...)
#endif

constexpr int C2_RING_NEED_M = 4;      // M words a move consumes, always (Main.java:199-208)
constexpr int C2_RING_NEED_R = 8;      // R words guaranteed before a move is posted: angle double + up to 6 nextInt(3) tries

struct __align__(16) C2Cand {          // one drawn move, the S-stream part (Space.randMolecule + the translation vector)
    double va, vb, vc;
    int m;
    int info;                          // site count | rot << 8
};

struct __align__(16) C2Plan {
    // ---- written by control warp C1
    C2Cand cand[2];                    // PLAN_AUTO: successor if no Metropolis double was drawn / if one was; PLAN_MOVE: cand[0]
    int mode, flags, refill, snap_slot;
    int movie_slot, endS0, endS1, pad0_;     // endS: S-ring position after each candidate's own draws (where its Metropolis double sits)
    double kBt, pad1_;
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

struct __align__(16) C2Chain {
    ChainCtrl ctrl;
    tr_schedule sch;
    uint32_t *mt;                      // global: [3][2][624] ping-pong blocks of this chain
    uint32_t ring[3][CREW_RING];
    C2Plan plan[2];
    CrewResult result[2];
    double etot;
    double bound;
    int rd[3];
    int pad0_;
    int wr[3];
    int pad1_;
};

// Chain-independent tables at compile-time offsets; the chain regions follow at a runtime stride.
template <int L, int M, int SS>
struct C2Tables {
    static constexpr int SPL = M * SS;                 // site slots per lane
    static constexpr int NPOS = L * SPL;
    static constexpr int NMOL = L * M;                 // particle slots
    static constexpr int OFF_MOL_NS = 16;                                      // [0..15]: block-wide scalars
    static constexpr int OFF_MOVEABLE = OFF_MOL_NS + align16(NMOL * 4);
    static constexpr int OFF_POS_INFO = OFF_MOVEABLE + align16(NMOL * 4);
    static constexpr int OFF_QPOS = OFF_POS_INFO + align16(NPOS * 4);
    static constexpr int OFF_CD = OFF_QPOS + NPOS * 8;

    __host__ __device__ static int table_bytes(int n_types, bool has_exp) { return (OFF_CD + n_types * (n_types + 1) * 16 * (has_exp ? 2 : 1) + 127) & ~127; }
    __host__ __device__ static int tri_count(int n_mol) { return (n_mol * (n_mol - 1) / 2 + 1) & ~1; }
    // chain region: fixed part, stage, com, coordinates, cache; padded to 64 mod 128 bytes so that the two teams a
    // 64-bit shared-memory access serves in one pass (8 lanes x 8 bytes each) fall into different banks
    __host__ __device__ static int chain_bytes(int n_mol)
    {
        int b = align16((int)sizeof(C2Chain)) + SS * (int)sizeof(C2Stage) + align16(n_mol * 24) + NPOS * 24 + tri_count(n_mol) * 8;
        b = (b + 127) & ~127;
        return b + 64;
    }

    __device__ static __forceinline__ int &alive(unsigned char *smem, int round) { return ((int *)smem)[round & 1]; }
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
    double *com, *xs, *ys, *zs, *cache;
};

template <int L, int M, int SS>
__device__ __forceinline__ C2View c2_view(unsigned char *chains, int c, int stride, int n_mol)
{
    constexpr int NPOS = L * M * SS;
    unsigned char *b = chains + (size_t)c * stride;
    C2View v;
    v.cs = (C2Chain *)b;
    b += align16((int)sizeof(C2Chain));
    v.stage = (C2Stage *)b;
    b += SS * (int)sizeof(C2Stage);
    v.com = (double *)b;
    b += align16(n_mol * 24);
    v.xs = (double *)b; v.ys = v.xs + NPOS; v.zs = v.ys + NPOS;
    v.cache = v.zs + NPOS;
    return v;
}

__device__ __forceinline__ int tri_index(int lo, int hi) { return ((hi * (hi - 1)) >> 1) + lo; }

// ---- ring upkeep (the three streams' windows), as in crew.cuh but on C2Chain ------------------------------------------

template <int L>
__device__ __forceinline__ CrewWindow<L> c2_ring_begin(C2Chain *cs, int s, int tl)
{
    CrewWindow<L> w;
    w.s = s > 2 ? -1 : s;
    w.par = w.gen = w.fed = w.count = w.load = w.wr = 0;
#pragma unroll
    for (int j = 0; j < 32 / L; j++) w.a[j] = w.b[j] = w.far[j] = 0u;
    if (w.s < 0) return w;
    const MtCursor cur = cs->ctrl.rng[s];
    uint32_t *mt = cs->mt + s * 2 * MT_N;
    int par = cur.par, gen = cur.gen, fed = cur.fed;
    if (fed >= MT_N) { par ^= 1; gen = 0; fed = 0; }
    const uint32_t *B = mt + par * MT_N, *O = mt + (par ^ 1) * MT_N;
    w.load = fed < gen;
    const int limit = w.load ? gen : MT_N;
    w.count = (limit - fed < 32) ? limit - fed : 32;
    w.par = par; w.gen = gen; w.fed = fed; w.wr = cs->wr[s];
#pragma unroll
    for (int j = 0; j < 32 / L; j++) {
        const int i = tl + j * L;
        if (i < w.count) {
            const int k = fed + i;
            if (w.load) w.a[j] = B[k];
            else {
                w.a[j] = O[k];
                w.b[j] = (k == MT_N - 1) ? B[0] : O[k + 1];
                w.far[j] = (k < MT_N - MT_M) ? O[k + MT_M] : B[k + MT_M - MT_N];
            }
        }
    }
    return w;
}

template <int L>
__device__ __forceinline__ void c2_ring_end(C2Chain *cs, const CrewWindow<L> &w, int tl, unsigned mask)
{
    if (w.s < 0) return;
    uint32_t *mt = cs->mt + w.s * 2 * MT_N;
#pragma unroll
    for (int j = 0; j < 32 / L; j++) {
        const int i = tl + j * L;
        if (i < w.count) {
            uint32_t v = w.a[j];
            if (!w.load) {
                const uint32_t y = (w.a[j] & 0x80000000u) | (w.b[j] & 0x7fffffffu);
                v = w.far[j] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
                mt[w.par * MT_N + w.fed + i] = v;
            }
            cs->ring[w.s][(w.wr + i) & (CREW_RING - 1)] = mt_temper(v);
        }
    }
    __syncwarp(mask);
    if (tl == 0) {
        MtCursor c;
        c.par = w.par; c.gen = w.load ? w.gen : w.fed + w.count; c.fed = w.fed + w.count;
        cs->ctrl.rng[w.s] = c;
        __threadfence_block();                          // the control lanes read ahead as soon as they see the new cursor
        cs->wr[w.s] = w.wr + w.count;
    }
    __syncwarp(mask);
}

template <int L>
__device__ __forceinline__ void c2_ring_fill(C2Chain *cs, int tl, unsigned mask)
{
    for (int s = 0; s < 3; s++) {
        while (cs->wr[s] - cs->rd[s] <= CREW_RING_LOW) {
            const CrewWindow<L> w = c2_ring_begin<L>(cs, s, tl);
            c2_ring_end<L>(cs, w, tl, mask);
        }
    }
}

// ---- energy -------------------------------------------------------------------------------------------------------------

// One staged site (trial position) against the SS site slots of ONE particle slot of this lane: SS evaluations written
// operation by operation across the group (independent dependency chains for the 8-cycle FP64 latency), summed into e.
// MODE 0: Coulomb only; 1: + Lennard-Jones on the slots named by `ljb`; 2: + Born-Mayer on every slot.
template <int SS, int MODE>
__device__ __forceinline__ double c2_group(const C2Stage &st, const unsigned char *tab_cd, const unsigned char *tab_ab, const double *x, const double *y,
                                           const double *z, const double *qm, const int *col, unsigned ljb, double e)
{
    double r2[SS], y0[SS], t[SS], u[SS], p[SS], ri[SS];
#pragma unroll
    for (int b = 0; b < SS; b++) { const double d = x[b] - st.nx; r2[b] = d * d; }
#pragma unroll
    for (int b = 0; b < SS; b++) { const double d = y[b] - st.ny; r2[b] = fma(d, d, r2[b]); }
#pragma unroll
    for (int b = 0; b < SS; b++) { const double d = z[b] - st.nz; r2[b] = fma(d, d, r2[b]); }
#pragma unroll
    for (int b = 0; b < SS; b++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[b]) : "d"(r2[b]));
#pragma unroll
    for (int b = 0; b < SS; b++) t[b] = r2[b] * y0[b];
#pragma unroll
    for (int b = 0; b < SS; b++) u[b] = fma(-t[b], y0[b], 1.0);
#pragma unroll
    for (int b = 0; b < SS; b++) p[b] = fma(u[b], 0.375, 0.5);
#pragma unroll
    for (int b = 0; b < SS; b++) u[b] = y0[b] * u[b];
#pragma unroll
    for (int b = 0; b < SS; b++) ri[b] = fma(p[b], u[b], y0[b]);
#pragma unroll
    for (int b = 0; b < SS; b++) e = fma(st.kq * qm[b], ri[b], e);
    if (MODE >= 1) {
#pragma unroll
        for (int b = 0; b < SS; b++) {
            if (MODE >= 2 || ((ljb >> b) & 1u)) {
                const double2 cd = *(const double2 *)(tab_cd + st.row + col[b]);
                const double i2 = ri[b] * ri[b];
                const double i6 = i2 * i2 * i2;
                e = fma(fma(cd.y, i6, -cd.x), i6, e);                  // + D/r^12 - C/r^6
                if (MODE >= 2) {
                    const double2 ab = *(const double2 *)(tab_ab + st.row + col[b]);
                    e = fma(ab.x, exp(-ab.y * (r2[b] * ri[b])), e);
                }
            }
        }
    }
    return e;
}

// What a team keeps in registers between rounds to commit the move it staged last (setTemps, Molecule.java:128-135).
template <int M>
struct C2Pending {
    double tx, ty, tz;         // lane a < s_m: trial coordinates of site a
    double cx, cy, cz;         // lane 0: the particle's centre of mass if the move is committed
    double e[M];               // per-neighbour energies at the trial position (what the cache takes)
    int cidx[M];               // their cache slots, or -1
    int pp;                    // position of this lane's staged site, or -1
    int pm;                    // lane 0: particle whose centre of mass moves, or -1 (rotation)
};

// One posted move for one chain, one team (Space.java:660-734 after the draws): trial geometry with the 0.75 * spaceLen box
// test, per-neighbour energies at the trial position, dE against the cached energies.
template <int L, int M, int SS, bool HAS_EXP, bool TRACE>
__device__ __forceinline__ void c2_move(const unsigned char *smem, const C2View &v, const DevSystem &sys, const C2Cand &d, const C2Plan *pl,
                                        unsigned ljb, int tl, unsigned mask, CrewResult *res, C2Pending<M> &pend, double &dE_out, int &oob_out)
{
    using TL = C2Tables<L, M, SS>;
    C2Chain *cs = v.cs;
    const int m = d.m, sm = d.info & 0xff, rot = (d.info >> 8) & 1;
    bool oob = false;
    pend.pp = -1;
    if (tl < sm) {
        const int p = TL::pos_of(m, tl);
        const int inf = TL::pos_info(smem, p);
        double ox = v.xs[p], oy = v.ys[p], oz = v.zs[p], tx, ty, tz;
        if (rot) {
            // Molecule.java:70-95, operation for operation, no contraction
            const double cx = v.com[3 * m], cy = v.com[3 * m + 1], cz = v.com[3 * m + 2];
            const double va = pl->sn, vb = pl->cs, nsn = __dmul_rn(-1.0, va);
            const int axis = pl->axis;
            const double ax = __dsub_rn(ox, cx), ay = __dsub_rn(oy, cy), az = __dsub_rn(oz, cz);
            if (axis == 0) {
                tx = ax;
                ty = __dadd_rn(__dmul_rn(vb, ay), __dmul_rn(va, az));
                tz = __dadd_rn(__dmul_rn(nsn, ay), __dmul_rn(vb, az));
            } else if (axis == 1) {
                tx = __dsub_rn(__dmul_rn(vb, ax), __dmul_rn(va, az));
                ty = ay;
                tz = __dadd_rn(__dmul_rn(va, ax), __dmul_rn(vb, az));
            } else {
                tx = __dadd_rn(__dmul_rn(vb, ax), __dmul_rn(va, ay));
                ty = __dadd_rn(__dmul_rn(nsn, ax), __dmul_rn(vb, ay));
                tz = az;
            }
            // a.x += x: the committed coordinate keeps the (a.x - x) + x round trip even if the move is rejected
            v.xs[p] = __dadd_rn(ax, cx); v.ys[p] = __dadd_rn(ay, cy); v.zs[p] = __dadd_rn(az, cz);
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
        const double b = cs->bound;                  // t > L*0.75 || t < L*-0.75; the lower bound is the exact negative
        oob = (fabs(tx) > b) | (fabs(ty) > b) | (fabs(tz) > b);
        pend.tx = tx; pend.ty = ty; pend.tz = tz; pend.pp = p;
    }
    if (tl == 0) {
        pend.pm = rot ? -1 : m;
        if (!rot) {
            pend.cx = __dadd_rn(v.com[3 * m], d.va);
            pend.cy = __dadd_rn(v.com[3 * m + 1], d.vb);
            pend.cz = __dadd_rn(v.com[3 * m + 2], d.vc);
        }
    }
    const bool any_oob = (__ballot_sync(mask, oob) & mask) != 0u;      // (also orders the staging stores before the reads below)
    __syncwarp(mask);
    if (any_oob) {
        if (tl == 0) res->oob = 1;
        dE_out = 0.0; oob_out = 1;
        return;
    }
    // This lane's particles, one particle slot at a time against every staged site.  m itself and empty slots are taken
    // out by a far-away x, a zero charge and the all-zero table column (no r = 0, no branch).
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    const int nmol = sys.n_mol;
    double dl = 0.0, sl = 0.0, el = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
        const int n = tl + L * k;
        const bool on = (n != m) & (n < nmol);
        const int lo = n < m ? n : m, hi = n < m ? m : n;
        const int ci = on ? tri_index(lo, hi) : -1;
        const double eold = on ? v.cache[on ? ci : 0] : 0.0;
        double x[SS], y[SS], z[SS], qm[SS];
        int cl[SS];
#pragma unroll
        for (int b = 0; b < SS; b++) {
            const int p = (k * SS + b) * L + tl;
            x[b] = on ? v.xs[p] : FAR_AWAY; y[b] = v.ys[p]; z[b] = v.zs[p];
            qm[b] = on ? TL::qpos(smem, p) : 0.0;
            cl[b] = on ? info_col(TL::pos_info(smem, p)) : sys.zero_col;
        }
        double e = 0.0;
        for (int a = 0; a < sm; a++) {
            const C2Stage st = v.stage[a];
            if (HAS_EXP) e = c2_group<SS, 2>(st, tab_cd, tab_ab, x, y, z, qm, cl, ljb, e);
            else if (st.lj) e = c2_group<SS, 1>(st, tab_cd, tab_ab, x, y, z, qm, cl, ljb, e);
            else e = c2_group<SS, 0>(st, tab_cd, tab_ab, x, y, z, qm, cl, ljb, e);
        }
        pend.cidx[k] = ci;
        pend.e[k] = e;
        dl += e - eold;                                   // eEnd - eStart from per-neighbour differences
        if (TRACE) { sl += eold; el += e; }
    }
    const double dE = crew_team_sum<L>(dl, mask);
    if (TRACE) { sl = crew_team_sum<L>(sl, mask); el = crew_team_sum<L>(el, mask); }
    if (tl == 0) {
        res->oob = 0;
        res->dE = dE;
        if (TRACE) { res->eS = sl; res->eE = el; }
    }
    dE_out = dE; oob_out = 0;
}

// Space.calcEnergy (Space.java:648-658) of the committed structure, one team; every particle-pair energy is stored into
// the cache on the way (this = the lower-indexed particle, as Space.calcEnergy has it).
template <int L, int M, int SS, bool HAS_EXP>
__device__ __noinline__ double c2_total_energy(const unsigned char *smem, const C2View &v, const DevSystem &sys, int tl, unsigned mask)
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
            const double xi = v.xs[p], yi = v.ys[p], zi = v.zs[p];
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
                        e[k] = pair_accumulate<HAS_EXP ? 2 : 1>(e[k], v.xs[pj] - xi, v.ys[pj] - yi, v.zs[pj] - zi, 0.0, kqi * TL::qpos(smem, pj), cd, ab);
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
#pragma unroll
    for (int j = 0; j < M * SS; j++) {
        const int p = j * L + tl, i = sys.pos_site[p];
        if (i >= 0) { dst[3 * i] = v.xs[p]; dst[3 * i + 1] = v.ys[p]; dst[3 * i + 2] = v.zs[p]; }
    }
}

template <int L, int M, int SS>
__device__ __forceinline__ void c2_sites_from_global(const C2View &v, const double *src, const DevSystem &sys, int tl)
{
#pragma unroll
    for (int j = 0; j < M * SS; j++) {
        const int p = j * L + tl, i = sys.pos_site[p];
        v.xs[p] = i >= 0 ? src[3 * i] : FAR_AWAY;
        v.ys[p] = i >= 0 ? src[3 * i + 1] : FAR_AWAY;
        v.zs[p] = i >= 0 ? src[3 * i + 2] : FAR_AWAY;
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------

struct C2Pick {                      // the S-stream part of a move starting at a given ring position
    double va, vb, vc;
    int m, sm, rot, end;             // end: ring position after its draws
    bool ok;
};

template <int L, int M, int SS, bool HAS_EXP, bool TRACE, int EW>
__global__ void __launch_bounds__(32 * (2 + EW), 1)
crew2_kernel(const __grid_constant__ KParams P, const int C, const int stride)
{
    using TL = C2Tables<L, M, SS>;
    constexpr int TPW = 32 / L;                    // chains an energy warp evaluates side by side
    constexpr int NT = 32 * (2 + EW);
    static_assert(EW * TPW <= 32, "one control lane per chain");
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
    for (int c0 = warp * TPW; c0 < nloc; c0 += (2 + EW) * TPW) {
        const int c = c0 + team;
        if (c < nloc) {
            const C2View v = c2_view<L, M, SS>(chains, c, stride, sys.n_mol);
            const long long ch = chain0 + c;
            const int *src = (const int *)(P.ctrl + ch);
            int *dst = (int *)&v.cs->ctrl;
            for (int i = tl; i < (int)(sizeof(ChainCtrl) / 4); i += L) dst[i] = src[i];
            const double *cg = P.com + (size_t)ch * sys.n_mol * 3;
            for (int i = tl; i < sys.n_mol * 3; i += L) v.com[i] = cg[i];
            const double *pc = P.pair_cache + (size_t)ch * ntri;
            for (int i = tl; i < ntri; i += L) v.cache[i] = pc[i];
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
            __syncwarp(mask);
            if (!v.cs->ctrl.done) c2_ring_fill<L>(v.cs, tl, mask);
        }
    }
    __syncthreads();

    if (warp == 0) {
        // ================================================================ control warp C1: lane c <-> chain c
        // S stream, bookkeeping, modes.  Iteration r runs while the teams execute round r: it settles what they did in round
        // r - 1, adopts the move of round r, and writes its part of plan[(r + 1) & 1].
        const bool mine = lane < nloc;
        const C2View v = c2_view<L, M, SS>(chains, mine ? lane : 0, stride, sys.n_mol);
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
        bool total_for_init = false, filled = false;
        int commit_pending = 0;
        int refill_cur = 3;
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
        auto choose_refill = [&]() {
            int aM = wrM - rdM, aS = wrS - rdS, aR = wrR - rdR;
            if (refill_cur == STREAM_M) aM = CREW_RING;
            if (refill_cur == STREAM_S) aS = CREW_RING;
            if (refill_cur == STREAM_R) aR = CREW_RING;
            int s = STREAM_M, amin = aM;
            if (aS < amin) { amin = aS; s = STREAM_S; }
            if (aR < amin) { amin = aR; s = STREAM_R; }
            return amin <= CREW_RING_LOW ? s : 3;
        };
        if (live) load_hot();
        C2TM(long long tm_work = 0, tm_wait = 0, tm_rounds = 0;)

        for (int r = -1;; r++) {
            C2TM(const long long tm0 = clock64();)
            C2Plan *pl = &cs->plan[(r + 1) & 1];
            const C2Plan *plc = &cs->plan[r & 1];
            int mode = PLAN_IDLE;
            bool bubble = false;                                  // chain state changed under C2's eyes: post nothing this round
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
                    bubble = true;
                }
                if (mode_cur == PLAN_MOVE) adopt(cA, plc);
                metro_ok = false;

                // ---------------------------------------------- plan round r + 1
                int flags = 0, refill = 3, snap_slot = -1, movie_slot = -1;
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
                        refill = choose_refill();
                        budget--;
                    }
                    prev_info = cur_info;
                } else if (mode_cur == PLAN_IDLE && !bubble) {
                    // nothing is running: everything that is not "next move of the same point" is decided here
                    bool fill = false;
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
                            mode = PLAN_MOVE; budget--; filled = false;
                        } else if (!filled) { fill = true; filled = true; }     // the rings ran low: one blocking fill, then retry
                        else c->done = CHAIN_ABORT_RNG;
                    }
                    if (fill || (mode == PLAN_IDLE && commit_pending)) mode = PLAN_SYNC;   // (parked with an accepted move not yet committed)
                    if (mode != PLAN_IDLE) {
                        flags = (commit_pending ? PF_COMMIT : 0) | (fill ? PF_FILL : 0);
                        commit_pending = 0;
                        refill = fill ? 3 : choose_refill();
                        if (fill) { cs->rd[STREAM_M] = rdM; cs->rd[STREAM_S] = rdS; cs->rd[STREAM_R] = rdR; }
                    }
                    if (c->done && mode == PLAN_IDLE && !bubble) live = false;
                }
                // (PLAN_TOTAL / PLAN_SYNC running now, or a bubble: wait)
                pl->mode = mode; pl->flags = flags; pl->refill = refill; pl->snap_slot = snap_slot; pl->movie_slot = movie_slot;
                pl->kBt = kBt;
                refill_cur = refill;
            } else if (mine) pl->mode = PLAN_IDLE;
            const unsigned any = __ballot_sync(0xffffffffu, mode != PLAN_IDLE || mode_cur != PLAN_IDLE || (live && bubble));
            mode_prev = mode_cur; mode_cur = mode;
            if (lane == 0) TL::alive(smem, r + 1) = (int)any;
            C2TM(const long long tm1 = clock64(); tm_work += tm1 - tm0; tm_rounds++;)
            crew_bar(NT);
            C2TM(tm_wait += clock64() - tm1;)
            if (!any) break;
        }
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("C1: rounds %lld work %lld wait %lld\n", tm_rounds, tm_work / tm_rounds, tm_wait / tm_rounds);)
        // park the chain: counters into c, unread ring words back to the stream cursors
        if (mine) {
            if (hot) flush_hot();
            mt_cursor_rewind(c->rng[STREAM_M], cs->wr[STREAM_M] - rdM);
            mt_cursor_rewind(c->rng[STREAM_S], cs->wr[STREAM_S] - rdS);
            mt_cursor_rewind(c->rng[STREAM_R], cs->wr[STREAM_R] - rdR);
        }
    } else if (warp == 1) {
        // ================================================================ control warp C2: lane c <-> chain c
        // Everything transcendental.  Iteration r: (a) which move is in flight in round r follows from the plan C1 posted
        // for it and from the sign of the dE of round r - 1; (b) its Metropolis double and the threshold -kT log(rnd) on dE;
        // (c) the rotation draw of the NEXT move (Molecule.java:67-68) with sincos.  Both go into plan[(r + 1) & 1].
        const bool mine = lane < nloc;
        const C2View v = c2_view<L, M, SS>(chains, mine ? lane : 0, stride, sys.n_mol);
        C2Chain *cs = v.cs;
        const ChainCtrl *c = &cs->ctrl;
        const FastMod three = fastmod_make(3);
        int rdM = 0, rdR = 0, nR_next = 0;                       // nR_next: R words the rotation draw posted last consumes
        auto ringM = [&](int pos) { return cs->ring[STREAM_M][pos & (CREW_RING - 1)]; };
        auto ringS = [&](int pos) { return cs->ring[STREAM_S][pos & (CREW_RING - 1)]; };
        auto ringR = [&](int pos) { return cs->ring[STREAM_R][pos & (CREW_RING - 1)]; };
        C2TM(long long tm_work = 0, tm_wait = 0, tm_rounds = 0;)
        for (int r = -1;; r++) {
            C2TM(const long long tm0 = clock64();)
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
            C2TM(const long long tm1 = clock64(); tm_work += tm1 - tm0; tm_rounds++;)
            crew_bar(NT);
            C2TM(tm_wait += clock64() - tm1;)
            if (TL::alive(smem, r + 1) == 0) break;
        }
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("C2: rounds %lld work %lld wait %lld\n", tm_rounds, tm_work / tm_rounds, tm_wait / tm_rounds);)
    } else {
        // ================================================================ energy warps: one team <-> one chain
        const int c = (warp - 2) * TPW + team;
        const bool has = c < nloc;
        const C2View v = c2_view<L, M, SS>(chains, has ? c : 0, stride, sys.n_mol);
        C2Chain *cs = v.cs;
        unsigned ljb = 0;                                   // site slots b that hold a Lennard-Jones site in ANY particle (block-uniform)
        for (int p = 0; p < TL::NPOS; p++)
            if (info_valid(TL::pos_info(smem, p)) && info_lj(TL::pos_info(smem, p))) ljb |= 1u << ((p / L) % SS);
        C2Pending<M> pend;
        pend.tx = pend.ty = pend.tz = pend.cx = pend.cy = pend.cz = 0.0;
        pend.pp = -1; pend.pm = -1;
#pragma unroll
        for (int k = 0; k < M; k++) { pend.e[k] = 0.0; pend.cidx[k] = -1; }
        double last_dE = 0.0;
        int last_oob = 0;
        crew_bar(NT);                                       // plan[0] is written
        C2TM(long long tm_work = 0, tm_wait = 0, tm_rounds = 0, tm_a = 0, tm_b = 0;)
        for (int r = 0;; r++) {
            if (TL::alive(smem, r) == 0) break;
            C2TM(const long long tm0 = clock64();)
            if (has) {
                const C2Plan *pl = &cs->plan[r & 1];
                const int mode = pl->mode;
                if (mode != PLAN_IDLE) {
                    const int flags = pl->flags;
                    const CrewWindow<L> w = c2_ring_begin<L>(cs, pl->refill, tl);    // loads fly while the move is evaluated
                    int commit = flags & PF_COMMIT, which = 0;
                    if (mode == PLAN_AUTO) {
                        int acc, drew;
                        crew_metropolis(last_oob, last_dE, (flags & PF_TZERO) != 0, true, pl->rnd, pl->thr, pl->grd, pl->kBt, acc, drew);
                        commit = acc; which = drew;
                    }
                    if (commit) {
                        // setTemps (Molecule.java:128-135) of the move staged last, and its energies into the pair cache
                        if (pend.pp >= 0) { v.xs[pend.pp] = pend.tx; v.ys[pend.pp] = pend.ty; v.zs[pend.pp] = pend.tz; }
                        if (tl == 0 && pend.pm >= 0) { v.com[3 * pend.pm] = pend.cx; v.com[3 * pend.pm + 1] = pend.cy; v.com[3 * pend.pm + 2] = pend.cz; }
#pragma unroll
                        for (int k = 0; k < M; k++)
                            if (pend.cidx[k] >= 0) v.cache[pend.cidx[k]] = pend.e[k];
                    }
                    __syncwarp(mask);
                    C2TM(const long long tma = clock64(); tm_a += tma - tm0;)
                    if (mode == PLAN_AUTO || mode == PLAN_MOVE) {
                        const C2Cand d = pl->cand[which];
                        c2_move<L, M, SS, HAS_EXP, TRACE>(smem, v, sys, d, pl, ljb, tl, mask, &cs->result[r & 1], pend, last_dE, last_oob);
                    } else if (mode == PLAN_TOTAL) {
                        const double e = c2_total_energy<L, M, SS, HAS_EXP>(smem, v, sys, tl, mask);
                        if (tl == 0) cs->etot = e;
                        const int k = pl->snap_slot, mv = pl->movie_slot;
                        if (k >= 0)
                            c2_sites_to_global<L, M, SS>(v, P.snap_sites + ((size_t)(chain0 + c) * P.max_snaps + k) * (size_t)sys.n_sites * 3, sys, tl);
                        if (mv >= 0)
                            c2_sites_to_global<L, M, SS>(v, P.movie_sites + ((size_t)(chain0 + c) * P.max_points + mv) * (size_t)sys.n_sites * 3, sys, tl);
                    }
                    C2TM(const long long tmb = clock64(); tm_b += tmb - tma;)
                    c2_ring_end<L>(cs, w, tl, mask);
                    if (flags & PF_FILL) c2_ring_fill<L>(cs, tl, mask);
                }
            }
            C2TM(const long long tm1 = clock64(); tm_work += tm1 - tm0; tm_rounds++;)
            crew_bar(NT);
            C2TM(tm_wait += clock64() - tm1;)
        }
        C2TM(if (blockIdx.x == 3 && lane == 0) printf("E%d: rounds %lld work %lld (plan+commit %lld, move %lld) wait %lld\n", warp, tm_rounds, tm_work / tm_rounds, tm_a / tm_rounds, tm_b / tm_rounds, tm_wait / tm_rounds);)
    }
    __syncthreads();

    // chain state -> global memory
    for (int c0 = warp * TPW; c0 < nloc; c0 += (2 + EW) * TPW) {
        const int c = c0 + team;
        if (c >= nloc) continue;
        const C2View v = c2_view<L, M, SS>(chains, c, stride, sys.n_mol);
        const long long ch = chain0 + c;
        int *dst = (int *)(P.ctrl + ch);
        const int *src = (const int *)&v.cs->ctrl;
        for (int i = tl; i < (int)(sizeof(ChainCtrl) / 4); i += L) dst[i] = src[i];
        double *cg = P.com + (size_t)ch * sys.n_mol * 3;
        for (int i = tl; i < sys.n_mol * 3; i += L) cg[i] = v.com[i];
        double *pc = P.pair_cache + (size_t)ch * ntri;
        for (int i = tl; i < ntri; i += L) pc[i] = v.cache[i];
        c2_sites_to_global<L, M, SS>(v, P.sites + (size_t)ch * sys.n_sites * 3, sys, tl);
    }
}

}  // namespace tr
