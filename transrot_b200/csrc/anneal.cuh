// The annealing hot path on the device: Main.sawtoothAnneal (reference Main.java:177-263),
// Space.move / Space.rotate (Space.java:660-734), Molecule.rotateTemp (Molecule.java:64-97),
// the Atom pair potential (Atom.java:91-119) and Space.calcEnergy (Space.java:648-658).
//
// Mapping.  A chain (one independent annealing run = one reference JVM process) is owned by
// a TEAM: L = 16 or 32 lanes of one warp (several chains per block; L = 16 puts two chains in
// a warp, which shares the per-move scalar work and leaves 146 registers per thread at the
// occupancy that keeps 4096 chains resident in one wave), or W > 1 whole warps (block per
// chain, for clusters of hundreds of sites).
// Data.  A chain's site coordinates live in SHARED MEMORY as {x, y, z, q} records in a
// host-chosen position order: position p belongs to team thread (p mod TPC), slot (p div TPC),
// and the sites that have Lennard-Jones parameters come first, so whole slots are Coulomb-only
// and skip the r^-6 / r^-12 arithmetic.  Per move, lanes 0..s_m-1 build the trial geometry of the
// moved particle's sites and stage old + trial coordinates (<= TR_MAX_MOL_SITES entries, read
// back as broadcasts); every thread then evaluates the site pairs between its own positions and
// the staged sites in FP64 (pairs inside the moved particle are neutralised by a zero charge and
// a unit offset on r^2 rather than by branches), and eEnd - eStart is reduced with shuffles
// (plus one shared-memory hop when W > 1).  The three MT19937 streams stay in global memory
// (L2-resident) behind 64-word rings of tempered words in shared memory (mt19937.cuh).  The
// schedule cursor lives in shared memory, so a launch can stop after any move and the next one
// resumes bit-identically.
#pragma once
#include <math.h>
#include "device_types.cuh"
#include "mt19937.cuh"

namespace tr {

template <int L, int W>
struct Team {
    static_assert(W == 1 || L == 32, "multi-warp teams use whole warps");
    static constexpr int TPC = (W > 1) ? 32 * W : L;     // threads per chain
    static constexpr int CPB = (W > 1) ? 1 : 4;          // chains per block
    static constexpr int BLOCK = TPC * CPB;
};

struct __align__(16) StageSite {
    double ox, oy, oz;   // committed coordinates (after the rotateTemp round trip)
    double nx, ny, nz;   // trial coordinates (Atom.tempx/y/z)
    double kq;           // Atom.kTimesQ = K_VALUE * q
    int row;             // byte offset of this site type's row in the pair tables
    int lj;              // type has Lennard-Jones parameters with some type
};

struct MoveDesc {        // W > 1 only: what warp 0 decided, broadcast through shared memory
    double a, b, c;      // translation vector, or (sin, cos, unused)
    int m;
    int rot;
    int axis;
    int magwalk;
};

// Fixed-size part of a chain's shared-memory region: every field sits at a compile-time offset
// from one base register.  com[N*3] and sites[n_pos] follow it.
struct __align__(16) ChainSmem {
    StageSite stage[TR_MAX_MOL_SITES];
    ChainCtrl ctrl;
    tr_schedule sch;                         // this chain's parameter set
    uint32_t *mt;                            // global: [3][2][624] ping-pong blocks of this chain
    uint32_t ring[3][MT_RING];
    double part[8];
    MoveDesc desc;
    int flag[2];
    unsigned long long thr_rot, thr_trans;   // u >= magwalkProb as 52-bit integer compares
    double bound_hi, pad_;                   // 0.75 * spaceLen (Space.java:663,706)
};

struct __align__(16) SiteRec { double x, y, z, q; };   // one position of a chain

// Per-position info word: mol | type << 16 | index-in-particle << 24 | has-LJ << 29.  An empty position has
// mol = INFO_UNUSED, type = n_types (the all-zero table column), q = 0 and far-away coordinates, so it adds
// exactly +0 to every sum without any test.
constexpr int INFO_UNUSED = 0xffff;
constexpr double FAR_AWAY = 1.0e9;
__device__ __forceinline__ int info_mol(int info) { return info & 0xffff; }
__device__ __forceinline__ bool info_valid(int info) { return (info & 0xffff) != INFO_UNUSED; }
__device__ __forceinline__ int info_type(int info) { return (info >> 16) & 0xff; }
__device__ __forceinline__ int info_atom(int info) { return (info >> 24) & 0x1f; }
__device__ __forceinline__ int info_lj(int info) { return (info >> 29) & 1; }
__device__ __forceinline__ int info_col(int info) { return (info >> 12) & 0xff0; }   // type * sizeof(double2)

// ---- the pair potential --------------------------------------------------------------------
// Atom.java:102-103: (kTimesQ*q/r) + A*exp(-1*B*r) - (C/pow(r,6)) + (D/pow(r,12)).  Evaluated
// through 1/r so no divide or sqrt is issued: rsqrt seed (MUFU.RSQ64H, ~22 bits) + one cubic
// Newton step (error ~e^3, below half an ulp), then r^-6 by multiplication.  Every operation
// is FP64; the result differs from the reference's pow/sqrt/divide form by a few ulp per term
// (bar: 1e-12 relative).  MODE 0: Coulomb only (C = D = A = 0 for the whole slot), 12 FP64
// instructions; MODE 1: + Lennard-Jones, 17; MODE 2: + Born-Mayer exp.
// `off` is 0 for a pair that counts and 1 for one that does not (own particle, empty slot):
// with kqq = 0 and zero table entries the latter adds exactly +0 and never sees r = 0.
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
__device__ __forceinline__ double pair_accumulate(double acc, double dx, double dy, double dz, double off, double kqq,
                                                  double2 cd, double2 ab)
{
    const double r2 = fma(dz, dz, fma(dy, dy, fma(dx, dx, off)));
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

template <int L, int W>
__device__ __forceinline__ unsigned team_mask()
{
    if (W > 1 || L == 32) return 0xffffffffu;
    constexpr unsigned lanes = (L >= 32) ? 0xffffffffu : ((1u << (L & 31)) - 1u);
    return lanes << ((threadIdx.x & 31) & ~(L - 1));
}

template <int L, int W>
__device__ __forceinline__ void team_sync(unsigned mask)
{
    if (W == 1) __syncwarp(mask); else __syncthreads();
}

// Sum over the team; every thread gets bitwise the same result.  `part` holds W doubles and
// may be reused only after another team barrier.
template <int L, int W>
__device__ __forceinline__ double team_sum(double a, double *part, int warp, int lane, unsigned mask)
{
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) a += __shfl_xor_sync(mask, a, o);
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

template <int L, int W>
__device__ __forceinline__ bool team_any(bool p, unsigned mask)
{
    if (W > 1) return __syncthreads_or(p) != 0;
    return __ballot_sync(mask, p) != 0u;
}

// ---- shared-memory layout ----------------------------------------------------------------------
// Everything sits at a COMPILE-TIME offset from the start of dynamic shared memory, so no base
// pointers have to live in registers (or be recomputed from kernel parameters):
//   CPB x { ChainSmem, com[NPOS x 3], sites[NPOS] }       chain c at c * CHAIN_BYTES
//   mol_off[NPOS + 1]  moveable[NPOS]  site_pos[NPOS]  pos_info[NPOS]      (int)
//   cd[T x (T+1)] double2, last column zero; then ab[...] when a Born-Mayer term exists (runtime size)

__host__ __device__ constexpr int align16(int v) { return (v + 15) & ~15; }

// PC0 > 0 (pair-energy cache form): particle-major positions - thread n % TPC owns particle n as its first (n < TPC) or second
// particle; the first has PC0 site slots (slots 0 .. PC0-1 of the thread), the second the remaining SPL - PC0.
template <int L, int W, int SPL, int PC0 = 0>
struct Layout {
    static constexpr int TPC = Team<L, W>::TPC;
    static constexpr int CPB = Team<L, W>::CPB;
    static constexpr int NPOS = TPC * SPL;
    static constexpr int NCOM = PC0 > 0 ? TPC * 2 : NPOS;        // particles a chain can have
    static constexpr int OFF_COM = align16((int)sizeof(ChainSmem));
    static constexpr int OFF_SITES = OFF_COM + align16(NCOM * 24);
    static constexpr int CHAIN_BYTES = OFF_SITES + NPOS * (int)sizeof(SiteRec);
    static constexpr int OFF_MOL_OFF = CPB * CHAIN_BYTES;
    static constexpr int OFF_MOVEABLE = OFF_MOL_OFF + align16((NPOS + 1) * 4);
    static constexpr int OFF_SITE_POS = OFF_MOVEABLE + align16(NPOS * 4);
    static constexpr int OFF_POS_INFO = OFF_SITE_POS + align16(NPOS * 4);
    static constexpr int OFF_CD = OFF_POS_INFO + align16(NPOS * 4);

    __host__ __device__ static int total_bytes(int n_types, bool has_exp) { return OFF_CD + n_types * (n_types + 1) * 16 * (has_exp ? 2 : 1); }

    __device__ static __forceinline__ ChainSmem *chain(unsigned char *smem, int team) { return (ChainSmem *)(smem + team * CHAIN_BYTES); }
    __device__ static __forceinline__ double *com(ChainSmem *cs) { return (double *)((unsigned char *)cs + OFF_COM); }
    __device__ static __forceinline__ SiteRec *sites(ChainSmem *cs) { return (SiteRec *)((unsigned char *)cs + OFF_SITES); }
    __device__ static __forceinline__ int mol_off(const unsigned char *smem, int m) { return *(const int *)(smem + OFF_MOL_OFF + 4 * m); }
    __device__ static __forceinline__ int moveable(const unsigned char *smem, int i) { return *(const int *)(smem + OFF_MOVEABLE + 4 * i); }
    __device__ static __forceinline__ int site_pos(const unsigned char *smem, int i) { return *(const int *)(smem + OFF_SITE_POS + 4 * i); }
    __device__ static __forceinline__ int pos_info(const unsigned char *smem, int p) { return *(const int *)(smem + OFF_POS_INFO + 4 * p); }
    __device__ static __forceinline__ const unsigned char *cd(const unsigned char *smem) { return smem + OFF_CD; }
    __device__ static __forceinline__ const unsigned char *ab(const unsigned char *smem, const DevSystem &sys)
    {
        return smem + OFF_CD + (sys.has_exp ? sys.n_types * sys.row_stride : 0);
    }

    __device__ static __forceinline__ void load_tables(unsigned char *smem, const DevSystem &sys)
    {
        const int T = sys.n_types, tab = T * (T + 1) * 16;
        double2 *cdp = (double2 *)(smem + OFF_CD), *abp = (double2 *)(smem + OFF_CD + tab);
        for (int i = threadIdx.x; i < T * (T + 1); i += blockDim.x) {
            const int r = i / (T + 1), c = i % (T + 1);
            cdp[i] = (c < T) ? sys.pair_cd[r * T + c] : make_double2(0.0, 0.0);
            if (sys.has_exp) abp[i] = (c < T) ? sys.pair_ab[r * T + c] : make_double2(0.0, 0.0);
        }
        int *mo = (int *)(smem + OFF_MOL_OFF), *mv = (int *)(smem + OFF_MOVEABLE);
        int *sp = (int *)(smem + OFF_SITE_POS), *pi = (int *)(smem + OFF_POS_INFO);
        for (int i = threadIdx.x; i <= sys.n_mol; i += blockDim.x) mo[i] = sys.mol_off[i];
        for (int i = threadIdx.x; i < sys.n_moveable; i += blockDim.x) mv[i] = sys.moveable[i];
        for (int i = threadIdx.x; i < sys.n_sites; i += blockDim.x) sp[i] = sys.site_pos[i];
        for (int i = threadIdx.x; i < NPOS; i += blockDim.x) pi[i] = sys.pos_info[i];
    }
};

// ---- energies ----------------------------------------------------------------------------------

// NS consecutive slots of this thread against one staged site: 2*NS pair evaluations (old and trial
// position) written operation by operation across the group, so that each dependent FP64 chain has
// 2*NS - 1 independent instructions between its links (FP64 issues every 2 cycles per warp and a
// dependent result is ready after ~8: a group of four saturates the pipe from a single warp).
// The first NLJ slots of the group take the MODE form (1: + Lennard-Jones, 2: + Born-Mayer), the rest are
// Coulomb-only, so a Lennard-Jones slot and a bare-charge slot still share one interleaved group.
template <int NS, int NLJ, int MODE>
__device__ __forceinline__ void group_energy(const StageSite &st, const unsigned char *tab_cd, const unsigned char *tab_ab,
                                             const double *x, const double *y, const double *z, const double *qm, const int *col,
                                             double &eS, double &eE)
{
    constexpr int G = 2 * NS;
    double kqq[NS], r2[G], y0[G], t[G], e[G], p[G], ri[G];
    double2 cd[NS], ab[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        kqq[s] = st.kq * qm[s];                              // kTimesQ * atom.q
        if (MODE >= 1 && s < NLJ) cd[s] = *(const double2 *)(tab_cd + st.row + col[s]);
        if (MODE >= 2 && s < NLJ) ab[s] = *(const double2 *)(tab_ab + st.row + col[s]);
    }
    {
        double dx[G], dy[G], dz[G];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            dx[2 * s] = x[s] - st.ox; dx[2 * s + 1] = x[s] - st.nx;
            dy[2 * s] = y[s] - st.oy; dy[2 * s + 1] = y[s] - st.ny;
            dz[2 * s] = z[s] - st.oz; dz[2 * s + 1] = z[s] - st.nz;
        }
#pragma unroll
        for (int g = 0; g < G; g++) r2[g] = dx[g] * dx[g];
#pragma unroll
        for (int g = 0; g < G; g++) r2[g] = fma(dy[g], dy[g], r2[g]);
#pragma unroll
        for (int g = 0; g < G; g++) r2[g] = fma(dz[g], dz[g], r2[g]);
    }
    // 1/r: MUFU.RSQ64H seed (~22 bits) + one cubic Newton step, y (1 + e/2 + 3e^2/8), e = 1 - r2 y^2
#pragma unroll
    for (int g = 0; g < G; g++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[g]) : "d"(r2[g]));
#pragma unroll
    for (int g = 0; g < G; g++) t[g] = r2[g] * y0[g];
#pragma unroll
    for (int g = 0; g < G; g++) e[g] = fma(-t[g], y0[g], 1.0);
#pragma unroll
    for (int g = 0; g < G; g++) p[g] = fma(e[g], 0.375, 0.5);
#pragma unroll
    for (int g = 0; g < G; g++) e[g] = y0[g] * e[g];
#pragma unroll
    for (int g = 0; g < G; g++) ri[g] = fma(p[g], e[g], y0[g]);
    if (MODE >= 1 && NLJ > 0) {
        constexpr int GL = 2 * NLJ;
        double r6[GL > 0 ? GL : 1], lj[GL > 0 ? GL : 1];
#pragma unroll
        for (int g = 0; g < GL; g++) t[g] = ri[g] * ri[g];
#pragma unroll
        for (int g = 0; g < GL; g++) r6[g] = t[g] * t[g];
#pragma unroll
        for (int g = 0; g < GL; g++) r6[g] = r6[g] * t[g];
#pragma unroll
        for (int g = 0; g < GL; g++) lj[g] = fma(cd[g / 2].y, r6[g], -cd[g / 2].x);      // D/r^6 - C
#pragma unroll
        for (int s = 0; s < NLJ; s++) { eS = fma(lj[2 * s], r6[2 * s], eS); eE = fma(lj[2 * s + 1], r6[2 * s + 1], eE); }
    }
    if (MODE >= 2) {
#pragma unroll
        for (int s = 0; s < NLJ; s++) {
            eS = fma(ab[s].x, exp(-ab[s].y * (r2[2 * s] * ri[2 * s])), eS);
            eE = fma(ab[s].x, exp(-ab[s].y * (r2[2 * s + 1] * ri[2 * s + 1])), eE);
        }
    }
#pragma unroll
    for (int s = 0; s < NS; s++) { eS = fma(kqq[s], ri[2 * s], eS); eE = fma(kqq[s], ri[2 * s + 1], eE); }
}

// One staged site against all SPL positions of this thread, in groups of two slots (PAIRS) or slot by slot.  The
// first LJN slots take the Lennard-Jones form (MODE_LJ), the rest are Coulomb-only.  Slot by slot keeps only two
// evaluations in flight: fewer live registers, and enough when several warps share the FP64 pipe in the same phase
// (crew kernel: +2.8 % on (H2O)20; the team kernels lose 1-4 % with it).
template <int SPL, int LJN, int MODE_LJ, bool PAIRS = true>
__device__ __forceinline__ void stage_site_energy(const StageSite &st, const unsigned char *tab_cd, const unsigned char *tab_ab,
                                                  const double (&x)[SPL], const double (&y)[SPL], const double (&z)[SPL],
                                                  const double (&qm)[SPL], const int (&col)[SPL], double &eS, double &eE)
{
    if constexpr (!PAIRS) {
#pragma unroll
        for (int k = 0; k < SPL; k++) {
            if (k < LJN) group_energy<1, 1, MODE_LJ>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
            else group_energy<1, 0, 0>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SPL; k += 2) {
            if (k + 1 < SPL) {
                if (k + 1 < LJN) group_energy<2, 2, MODE_LJ>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
                else if (k < LJN) group_energy<2, 1, MODE_LJ>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
                else group_energy<2, 0, 0>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
            } else {
                if (k < LJN) group_energy<1, 1, MODE_LJ>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
                else group_energy<1, 0, 0>(st, tab_cd, tab_ab, x + k, y + k, z + k, qm + k, col + k, eS, eE);
            }
        }
    }
}

// Per-thread partial sums of eStart / eEnd (Space.java:711-718) for the particle staged in
// cs->stage (sm sites, index m).  `mine` = this thread's first position record; slot k is at
// mine[k * TPC]; info[k] its info word.  LJS = slots that may hold Lennard-Jones sites
// (compile time, >= the system's actual count).  Positions of m itself are taken out by a zero
// charge, the all-zero table column and a far-away x coordinate (no r = 0, no branch).
template <int TPC, int SPL, int LJS, bool HAS_EXP>
__device__ __forceinline__ void move_energy(const SiteRec *mine, const int (&info)[SPL], const ChainSmem *cs, int sm, int m,
                                            const unsigned char *tab_cd, const unsigned char *tab_ab, int zero_col,
                                            double &eS, double &eE)
{
    eS = 0.0; eE = 0.0;
    double x[SPL], y[SPL], z[SPL], qm[SPL];
    int col[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        const SiteRec r = mine[k * TPC];
        const bool on = info_mol(info[k]) != m;
        x[k] = on ? r.x : FAR_AWAY; y[k] = r.y; z[k] = r.z;
        qm[k] = on ? r.q : 0.0;
        col[k] = on ? info_col(info[k]) : zero_col;
    }
    for (int a = 0; a < sm; a++) {
        const StageSite st = cs->stage[a];
        if (HAS_EXP) stage_site_energy<SPL, SPL, 2>(st, tab_cd, tab_ab, x, y, z, qm, col, eS, eE);
        else if (LJS > 0 && st.lj) stage_site_energy<SPL, LJS, 1>(st, tab_cd, tab_ab, x, y, z, qm, col, eS, eE);
        else stage_site_energy<SPL, 0, 0>(st, tab_cd, tab_ab, x, y, z, qm, col, eS, eE);
    }
}

// Space.calcEnergy (Space.java:648-658): sum over particle pairs x < y of Molecule.calcEnergy(x -> y);
// the site of the lower-indexed particle is `this` (its kTimesQ multiplies the other's q).
template <int L, int W, int SPL, bool HAS_EXP>
__device__ __forceinline__ double team_total_energy(const unsigned char *smem, ChainSmem *cs, const DevSystem &sys,
                                                    int tid, int warp, int lane, unsigned mask)
{
    using Lay = Layout<L, W, SPL>;
    constexpr int TPC = Lay::TPC;
    const SiteRec *sites = Lay::sites(cs);
    const unsigned char *tab_cd = Lay::cd(smem), *tab_ab = Lay::ab(smem, sys);
    double acc = 0.0;
    double x[SPL], y[SPL], z[SPL], q[SPL];
    int info[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        const SiteRec r = sites[tid + k * TPC];
        x[k] = r.x; y[k] = r.y; z[k] = r.z; q[k] = r.q;
        info[k] = Lay::pos_info(smem, tid + k * TPC);
    }
    for (int p = 0; p < Lay::NPOS; p++) {
        const int infi = Lay::pos_info(smem, p);
        if (!info_valid(infi)) continue;
        const int mi = info_mol(infi), row = info_type(infi) * sys.row_stride;
        const SiteRec ri = sites[p];
        const double kqi = K_VALUE * ri.q;
#pragma unroll
        for (int k = 0; k < SPL; k++) {
            if (info_valid(info[k]) && mi < info_mol(info[k])) {
                const double2 cd = *(const double2 *)(tab_cd + row + info_col(info[k]));
                double2 ab = make_double2(0.0, 0.0);
                if (HAS_EXP) ab = *(const double2 *)(tab_ab + row + info_col(info[k]));
                acc = pair_accumulate<HAS_EXP ? 2 : 1>(acc, x[k] - ri.x, y[k] - ri.y, z[k] - ri.z, 0.0, kqi * q[k], cd, ab);
            }
        }
    }
    return team_sum<L, W>(acc, cs->part, warp, lane, mask);
}

// One staged site (trial position only) against SS site slots of one particle of this thread (particle-major layouts: the
// kernels with a pair-energy cache, crew2.cuh and the PC0 > 0 form of anneal_kernel): SS evaluations written
// operation by operation across the group (independent dependency chains for the 8-cycle FP64 latency), summed into e.
// `lj`: the staged site carries Lennard-Jones parameters in some team of the warp (warp-uniform); then the site slots named by
// `ljb` add the r^-6 / r^-12 terms (every slot, with the Born-Mayer term, when HAS_EXP).
template <int SS, bool HAS_EXP>
__device__ __forceinline__ double trial_site_energy(double nx, double ny, double nz, double kq, int row, bool lj, const unsigned char *tab_cd,
                                          const unsigned char *tab_ab, const double *x, const double *y, const double *z, const double *q,
                                          const int *info, unsigned ljb, double e)
{
    double r2[SS], y0[SS], t[SS], u[SS], p[SS], ri[SS];
#pragma unroll
    for (int j = 0; j < SS; j++) { const double d = x[j] - nx; r2[j] = d * d; }
#pragma unroll
    for (int j = 0; j < SS; j++) { const double d = y[j] - ny; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
    for (int j = 0; j < SS; j++) { const double d = z[j] - nz; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
    for (int j = 0; j < SS; j++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[j]) : "d"(r2[j]));
#pragma unroll
    for (int j = 0; j < SS; j++) t[j] = r2[j] * y0[j];
#pragma unroll
    for (int j = 0; j < SS; j++) u[j] = fma(-t[j], y0[j], 1.0);
#pragma unroll
    for (int j = 0; j < SS; j++) p[j] = fma(u[j], 0.375, 0.5);
#pragma unroll
    for (int j = 0; j < SS; j++) u[j] = y0[j] * u[j];
#pragma unroll
    for (int j = 0; j < SS; j++) ri[j] = fma(p[j], u[j], y0[j]);
#pragma unroll
    for (int j = 0; j < SS; j++) e = fma(kq * q[j], ri[j], e);
    if (HAS_EXP || lj) {
#pragma unroll
        for (int j = 0; j < SS; j++) {
            if (HAS_EXP || ((ljb >> j) & 1u)) {
                const int cj = info_col(info[j]);
                const double2 cd = *(const double2 *)(tab_cd + row + cj);
                const double i2 = ri[j] * ri[j];
                const double i6 = i2 * i2 * i2;
                e = fma(fma(cd.y, i6, -cd.x), i6, e);                  // + D/r^12 - C/r^6
                if (HAS_EXP) {
                    const double2 ab = *(const double2 *)(tab_ab + row + cj);
                    e = fma(ab.x, exp(-ab.y * (r2[j] * ri[j])), e);
                }
            }
        }
    }
    return e;
}

// ---- the pair-energy cache form of the team kernel (PC0 > 0: two particles per thread) ------------------------------------------
// The reference recomputes eStart = sum_n m.calcEnergy(n) for every move (Space.java:711-713) although nothing moved since
// those pair energies were last formed.  With particle-major positions every particle-pair energy E(m, n) is a plain
// per-thread sum (the thread that owns n), so a chain keeps all of them (packed lower triangle, global memory / L2,
// KParams::pair_cache): a move evaluates the TRIAL positions only, reads eStart from the cache, and an accepted move writes its
// per-neighbour energies back.  Entries are overwritten, never accumulated; every Space.calcEnergy rebuilds them.  A rejected
// rotation still applies rotateTemp's (a.x - x) + x round trip to the committed coordinates (Molecule.java:70-72,90-92): the
// entries of that particle then describe coordinates that differ by at most one ulp, once (the round trip is idempotent).
__host__ __device__ __forceinline__ int pair_cache_count(int n_mol) { return (n_mol * (n_mol - 1) / 2 + 1) & ~1; }
__device__ __forceinline__ int pair_cache_index(int m, int n) { return n < m ? ((m * (m - 1)) >> 1) + n : ((n * (n - 1)) >> 1) + m; }

// All N0 + N1 site slots of this thread (its first particle's N0, its second particle's N1) against the staged sites, one staged
// site at a time in a real loop with N0 + N1 independent evaluations in flight; one accumulator per site slot, summed per
// particle at the end.  `ljb`: site slots that hold a Lennard-Jones site in any particle (team-uniform).
template <int N0, int N1, bool HAS_EXP>
__device__ __forceinline__ void trial_energy_two(const StageSite *stage, int sm, unsigned ljb, const unsigned char *tab_cd, const unsigned char *tab_ab,
                                                 const double (&x)[N0 + N1], const double (&y)[N0 + N1], const double (&z)[N0 + N1],
                                                 const double (&q)[N0 + N1], const int (&col)[N0 + N1], double &e0, double &e1)
{
    constexpr int N = N0 + N1;
    double es[N];
#pragma unroll
    for (int j = 0; j < N; j++) es[j] = 0.0;
#pragma unroll 1
    for (int a = 0; a < sm; a++) {
        const StageSite *st = stage + a;
        const double nx = st->nx, ny = st->ny, nz = st->nz, kq = st->kq;
        double r2[N], y0[N], t[N], u[N], p[N], ri[N];
#pragma unroll
        for (int j = 0; j < N; j++) { const double d = x[j] - nx; r2[j] = d * d; }
#pragma unroll
        for (int j = 0; j < N; j++) { const double d = y[j] - ny; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
        for (int j = 0; j < N; j++) { const double d = z[j] - nz; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
        for (int j = 0; j < N; j++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[j]) : "d"(r2[j]));
#pragma unroll
        for (int j = 0; j < N; j++) t[j] = r2[j] * y0[j];
#pragma unroll
        for (int j = 0; j < N; j++) u[j] = fma(-t[j], y0[j], 1.0);
#pragma unroll
        for (int j = 0; j < N; j++) p[j] = fma(u[j], 0.375, 0.5);
#pragma unroll
        for (int j = 0; j < N; j++) u[j] = y0[j] * u[j];
#pragma unroll
        for (int j = 0; j < N; j++) ri[j] = fma(p[j], u[j], y0[j]);
#pragma unroll
        for (int j = 0; j < N; j++) es[j] = fma(kq * q[j], ri[j], es[j]);
        if (HAS_EXP || st->lj) {
            const int row = st->row;
#pragma unroll
            for (int j = 0; j < N; j++) {
                if (HAS_EXP || ((ljb >> j) & 1u)) {
                    const double2 cd = *(const double2 *)(tab_cd + row + col[j]);
                    const double i2 = ri[j] * ri[j];
                    const double i6 = i2 * i2 * i2;
                    es[j] = fma(fma(cd.y, i6, -cd.x), i6, es[j]);                  // + D/r^12 - C/r^6
                    if (HAS_EXP) {
                        const double2 ab = *(const double2 *)(tab_ab + row + col[j]);
                        es[j] = fma(ab.x, exp(-ab.y * (r2[j] * ri[j])), es[j]);
                    }
                }
            }
        }
    }
    e0 = 0.0; e1 = 0.0;
#pragma unroll
    for (int j = 0; j < N0; j++) e0 += es[j];
#pragma unroll
    for (int j = N0; j < N; j++) e1 += es[j];
}

// This thread's parts of eStart (from the cache) and eEnd (evaluated) for the particle staged in cs->stage; `pend` keeps the
// per-neighbour energies for the commit.  The thread that owns m itself evaluates r = 0 garbage for that particle's slots and
// discards it; empty site slots (a shorter particle, no particle) sit far away with a zero charge and the all-zero column.
template <int TPC, int N0, int N1, bool HAS_EXP>
__device__ __forceinline__ void move_energy_cached(const SiteRec *mine, const int (&col)[N0 + N1], const ChainSmem *cs, int sm, int m, int n_mol,
                                                   unsigned ljb, const unsigned char *tab_cd, const unsigned char *tab_ab, const double *cache, int tid,
                                                   double (&pend)[2], double &eS, double &eE)
{
    // the two cached energies: requested first, consumed after the evaluation (measured against prefetch.global.L1 at the time the
    // move is drawn + late loads: the early loads are 3-12 % faster)
    const bool on0 = (tid != m) & (tid < n_mol), on1 = (tid + TPC != m) & (tid + TPC < n_mol);
    const double old0 = cache[on0 ? pair_cache_index(m, tid) : 0];
    const double old1 = cache[on1 ? pair_cache_index(m, tid + TPC) : 0];
    double x[N0 + N1], y[N0 + N1], z[N0 + N1], q[N0 + N1];
#pragma unroll
    for (int j = 0; j < N0 + N1; j++) {
        const SiteRec r = mine[j * TPC];
        x[j] = r.x; y[j] = r.y; z[j] = r.z; q[j] = r.q;
    }
    trial_energy_two<N0, N1, HAS_EXP>(cs->stage, sm, ljb, tab_cd, tab_ab, x, y, z, q, col, pend[0], pend[1]);
    eE = (on0 ? pend[0] : 0.0) + (on1 ? pend[1] : 0.0);
    eS = (on0 ? old0 : 0.0) + (on1 ? old1 : 0.0);
}

// Space.calcEnergy (Space.java:648-658) particle pair by particle pair; every pair energy goes into the cache on the way
// (this = the lower-indexed particle, as Space.calcEnergy has it).
template <int L, int W, int SPL, int PC0, bool HAS_EXP>
__device__ __forceinline__ double team_total_energy_cached(const unsigned char *smem, ChainSmem *cs, const DevSystem &sys, double *cache,
                                                           int tid, int warp, int lane, unsigned mask)
{
    using Lay = Layout<L, W, SPL, PC0>;
    constexpr int TPC = Lay::TPC;
    const SiteRec *sites = Lay::sites(cs);
    const unsigned char *tab_cd = Lay::cd(smem), *tab_ab = Lay::ab(smem, sys);
    const int nmol = sys.n_mol;
    double acc = 0.0;
    for (int xm = 0; xm < nmol; xm++) {
        const int off = Lay::mol_off(smem, xm), sx = Lay::mol_off(smem, xm + 1) - off;
        double e[2] = {0.0, 0.0};
        for (int a = 0; a < sx; a++) {
            const int p = Lay::site_pos(smem, off + a);
            const int infi = Lay::pos_info(smem, p);
            const int row = info_type(infi) * sys.row_stride;
            const SiteRec ri = sites[p];
            const double kqi = K_VALUE * ri.q;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int n = tid + TPC * k;
                if (n > xm && n < nmol) {
#pragma unroll
                    for (int j = (k ? PC0 : 0); j < (k ? SPL : PC0); j++) {
                        const int pj = j * TPC + tid;
                        const int cj = info_col(Lay::pos_info(smem, pj));
                        const SiteRec rj = sites[pj];
                        const double2 cd = *(const double2 *)(tab_cd + row + cj);
                        double2 ab = make_double2(0.0, 0.0);
                        if (HAS_EXP) ab = *(const double2 *)(tab_ab + row + cj);
                        e[k] = pair_accumulate<HAS_EXP ? 2 : 1>(e[k], rj.x - ri.x, rj.y - ri.y, rj.z - ri.z, 0.0, kqi * rj.q, cd, ab);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int n = tid + TPC * k;
            if (n > xm && n < nmol) { cache[pair_cache_index(xm, n)] = e[k]; acc += e[k]; }
        }
    }
    return team_sum<L, W>(acc, cs->part, warp, lane, mask);
}

// global [S][3] (site order) <-> shared SiteRec[NPOS] (position order)
template <int TPC, int NPOS>
__device__ __forceinline__ void sites_to_global(const SiteRec *sites, double *dst, const DevSystem &sys, int tid)
{
    for (int p = tid; p < NPOS; p += TPC) {
        const int i = sys.pos_site[p];
        if (i >= 0) { const SiteRec r = sites[p]; dst[3 * i] = r.x; dst[3 * i + 1] = r.y; dst[3 * i + 2] = r.z; }
    }
}

template <int TPC, int NPOS>
__device__ __forceinline__ void sites_from_global(SiteRec *sites, const double *src, const DevSystem &sys, int tid)
{
    for (int p = tid; p < NPOS; p += TPC) {
        const int i = sys.pos_site[p];
        SiteRec r;
        if (i >= 0) { r.x = src[3 * i]; r.y = src[3 * i + 1]; r.z = src[3 * i + 2]; r.q = sys.site_q[i]; }
        else { r.x = r.y = r.z = FAR_AWAY; r.q = 0.0; }
        sites[p] = r;
    }
}

// ---- schedule bookkeeping -----------------------------------------------------------------------

// write(n) (Space.java:550-623 as far as the device is concerned): calcEnergy, keep the
// structure, track the minimum (Space.java:614-617).
// Space.calcEnergy of the committed structure, in the form the kernel variant keeps its positions in.
template <int L, int W, int SPL, bool HAS_EXP, int PC0>
__device__ __forceinline__ double chain_total_energy(const KParams &P, unsigned char *smem, long long chain, int team, int tid, int warp, int lane,
                                                     unsigned mask)
{
    ChainSmem *cs = Layout<L, W, SPL, PC0>::chain(smem, team);
    if constexpr (PC0 > 0)
        return team_total_energy_cached<L, W, SPL, PC0, HAS_EXP>(smem, cs, P.sys, P.pair_cache + (size_t)chain * pair_cache_count(P.sys.n_mol), tid, warp,
                                                                 lane, mask);
    else return team_total_energy<L, W, SPL, HAS_EXP>(smem, cs, P.sys, tid, warp, lane, mask);
}

template <int L, int W, int SPL, bool HAS_EXP, int PC0 = 0>
__device__ __noinline__ double snapshot(const KParams &P, unsigned char *smem, long long chain, int team, int n,
                                        int tid, int warp, int lane, unsigned mask)
{
    using Lay = Layout<L, W, SPL, PC0>;
    ChainSmem *cs = Lay::chain(smem, team);
    const double e = chain_total_energy<L, W, SPL, HAS_EXP, PC0>(P, smem, chain, team, tid, warp, lane, mask);
    ChainCtrl *c = &cs->ctrl;
    const int k = c->n_snaps;
    if (k < P.max_snaps) {
        if (tid == 0) {
            P.snap_energy[chain * P.max_snaps + k] = e;
            P.snap_tooth[chain * P.max_snaps + k] = n;
        }
        if (P.snap_sites)
            sites_to_global<Lay::TPC, Lay::NPOS>(Lay::sites(cs), P.snap_sites + ((size_t)chain * P.max_snaps + k) * (size_t)P.sys.n_sites * 3,
                                                 P.sys, tid);
    }
    team_sync<L, W>(mask);
    if (tid == 0) {
        c->n_snaps = k + 1;
        if (e < c->min_energy) { c->min_energy = e; c->min_tooth = n; }
    }
    team_sync<L, W>(mask);
    return e;
}

__device__ __forceinline__ void refresh_derived(ChainSmem *cs)
{
    cs->thr_rot = double_threshold52(cs->ctrl.probRot);
    cs->thr_trans = double_threshold52(cs->ctrl.probTrans);
    cs->bound_hi = cs->ctrl.space_len * 0.75;    // Space.java:663,706; the lower bound spaceLen * -0.75 is its exact negative
}

// End of a point (Main.java:225-240) and, when it was the last one, end of the tooth (:242-257).
// Returns true when the chain has finished.  Rare (once per movePerPt moves): kept out of line.
template <int L, int W, int SPL, bool HAS_EXP, int PC0 = 0>
__device__ __noinline__ bool end_of_point(const KParams &P, unsigned char *smem, long long chain, int team,
                                          int tid, int warp, int lane, unsigned mask)
{
    using Lay = Layout<L, W, SPL, PC0>;
    ChainSmem *cs = Lay::chain(smem, team);
    ChainCtrl *c = &cs->ctrl;
    const tr_schedule &sch = cs->sch;
    const int numTeeth = sch.static_temp ? 1 : sch.num_teeth;   // Main.java:74-76
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
        if (movie && P.points) me = chain_total_energy<L, W, SPL, HAS_EXP, PC0>(P, smem, chain, team, tid, warp, lane, mask);
        if (movie && P.movie_sites && pi < P.max_points)          // the frame writeMovie appends (Space.java:740-790)
            sites_to_global<Lay::TPC, Lay::NPOS>(Lay::sites(cs), P.movie_sites + ((size_t)chain * P.max_points + pi) * (size_t)P.sys.n_sites * 3,
                                                 P.sys, tid);
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
        team_sync<L, W>(mask);
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
            refresh_derived(cs);
            // x-- then x++: the finale repeats the same tooth index
        } else {
            c->x = x + 1;
        }
        c->y = 0; c->accepted = 0; c->total = 0;
        c->delT = c->t / (double)(c->ptsPerTooth - 1);
    }
    team_sync<L, W>(mask);
    if (!start_finale) {
        snapshot<L, W, SPL, HAS_EXP, PC0>(P, smem, chain, team, x + 1, tid, warp, lane, mask);   // write(x+1)
        if (c->x >= numTeeth) {
            if (tid == 0) c->done = 1;
            team_sync<L, W>(mask);
            return true;
        }
    }
    return false;
}

// ---- the kernel ------------------------------------------------------------------------------

// Ring refills happen once per ~3 moves and are kept out of the move loop's register budget.
template <int L>
__device__ __noinline__ int refill_stream(ChainSmem *cs, int stream, int wr, int tl, unsigned mask)
{
    return mt_ring_refill_words<L>(cs->mt + stream * 2 * MT_N, cs->ring[stream], &cs->ctrl.rng[stream], wr, tl, mask);
}

template <int L, int W, int SPL, int LJS, bool HAS_EXP, bool TRACE, int MINB, int PC0 = 0>
__global__ void __launch_bounds__(Team<L, W>::BLOCK, MINB) anneal_kernel(const __grid_constant__ KParams P)
{
    using Lay = Layout<L, W, SPL, PC0>;
    static_assert(PC0 >= 0 && PC0 < SPL, "site slots of the first particle, of the second");
    constexpr int TPC = Lay::TPC;
    constexpr int CPB = Lay::CPB;
    extern __shared__ __align__(16) unsigned char smem[];
    const DevSystem &sys = P.sys;
    Lay::load_tables(smem, sys);

    const int lane = threadIdx.x & 31;
    const int team = (W == 1) ? (int)threadIdx.x / TPC : 0;
    const int warp = (W == 1) ? 0 : (int)threadIdx.x >> 5;            // warp index inside the team
    const int tid = (W == 1) ? (int)threadIdx.x % TPC : (int)threadIdx.x;   // thread index inside the team
    const unsigned mask = team_mask<L, W>();
    const long long chain = (long long)blockIdx.x * CPB + team;
    ChainSmem *cs = Lay::chain(smem, team);
    double *com = Lay::com(cs);
    SiteRec *sites = Lay::sites(cs);
    const bool active = chain < P.n_chains;
    const long long ch = active ? chain : 0;                 // idle teams shadow chain 0 read-only

    // chain state -> shared memory
    ChainCtrl *c = &cs->ctrl;
    {
        const int *src = (const int *)(P.ctrl + ch);
        int *dst = (int *)c;
        for (int i = tid; i < (int)(sizeof(ChainCtrl) / 4); i += TPC) dst[i] = src[i];
        const double *cg = P.com + (size_t)ch * sys.n_mol * 3;
        for (int i = tid; i < sys.n_mol * 3; i += TPC) com[i] = cg[i];
        sites_from_global<TPC, Lay::NPOS>(sites, P.sites + (size_t)ch * sys.n_sites * 3, sys, tid);
    }
    team_sync<L, W>(mask);
    if (tid == 0) {
        cs->sch = P.sched[c->sched];
        cs->mt = P.mt + (size_t)ch * 3 * 2 * MT_N;
        refresh_derived(cs);
    }
    __syncthreads();
    if (!active || c->done) return;     // whole teams leave together (W > 1: the whole block)

    const tr_schedule &sch = cs->sch;
    const int S = sys.n_sites;
    const FastMod three = fastmod_make(3);
    const SiteRec *mine = sites + tid;            // this thread's positions: mine[k * TPC]
    const unsigned char *tab_cd = Lay::cd(smem), *tab_ab = Lay::ab(smem, sys);
    int info[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) info[k] = Lay::pos_info(smem, tid + k * TPC);
    // pair-energy cache form: this chain's cache, the table columns of this thread's site slots, the site slots that hold
    // Lennard-Jones sites in any particle (team-uniform), and the per-neighbour energies of the move in flight
    double *cache = nullptr;
    unsigned ljb = 0;
    double pend[2] = {0.0, 0.0};
    int col[SPL];
    if constexpr (PC0 > 0) {
        cache = P.pair_cache + (size_t)ch * pair_cache_count(sys.n_mol);
#pragma unroll
        for (int k = 0; k < SPL; k++) col[k] = info_col(info[k]);
        for (int p = 0; p < Lay::NPOS; p++) {
            const int inf = Lay::pos_info(smem, p);
            if (info_valid(inf) && info_lj(inf)) ljb |= 1u << (p / TPC);
        }
    }
    // ring cursors: (read slot, unread words) per stream, team-uniform
    int rdM = 0, avM = 0, rdS = 0, avS = 0, rdR = 0, avR = 0;

    // Main.java:64,71: write(0) and startingEnergy = calcEnergy()
    if (!c->started) {
        const double e0 = snapshot<L, W, SPL, HAS_EXP, PC0>(P, smem, chain, team, 0, tid, warp, lane, mask);
        if (tid == 0) {
            c->start_energy = e0;
            c->running = e0;
            c->started = 1;
            c->delT = c->t / (double)(c->ptsPerTooth - 1);      // Main.java:188 for tooth 0
        }
        team_sync<L, W>(mask);
    }

    long long budget = P.max_moves > 0 ? P.max_moves : 0x7fffffffffffffffLL;

// one tempered word / the 52-bit integer of nextDouble / nextDouble itself from stream X (M, S or R)
#define RING_NEXT32(X) (av##X--, rd##X = (rd##X + 1) & (MT_RING - 1), cs->ring[STREAM_##X][(rd##X - 1) & (MT_RING - 1)])
#define RING_RESERVE(X, need) while (av##X < (need)) av##X += refill_stream<L>(cs, STREAM_##X, rd##X + av##X, tid, mask)

    while (budget > 0) {
        // ---------------------------------------------------------------- z-loop (Main.java:195)
        int z = c->z;
        int nz = sch.moves_per_point - z;
        if ((long long)nz > budget) nz = (int)budget;
        if (nz > (1 << 20)) nz = 1 << 20;         // keeps the 32-bit per-segment counters below in range
        const long long moves_before = c->moves;
        const double t = c->t;
        // loop-carried statistics live in registers of every thread (team-uniform values)
        double running = c->running, totalE = c->totalE, totalSqE = c->totalSqE;
        int accepted = 0, evaluated = 0;
        unsigned sum_sm = 0, sum_sm2 = 0;         // pair evaluations = 2 * (S * sum_sm - sum_sm2)

        for (int it = 0; it < nz; it++, z++) {
            // -- choose the move: Main.java:197-213 (warp 0 of the team owns the RNG)
            int m = 0, rot = 0, axis = -1, magwalk = 0;
            double va = 0.0, vb = 0.0, vc = 0.0;       // translation vector or (sin, cos)
            if (warp == 0) {
                RING_RESERVE(S, 9);                    // pick + 3 doubles + Metropolis double
                RING_RESERVE(M, 4);                    // two doubles, always (Main.java:199-208)
                RING_RESERVE(R, 3);                    // angle double + axis
                int pick;                              // Space.java:736: r.nextInt(moveableMolecules.size())
                {
                    const int n = (int)sys.pick_mod.n;
                    if ((n & -n) == n) pick = (int)(((long long)n * (long long)(RING_NEXT32(S) >> 1)) >> 31);
                    else for (;;) {
                        const int bits = (int)(RING_NEXT32(S) >> 1);
                        pick = (int)fastmod(sys.pick_mod, (uint32_t)bits);
                        if ((int)((uint32_t)bits - (uint32_t)pick + (uint32_t)(n - 1)) >= 0) break;
                        RING_RESERVE(S, 9);
                    }
                }
                m = Lay::moveable(smem, pick);
                const int sm0 = Lay::mol_off(smem, m + 1) - Lay::mol_off(smem, m);
                // r.nextDouble() >= 0.5 (Main.java:199) is the top bit of the first word; both words are consumed.
                // The second double (:200 or :208) is compared as a 52-bit integer against ceil(prob * 2^52).
                const uint32_t w0 = RING_NEXT32(M);
                (void)RING_NEXT32(M);
                const uint32_t h2 = RING_NEXT32(M) >> 6, l2 = RING_NEXT32(M) >> 6;
                const unsigned long long u2 = ((unsigned long long)h2 << 26) | l2;
                rot = ((w0 >> 31) && sm0 > 1) ? 1 : 0;
                if (rot) {
                    magwalk = (u2 >= cs->thr_rot) ? 0 : 1;
                    const double maxR = magwalk ? JAVA_TWO_PI : c->maxRot;
                    const uint32_t hr = RING_NEXT32(R) >> 6, lr = RING_NEXT32(R) >> 6;               // Molecule.java:67
                    const double u = (double)(long long)(((unsigned long long)hr << 26) | lr) * 0x1.0p-52;
                    const double rotAmount = __dmul_rn(__dmul_rn(__dsub_rn(u, 0.5), 2.0), maxR);
                    for (;;) {                                                                    // Molecule.java:68: nextInt(3)
                        const int bits = (int)(RING_NEXT32(R) >> 1);
                        axis = (int)fastmod(three, (uint32_t)bits);
                        if ((int)((uint32_t)bits - (uint32_t)axis + 2u) >= 0) break;
                        RING_RESERVE(R, 1);
                    }
                    sincos(rotAmount, &va, &vb);
                } else {
                    magwalk = (u2 >= cs->thr_trans) ? 0 : 1;
                    const double maxD = c->maxD;
                    const double maxT = magwalk ? __dmul_rn(maxD, sch.magwalk_factor_trans) : maxD;
                    const uint32_t hx = RING_NEXT32(S) >> 6, lx = RING_NEXT32(S) >> 6;               // Space.java:694-696
                    const uint32_t hy = RING_NEXT32(S) >> 6, ly = RING_NEXT32(S) >> 6;
                    const uint32_t hz = RING_NEXT32(S) >> 6, lz = RING_NEXT32(S) >> 6;
                    const double ux = (double)(long long)(((unsigned long long)hx << 26) | lx) * 0x1.0p-52;
                    const double uy = (double)(long long)(((unsigned long long)hy << 26) | ly) * 0x1.0p-52;
                    const double uz = (double)(long long)(((unsigned long long)hz << 26) | lz) * 0x1.0p-52;
                    va = __dsub_rn(__dmul_rn(__dmul_rn(ux, 2.0), maxT), maxT);
                    vb = __dsub_rn(__dmul_rn(__dmul_rn(uy, 2.0), maxT), maxT);
                    vc = __dsub_rn(__dmul_rn(__dmul_rn(uz, 2.0), maxT), maxT);
                }
                if (W > 1 && lane == 0) {
                    MoveDesc d; d.a = va; d.b = vb; d.c = vc; d.m = m; d.rot = rot; d.axis = axis; d.magwalk = magwalk;
                    cs->desc = d;
                }
            }
            if (W > 1) {
                __syncthreads();
                const MoveDesc d = cs->desc;
                va = d.a; vb = d.b; vc = d.c; m = d.m; rot = d.rot; axis = d.axis; magwalk = d.magwalk;
            }
            const int off = Lay::mol_off(smem, m);
            const int sm = Lay::mol_off(smem, m + 1) - off;

            // -- trial geometry: thread a < sm handles site a of m; stage old + trial coordinates
            bool oob = false;
            int my_pos = 0;
            if (tid < sm) {
                my_pos = Lay::site_pos(smem, off + tid);
                const SiteRec r = sites[my_pos];
                const int inf = Lay::pos_info(smem, my_pos);
                double ox = r.x, oy = r.y, oz = r.z, tx, ty, tz;
                if (rot) {
                    // Molecule.java:70-95, operation for operation, no contraction (va = sin, vb = cos)
                    const double cx = com[3 * m], cy = com[3 * m + 1], cz = com[3 * m + 2];
                    const double nsn = __dmul_rn(-1.0, va);
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
                    ox = __dadd_rn(ax, cx); oy = __dadd_rn(ay, cy); oz = __dadd_rn(az, cz);
                    sites[my_pos].x = ox; sites[my_pos].y = oy; sites[my_pos].z = oz;
                    tx = __dadd_rn(tx, cx); ty = __dadd_rn(ty, cy); tz = __dadd_rn(tz, cz);
                } else {
                    tx = __dadd_rn(ox, va); ty = __dadd_rn(oy, vb); tz = __dadd_rn(oz, vc);
                }
                StageSite st;
                st.ox = ox; st.oy = oy; st.oz = oz;
                st.nx = tx; st.ny = ty; st.nz = tz;
                st.kq = __dmul_rn(K_VALUE, r.q);
                st.row = info_type(inf) * sys.row_stride; st.lj = info_lj(inf);
                cs->stage[tid] = st;
                const double b = cs->bound_hi;      // Space.java:663,706: t > L*0.75 || t < L*-0.75
                oob = (fabs(tx) > b) | (fabs(ty) > b) | (fabs(tz) > b);
            }
            const bool any_oob = team_any<L, W>(oob, mask);
            if (W == 1) __syncwarp(mask);

            // -- energy, Metropolis, commit: Space.java:668-689 / :711-733
            int acc = 0, drew = 0;
            double eS = 0.0, eE = 0.0, dE = 0.0, rnd = 0.0;
            if (!any_oob) {
                if constexpr (PC0 > 0)
                    move_energy_cached<TPC, PC0, SPL - PC0, HAS_EXP>(mine, col, cs, sm, m, sys.n_mol, ljb, tab_cd, tab_ab, cache, tid, pend, eS, eE);
                else move_energy<TPC, SPL, LJS, HAS_EXP>(mine, info, cs, sm, m, tab_cd, tab_ab, sys.zero_col, eS, eE);
                // eEnd - eStart, differenced per thread before the reduction (less cancellation than
                // reducing the two ~100x larger sums; the reference's own value carries the same rounding noise)
                dE = team_sum<L, W>(eE - eS, cs->part, warp, lane, mask);
                if (TRACE) {
                    team_sync<L, W>(mask);
                    eS = team_sum<L, W>(eS, cs->part, warp, lane, mask);
                    team_sync<L, W>(mask);
                    eE = team_sum<L, W>(eE, cs->part, warp, lane, mask);
                }
                acc = 1;
                if (dE > 0.0) {
                    if (warp == 0) {
                        const uint32_t hm = RING_NEXT32(S) >> 6, lm = RING_NEXT32(S) >> 6;           // drawn before the T == 0 test
                        rnd = (double)(long long)(((unsigned long long)hm << 26) | lm) * 0x1.0p-52;
                        if (t == 0.0) acc = 0;
                        else {
                            const double test = exp(__ddiv_rn(__dmul_rn(-1.0, dE), __dmul_rn(BOLTZMANN_CONSTANT, t)));
                            if (rnd > test) acc = 0;
                        }
                        if (W > 1 && lane == 0) cs->flag[0] = acc;
                    }
                    drew = 1;
                    if (W > 1) { __syncthreads(); acc = cs->flag[0]; }
                }
                if (acc) {
                    // every thread read its own positions before the reduction it just took part in, so m's
                    // records can be overwritten now: setTemps (Molecule.java:128-135)
                    if (tid < sm) {
                        const StageSite *st = &cs->stage[tid];
                        sites[my_pos].x = st->nx; sites[my_pos].y = st->ny; sites[my_pos].z = st->nz;
                    }
                    if (!rot && tid == 0) {   // m.x = m.tempx
                        com[3 * m] = __dadd_rn(com[3 * m], va);
                        com[3 * m + 1] = __dadd_rn(com[3 * m + 1], vb);
                        com[3 * m + 2] = __dadd_rn(com[3 * m + 2], vc);
                    }
                    if constexpr (PC0 > 0) {  // the pair energies the accepted move left behind
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const int n = tid + TPC * k;
                            if ((n != m) & (n < sys.n_mol)) cache[pair_cache_index(m, n)] = pend[k];
                        }
                    }
                }
                evaluated++;
                sum_sm += (unsigned)sm;
                sum_sm2 += (unsigned)(sm * sm);
            }
            team_sync<L, W>(mask);   // sites / stage / com / flag are settled for the next move

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
            c->pair_evals += (PC0 > 0 ? 1LL : 2LL) * ((long long)S * sum_sm - sum_sm2);     // (eStart comes from the cache)
        }
        team_sync<L, W>(mask);
        if (z < sch.moves_per_point) continue;    // budget exhausted (loop ends) or a capped segment (loop goes on)
        if (end_of_point<L, W, SPL, HAS_EXP, PC0>(P, smem, chain, team, tid, warp, lane, mask)) break;
    }
#undef RING_NEXT32
#undef RING_RESERVE

    // chain state -> global memory: unread ring words go back to the cursors
    if (warp == 0) {
        __syncwarp(mask);
        if (tid == 0) {
            mt_cursor_rewind(c->rng[STREAM_M], avM);
            mt_cursor_rewind(c->rng[STREAM_S], avS);
            mt_cursor_rewind(c->rng[STREAM_R], avR);
        }
    }
    team_sync<L, W>(mask);
    {
        int *dst = (int *)(P.ctrl + chain);
        const int *src = (const int *)c;
        for (int i = tid; i < (int)(sizeof(ChainCtrl) / 4); i += TPC) dst[i] = src[i];
        double *cg = P.com + (size_t)chain * sys.n_mol * 3;
        for (int i = tid; i < sys.n_mol * 3; i += TPC) cg[i] = com[i];
        sites_to_global<TPC, Lay::NPOS>(sites, P.sites + (size_t)chain * S * 3, sys, tid);
    }
}

}  // namespace tr
