// The annealing hot path, warp-specialised ("crew" kernel).  Same reference code as anneal.cuh:
// Main.sawtoothAnneal (Main.java:177-263), Space.move / Space.rotate (Space.java:660-734),
// Molecule.rotateTemp (Molecule.java:64-97), Atom.calcEnergy (Atom.java:91-119), Space.calcEnergy
// (Space.java:648-658).
//
// Why.  In the team kernel (anneal.cuh) every lane of a chain's team executes the move's serial
// control flow (ring draws, u -> double conversions, sincos, Metropolis exp, bookkeeping): ~500
// issue slots per move that serve ONE chain, as many as the pair arithmetic itself (ncu: FP64 pipe
// 38 % active, issue 53-68 %).  Here a block owns C chains and splits the work by warp role:
//   * warp 0, the CONTROL warp: lane c runs chain c's control flow, so those instructions serve up to
//     32 chains at once - and it runs ONE ROUND BEHIND the energy teams.  While the teams evaluate the
//     move of round r, lane c settles the move of round r - 1 (Metropolis test, the sawtoothAnneal
//     accumulators, trace record) and draws everything round r + 1 can need: the Metropolis double of
//     the move in flight with its threshold on dE, and BOTH possible successors (the S-stream position
//     of the next pick depends on whether that double gets drawn: Space.java:720-721), sincos included.
//     It posts them as a PLAN.  Everything that is not "next move of the same point" (end of point /
//     tooth, calcEnergy rounds, finale, budget, low rings) is an explicit plan, at the cost of a bubble
//     round for that chain only.
//   * warps 1..E, the ENERGY teams: L = 16 or 32 lanes per chain (two chains per warp for L = 16), lane =
//     position mod L.  A team settles its own previous move with the posted threshold (same code, same
//     bits as the control lane), commits it if accepted (setTemps), builds the trial geometry of the
//     chosen successor (lane a < s_m <-> site a) with the 0.75 * spaceLen box test, evaluates the staged
//     sites against its positions in FP64 (anneal.cuh: group_energy) and shuffle-reduces eEnd - eStart.
//     It also appends the 32-word MT19937 window the plan asks for (global loads issued before the pair
//     arithmetic, twisted and tempered after it) and serves calcEnergy / structure-dump rounds.
// One named barrier per round.  Plans and results are double-buffered by round parity: nobody reads
// what the other side is writing.
//
// Shared memory per chain: CrewChain (ChainCtrl, schedule, three 128-word rings of tempered words, two
// plans, two results) + stage[max s_m] + com[N][3] + x[NPOS] y[NPOS] z[NPOS] (SoA: a team lane reads
// position lane + L k without bank conflicts; charges and table columns are chain-independent and live
// in the team lanes' registers).  Positions are ordered Lennard-Jones sites first, as in anneal.cuh.
#pragma once
#include "anneal.cuh"

namespace tr {

enum { REQ_IDLE = 0, REQ_MOVE = 1, REQ_TOTAL = 2, REQ_SYNC = 3 };
enum { CHAIN_ABORT_RNG = 2 };      // ChainCtrl.done: a nextInt rejection loop outran a freshly filled ring (p < 1e-50 per move)

// What a control lane posts for its chain for the NEXT round.
enum { PLAN_IDLE = 0, PLAN_AUTO = 1, PLAN_MOVE = 2, PLAN_TOTAL = 3, PLAN_SYNC = 4 };
constexpr int PF_COMMIT = 1;             // explicit modes: commit the move staged before (setTemps) first
constexpr int PF_FILL = 2;               // bring every ring above CREW_RING_LOW unread words before returning (rare)
constexpr int PF_TZERO = 4;              // PLAN_AUTO: the temperature is exactly 0 (every uphill move is rejected)

constexpr int CREW_RING = 128;           // tempered words per stream kept in shared memory
constexpr int CREW_RING_LOW = CREW_RING - 32;   // a 32-word window may be appended while at most this many are unread
constexpr int CREW_SETTLE_SLACK = 2;            // S words that must be in the ring beyond a posted move's own draws: its Metropolis double

struct __align__(16) CrewDesc {    // one drawn move (Space.move / Space.rotate arguments after the draws)
    double va, vb, vc;             // translation vector, or (sin, cos, -) of the rotation angle
    int m, off;                    // particle, its first site
    int info;                      // site count | rot << 8 | (axis + 1) << 12
    int pad_;
};
__device__ __forceinline__ int desc_sm(int info) { return info & 0xff; }
__device__ __forceinline__ int desc_rot(int info) { return (info >> 8) & 1; }
__device__ __forceinline__ int desc_axis(int info) { return ((info >> 12) & 7) - 1; }

// PLAN_AUTO: the team first settles the move it evaluated last round (Metropolis test with the pre-drawn double and
// its dE threshold, commit if accepted), then evaluates cand[drew] - the successor for "no Metropolis double was
// drawn" / "one was drawn".  PLAN_MOVE: evaluate cand[0].  Two plans alternate: the control lane writes one while the
// team reads the other.
struct __align__(16) CrewPlan {
    CrewDesc cand[2];
    double rnd, thr, grd, kBt;     // Space.r.nextDouble() of the Metropolis test, -kT log(rnd), its guard band, kB * T
    int mode;                      // PLAN_*
    int flags;                     // PF_*
    int refill;                    // stream whose ring gets the next window, 3 = none
    int snap_slot;                 // PLAN_TOTAL: write(n) slot to dump the structure into, or -1
    int movie_slot;                // PLAN_TOTAL: point whose movie frame is dumped, or -1
    int pad_[3];
};

struct __align__(16) CrewResult {  // what a team reports for the move it evaluated
    double dE, eS, eE;             // eEnd - eStart; the two sums (tracing only)
    int oob;                       // rejected by the 0.75 * spaceLen box, no energy evaluated
    int pad_;
};

struct __align__(16) CrewChain {
    ChainCtrl ctrl;
    tr_schedule sch;
    uint32_t *mt;                  // global: [3][2][624] ping-pong blocks of this chain
    uint32_t ring[3][CREW_RING];
    CrewPlan plan[2];              // control lane -> team
    CrewResult result[2];          // team -> control lane
    double etot;                   // PLAN_TOTAL: Space.calcEnergy
    double bound;                  // 0.75 * spaceLen (Space.java:663,706)
    int rd[3];                     // ring words consumed so far (published with PF_FILL only)
    int pad0_;
    int wr[3];                     // ring words appended so far
    int pad1_;
    // team-private between rounds: what a commit of the move staged last needs
    int prev_p[TR_MAX_MOL_SITES];  // positions of its sites (-1 beyond s_m)
    double prev_v[3];              // its centre of mass if it is committed (m.x = m.tempx)
    int prev_m;                    // its particle, or -1 for a rotation (Molecule.rotateTemp leaves the centre of mass alone)
    int pad2_;
};

// The Metropolis test of Space.java:676-687 / :720-731 on a finished move.  Runs identically in the team that
// evaluated the move (PLAN_AUTO) and in the chain's control lane (bookkeeping): same inputs, same code, same bits.
// rand > exp(-dE / kT)  <=>  dE > -kT log(rand) is decided on the pre-computed threshold when dE is away from the
// rounding noise of either form, else by the reference's own expression.
// The reference's own expression, rand > Math.pow(Math.E, -dE / (kB T)) (Space.java:683-687); out of line: it is only needed
// inside the guard band and when no threshold was prepared, and the hot loops have to stay small (instruction cache).
static __device__ __noinline__ int crew_metropolis_exact(double dE, double kBt, double rnd)
{
    const double test = exp(__ddiv_rn(__dmul_rn(-1.0, dE), kBt));
    return rnd > test ? 0 : 1;
}

__device__ __forceinline__ void crew_metropolis(int oob, double dE, bool tzero, bool metro_ok, double rnd, double thr, double grd,
                                                double kBt, int &acc, int &drew)
{
    acc = oob ? 0 : 1; drew = 0;
    if (!oob && dE > 0.0) {
        drew = 1;                                       // the double is drawn before the T == 0 test
        if (tzero) acc = 0;
        else if (metro_ok && dE > thr + grd) acc = 0;
        else if (metro_ok && dE < thr - grd) acc = 1;
        else acc = crew_metropolis_exact(dE, kBt, rnd);
    }
}

#ifdef TR_CREW_TIMING
#define TMG(i, expr) do { tmv[i] += (expr); } while (0)
#define TMCLK() clock64()
#define TMV_PARAM , long long *tmv
#define TMV_ARG , tmv
#else
#define TMG(i, expr) do { } while (0)
#define TMCLK() 0LL
#define TMV_PARAM
#define TMV_ARG
#endif

constexpr int CREW_BAR = 1;        // named barrier used by both roles

__device__ __forceinline__ void crew_bar(int nthreads) { asm volatile("bar.sync %0, %1;" ::"n"(CREW_BAR), "r"(nthreads) : "memory"); }

// Chain-independent tables at compile-time offsets from the start of dynamic shared memory; the C
// chain regions follow at a runtime stride.
template <int L, int SPL>
struct CrewTables {
    static constexpr int NPOS = L * SPL;
    static constexpr int OFF_MOL_OFF = 16;                                       // [0..15]: block-wide scalars
    static constexpr int OFF_MOVEABLE = OFF_MOL_OFF + align16((NPOS + 1) * 4);
    static constexpr int OFF_SITE_POS = OFF_MOVEABLE + align16(NPOS * 4);
    static constexpr int OFF_POS_INFO = OFF_SITE_POS + align16(NPOS * 4);
    static constexpr int OFF_QPOS = OFF_POS_INFO + align16(NPOS * 4);
    static constexpr int OFF_CD = OFF_QPOS + NPOS * 8;

    __host__ __device__ static int table_bytes(int n_types, bool has_exp) { return (OFF_CD + n_types * (n_types + 1) * 16 * (has_exp ? 2 : 1) + 127) & ~127; }
    // fixed part + stage + com + coordinates, padded to 16 mod 128 so that lanes of the control warp
    // (one chain each) hit different banks when they touch the same field
    __host__ __device__ static int chain_bytes(int n_mol, int max_mol_sites)
    {
        int b = align16((int)sizeof(CrewChain)) + max_mol_sites * (int)sizeof(StageSite) + align16(n_mol * 24) + NPOS * 24;
        b = (b + 127) & ~127;
        return b + 16;
    }

    // "some chain still has work" for round r, written by the control warp one round ahead: two slots by round
    // parity, so the warps that read it at the top of round r never race with the write for round r + 1
    __device__ static __forceinline__ int &alive(unsigned char *smem, int round) { return ((int *)smem)[round & 1]; }
    __device__ static __forceinline__ int mol_off(const unsigned char *smem, int m) { return *(const int *)(smem + OFF_MOL_OFF + 4 * m); }
    __device__ static __forceinline__ int moveable(const unsigned char *smem, int i) { return *(const int *)(smem + OFF_MOVEABLE + 4 * i); }
    __device__ static __forceinline__ int site_pos(const unsigned char *smem, int i) { return *(const int *)(smem + OFF_SITE_POS + 4 * i); }
    __device__ static __forceinline__ int pos_info(const unsigned char *smem, int p) { return *(const int *)(smem + OFF_POS_INFO + 4 * p); }
    __device__ static __forceinline__ double qpos(const unsigned char *smem, int p) { return *(const double *)(smem + OFF_QPOS + 8 * p); }
    __device__ static __forceinline__ const unsigned char *cd(const unsigned char *smem) { return smem + OFF_CD; }
    __device__ static __forceinline__ const unsigned char *ab(const unsigned char *smem, const DevSystem &sys)
    {
        return smem + OFF_CD + (sys.has_exp ? sys.n_types * sys.row_stride : 0);
    }

    __device__ static void load(unsigned char *smem, const DevSystem &sys)
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
        double *qp = (double *)(smem + OFF_QPOS);
        for (int i = threadIdx.x; i <= sys.n_mol; i += blockDim.x) mo[i] = sys.mol_off[i];
        for (int i = threadIdx.x; i < sys.n_moveable; i += blockDim.x) mv[i] = sys.moveable[i];
        for (int i = threadIdx.x; i < sys.n_sites; i += blockDim.x) sp[i] = sys.site_pos[i];
        for (int i = threadIdx.x; i < NPOS; i += blockDim.x) {
            pi[i] = sys.pos_info[i];
            const int s = sys.pos_site[i];
            qp[i] = s >= 0 ? sys.site_q[s] : 0.0;
        }
    }
};

// Pointers into one chain's region.
struct CrewView {
    CrewChain *cs;
    StageSite *stage;
    double *com, *xs, *ys, *zs;
};

template <int L, int SPL>
__device__ __forceinline__ CrewView crew_view(unsigned char *chains, int c, int stride, int n_mol, int max_mol_sites)
{
    constexpr int NPOS = L * SPL;
    unsigned char *b = chains + (size_t)c * stride;
    CrewView v;
    v.cs = (CrewChain *)b;
    b += align16((int)sizeof(CrewChain));
    v.stage = (StageSite *)b;
    b += max_mol_sites * (int)sizeof(StageSite);
    v.com = (double *)b;
    b += align16(n_mol * 24);
    v.xs = (double *)b; v.ys = v.xs + NPOS; v.zs = v.ys + NPOS;
    return v;
}

// Sum over the L lanes of a team (a whole warp or one half of it); every lane gets the same bits.
template <int L>
__device__ __forceinline__ double crew_team_sum(double a, unsigned mask)
{
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) a += __shfl_xor_sync(mask, a, o);
    return a;
}

// ---- energy-warp services -----------------------------------------------------------------------

// Ring upkeep.  The control lane reads ahead of the published `rd` cursors while the energy teams append
// behind `wr`, so a 32-word window is only ever appended when at most CREW_RING_LOW words are unread.
// The control lane decides which stream gets the next window (it sees both cursors) and says so in the order
// word; crew_ring_begin issues the global loads of that window, the pair arithmetic of the round hides their L2
// latency, crew_ring_end twists, tempers and appends.  A team of L = 16 lanes takes two words per lane.

template <int L>
struct CrewWindow {
    int s, par, gen, fed, count, load, wr;
    uint32_t a[32 / L], b[32 / L], far[32 / L];
};

template <int L>
__device__ __forceinline__ CrewWindow<L> crew_ring_begin(CrewChain *cs, int s, int tl)
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
__device__ __forceinline__ void crew_ring_end(CrewChain *cs, const CrewWindow<L> &w, int tl, unsigned mask)
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
        __threadfence_block();                          // the control lane reads ahead as soon as it sees the new cursor
        cs->wr[w.s] = w.wr + w.count;
    }
    __syncwarp(mask);
}

// Blocking fill of every ring to more than CREW_RING_LOW unread words (kernel prologue; ORD_FILL).
template <int L>
__device__ __forceinline__ void crew_ring_fill(CrewChain *cs, int tl, unsigned mask)
{
    for (int s = 0; s < 3; s++) {
        while (cs->wr[s] - cs->rd[s] <= CREW_RING_LOW) {
            const CrewWindow<L> w = crew_ring_begin<L>(cs, s, tl);
            crew_ring_end<L>(cs, w, tl, mask);
        }
    }
}

// One posted move for one chain, one team: trial geometry of the s_m sites (lane a < s_m <-> site a) into the
// staging area with the 0.75 * spaceLen box test (Space.java:663,706), then eEnd - eStart (Space.java:711-718).
// Also leaves what a later commit needs (prev_p, prev_m, prev_v).
template <int L, int SPL, int LJS, bool HAS_EXP, bool TRACE>
__device__ __forceinline__ void crew_move(const unsigned char *smem, const CrewView &v, const DevSystem &sys, const CrewDesc &d,
                                          const int (&info)[SPL], const double (&q)[SPL], int tl, unsigned mask, CrewResult *res,
                                          double &dE_out, int &oob_out TMV_PARAM)
{
    using TL = CrewTables<L, SPL>;
    static_assert(TR_MAX_MOL_SITES <= L, "one lane per site of the moved particle");
    CrewChain *cs = v.cs;
    const int m = d.m, sm = desc_sm(d.info), rot = desc_rot(d.info), axis = desc_axis(d.info);
    [[maybe_unused]] const long long tq0 = TMCLK();
    // this lane's positions; those of m itself are taken out by a far-away x, a zero charge and the all-zero table column
    double x[SPL], y[SPL], z[SPL], qm[SPL];
    int col[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        const bool on = info_mol(info[k]) != m;
        x[k] = on ? v.xs[tl + L * k] : FAR_AWAY; y[k] = v.ys[tl + L * k]; z[k] = v.zs[tl + L * k];
        qm[k] = on ? q[k] : 0.0;
        col[k] = on ? info_col(info[k]) : sys.zero_col;
    }
    bool oob = false;
    if (tl < TR_MAX_MOL_SITES) {
        int p = -1;
        if (tl < sm) {
            p = TL::site_pos(smem, d.off + tl);
            const int inf = TL::pos_info(smem, p);
            double ox = v.xs[p], oy = v.ys[p], oz = v.zs[p], tx, ty, tz;
            if (rot) {
                // Molecule.java:70-95, operation for operation, no contraction (va = sin, vb = cos)
                const double cx = v.com[3 * m], cy = v.com[3 * m + 1], cz = v.com[3 * m + 2];
                const double va = d.va, vb = d.vb, nsn = __dmul_rn(-1.0, d.va);
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
                v.xs[p] = ox; v.ys[p] = oy; v.zs[p] = oz;
                tx = __dadd_rn(tx, cx); ty = __dadd_rn(ty, cy); tz = __dadd_rn(tz, cz);
            } else {
                tx = __dadd_rn(ox, d.va); ty = __dadd_rn(oy, d.vb); tz = __dadd_rn(oz, d.vc);
            }
            StageSite st;
            st.ox = ox; st.oy = oy; st.oz = oz;
            st.nx = tx; st.ny = ty; st.nz = tz;
            st.kq = __dmul_rn(K_VALUE, TL::qpos(smem, p));
            st.row = info_type(inf) * sys.row_stride; st.lj = info_lj(inf);
            v.stage[tl] = st;
            const double b = cs->bound;                  // t > L*0.75 || t < L*-0.75; the lower bound is the exact negative
            oob = (fabs(tx) > b) | (fabs(ty) > b) | (fabs(tz) > b);
        }
        cs->prev_p[tl] = p;
        if (tl == 0) {
            cs->prev_m = rot ? -1 : m;
            if (!rot) {
                cs->prev_v[0] = __dadd_rn(v.com[3 * m], d.va);
                cs->prev_v[1] = __dadd_rn(v.com[3 * m + 1], d.vb);
                cs->prev_v[2] = __dadd_rn(v.com[3 * m + 2], d.vc);
            }
        }
    }
    const bool any_oob = (__ballot_sync(mask, oob) & mask) != 0u;
    __syncwarp(mask);                                       // staging stores before the broadcast reads below
    [[maybe_unused]] const long long tq1 = TMCLK();
    TMG(0, tq1 - tq0);
    if (any_oob) {
        if (tl == 0) res->oob = 1;
        dE_out = 0.0; oob_out = 1;
        return;
    }
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    double eS = 0.0, eE = 0.0;
    for (int a = 0; a < sm; a++) {
        const StageSite st = v.stage[a];
        if (HAS_EXP) stage_site_energy<SPL, SPL, 2, false>(st, tab_cd, tab_ab, x, y, z, qm, col, eS, eE);
        else if (LJS > 0 && st.lj) stage_site_energy<SPL, LJS, 1, false>(st, tab_cd, tab_ab, x, y, z, qm, col, eS, eE);
        else stage_site_energy<SPL, 0, 0, false>(st, tab_cd, tab_ab, x, y, z, qm, col, eS, eE);
    }
    [[maybe_unused]] const long long tq2 = TMCLK();
    TMG(1, tq2 - tq1);
    // differenced per lane before the reduction (less cancellation than reducing the two ~100x larger sums)
    const double dE = crew_team_sum<L>(eE - eS, mask);
    if (TRACE) { eS = crew_team_sum<L>(eS, mask); eE = crew_team_sum<L>(eE, mask); }
    if (tl == 0) {
        res->oob = 0;
        res->dE = dE;
        if (TRACE) { res->eS = eS; res->eE = eE; }
    }
    dE_out = dE; oob_out = 0;
    TMG(2, TMCLK() - tq2);
}

// Space.calcEnergy (Space.java:648-658) of the chain's committed structure, one team.
template <int L, int SPL, bool HAS_EXP>
__device__ __forceinline__ double crew_total_energy(const unsigned char *smem, const CrewView &v, const DevSystem &sys, const int (&info)[SPL],
                                                    const double (&q)[SPL], int tl, unsigned mask)
{
    using TL = CrewTables<L, SPL>;
    const unsigned char *tab_cd = TL::cd(smem), *tab_ab = TL::ab(smem, sys);
    double x[SPL], y[SPL], z[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) { x[k] = v.xs[tl + L * k]; y[k] = v.ys[tl + L * k]; z[k] = v.zs[tl + L * k]; }
    double acc = 0.0;
    for (int p = 0; p < TL::NPOS; p++) {
        const int infi = TL::pos_info(smem, p);
        if (!info_valid(infi)) continue;
        const int mi = info_mol(infi), row = info_type(infi) * sys.row_stride;
        const double xi = v.xs[p], yi = v.ys[p], zi = v.zs[p];
        const double kqi = K_VALUE * TL::qpos(smem, p);
#pragma unroll
        for (int k = 0; k < SPL; k++) {
            if (info_valid(info[k]) && mi < info_mol(info[k])) {
                const double2 cd = *(const double2 *)(tab_cd + row + info_col(info[k]));
                double2 ab = make_double2(0.0, 0.0);
                if (HAS_EXP) ab = *(const double2 *)(tab_ab + row + info_col(info[k]));
                acc = pair_accumulate<HAS_EXP ? 2 : 1>(acc, x[k] - xi, y[k] - yi, z[k] - zi, 0.0, kqi * q[k], cd, ab);
            }
        }
    }
    return crew_team_sum<L>(acc, mask);
}

// shared SoA (position order) <-> global [S][3] (site order), one team
template <int L, int SPL>
__device__ __forceinline__ void crew_sites_to_global(const CrewView &v, double *dst, const DevSystem &sys, int tl)
{
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        const int p = tl + L * k, i = sys.pos_site[p];
        if (i >= 0) { dst[3 * i] = v.xs[p]; dst[3 * i + 1] = v.ys[p]; dst[3 * i + 2] = v.zs[p]; }
    }
}

template <int L, int SPL>
__device__ __forceinline__ void crew_sites_from_global(const CrewView &v, const double *src, const DevSystem &sys, int tl)
{
#pragma unroll
    for (int k = 0; k < SPL; k++) {
        const int p = tl + L * k, i = sys.pos_site[p];
        v.xs[p] = i >= 0 ? src[3 * i] : FAR_AWAY;
        v.ys[p] = i >= 0 ? src[3 * i + 1] : FAR_AWAY;
        v.zs[p] = i >= 0 ? src[3 * i + 2] : FAR_AWAY;
    }
}

// ---- control-lane pieces ---------------------------------------------------------------------------

// End of a point (Main.java:225-240) and, when it was the tooth's last one, end of the tooth (:242-257), for
// the chain of this lane.  `etot` = Space.calcEnergy of the current structure when a REQ_TOTAL round
// preceded the call (movie energy and write(x+1) see the same structure), else unused.
// Returns with c->done set when the chain has finished.
static __device__ __noinline__ void crew_finish_point(const KParams &P, ChainCtrl *c, const tr_schedule &sch, long long chain, double etot, bool have_etot)
{
    const int numTeeth = sch.static_temp ? 1 : sch.num_teeth;   // Main.java:74-76
    const double t = c->t;
    const int pi = c->n_points;
    bool movie = true;
    if (sch.write_conf_heat_capacities && t < 0.005 && t > -0.005) movie = false;   // the `continue` at :228
    double t_next = t;
    if (movie) {
        if (!sch.static_temp) t_next = __dsub_rn(t, c->delT);
        if (t_next < 0.0) t_next = 0.0;
    }
    if (P.points && pi < P.max_points) {
        tr_point_rec r;
        r.tooth = c->x; r.point = c->y; r.accepted = c->accepted; r.total = c->total;
        r.movie_written = movie ? 1 : 0; r.finale = c->isDoubleCycle;
        r.temp = t; r.total_energy = c->totalE; r.total_squared_energy = c->totalSqE;
        r.movie_energy = (movie && have_etot) ? etot : __longlong_as_double(0x7ff8000000000000LL);
        P.points[chain * P.max_points + pi] = r;
    }
    c->n_points = pi + 1;
    c->t = t_next;
    c->y += 1; c->z = 0;
    c->totalE = 0.0; c->totalSqE = 0.0;
    if (c->y < c->ptsPerTooth) return;

    const int x = c->x;
    const bool start_finale = (x == numTeeth - 1) && !c->isDoubleCycle && sch.finale;
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
    if (!start_finale) {
        // write(x+1) (Space.java:550-623): energy + minimum tracking here, the structure was dumped by the energy warp
        const int k = c->n_snaps;
        if (k < P.max_snaps) {
            P.snap_energy[chain * P.max_snaps + k] = etot;
            P.snap_tooth[chain * P.max_snaps + k] = x + 1;
        }
        c->n_snaps = k + 1;
        if (etot < c->min_energy) { c->min_energy = etot; c->min_tooth = x + 1; }
        if (c->x >= numTeeth) c->done = 1;
    }
}

// Does the point that just ended need Space.calcEnergy (movie frame energy or write(x+1))?  Returns the
// dump slot for the energy team in `snap_slot`.
__device__ __forceinline__ bool crew_point_needs_total(const KParams &P, const ChainCtrl *c, const tr_schedule &sch, int &snap_slot, int &movie_slot)
{
    const int numTeeth = sch.static_temp ? 1 : sch.num_teeth;
    const double t = c->t;
    const bool movie = !(sch.write_conf_heat_capacities && t < 0.005 && t > -0.005);
    const bool tooth_end = c->y + 1 >= c->ptsPerTooth;
    const bool start_finale = tooth_end && (c->x == numTeeth - 1) && !c->isDoubleCycle && sch.finale;
    const bool snap = tooth_end && !start_finale;
    snap_slot = (snap && P.snap_sites && c->n_snaps < P.max_snaps) ? c->n_snaps : -1;
    movie_slot = (movie && P.movie_sites && c->n_points < P.max_points) ? c->n_points : -1;
    return snap || (movie && P.points != nullptr) || movie_slot >= 0;
}

// ---- the kernel ---------------------------------------------------------------------------------------

// BitsStreamGenerator.nextDouble from its two words, ((long)(hi >>> 6) << 26 | lo >>> 6) * 2^-52, without an
// integer-to-double conversion: 1 + u 2^-52 is exact in [1, 2), and so is the subtraction.
__device__ __forceinline__ double crew_next_double(uint32_t hi, uint32_t lo)
{
    const unsigned long long u = ((unsigned long long)(hi >> 6) << 26) | (unsigned long long)(lo >> 6);
    return __longlong_as_double((long long)(0x3ff0000000000000ULL | u)) - 1.0;
}

// What the M and R streams say about the next move (Main.java:199-213, Molecule.java:67-68), drawn ahead of
// time: both are consumed at positions that do not depend on the outcome of the move in flight.
struct CrewCommon {
    double sn, cs;            // sin / cos of rotAmount, if the move turns out to be a rotation
    int kindbit;              // r.nextDouble() >= 0.5
    int magw_rot, magw_trans; // magwalk flag of a rotation / of a translation
    int axis, nR;             // rotation axis, words it takes from R
    bool ok;
};

// The S-stream part of a move starting at ring position `pos`: Space.randMolecule (Space.java:735-738) and, for a
// translation, the three components (Space.java:694-696).
struct CrewPick {
    double va, vb, vc;
    int m, off, sm, rot, nS;
    bool ok;
};

template <int L, int SPL, int LJS, bool HAS_EXP, bool TRACE, int EW, int MINB>
__global__ void __launch_bounds__(32 * (1 + EW), MINB)
crew_kernel(const __grid_constant__ KParams P, const int C, const int stride, const int max_mol_sites)
{
    using TL = CrewTables<L, SPL>;
    constexpr int TPW = 32 / L;                    // chains an energy warp evaluates side by side
    constexpr int NT = 32 * (1 + EW);
    extern __shared__ __align__(16) unsigned char smem[];
    const DevSystem &sys = P.sys;
    TL::load(smem, sys);
    unsigned char *chains = smem + TL::table_bytes(sys.n_types, sys.has_exp);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long chain0 = (long long)blockIdx.x * C;            // C <= EW * TPW: every chain has its own team
    const int nloc = (int)((P.n_chains - chain0 < C) ? P.n_chains - chain0 : C);
    const int tl = lane % L;                                         // lane inside its team
    const int team = lane / L;
    const unsigned mask = team_mask<L, 1>();

    // chain state -> shared memory (every team takes whole chains)
    for (int c0 = warp * TPW; c0 < nloc; c0 += (1 + EW) * TPW) {
        const int c = c0 + team;
        if (c < nloc) {
            const CrewView v = crew_view<L, SPL>(chains, c, stride, sys.n_mol, max_mol_sites);
            const long long ch = chain0 + c;
            const int *src = (const int *)(P.ctrl + ch);
            int *dst = (int *)&v.cs->ctrl;
            for (int i = tl; i < (int)(sizeof(ChainCtrl) / 4); i += L) dst[i] = src[i];
            const double *cg = P.com + (size_t)ch * sys.n_mol * 3;
            for (int i = tl; i < sys.n_mol * 3; i += L) v.com[i] = cg[i];
            crew_sites_from_global<L, SPL>(v, P.sites + (size_t)ch * sys.n_sites * 3, sys, tl);
            __syncwarp(mask);
            if (tl == 0) {
                v.cs->sch = P.sched[v.cs->ctrl.sched];
                v.cs->mt = P.mt + (size_t)ch * 3 * 2 * MT_N;
                v.cs->plan[0].mode = PLAN_IDLE; v.cs->plan[1].mode = PLAN_IDLE; v.cs->prev_m = -1;
                v.cs->bound = v.cs->ctrl.space_len * 0.75;
            }
            if (tl < 3) { v.cs->wr[tl] = 0; v.cs->rd[tl] = 0; }
            if (tl < TR_MAX_MOL_SITES) v.cs->prev_p[tl] = -1;
            __syncwarp(mask);
            if (!v.cs->ctrl.done) crew_ring_fill<L>(v.cs, tl, mask);
        }
    }
    __syncthreads();

#ifdef TR_CREW_TIMING
    long long tm_a = 0, tm_b = 0, tm_c = 0, tm_rounds = 0;
#define TM(...) __VA_ARGS__
#else
#define TM(...)
#endif
    if (warp == 0) {
        // ================================================================ control warp: lane c <-> chain c
        // Iteration r runs while the teams execute round r: it settles what they did in round r - 1 and writes
        // plan[(r + 1) & 1], what they do in round r + 1.  One barrier per round.
        const bool mine = lane < nloc;
        const CrewView v = crew_view<L, SPL>(chains, mine ? lane : 0, stride, sys.n_mol, max_mol_sites);
        CrewChain *cs = v.cs;
        ChainCtrl *c = &cs->ctrl;
        const tr_schedule &sch = cs->sch;
        const long long chain = chain0 + lane;
        const FastMod three = fastmod_make(3);
        const int S = sys.n_sites;

        // hot state in registers; everything else stays in c (shared memory)
        double t = 0.0, kBt = 0.0, running = 0.0, totalE = 0.0, totalSqE = 0.0, maxD = 0.0, maxRot = 0.0, maxTmag = 0.0;
        unsigned long long thr_rot = 0, thr_trans = 0;
        int z = 0, mpp = 1;
        int seg_moves = 0, seg_acc = 0, seg_eval = 0;            // since the last flush into c
        unsigned seg_sm = 0, seg_sm2 = 0;                         // pair evaluations = 2 * (S * sum s_m - sum s_m^2)
        int rdM = 0, rdS = 0, rdR = 0;                            // ring positions consumed so far
        int wrM = 0, wrS = 0, wrR = 0;                            // snapshot of the producers' cursors
        long long budget = P.max_moves > 0 ? P.max_moves : 0x7fffffffffffffffLL;
        bool live = mine && !c->done;
        bool hot = false;
        // pipeline state
        int mode_prev = PLAN_IDLE, mode_cur = PLAN_IDLE;         // what the team did in round r - 1 / does in round r
        bool total_for_init = false, filled = false;
        int commit_pending = 0;
        int refill_cur = 3;                                      // stream whose window is in flight in round r
        int cur_info = 0;                                        // move of round r: m | sm << 16 | rot << 21 | magwalk << 22 | (axis + 1) << 23
        int prev_info = 0;                                       // move of round r - 1
        bool metro_ok = false;                                   // (rnd, thr, grd) belong to the move of round r - 1 when it is settled
        double rnd = 0.0, thr = 0.0, grd = 0.0;
        CrewCommon cm; cm.sn = cm.cs = 0.0; cm.kindbit = cm.magw_rot = cm.magw_trans = 0; cm.axis = -1; cm.nR = 0; cm.ok = false;
        CrewPick cA, cB;                                         // successor if no Metropolis double is drawn / if one is
        cA.va = cA.vb = cA.vc = 0.0; cA.m = cA.off = cA.sm = cA.rot = cA.nS = 0; cA.ok = false;
        cB = cA;

        auto load_hot = [&]() {
            t = c->t; kBt = __dmul_rn(BOLTZMANN_CONSTANT, c->t);
            running = c->running; totalE = c->totalE; totalSqE = c->totalSqE;
            maxD = c->maxD; maxRot = c->maxRot; maxTmag = __dmul_rn(c->maxD, sch.magwalk_factor_trans);
            thr_rot = double_threshold52(c->probRot); thr_trans = double_threshold52(c->probTrans);
            z = c->z; mpp = sch.moves_per_point; hot = true;
        };
        auto flush_hot = [&]() {
            c->z = z; c->running = running; c->totalE = totalE; c->totalSqE = totalSqE;
            c->total += seg_moves;                 // total++ per move (Main.java:196)
            c->accepted += seg_acc; c->accepted_all += seg_acc;
            c->moves += seg_moves; c->evaluated += seg_eval;
            c->pair_evals += 2LL * ((long long)S * seg_sm - seg_sm2);
            seg_moves = seg_acc = seg_eval = 0; seg_sm = seg_sm2 = 0;
        };
        auto ringM = [&](int pos) { return cs->ring[STREAM_M][pos & (CREW_RING - 1)]; };
        auto ringS = [&](int pos) { return cs->ring[STREAM_S][pos & (CREW_RING - 1)]; };
        auto ringR = [&](int pos) { return cs->ring[STREAM_R][pos & (CREW_RING - 1)]; };
        // Main.java:199-213 + Molecule.java:67-68 at the current M / R positions (nothing is consumed here)
        auto draw_common = [&]() {
            cm.ok = false; cm.nR = 0; cm.axis = -1;
            if (wrM - rdM < 4) return;
            // r.nextDouble() >= 0.5 (Main.java:199) is the top bit of the first word; both words are consumed.
            // The second double (:200 or :208) is compared as a 52-bit integer against ceil(prob * 2^52).
            const uint32_t w0 = ringM(rdM);
            const unsigned long long u2 = ((unsigned long long)(ringM(rdM + 2) >> 6) << 26) | (ringM(rdM + 3) >> 6);
            cm.kindbit = (int)(w0 >> 31);
            cm.magw_rot = (u2 >= thr_rot) ? 0 : 1;
            cm.magw_trans = (u2 >= thr_trans) ? 0 : 1;
            if (cm.kindbit) {
                if (wrR - rdR < 3) return;
                const double maxR = cm.magw_rot ? JAVA_TWO_PI : maxRot;
                const double u = crew_next_double(ringR(rdR), ringR(rdR + 1));                       // Molecule.java:67
                const double rotAmount = __dmul_rn(__dmul_rn(__dsub_rn(u, 0.5), 2.0), maxR);
                int n = 2;
                for (;;) {                                                                          // Molecule.java:68: nextInt(3)
                    if (rdR + n >= wrR) return;
                    const int bits = (int)(ringR(rdR + n) >> 1);
                    n++;
                    cm.axis = (int)fastmod(three, (uint32_t)bits);
                    if ((int)((uint32_t)bits - (uint32_t)cm.axis + 2u) >= 0) break;
                }
                cm.nR = n;
                sincos(rotAmount, &cm.sn, &cm.cs);
            }
            cm.ok = true;
        };
        auto draw_pick = [&](CrewPick &k, int pos) {
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
            k.off = TL::mol_off(smem, k.m);
            k.sm = TL::mol_off(smem, k.m + 1) - k.off;
            k.rot = (cm.kindbit && k.sm > 1) ? 1 : 0;
            if (!k.rot) {
                if (p + 6 + CREW_SETTLE_SLACK > wrS) return;
                const double maxT = cm.magw_trans ? maxTmag : maxD;
                const double ux = crew_next_double(ringS(p), ringS(p + 1));                          // Space.java:694-696
                const double uy = crew_next_double(ringS(p + 2), ringS(p + 3));
                const double uz = crew_next_double(ringS(p + 4), ringS(p + 5));
                k.va = __dsub_rn(__dmul_rn(__dmul_rn(ux, 2.0), maxT), maxT);
                k.vb = __dsub_rn(__dmul_rn(__dmul_rn(uy, 2.0), maxT), maxT);
                k.vc = __dsub_rn(__dmul_rn(__dmul_rn(uz, 2.0), maxT), maxT);
                p += 6;
            }
            if (p + CREW_SETTLE_SLACK > wrS) return;      // the move's own Metropolis double must already be in the ring
            k.nS = p - pos;
            k.ok = true;
        };
        // The descriptor the energy team reads for a drawn move.
        auto write_desc = [&](const CrewPick &k, CrewDesc *dst) {
            CrewDesc d;
            const int axis = k.rot ? cm.axis : -1;
            d.va = k.rot ? cm.sn : k.va; d.vb = k.rot ? cm.cs : k.vb; d.vc = k.rot ? 0.0 : k.vc;
            d.m = k.m; d.off = k.off; d.info = k.sm | (k.rot << 8) | ((axis + 1) << 12); d.pad_ = 0;
            *dst = d;
        };
        // The stream cursors move past a drawn move's words; remember what bookkeeping and the trace need of it.
        auto adopt = [&](const CrewPick &k) {
            const int magwalk = k.rot ? cm.magw_rot : cm.magw_trans;
            const int axis = k.rot ? cm.axis : -1;
            cur_info = k.m | (k.sm << 16) | (k.rot << 21) | (magwalk << 22) | ((axis + 1) << 23);
            rdS += k.nS; rdM += 4; rdR += k.rot ? cm.nR : 0;
        };
        // Which ring gets the window of round r + 1: the one with the fewest unread words if that is few enough, never
        // the one whose window is in flight in round r (its cursor is not final yet); 3 = none.
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

        TM(const long long tm_start = clock64();)
        for (int r = -1;; r++) {
            TM(const long long tm0 = clock64();)
            CrewPlan *pl = &cs->plan[(r + 1) & 1];
            int mode = PLAN_IDLE;
            if (live) {
                wrM = cs->wr[STREAM_M]; wrS = cs->wr[STREAM_S]; wrR = cs->wr[STREAM_R];
                // ---------------------------------------------- settle round r - 1
                if (mode_prev == PLAN_AUTO || mode_prev == PLAN_MOVE) {
                    // Metropolis: Space.java:668-689 / :711-733 (the team took the same decision if round r is PLAN_AUTO)
                    const CrewResult *res = &cs->result[(r - 1) & 1];
                    const int oob = res->oob;
                    const double dE = oob ? 0.0 : res->dE;
                    int acc, drew;
                    // (the words are there: a move is only ever posted with at least 16 unread S words beyond its own draws,
                    // see draw_pick's availability tests and CREW_SETTLE_SLACK)
                    if (!metro_ok && !oob && dE > 0.0) rnd = crew_next_double(ringS(rdS), ringS(rdS + 1));
                    crew_metropolis(oob, dE, t == 0.0, metro_ok, rnd, thr, grd, kBt, acc, drew);
                    if (acc) commit_pending = 1;
                    rdS += 2 * drew;
                    // Main.java:215-223
                    running = __dadd_rn(running, acc ? dE : 0.0);
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
                        CrewPick k = cA;
                        if (drew) k = cB;
                        adopt(k);
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
                }
                metro_ok = false;

                // ---------------------------------------------- plan round r + 1
                int flags = 0, refill = 3, snap_slot = -1, movie_slot = -1;
                if (mode_cur == PLAN_AUTO || mode_cur == PLAN_MOVE) {
                    // a move is being evaluated now: draw what its Metropolis test will need and both possible successors;
                    // if all of that exists, the team may go on by itself
                    bool chain_ok = false;
                    if (wrS - rdS >= 2) {
                        rnd = crew_next_double(ringS(rdS), ringS(rdS + 1));
                        if (t != 0.0) {
                            thr = __dmul_rn(-kBt, log(rnd));
                            grd = (kBt + fabs(thr)) * 1e-11;
                        }
                        metro_ok = true;
                        if (z + 1 < mpp && budget > 0) {         // the move in flight has a successor inside this point and this launch
                            draw_common();
                            if (cm.ok) { draw_pick(cA, rdS); draw_pick(cB, rdS + 2); }
                            chain_ok = cm.ok && cA.ok && cB.ok;
                        }
                    }
                    if (chain_ok) {
                        mode = PLAN_AUTO;
                        write_desc(cA, &pl->cand[0]); write_desc(cB, &pl->cand[1]);
                        pl->rnd = rnd; pl->thr = thr; pl->grd = grd; pl->kBt = kBt;
                        flags = (t == 0.0) ? PF_TZERO : 0;
                        refill = choose_refill();
                        budget--;
                    }
                    prev_info = cur_info;
                } else if (mode_cur == PLAN_IDLE) {
                    // nothing is running: everything that is not "next move of the same point" is decided here
                    bool fill = false;
                    if (!c->done && !c->started) {
                        mode = PLAN_TOTAL; total_for_init = true;
                        snap_slot = (P.snap_sites && c->n_snaps < P.max_snaps) ? c->n_snaps : -1;
                    } else if (!c->done && z >= mpp) {
                        // end of the point: with a calcEnergy round first when its result is needed
                        flush_hot();
                        if (crew_point_needs_total(P, c, sch, snap_slot, movie_slot)) mode = PLAN_TOTAL;
                        else {
                            crew_finish_point(P, c, sch, chain, 0.0, false);
                            load_hot();
                        }
                    }
                    if (!c->done && mode == PLAN_IDLE && budget > 0) {
                        draw_common();
                        if (cm.ok) draw_pick(cA, rdS);
                        if (cm.ok && cA.ok) {
                            write_desc(cA, &pl->cand[0]);
                            adopt(cA);
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
                    if (c->done && mode == PLAN_IDLE) live = false;
                }
                // (PLAN_TOTAL / PLAN_SYNC running now: wait for it)
                pl->mode = mode; pl->flags = flags; pl->refill = refill; pl->snap_slot = snap_slot; pl->movie_slot = movie_slot;
                refill_cur = refill;
            } else if (mine) pl->mode = PLAN_IDLE;
            const unsigned any = __ballot_sync(0xffffffffu, mode != PLAN_IDLE || mode_cur != PLAN_IDLE);
            mode_prev = mode_cur; mode_cur = mode;
            if (lane == 0) TL::alive(smem, r + 1) = (int)any;
            TM(const long long tm1 = clock64(); tm_a += tm1 - tm0; tm_rounds++;)
            crew_bar(NT);
            TM(tm_c += clock64() - tm1;)
            if (!any) break;
        }
        TM(if (blockIdx.x == 3 && lane == 0) printf("block 3 control: rounds %lld work %lld wait %lld total %lld cycles/round\n", tm_rounds, tm_a / tm_rounds, tm_c / tm_rounds, (clock64() - tm_start) / tm_rounds);)
        // park the chain: counters into c, unread ring words back to the stream cursors
        if (mine) {
            if (hot) flush_hot();
            mt_cursor_rewind(c->rng[STREAM_M], cs->wr[STREAM_M] - rdM);
            mt_cursor_rewind(c->rng[STREAM_S], cs->wr[STREAM_S] - rdS);
            mt_cursor_rewind(c->rng[STREAM_R], cs->wr[STREAM_R] - rdR);
        }
    } else {
        // ================================================================ energy warps: one team <-> one chain
        const int c = (warp - 1) * TPW + team;
        const bool has = c < nloc;
        const CrewView v = crew_view<L, SPL>(chains, has ? c : 0, stride, sys.n_mol, max_mol_sites);
        CrewChain *cs = v.cs;
        int info[SPL];
        double q[SPL];
#pragma unroll
        for (int k = 0; k < SPL; k++) { info[k] = TL::pos_info(smem, tl + L * k); q[k] = TL::qpos(smem, tl + L * k); }
        TM(long long tmv[8] = {0, 0, 0, 0, 0, 0, 0, 0};)
        double last_dE = 0.0;                               // result of the move this team evaluated last
        int last_oob = 0;
        crew_bar(NT);                                       // plan[0] is written
        for (int r = 0;; r++) {
            if (TL::alive(smem, r) == 0) break;
            TM(const long long tm0 = clock64(); tm_rounds++;)
            if (has) {
                const CrewPlan *pl = &cs->plan[r & 1];
                const int mode = pl->mode;
                // what a commit needs sits at fixed addresses: these loads do not wait for the plan
                int pp = -1, pm = -1;
                double cnx = 0.0, cny = 0.0, cnz = 0.0, cmx = 0.0, cmy = 0.0, cmz = 0.0;
                if (tl < TR_MAX_MOL_SITES) {
                    pp = cs->prev_p[tl];
                    const StageSite *st = &v.stage[tl < max_mol_sites ? tl : 0];
                    cnx = st->nx; cny = st->ny; cnz = st->nz;
                }
                if (tl == 0) { pm = cs->prev_m; cmx = cs->prev_v[0]; cmy = cs->prev_v[1]; cmz = cs->prev_v[2]; }
                if (mode != PLAN_IDLE) {
                    TM(const long long tm1 = clock64();)
                    const int flags = pl->flags;
                    const CrewWindow<L> w = crew_ring_begin<L>(cs, pl->refill, tl);    // loads fly while the move is evaluated
                    TMG(3, TMCLK() - tm1);
                    int commit = flags & PF_COMMIT, which = 0;
                    if (mode == PLAN_AUTO) {
                        int acc, drew;
                        crew_metropolis(last_oob, last_dE, (flags & PF_TZERO) != 0, true, pl->rnd, pl->thr, pl->grd, pl->kBt, acc, drew);
                        commit = acc; which = drew;
                    }
                    if (commit) {
                        // setTemps (Molecule.java:128-135) of the move staged last
                        if (pp >= 0) { v.xs[pp] = cnx; v.ys[pp] = cny; v.zs[pp] = cnz; }
                        if (pm >= 0) { v.com[3 * pm] = cmx; v.com[3 * pm + 1] = cmy; v.com[3 * pm + 2] = cmz; }    // m.x = m.tempx
                        __syncwarp(mask);
                    }
                    TMG(4, TMCLK() - tm1);
                    if (mode == PLAN_AUTO || mode == PLAN_MOVE) {
                        const CrewDesc d = pl->cand[which];
                        crew_move<L, SPL, LJS, HAS_EXP, TRACE>(smem, v, sys, d, info, q, tl, mask, &cs->result[r & 1], last_dE, last_oob TMV_ARG);
                    } else if (mode == PLAN_TOTAL) {
                        const double e = crew_total_energy<L, SPL, HAS_EXP>(smem, v, sys, info, q, tl, mask);
                        if (tl == 0) cs->etot = e;
                        const int k = pl->snap_slot, mv = pl->movie_slot;
                        if (k >= 0)
                            crew_sites_to_global<L, SPL>(v, P.snap_sites + ((size_t)(chain0 + c) * P.max_snaps + k) * (size_t)sys.n_sites * 3, sys, tl);
                        if (mv >= 0)
                            crew_sites_to_global<L, SPL>(v, P.movie_sites + ((size_t)(chain0 + c) * P.max_points + mv) * (size_t)sys.n_sites * 3, sys, tl);
                    }
                    TM(const long long tm2 = clock64(); tm_a += tm2 - tm1;)
                    crew_ring_end<L>(cs, w, tl, mask);
                    if (flags & PF_FILL) crew_ring_fill<L>(cs, tl, mask);
                    TM(tm_b += clock64() - tm2;)
                }
            }
            TM(tm_c += clock64() - tm0;)
            crew_bar(NT);
        }
        TM(if (blockIdx.x == 3 && lane == 0) printf("block 3 warp %d per round: geometry %lld pair %lld reduce %lld ring_begin %lld (begin+commit %lld)\n", warp, tmv[0] / tm_rounds, tmv[1] / tm_rounds, tmv[2] / tm_rounds, tmv[3] / tm_rounds, tmv[4] / tm_rounds);)
        TM(if (blockIdx.x == 3 && lane == 0) printf("block 3 warp %d: rounds %lld busy %lld (work %lld ring_end %lld) cycles/round\n", warp, tm_rounds, tm_c / tm_rounds, tm_a / tm_rounds, tm_b / tm_rounds);)
    }
    __syncthreads();

    // chain state -> global memory
    for (int c0 = warp * TPW; c0 < nloc; c0 += (1 + EW) * TPW) {
        const int c = c0 + team;
        if (c >= nloc) continue;
        const CrewView v = crew_view<L, SPL>(chains, c, stride, sys.n_mol, max_mol_sites);
        const long long ch = chain0 + c;
        int *dst = (int *)(P.ctrl + ch);
        const int *src = (const int *)&v.cs->ctrl;
        for (int i = tl; i < (int)(sizeof(ChainCtrl) / 4); i += L) dst[i] = src[i];
        double *cg = P.com + (size_t)ch * sys.n_mol * 3;
        for (int i = tl; i < sys.n_mol * 3; i += L) cg[i] = v.com[i];
        crew_sites_to_global<L, SPL>(v, P.sites + (size_t)ch * sys.n_sites * 3, sys, tl);
    }
#undef TM
}

}  // namespace tr
