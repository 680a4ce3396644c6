// Kernels around the hot loop: chain initialisation (seed fan-out, propagate), centres of
// mass, parity probes, RNG probe, minima packing, and the DFMA peak micro-benchmark.
#pragma once
#include "anneal.cuh"

namespace tr {

struct InitParams {
    DevSystem sys;
    const tr_schedule *sched;
    ChainCtrl *ctrl;
    double *sites;
    double *com;
    uint32_t *mt;
    const long long *seeds;        // may be null when explicit RNG states were uploaded
    const int *sched_idx;          // may be null
    const int *mt_pos;             // [n_chains][3] explicit mti, or null
    long long n_chains;
    double space_len;
    int max_prop_failures;
    int propagate;                 // 1: Space.propagate on the device; 0: sites/com already set
    int n_sched;
};

// Molecule.setCenterOfMass (Molecule.java:137-156), same accumulation order, no contraction.
__device__ inline void center_of_mass(const DevSystem &sys, const double *sites, int m, double *out)
{
    double tot = 0.0, px = 0.0, py = 0.0, pz = 0.0;
    for (int s = sys.mol_off[m]; s < sys.mol_off[m + 1]; s++) {
        if (sys.site_ghost[s]) continue;
        const double w = sys.site_mass[s];
        tot = __dadd_rn(tot, w);
        px = __dadd_rn(px, __dmul_rn(sites[3 * s], w));
        py = __dadd_rn(py, __dmul_rn(sites[3 * s + 1], w));
        pz = __dadd_rn(pz, __dmul_rn(sites[3 * s + 2], w));
    }
    out[0] = __ddiv_rn(px, tot);
    out[1] = __ddiv_rn(py, tot);
    out[2] = __ddiv_rn(pz, tot);
}

__global__ void com_kernel(DevSystem sys, const double *sites, double *com, long long n_chains)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains * sys.n_mol) return;
    const long long chain = i / sys.n_mol;
    const int m = (int)(i % sys.n_mol);
    center_of_mass(sys, sites + (size_t)chain * sys.n_sites * 3, m, com + ((size_t)chain * sys.n_mol + m) * 3);
}

// ---- chain initialisation: one WARP per chain --------------------------------------------------------------------------------
// Main.java:34-37 (seed fan-out), Main.java:58-60 + Space.java:72-117 (propagate with box growth, then the random initial
// rotation), and the schedule locals of Main.java:178-186.  The initial calcEnergy / write(0) is done by the first anneal launch.
//
// What is serial stays on one lane (MersenneTwister.setSeed is a recurrence; the order of the draws is the reference's), what
// is not is spread over the warp: the three stream seedings run side by side on three lanes, a 624-word block is regenerated
// 32 words at a time, an attempted position is tested against the particles placed so far one particle per lane, sites are
// put and rotated one site per lane, and the stream states leave for global memory in coalesced rows.  The streams live in
// shared memory while the chain is set up.

constexpr int INIT_WARPS = 4;            // chains per block

// mt[k] of the next block from the current one, in place, by the whole warp: new[k] = (k < 227 ? old[k+397] : new[k-227]) ^
// twist(old[k], k == 623 ? new[0] : old[k+1]).  Chunks of 32 in ascending order: everything a chunk reads is either still old
// (k + 1, k + 397 lie in later chunks) or already new (k - 227 lies at least seven chunks back; new[0] for k = 623).
__device__ __forceinline__ void warp_mt_regen(uint32_t *mt, int lane)
{
    for (int base = 0; base < MT_N; base += 32) {
        const int k = base + lane;
        uint32_t v = 0;
        if (k < MT_N) {
            const uint32_t a = mt[k];
            const uint32_t b = mt[k == MT_N - 1 ? 0 : k + 1];
            const uint32_t far = mt[k < MT_N - MT_M ? k + MT_M : k + MT_M - MT_N];
            const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
            v = far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (k < MT_N) mt[k] = v;
        __syncwarp();
    }
}

// MersenneTwister.next(32) for the whole warp (every lane gets the word; mti is warp-uniform).
__device__ __forceinline__ uint32_t warp_mt_next32(uint32_t *mt, int &mti, int lane)
{
    if (mti >= MT_N) { warp_mt_regen(mt, lane); mti = 0; }
    return mt_temper(mt[mti++]);
}

__device__ __forceinline__ double warp_mt_next_double(uint32_t *mt, int &mti, int lane)
{
    const uint64_t hi = (uint64_t)(warp_mt_next32(mt, mti, lane) >> 6) << 26;
    const uint64_t lo = warp_mt_next32(mt, mti, lane) >> 6;
    return (double)(int64_t)(hi | lo) * 0x1.0p-52;
}

__device__ __forceinline__ int warp_mt_next_int(uint32_t *mt, int &mti, int n, int lane)
{
    if ((n & -n) == n) return (int)(((int64_t)n * (int64_t)(warp_mt_next32(mt, mti, lane) >> 1)) >> 31);
    int bits, val;
    do {
        bits = (int)(warp_mt_next32(mt, mti, lane) >> 1);
        val = bits % n;
    } while ((int)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
    return val;
}

__global__ void __launch_bounds__(32 * INIT_WARPS) init_chains_kernel(InitParams ip)
{
    __shared__ uint32_t s_mt[INIT_WARPS][3][MT_N];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chain = (long long)blockIdx.x * INIT_WARPS + warp;
    if (chain >= ip.n_chains) return;
    const DevSystem &sys = ip.sys;
    uint32_t *mtM = s_mt[warp][STREAM_M], *mtS = s_mt[warp][STREAM_S], *mtR = s_mt[warp][STREAM_R];
    // buffer 0 of each stream's ping-pong pair in global memory receives the reference-form state at the end
    uint32_t *gM = ip.mt + ((size_t)chain * 3 + STREAM_M) * 2 * MT_N;
    uint32_t *gS = ip.mt + ((size_t)chain * 3 + STREAM_S) * 2 * MT_N;
    uint32_t *gR = ip.mt + ((size_t)chain * 3 + STREAM_R) * 2 * MT_N;
    int posM = MT_N, posS = MT_N, posR = MT_N;
    if (ip.seeds) {
        // seedGenerator (Main.java:24,34): one lane seeds it in the M buffer, the warp regenerates, three nextLong()
        if (lane == 0) mt_std_seed_long(mtM, ip.seeds[chain]);
        __syncwarp();
        int mti = MT_N;
        long long sd[3];
        for (int i = 0; i < 3; i++) {
            const uint64_t hi = (uint64_t)warp_mt_next32(mtM, mti, lane) << 32;
            const uint64_t lo = warp_mt_next32(mtM, mti, lane);
            sd[i] = (long long)(hi | lo);
        }
        __syncwarp();
        // Main.setSeed / Space.setSeed / Molecule.setSeed (Main.java:35-37), side by side
        if (lane < 3) mt_std_seed_long(s_mt[warp][lane], lane == 0 ? sd[0] : (lane == 1 ? sd[1] : sd[2]));
        __syncwarp();
    } else {
        for (int s = 0; s < 3; s++)
            for (int k = lane; k < MT_N; k += 32) s_mt[warp][s][k] = ip.mt[((size_t)chain * 3 + s) * 2 * MT_N + k];
        if (ip.mt_pos) {
            posM = ip.mt_pos[chain * 3 + STREAM_M];
            posS = ip.mt_pos[chain * 3 + STREAM_S];
            posR = ip.mt_pos[chain * 3 + STREAM_R];
        }
        __syncwarp();
    }

    double L = ip.space_len;
    int retries = 0;
    double *sites = ip.sites + (size_t)chain * sys.n_sites * 3;
    double *com = ip.com + (size_t)chain * sys.n_mol * 3;
    if (ip.propagate) {
        const int N = sys.n_mol;
        for (;;) {                                        // while (!s.propagate()) spaceLen *= 1.1
            bool failed = false;
            for (int li = 0; li < N && !failed; li++) {
                int c = 0;
                const double rad = sys.mol_radius[li];
                for (;;) {
                    const double half = __ddiv_rn(L, 2.0);
                    const double x = __dsub_rn(__dmul_rn(warp_mt_next_double(mtS, posS, lane), L), half);
                    const double y = __dsub_rn(__dmul_rn(warp_mt_next_double(mtS, posS, lane), L), half);
                    const double z = __dsub_rn(__dmul_rn(warp_mt_next_double(mtS, posS, lane), L), half);
                    // Space.java:84-93: too close to a particle placed earlier?  (the reference stops at the first hit; any hit
                    // fails the attempt, so the particles are tested one per lane)
                    bool hit = false;
                    for (int n = lane; n < li; n += 32) {
                        const double mind = __dadd_rn(rad, sys.mol_radius[n]);
                        const double dx = __dsub_rn(com[3 * n], x), dy = __dsub_rn(com[3 * n + 1], y), dz = __dsub_rn(com[3 * n + 2], z);
                        const double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
                        hit |= d < mind;
                    }
                    if (!__any_sync(0xffffffffu, hit)) {
                        // Molecule.put (Molecule.java:51-62), one site per lane
                        const int s0 = sys.mol_off[li], s1 = sys.mol_off[li + 1];
                        for (int s = s0 + lane; s < s1; s += 32) {
                            sites[3 * s] = __dadd_rn(sys.site_def[3 * s], x);
                            sites[3 * s + 1] = __dadd_rn(sys.site_def[3 * s + 1], y);
                            sites[3 * s + 2] = __dadd_rn(sys.site_def[3 * s + 2], z);
                        }
                        __syncwarp();
                        if (lane == 0) center_of_mass(sys, sites, li, com + 3 * li);
                        __syncwarp();
                        break;
                    }
                    c++;
                    if (c >= ip.max_prop_failures) { failed = true; break; }
                }
            }
            if (!failed) break;
            L = __dmul_rn(L, 1.1);
            retries++;
        }
        // Space.java:110-114: rotateTemp(2*PI) + setTemps for every particle, in space order; one site per lane
        for (int m = 0; m < N; m++) {
            const int s0 = sys.mol_off[m], s1 = sys.mol_off[m + 1];
            if (s1 - s0 <= 1) continue;
            const double u = warp_mt_next_double(mtR, posR, lane);
            const double rotAmount = __dmul_rn(__dmul_rn(__dsub_rn(u, 0.5), 2.0), JAVA_TWO_PI);
            const int axis = warp_mt_next_int(mtR, posR, 3, lane);
            double sn, cs;
            sincos(rotAmount, &sn, &cs);
            const double nsn = __dmul_rn(-1.0, sn);
            const double cx = com[3 * m], cy = com[3 * m + 1], cz = com[3 * m + 2];
            for (int s = s0 + lane; s < s1; s += 32) {
                const double ax = __dsub_rn(sites[3 * s], cx), ay = __dsub_rn(sites[3 * s + 1], cy), az = __dsub_rn(sites[3 * s + 2], cz);
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
                sites[3 * s] = __dadd_rn(tx, cx);
                sites[3 * s + 1] = __dadd_rn(ty, cy);
                sites[3 * s + 2] = __dadd_rn(tz, cz);
            }
        }
    }
    __syncwarp();
    // the streams leave for global memory in coalesced rows
    for (int k = lane; k < MT_N; k += 32) { gM[k] = mtM[k]; gS[k] = mtS[k]; gR[k] = mtR[k]; }

    if (lane == 0) {
        int si = ip.sched_idx ? ip.sched_idx[chain] : 0;
        if (si < 0 || si >= ip.n_sched) si = 0;
        const tr_schedule sch = ip.sched[si];
        ChainCtrl c;
        memset(&c, 0, sizeof c);
        c.t = sch.max_temp; c.saveT = sch.max_temp; c.delT = 0.0;
        c.maxD = sch.max_trans; c.maxRot = sch.max_rot;
        c.probRot = sch.magwalk_prob_rot; c.probTrans = sch.magwalk_prob_trans;
        c.space_len = L;
        c.min_energy = 1.7976931348623157e308;          // Double.MAX_VALUE (Space.java:23)
        c.ptsPerTooth = sch.static_temp ? 1 : sch.points_per_tooth;   // Main.java:74-76
        c.sched = si;
        c.prop_retries = retries;
        c.rng[STREAM_M].par = 0; c.rng[STREAM_M].gen = MT_N; c.rng[STREAM_M].fed = posM;
        c.rng[STREAM_S].par = 0; c.rng[STREAM_S].gen = MT_N; c.rng[STREAM_S].fed = posS;
        c.rng[STREAM_R].par = 0; c.rng[STREAM_R].gen = MT_N; c.rng[STREAM_R].fed = posR;
        ip.ctrl[chain] = c;
    }
}

// ---- probes -----------------------------------------------------------------------------------

// Space.calcEnergy on supplied structures: one team per structure, the same device routine
// the anneal kernel uses for write(n) / writeMovie energies.
template <int L, int W, int SPL, bool HAS_EXP>
__global__ void __launch_bounds__(Team<L, W>::TPC) total_energy_kernel(DevSystem sys, const double *sites_g, long long n, double *out)
{
    using Lay = Layout<L, W, SPL>;
    constexpr int TPC = Lay::TPC;
    extern __shared__ __align__(16) unsigned char smem[];
    Lay::load_tables(smem, sys);
    ChainSmem *cs = Lay::chain(smem, 0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mask = team_mask<L, W>();
    __syncthreads();
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
        sites_from_global<TPC, Lay::NPOS>(Lay::sites(cs), sites_g + (size_t)i * sys.n_sites * 3, sys, tid);
        __syncthreads();
        const double e = team_total_energy<L, W, SPL, HAS_EXP>(smem, cs, sys, tid, warp, lane, mask);
        if (tid == 0) out[i] = e;
        __syncthreads();
    }
}

// eStart / eEnd (Space.java:711-718) for supplied trial coordinates of one particle: the
// staging + move_energy + reduction path of the anneal kernel.
template <int L, int W, int SPL, bool HAS_EXP>
__global__ void __launch_bounds__(Team<L, W>::TPC) delta_energy_kernel(DevSystem sys, const double *sites_g, const int *mol, const double *trial,
                                                                       long long n, double *e_start, double *e_end, double *d_e)
{
    using Lay = Layout<L, W, SPL>;
    constexpr int TPC = Lay::TPC;
    extern __shared__ __align__(16) unsigned char smem[];
    Lay::load_tables(smem, sys);
    ChainSmem *cs = Lay::chain(smem, 0);
    SiteRec *sites = Lay::sites(cs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned mask = team_mask<L, W>();
    const unsigned char *tab_cd = Lay::cd(smem), *tab_ab = Lay::ab(smem, sys);
    __syncthreads();
    int info[SPL];
#pragma unroll
    for (int k = 0; k < SPL; k++) info[k] = Lay::pos_info(smem, tid + k * TPC);
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
        sites_from_global<TPC, Lay::NPOS>(sites, sites_g + (size_t)i * sys.n_sites * 3, sys, tid);
        __syncthreads();
        const int m = mol[i];
        const int off = Lay::mol_off(smem, m), sm = Lay::mol_off(smem, m + 1) - off;
        const double *tr_i = trial + (size_t)i * TR_MAX_MOL_SITES * 3;
        if (tid < sm) {
            const int p = Lay::site_pos(smem, off + tid);
            const SiteRec r = sites[p];
            const int inf = Lay::pos_info(smem, p);
            StageSite st;
            st.ox = r.x; st.oy = r.y; st.oz = r.z;
            st.nx = tr_i[3 * tid]; st.ny = tr_i[3 * tid + 1]; st.nz = tr_i[3 * tid + 2];
            st.kq = __dmul_rn(K_VALUE, r.q);
            st.row = info_type(inf) * sys.row_stride; st.lj = info_lj(inf);
            cs->stage[tid] = st;
        }
        __syncthreads();
        double eS, eE;
        move_energy<TPC, SPL, SPL, HAS_EXP>(sites + tid, info, cs, sm, m, tab_cd, tab_ab, sys.zero_col, eS, eE);
        const double dE = team_sum<L, W>(eE - eS, cs->part, warp, lane, mask);
        __syncthreads();
        eS = team_sum<L, W>(eS, cs->part, warp, lane, mask);
        __syncthreads();
        eE = team_sum<L, W>(eE, cs->part, warp, lane, mask);
        if (tid == 0) { e_start[i] = eS; e_end[i] = eE; d_e[i] = dE; }
        __syncthreads();
    }
}

// First k outputs of the three streams of each seed, drawn through the ring generator the anneal
// kernel uses (one warp per seed; `mt` is scratch [n][3][2][624]), in bursts of the sizes a move uses.
template <int L>
__global__ void __launch_bounds__(4 * L) rng_words_kernel(const long long *seeds, long long n, int k, uint32_t *mt, uint32_t *out)
{
    __shared__ uint32_t s_ring[4][MT_RING];
    __shared__ MtCursor s_cur[4];
    const int w = threadIdx.x / L;
    const long long i = (long long)blockIdx.x * 4 + w;
    const int lane = threadIdx.x % L;
    const unsigned mask = team_mask<L, 1>();
    if (i >= n) return;
    uint32_t *st = mt + (size_t)i * 3 * 2 * MT_N;
    if (lane == 0) {
        mt_std_seed_long(st, seeds[i]);
        int mti = MT_N;
        const int64_t s0 = mt_std_next_long(st, mti), s1 = mt_std_next_long(st, mti), s2 = mt_std_next_long(st, mti);
        mt_std_seed_long(st, s0);
        mt_std_seed_long(st + 2 * MT_N, s1);
        mt_std_seed_long(st + 4 * MT_N, s2);
    }
    __syncwarp(mask);
    for (int s = 0; s < 3; s++) {
        if (lane == 0) { s_cur[w].par = 0; s_cur[w].gen = MT_N; s_cur[w].fed = MT_N; }
        __syncwarp(mask);
        MtRing g;
        mt_ring_open(g, st + s * 2 * MT_N, s_ring[w], &s_cur[w]);
        int j = 0, burst = 1;
        bool resumed = false;
        while (j < k) {
            const int need = (k - j < burst) ? k - j : burst;
            mt_ring_reserve<L>(g, need, lane, mask);
            for (int b = 0; b < need; b++, j++) {
                const uint32_t v = mt_ring_next32(g);
                if (lane == 0) out[((size_t)i * 3 + s) * k + j] = v;
            }
            burst = burst % 11 + 1;      // 1..11 words per reservation, like the moves
            if (!resumed && j >= k / 2) {     // suspend / resume, as between two launches
                mt_ring_close(g, lane, mask);
                mt_ring_open(g, st + s * 2 * MT_N, s_ring[w], &s_cur[w]);
                resumed = true;
            }
        }
        __syncwarp(mask);
    }
}

// The reference's view of every stream (624 words of the current block + mti).  Completes the
// block in place first (harmless: the cursor records it), so chains stay resumable.
__global__ void __launch_bounds__(128) rng_export_kernel(ChainCtrl *ctrl, uint32_t *mt, long long n_chains, uint32_t *out)
{
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_chains) return;
    for (int s = 0; s < 3; s++) {
        MtCursor c = ctrl[i].rng[s];
        uint32_t *B = mt + (((size_t)i * 3 + s) * 2 + c.par) * MT_N;
        const uint32_t *O = mt + (((size_t)i * 3 + s) * 2 + (c.par ^ 1)) * MT_N;
        for (int base = c.gen; base < MT_N; base += 32) {       // gen is window-aligned whenever it is < 624
            const int k = base + lane;
            if (k < MT_N) {
                const uint32_t a = O[k];
                const uint32_t b = (k == MT_N - 1) ? B[0] : O[k + 1];
                const uint32_t far = (k < MT_N - MT_M) ? O[k + MT_M] : B[k + MT_M - MT_N];
                const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
                B[k] = far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncwarp();
        }
        uint32_t *dst = out + ((size_t)i * 3 + s) * MT_WORDS;
        for (int k = lane; k < MT_N; k += 32) dst[k] = B[k];
        if (lane == 0) { dst[MT_N] = (uint32_t)c.fed; ctrl[i].rng[s].gen = MT_N; }
        __syncwarp();
    }
}

__global__ void pack_minima_kernel(const ChainCtrl *ctrl, long long n_chains, long long first_chain, tr_minimum *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains) return;
    tr_minimum m;
    m.energy = ctrl[i].min_energy;
    m.chain = first_chain + i;
    m.tooth = ctrl[i].min_tooth;
    m.pad_ = 0;
    out[i] = m;
}

// The structure write(min_tooth) saw, per chain: [n_chains][row] (what writeMinEnergy copies, Space.java:836-844).
// The first snapshot with the lowest energy wins (strict < of Space.java:614).
__global__ void pack_min_structures_kernel(const double *snap_energy, const int *snap_tooth, const double *snap_sites, long long n_chains,
                                           int max_snaps, int row, double *out)
{
    const long long total = n_chains * row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long chain = i / row;
        const int j = (int)(i % row);
        int best = -1;
        double eb = 0.0;
        for (int k = 0; k < max_snaps; k++) {
            if (snap_tooth[chain * max_snaps + k] < 0) continue;
            const double e = snap_energy[chain * max_snaps + k];
            if (best < 0 || e < eb) { best = k; eb = e; }
        }
        out[i] = best >= 0 ? snap_sites[((size_t)chain * max_snaps + best) * row + j] : __longlong_as_double(0x7ff8000000000000LL);
    }
}

__global__ void summary_kernel(const ChainCtrl *ctrl, long long n_chains, tr_chain_summary *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains) return;
    const ChainCtrl &c = ctrl[i];
    tr_chain_summary s;
    s.start_energy = c.start_energy; s.running_energy = c.running; s.min_energy = c.min_energy;
    s.space_len = c.space_len;
    s.moves = c.moves; s.evaluated = c.evaluated; s.accepted = c.accepted_all; s.pair_evals = c.pair_evals;
    s.min_tooth = c.min_tooth; s.snapshots = c.n_snaps; s.points = c.n_points; s.done = c.done;
    s.prop_retries = c.prop_retries; s.pad_ = 0;
    out[i] = s;
}

// DFMA micro-benchmark: 8 independent dependent-FMA chains per thread.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 0.999999, c = 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace tr
