// C ABI of the B200 annealing backend (include/transrot_b200.h): context, device buffers,
// kernel dispatch.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/transrot_b200.h"
#include "context.cuh"
#include "anneal.cuh"
#include "setup_kernels.cuh"

using namespace tr;

namespace tr {

thread_local std::string g_create_error;

int fail(tr_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

}  // namespace tr

namespace {

// (lanes per chain, warps per chain, sites per thread) instantiated below; capacity = tpc * SPL sites.
// Order = preference of the automatic choice: half-warp teams while the cluster fits 64 sites, then a warp,
// then a block per chain.
const Variant kVariants[] = { {16, 1, 1}, {16, 1, 2}, {16, 1, 4}, {32, 1, 2}, {32, 1, 4}, {32, 1, 6},
                              {32, 4, 6}, {32, 8, 3}, {32, 8, 6} };

template <typename T>
int upload(tr_ctx *ctx, DevBuf<T> &buf, const T *src, size_t n)
{
    CK(ctx, buf.alloc(n));
    if (n) CK(ctx, cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return TR_OK;
}

int bind(tr_ctx *ctx)
{
    CK(ctx, cudaSetDevice(ctx->device));
    return TR_OK;
}

bool pick_variant(const tr_ctx *ctx, int n_sites, int tpc, Variant *out)
{
    for (const Variant &v : kVariants) {
        if (tpc > 0 && v.tpc() != tpc) continue;
        if (v.tpc() * v.SPL >= n_sites) { *out = v; return true; }
    }
    (void)ctx;
    return false;
}

bool same_variant(const Variant &a, const Variant &b) { return a.L == b.L && a.W == b.W && a.SPL == b.SPL; }

// Site -> (thread, slot) permutation of a kernel variant into `tabs`, and the DevSystem that carries it.
// layout 0 (anneal.cuh, crew.cuh): sites with Lennard-Jones parameters first, so that the trailing slots are Coulomb-only
// for every thread.  layout 1 (crew2.cuh): particle-major - lane n % L owns particle n in its slot n / L, SS site slots each:
// position of site a of particle n = ((n / L) * SS + a) * L + n % L.  layout 2 (anneal.cuh with the pair cache): thread
// n % T owns particle n as its first (n < T, site slots 0 .. N0-1) or second particle (site slots N0 ..): v = {T, N0, N1}.
int prepare_tables(tr_ctx *ctx, PosTables &tabs, int layout, const Variant &v, int n_pos, DevSystem *out)
{
    const int S = ctx->sys.n_sites, T = ctx->sys.n_types;
    if (!(tabs.layout == layout && same_variant(tabs.prepared, v))) {
        std::vector<int> pos((size_t)n_pos, -1), inv((size_t)S, 0), pinfo((size_t)n_pos, INFO_UNUSED | (T << 16));
        int n_lj = 0;
        if (layout == 0) {
            int p = 0;
            for (int i = 0; i < S; i++) if (ctx->h_site_lj[(size_t)i]) pos[(size_t)p++] = i;
            n_lj = p;
            for (int i = 0; i < S; i++) if (!ctx->h_site_lj[(size_t)i]) pos[(size_t)p++] = i;
        } else if (layout == 1) {
            const int L = v.L, SS = v.SPL;
            for (int n = 0; n < ctx->sys.n_mol; n++)
                for (int i = ctx->h_mol_off[(size_t)n]; i < ctx->h_mol_off[(size_t)n + 1]; i++)
                    pos[(size_t)(((n / L) * SS + (i - ctx->h_mol_off[(size_t)n])) * L + n % L)] = i;
        } else {
            const int T2 = v.L, N0 = v.W;              // (threads, site slots of the first particle, of the second)
            for (int n = 0; n < ctx->sys.n_mol; n++)
                for (int i = ctx->h_mol_off[(size_t)n]; i < ctx->h_mol_off[(size_t)n + 1]; i++)
                    pos[(size_t)(((n >= T2 ? N0 : 0) + (i - ctx->h_mol_off[(size_t)n])) * T2 + n % T2)] = i;
        }
        for (int q = 0; q < n_pos; q++)
            if (pos[(size_t)q] >= 0) { inv[(size_t)pos[(size_t)q]] = q; pinfo[(size_t)q] = ctx->h_site_info[(size_t)pos[(size_t)q]]; }
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        if (int r = upload(ctx, tabs.pos_site, pos.data(), pos.size())) return r;
        if (int r = upload(ctx, tabs.site_pos, inv.data(), inv.size())) return r;
        if (int r = upload(ctx, tabs.pos_info, pinfo.data(), pinfo.size())) return r;
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        tabs.prepared = v; tabs.layout = layout; tabs.n_pos = n_pos; tabs.n_lj_pos = n_lj;
    }
    *out = ctx->sys;
    out->pos_site = tabs.pos_site.p;
    out->site_pos = tabs.site_pos.p;
    out->pos_info = tabs.pos_info.p;
    out->n_pos = tabs.n_pos;
    out->n_lj_pos = tabs.n_lj_pos;
    return TR_OK;
}

int prepare_variant(tr_ctx *ctx, PosTables &tabs, const Variant &v, DevSystem *out)
{
    return prepare_tables(ctx, tabs, 0, v, v.tpc() * v.SPL, out);
}

int max_mol_sites_of(const tr_ctx *ctx)
{
    int m = 1;
    for (size_t i = 0; i + 1 < ctx->h_mol_off.size(); i++) m = std::max(m, ctx->h_mol_off[i + 1] - ctx->h_mol_off[i]);
    return m;
}

// ---- kernel dispatch -----------------------------------------------------------------------

// Blocks per SM the register budget is squeezed for.  Half-warp teams: 7 x 64 threads = 28 chains per SM lets
// 4096 chains be resident in one wave on 148 SMs with up to 146 registers per thread.
template <int L, int W, int SPL> struct MinBlocks { static constexpr int value = 1; };
template <int SPL> struct MinBlocks<16, 1, SPL> { static constexpr int value = 7; };
#ifndef TR_MINB_L32S2
#define TR_MINB_L32S2 4
#endif
template <> struct MinBlocks<32, 1, 2> { static constexpr int value = TR_MINB_L32S2; };
template <> struct MinBlocks<32, 1, 4> { static constexpr int value = 4; };
template <> struct MinBlocks<32, 1, 6> { static constexpr int value = 3; };
template <> struct MinBlocks<32, 4, 6> { static constexpr int value = 3; };
template <> struct MinBlocks<32, 8, 3> { static constexpr int value = 2; };

template <int L, int W, int SPL, int LJS, bool HAS_EXP>
cudaError_t launch_anneal_lj(tr_ctx *ctx, const KParams &P)
{
    constexpr int CPB = Team<L, W>::CPB;
    constexpr int MINB = MinBlocks<L, W, SPL>::value;
    const int smem = Layout<L, W, SPL>::total_bytes(P.sys.n_types, HAS_EXP);
    const unsigned blocks = (unsigned)((P.n_chains + CPB - 1) / CPB);
    cudaError_t e;
    if (P.trace) {
        auto kern = anneal_kernel<L, W, SPL, SPL, HAS_EXP, true, 1>;     // tracing: one generic variant
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
        kern<<<blocks, Team<L, W>::BLOCK, smem, ctx->stream>>>(P);
    } else {
        auto kern = anneal_kernel<L, W, SPL, LJS, HAS_EXP, false, MINB>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
        kern<<<blocks, Team<L, W>::BLOCK, smem, ctx->stream>>>(P);
    }
    return cudaGetLastError();
}

// The number of leading slots that can hold Lennard-Jones sites is a compile-time parameter of the kernel
// (straight-line inner loop); instantiated for 0, about a third, and all slots, rounded up.
template <int L, int W, int SPL, bool HAS_EXP>
cudaError_t launch_anneal_t(tr_ctx *ctx, const KParams &P)
{
    constexpr int TPC = Team<L, W>::TPC;
    constexpr int THIRD = (SPL + 2) / 3 < 1 ? 1 : (SPL + 2) / 3;
    const int lj_slots = (P.sys.n_lj_pos + TPC - 1) / TPC;
    if (HAS_EXP) return launch_anneal_lj<L, W, SPL, SPL, HAS_EXP>(ctx, P);
    if (lj_slots == 0) return launch_anneal_lj<L, W, SPL, 0, HAS_EXP>(ctx, P);
    if (lj_slots <= THIRD) return launch_anneal_lj<L, W, SPL, THIRD, HAS_EXP>(ctx, P);
    return launch_anneal_lj<L, W, SPL, SPL, HAS_EXP>(ctx, P);
}

template <int L, int W, int SPL, bool HAS_EXP>
cudaError_t smem_need_t(tr_ctx *ctx, int *bytes)
{
    *bytes = Layout<L, W, SPL>::total_bytes(ctx->sys.n_types, HAS_EXP);
    return cudaSuccess;
}

template <int L, int W, int SPL, bool HAS_EXP>
cudaError_t launch_total_t(tr_ctx *ctx, const DevSystem &dsys, const double *d_sites, long long n, double *d_out)
{
    auto kern = total_energy_kernel<L, W, SPL, HAS_EXP>;
    const int smem = Layout<L, W, SPL>::total_bytes(ctx->sys.n_types, HAS_EXP);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)(n < 4096 ? n : 4096);
    kern<<<blocks, Team<L, W>::TPC, smem, ctx->stream>>>(dsys, d_sites, n, d_out);
    return cudaGetLastError();
}

template <int L, int W, int SPL, bool HAS_EXP>
cudaError_t launch_delta_t(tr_ctx *ctx, const DevSystem &dsys, const double *d_sites, const int *d_mol, const double *d_trial, long long n,
                           double *d_es, double *d_ee, double *d_de)
{
    auto kern = delta_energy_kernel<L, W, SPL, HAS_EXP>;
    const int smem = Layout<L, W, SPL>::total_bytes(ctx->sys.n_types, HAS_EXP);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)(n < 4096 ? n : 4096);
    kern<<<blocks, Team<L, W>::TPC, smem, ctx->stream>>>(dsys, d_sites, d_mol, d_trial, n, d_es, d_ee, d_de);
    return cudaGetLastError();
}


#define TR_DISPATCH_CASES(CALL, HE)                                   \
    switch (key__) {                                                  \
    case 160101: e = CALL(16, 1, 1, HE); break;                       \
    case 160102: e = CALL(16, 1, 2, HE); break;                       \
    case 160104: e = CALL(16, 1, 4, HE); break;                       \
    case 320102: e = CALL(32, 1, 2, HE); break;                       \
    case 320104: e = CALL(32, 1, 4, HE); break;                       \
    case 320106: e = CALL(32, 1, 6, HE); break;                       \
    case 320406: e = CALL(32, 4, 6, HE); break;                       \
    case 320803: e = CALL(32, 8, 3, HE); break;                       \
    case 320806: e = CALL(32, 8, 6, HE); break;                       \
    default: e = cudaErrorInvalidValue;                               \
    }

#define TR_DISPATCH(v, has_exp, CALL)                                          \
    do {                                                                       \
        const int key__ = (v).L * 10000 + (v).W * 100 + (v).SPL;               \
        if (has_exp) { TR_DISPATCH_CASES(CALL, true) } else { TR_DISPATCH_CASES(CALL, false) } \
    } while (0)

int need_chains(tr_ctx *ctx)
{
    NEED(ctx, ctx->n_chains > 0, "Error: no chains initialised (call tr_init_chains_* first)");
    return TR_OK;
}

int alloc_chain_buffers(tr_ctx *ctx, long long n_chains, long long first_chain)
{
    NEED(ctx, ctx->have_system, "Error: tr_set_system must be called before initialising chains");
    NEED(ctx, !ctx->h_sched.empty(), "Error: tr_set_schedules must be called before initialising chains");
    NEED(ctx, n_chains > 0, "Error: n_chains must be positive");
    ctx->n_chains = 0;            // the batch only becomes live when its initialisation has completed (commit_chains)
    Variant v;
    if (!pick_variant(ctx, ctx->sys.n_sites, ctx->threads_per_chain, &v))
        return fail(ctx, TR_ELIMIT, "Error: %d sites do not fit %d threads per chain (limit %d sites)", ctx->sys.n_sites,
                    ctx->threads_per_chain, TR_MAX_SITES);
    ctx->variant = v;
    // the crew kernel wins while a chain fits a half-warp team (<= 64 sites: 1.3 vs 1.1 G moves/s on (H2O)20); for
    // 65..192 sites the warp-per-chain team kernel is ahead (0.55 vs 0.45 G moves/s on (NH4Cl)32); tr_set_kernel overrides
    ctx->use_crew = ctx->threads_per_chain == 0 && v.W == 1 && ctx->kernel_choice != TR_KERNEL_TEAM &&
                    (v.L == 16 || ctx->kernel_choice == TR_KERNEL_CREW);
    // the particle-major kernel with the pair-energy cache takes over wherever it has a variant (crew2.cuh)
    NEED(ctx, ctx->pair_table_symmetric || (ctx->kernel_choice != TR_KERNEL_CREW2 && ctx->kernel_choice != TR_KERNEL_TEAM_PC),
         "Error: the pair table is not symmetric; the kernels with a pair-energy cache (TR_KERNEL_CREW2, TR_KERNEL_TEAM_PC) need a symmetric one");
    ctx->use_crew2 = ctx->pair_table_symmetric && ctx->threads_per_chain == 0 && (ctx->kernel_choice == TR_KERNEL_CREW2 || (ctx->kernel_choice == TR_KERNEL_AUTO && TR_CREW2_AUTO_BIG) ||
                                                      (ctx->kernel_choice == TR_KERNEL_AUTO && ctx->sys.n_mol <= TR_CREW2_SMALL_MOL)) &&
                     crew2_pick(ctx, max_mol_sites_of(ctx), &ctx->crew2);
    NEED(ctx, ctx->use_crew2 || ctx->kernel_choice != TR_KERNEL_CREW2, "Error: no crew2 kernel variant for this system (%d particles, %d sites in the largest)",
         ctx->sys.n_mol, max_mol_sites_of(ctx));
    // above that, the team kernel on particle-major positions with the same cache in global memory (anneal.cuh, PC0 > 0): the
    // automatic choice where it measured faster than the plain team kernel - six site slots per thread, e.g. (NH4Cl)32 as 5 + 1
    // (594 vs 554 M moves/s) and (H2O)256 as 3 + 3 (144 vs 135); with eight (the mixed 64-particle cluster as 5 + 3: 409 vs 550)
    // the wider evaluation costs more registers than the halved arithmetic returns.  tr_set_kernel overrides.
    ctx->use_team_pc = false;
    if (!ctx->use_crew2 && ctx->pair_table_symmetric && ctx->threads_per_chain == 0 &&
        (ctx->kernel_choice == TR_KERNEL_TEAM_PC || ctx->kernel_choice == TR_KERNEL_AUTO) &&
        ctx->sys.n_mol > TR_CREW2_SMALL_MOL && team_pc_pick(ctx, &ctx->team_pc))
        ctx->use_team_pc = ctx->kernel_choice == TR_KERNEL_TEAM_PC || ctx->team_pc.N0 + ctx->team_pc.N1 <= TR_TEAM_PC_AUTO_SLOTS;
    NEED(ctx, ctx->use_team_pc || ctx->kernel_choice != TR_KERNEL_TEAM_PC, "Error: no pair-cache team kernel variant for this system (%d particles, %d sites in the largest)",
         ctx->sys.n_mol, max_mol_sites_of(ctx));
    if (ctx->use_crew2) {
        ctx->use_crew = false;
        const Crew2Variant &c2 = ctx->crew2;
        if (int r = prepare_tables(ctx, ctx->run_tabs, 1, Variant{c2.L, c2.M, c2.SS}, c2.L * c2.M * c2.SS, &ctx->run_sys)) return r;
    } else if (ctx->use_team_pc) {
        ctx->use_crew = false;
        const TeamPcVariant &pc = ctx->team_pc;        // (positions: thread n % threads, its first or second particle)
        if (int r = prepare_tables(ctx, ctx->run_tabs, 2, Variant{pc.threads(), pc.N0, pc.N1}, pc.threads() * (pc.N0 + pc.N1), &ctx->run_sys)) return r;
    } else if (int r = prepare_variant(ctx, ctx->run_tabs, v, &ctx->run_sys)) return r;
    {
        char nm[96];
        if (ctx->use_crew2) snprintf(nm, sizeof nm, "crew2_kernel<L=%d,M=%d,SS=%d>", ctx->crew2.L, ctx->crew2.M, ctx->crew2.SS);
        else if (ctx->use_team_pc) snprintf(nm, sizeof nm, "anneal_kernel<L=%d,W=%d,SPL=%d+%d,pair cache>", ctx->team_pc.L, ctx->team_pc.W, ctx->team_pc.N0, ctx->team_pc.N1);
        else if (ctx->use_crew) snprintf(nm, sizeof nm, "crew_kernel<L=%d,SPL=%d>", v.L, v.SPL);
        else snprintf(nm, sizeof nm, "anneal_kernel<L=%d,W=%d,SPL=%d>", v.L, v.W, v.SPL);
        ctx->kernel_name = nm;
    }
    int smem = 0, tri = 0;
    if (ctx->use_crew2) {
        if (int r = crew2_smem_bytes(ctx, n_chains, &smem, &tri)) return r;
    } else if (ctx->use_team_pc) {
        if (int r = team_pc_smem_bytes(ctx, &smem, &tri)) return r;
    } else if (ctx->use_crew) {
        if (int r = crew_smem_bytes(ctx, n_chains, &smem)) return r;
    } else {
        cudaError_t e;
#define CALL_SMEM(L, W, SPL, HE) smem_need_t<L, W, SPL, HE>(ctx, &smem)
        TR_DISPATCH(v, ctx->sys.has_exp, CALL_SMEM);
#undef CALL_SMEM
        if (e != cudaSuccess) return fail(ctx, TR_ELIMIT, "Error: no kernel variant for this system");
    }
    if (smem > ctx->max_smem_optin)
        return fail(ctx, TR_ELIMIT, "Error: system needs %d bytes of shared memory per block, device offers %d", smem,
                    ctx->max_smem_optin);
    // records: points and snapshots are sized for the largest schedule
    int max_points = 0, max_snaps = 0;
    for (const tr_schedule &s : ctx->h_sched) {
        long long pts = 0, p = s.static_temp ? 1 : s.points_per_tooth;
        const int teeth = s.static_temp ? 1 : s.num_teeth;
        for (int x = 0; x < teeth; x++) { pts += p; p += s.points_added_per_tooth; }
        if (s.finale) pts += p;
        if (pts > max_points) max_points = (int)pts;
        if (teeth + 1 > max_snaps) max_snaps = teeth + 1;
    }
    ctx->max_points = max_points;
    ctx->max_snaps = max_snaps;
    const size_t S = ctx->sys.n_sites, N = ctx->sys.n_mol, C = (size_t)n_chains;
    CK(ctx, ctx->d_ctrl.alloc(C));
    CK(ctx, ctx->d_sites.alloc(C * S * 3));
    CK(ctx, ctx->d_com.alloc(C * N * 3));
    CK(ctx, ctx->d_mt.alloc(C * 3 * 2 * MT_N + 64));   // + slack for the look-ahead prefetch of the last stream
    CK(ctx, ctx->d_snap_energy.alloc(C * max_snaps));
    CK(ctx, ctx->d_snap_tooth.alloc(C * max_snaps));
    CK(ctx, cudaMemsetAsync(ctx->d_snap_tooth.p, 0xff, C * max_snaps * sizeof(int), ctx->stream));
    if (ctx->record_snapshots) CK(ctx, ctx->d_snap_sites.alloc(C * max_snaps * S * 3)); else ctx->d_snap_sites.release();
    if (ctx->record_movie && ctx->record_points) {
        CK(ctx, ctx->d_movie_sites.alloc(C * (size_t)max_points * S * 3));
        CK(ctx, cudaMemsetAsync(ctx->d_movie_sites.p, 0, C * (size_t)max_points * S * 3 * sizeof(double), ctx->stream));
    } else ctx->d_movie_sites.release();
    if (ctx->record_points) {
        CK(ctx, ctx->d_points.alloc(C * max_points));
        CK(ctx, cudaMemsetAsync(ctx->d_points.p, 0, C * max_points * sizeof(tr_point_rec), ctx->stream));
    } else ctx->d_points.release();
    if (ctx->use_crew2 || ctx->use_team_pc) CK(ctx, ctx->d_pair_cache.alloc(C * (size_t)tri)); else ctx->d_pair_cache.release();
    if (ctx->trace_moves > 0) {
        CK(ctx, ctx->d_trace.alloc(C * (size_t)ctx->trace_moves));
        CK(ctx, cudaMemsetAsync(ctx->d_trace.p, 0, C * (size_t)ctx->trace_moves * sizeof(tr_trace_rec), ctx->stream));
    } else ctx->d_trace.release();
    ctx->pending_chains = n_chains;
    ctx->first_chain = first_chain;
    ctx->live_trace_moves = ctx->trace_moves;
    return TR_OK;
}

// The initialisation kernels have run and every staging copy has completed: the batch is live.
int commit_chains(tr_ctx *ctx)
{
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_chains = ctx->pending_chains;
    return TR_OK;
}

int launch_init(tr_ctx *ctx, const long long *d_seeds, const int *d_sched_idx, const int *d_mt_pos, double space_len,
                int max_prop_failures, int propagate)
{
    InitParams ip{};
    ip.sys = ctx->sys;
    ip.sched = ctx->d_sched.p;
    ip.ctrl = ctx->d_ctrl.p;
    ip.sites = ctx->d_sites.p;
    ip.com = ctx->d_com.p;
    ip.mt = ctx->d_mt.p;
    ip.seeds = d_seeds;
    ip.sched_idx = d_sched_idx;
    ip.mt_pos = d_mt_pos;
    ip.n_chains = ctx->pending_chains;
    ip.space_len = space_len;
    ip.max_prop_failures = max_prop_failures;
    ip.propagate = propagate;
    ip.n_sched = (int)ctx->h_sched.size();
    const unsigned blocks = (unsigned)((ctx->pending_chains + INIT_WARPS - 1) / INIT_WARPS);       // one warp per chain
    init_chains_kernel<<<blocks, 32 * INIT_WARPS, 0, ctx->stream>>>(ip);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return TR_OK;
}

}  // namespace

// =================================================================================== lifetime

extern "C" int tr_create(int device, tr_ctx **out)
{
    if (!out) return fail(nullptr, TR_EINVAL, "Error: tr_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, TR_ENODEVICE, "Error: no CUDA device available (%s); this backend has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(nullptr, TR_EINVAL, "Error: device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, TR_ECUDA, "Error: cannot open CUDA device %d: %s", device, cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, TR_ENODEVICE, "Error: device %d is sm_%d%d; this library holds sm_100a code only", device, prop.major,
                    prop.minor);
    tr_ctx *ctx = new (std::nothrow) tr_ctx();
    if (!ctx) return fail(nullptr, TR_ENOMEM, "Error: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, TR_ECUDA, "Error: cannot create stream/events: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return TR_OK;
}

extern "C" void tr_destroy(tr_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char *tr_last_error(const tr_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int tr_set_stream(tr_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return TR_OK;
}

extern "C" int tr_synchronize(tr_ctx *ctx)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

// ========================================================================= problem definition

extern "C" int tr_set_system(tr_ctx *ctx, const tr_system *s)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, s != nullptr, "Error: tr_set_system: sys is NULL");
    NEED(ctx, s->n_mol > 0 && s->n_sites > 0 && s->n_types > 0, "Error: empty system");
    NEED(ctx, s->mol_site_off && s->site_type && s->site_q && s->moveable_idx && s->pair_abcd && s->site_mass && s->site_ghost,
         "Error: tr_set_system: a required array is NULL");
    NEED(ctx, s->n_moveable > 0 && s->n_moveable <= s->n_mol, "Error: there must be between 1 and n_mol moveable particles");
    if (s->n_sites > TR_MAX_SITES) return fail(ctx, TR_ELIMIT, "Error: %d sites exceed the limit of %d", s->n_sites, TR_MAX_SITES);
    if (s->n_types > TR_MAX_TYPES) return fail(ctx, TR_ELIMIT, "Error: %d atom types exceed the limit of %d", s->n_types, TR_MAX_TYPES);
    if (s->n_mol > 65535) return fail(ctx, TR_ELIMIT, "Error: too many particles");
    NEED(ctx, s->mol_site_off[0] == 0 && s->mol_site_off[s->n_mol] == s->n_sites, "Error: mol_site_off must run from 0 to n_sites");
    std::vector<int> info((size_t)s->n_sites);
    for (int m = 0; m < s->n_mol; m++) {
        const int a = s->mol_site_off[m], b = s->mol_site_off[m + 1];
        NEED(ctx, b > a, "Error: particle %d has no sites", m);
        if (b - a > TR_MAX_MOL_SITES)
            return fail(ctx, TR_ELIMIT, "Error: particle %d has %d sites, limit is %d", m, b - a, TR_MAX_MOL_SITES);
        for (int i = a; i < b; i++) {
            NEED(ctx, s->site_type[i] >= 0 && s->site_type[i] < s->n_types, "Error: site %d has type %d outside 0..%d", i,
                 s->site_type[i], s->n_types - 1);
            info[(size_t)i] = m | (s->site_type[i] << 16) | ((i - a) << 24);
        }
    }
    for (int k = 0; k < s->n_moveable; k++)
        NEED(ctx, s->moveable_idx[k] >= 0 && s->moveable_idx[k] < s->n_mol, "Error: moveable_idx[%d] out of range", k);
    const size_t TT = (size_t)s->n_types * s->n_types;
    std::vector<double2> cd(TT), ab(TT);
    int has_exp = 0;
    for (size_t i = 0; i < TT; i++) {
        ab[i] = make_double2(s->pair_abcd[4 * i], s->pair_abcd[4 * i + 1]);
        cd[i] = make_double2(s->pair_abcd[4 * i + 2], s->pair_abcd[4 * i + 3]);
        if (s->pair_abcd[4 * i] != 0.0) has_exp = 1;   // NaN (an undefined pair) also keeps the full formula
    }
    // a type "has Lennard-Jones" when C or D is non-zero (or NaN) against any type; with a Born-Mayer term
    // anywhere every site takes the full formula
    std::vector<int> type_lj((size_t)s->n_types, 0);
    for (int t1 = 0; t1 < s->n_types; t1++)
        for (int t2 = 0; t2 < s->n_types; t2++) {
            // row OR column: the ABI accepts tables that are not symmetric, and a site's class decides both how it is
            // evaluated when staged (row) and which slot it occupies as "the other site" (column)
            const double2 v = cd[(size_t)t1 * s->n_types + t2], w = cd[(size_t)t2 * s->n_types + t1];
            if (!(v.x == 0.0) || !(v.y == 0.0) || !(w.x == 0.0) || !(w.y == 0.0) || has_exp) type_lj[(size_t)t1] = 1;
        }
    // the kernels with a pair-energy cache keep ONE energy per particle pair whichever of the two moved last: that is the
    // reference's value only when the table is symmetric (it always is when setupPairVals built it; NaN entries compare by bits)
    ctx->pair_table_symmetric = true;
    for (int t1 = 0; t1 < s->n_types; t1++)
        for (int t2 = 0; t2 < t1; t2++)
            if (memcmp(&s->pair_abcd[4 * ((size_t)t1 * s->n_types + t2)], &s->pair_abcd[4 * ((size_t)t2 * s->n_types + t1)], 4 * sizeof(double)) != 0)
                ctx->pair_table_symmetric = false;
    ctx->h_site_lj.assign((size_t)s->n_sites, 0);
    for (int i = 0; i < s->n_sites; i++) {
        ctx->h_site_lj[(size_t)i] = type_lj[(size_t)s->site_type[i]];
        info[(size_t)i] |= type_lj[(size_t)s->site_type[i]] << 29;
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_system = false;
    ctx->n_chains = 0;
    ctx->run_tabs.layout = -1; ctx->probe_tabs.layout = -1;
    ctx->h_mol_off.assign(s->mol_site_off, s->mol_site_off + s->n_mol + 1);
    ctx->h_site_info = info;
    if (int r = upload(ctx, ctx->d_mol_off, s->mol_site_off, (size_t)s->n_mol + 1)) return r;
    if (int r = upload(ctx, ctx->d_site_info, info.data(), info.size())) return r;
    if (int r = upload(ctx, ctx->d_site_q, s->site_q, (size_t)s->n_sites)) return r;
    if (int r = upload(ctx, ctx->d_moveable, s->moveable_idx, (size_t)s->n_moveable)) return r;
    if (int r = upload(ctx, ctx->d_site_mass, s->site_mass, (size_t)s->n_sites)) return r;
    if (int r = upload(ctx, ctx->d_site_ghost, s->site_ghost, (size_t)s->n_sites)) return r;
    if (int r = upload(ctx, ctx->d_pair_cd, cd.data(), TT)) return r;
    if (int r = upload(ctx, ctx->d_pair_ab, ab.data(), TT)) return r;
    ctx->have_propagate_data = s->mol_radius && s->site_def;
    if (ctx->have_propagate_data) {
        if (int r = upload(ctx, ctx->d_mol_radius, s->mol_radius, (size_t)s->n_mol)) return r;
        if (int r = upload(ctx, ctx->d_site_def, s->site_def, (size_t)s->n_sites * 3)) return r;
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));   // host staging vectors die here
    DevSystem &d = ctx->sys;
    d.n_mol = s->n_mol; d.n_sites = s->n_sites; d.n_types = s->n_types; d.n_moveable = s->n_moveable;
    d.has_exp = has_exp; d.pad_ = 0;
    d.n_pos = 0; d.n_lj_pos = 0; d.pos_site = nullptr; d.site_pos = nullptr; d.pos_info = nullptr;
    d.row_stride = (s->n_types + 1) * 16; d.zero_col = s->n_types * 16;
    d.pick_mod = fastmod_make((uint32_t)s->n_moveable);
    d.mol_off = ctx->d_mol_off.p; d.site_info = ctx->d_site_info.p; d.site_q = ctx->d_site_q.p;
    d.pair_cd = ctx->d_pair_cd.p; d.pair_ab = ctx->d_pair_ab.p; d.moveable = ctx->d_moveable.p;
    d.mol_radius = ctx->d_mol_radius.p; d.site_def = ctx->d_site_def.p; d.site_mass = ctx->d_site_mass.p;
    d.site_ghost = ctx->d_site_ghost.p;
    ctx->have_system = true;
    return TR_OK;
}

extern "C" int tr_set_schedules(tr_ctx *ctx, const tr_schedule *sched, int32_t n_sched)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, sched && n_sched > 0, "Error: tr_set_schedules: no schedules given");
    for (int i = 0; i < n_sched; i++) {
        NEED(ctx, sched[i].moves_per_point > 0, "Error: schedule %d: Moves per Point must be positive", i);
        NEED(ctx, sched[i].static_temp || (sched[i].points_per_tooth > 0 && sched[i].num_teeth > 0),
             "Error: schedule %d: Points per Tooth and Number of Teeth must be positive", i);
        NEED(ctx, sched[i].points_added_per_tooth >= 0, "Error: schedule %d: Points Added per Tooth must not be negative", i);
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h_sched.assign(sched, sched + n_sched);
    ctx->n_chains = 0;
    if (int r = upload(ctx, ctx->d_sched, ctx->h_sched.data(), (size_t)n_sched)) return r;
    return TR_OK;
}

extern "C" int tr_set_options(tr_ctx *ctx, int64_t trace_moves, int32_t record_points, int32_t record_snapshots,
                              int32_t threads_per_chain)
{
    if (!ctx) return TR_EINVAL;
    NEED(ctx, trace_moves >= 0, "Error: trace_moves must not be negative");
    NEED(ctx, threads_per_chain == 0 || threads_per_chain == 16 || threads_per_chain == 32 || threads_per_chain == 128 ||
                  threads_per_chain == 256,
         "Error: threads_per_chain must be 0 (auto), 16, 32, 128 or 256");
    ctx->trace_moves = trace_moves;
    ctx->record_points = record_points ? 1 : 0;
    ctx->record_snapshots = record_snapshots ? 1 : 0;
    ctx->threads_per_chain = threads_per_chain;
    return TR_OK;
}

extern "C" int tr_set_record_movie(tr_ctx *ctx, int32_t on)
{
    if (!ctx) return TR_EINVAL;
    ctx->record_movie = on ? 1 : 0;
    return TR_OK;
}

extern "C" int tr_init_chains_seeded(tr_ctx *ctx, int64_t n_chains, int64_t first_chain, const int64_t *seeds,
                                     const int32_t *sched_idx, double space_len, int32_t max_prop_failures)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, seeds != nullptr, "Error: tr_init_chains_seeded: seeds is NULL");
    NEED(ctx, ctx->have_system && ctx->have_propagate_data,
         "Error: a seeded start needs mol_radius and site_def in tr_set_system (Space.propagate places particles)");
    NEED(ctx, space_len > 0.0, "Error: Length of Cubic Space must be positive");
    if (int r = alloc_chain_buffers(ctx, n_chains, first_chain)) return r;
    DevBuf<long long> &d_seeds = ctx->d_seeds;
    DevBuf<int> &d_idx = ctx->d_sched_idx;
    CK(ctx, d_seeds.alloc((size_t)n_chains));
    CK(ctx, cudaMemcpyAsync(d_seeds.p, seeds, (size_t)n_chains * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (sched_idx) {
        CK(ctx, d_idx.alloc((size_t)n_chains));
        CK(ctx, cudaMemcpyAsync(d_idx.p, sched_idx, (size_t)n_chains * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (int r = launch_init(ctx, d_seeds.p, sched_idx ? d_idx.p : nullptr, nullptr, space_len, max_prop_failures, 1)) return r;
    return commit_chains(ctx);
}

extern "C" int tr_init_chains_explicit(tr_ctx *ctx, int64_t n_chains, int64_t first_chain, const double *sites, int64_t n_structs,
                                       const int64_t *seeds, const uint32_t *mt_state, const int32_t *sched_idx, double space_len)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, sites != nullptr, "Error: tr_init_chains_explicit: sites is NULL");
    NEED(ctx, n_structs == 1 || n_structs == n_chains, "Error: n_structs must be 1 or n_chains");
    NEED(ctx, seeds != nullptr || mt_state != nullptr, "Error: give seeds or explicit RNG states");
    NEED(ctx, space_len > 0.0, "Error: Length of Cubic Space must be positive");
    NEED(ctx, n_chains > 0, "Error: n_chains must be positive");
    if (mt_state)
        for (size_t i = 0; i < (size_t)n_chains * 3; i++)
            NEED(ctx, mt_state[i * MT_WORDS + MT_N] <= (uint32_t)MT_N, "Error: RNG state %zu has mti %u > 624", i,
                 mt_state[i * MT_WORDS + MT_N]);
    if (int r = alloc_chain_buffers(ctx, n_chains, first_chain)) return r;
    const size_t S3 = (size_t)ctx->sys.n_sites * 3;
    if (n_structs == n_chains) {
        CK(ctx, cudaMemcpyAsync(ctx->d_sites.p, sites, (size_t)n_chains * S3 * 8, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        std::vector<double> rep((size_t)n_chains * S3);
        for (int64_t c = 0; c < n_chains; c++) memcpy(rep.data() + (size_t)c * S3, sites, S3 * 8);
        CK(ctx, cudaMemcpyAsync(ctx->d_sites.p, rep.data(), rep.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    {
        const long long total = n_chains * ctx->sys.n_mol;
        com_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(ctx->sys, ctx->d_sites.p, ctx->d_com.p, n_chains);
        ctx->launches++;
        CK(ctx, cudaGetLastError());
    }
    DevBuf<long long> &d_seeds = ctx->d_seeds;
    DevBuf<int> &d_idx = ctx->d_sched_idx, &d_pos = ctx->d_mt_pos;
    std::vector<uint32_t> words;
    std::vector<int> pos;
    if (mt_state) {
        words.assign((size_t)n_chains * 3 * 2 * MT_N, 0u);
        pos.resize((size_t)n_chains * 3);
        for (size_t i = 0; i < (size_t)n_chains * 3; i++) {
            memcpy(words.data() + i * 2 * MT_N, mt_state + i * MT_WORDS, MT_N * 4);   // buffer 0 of the ping-pong pair
            pos[i] = (int)mt_state[i * MT_WORDS + MT_N];
        }
        CK(ctx, cudaMemcpyAsync(ctx->d_mt.p, words.data(), words.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, d_pos.alloc(pos.size()));
        CK(ctx, cudaMemcpyAsync(d_pos.p, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        CK(ctx, d_seeds.alloc((size_t)n_chains));
        CK(ctx, cudaMemcpyAsync(d_seeds.p, seeds, (size_t)n_chains * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (sched_idx) {
        CK(ctx, d_idx.alloc((size_t)n_chains));
        CK(ctx, cudaMemcpyAsync(d_idx.p, sched_idx, (size_t)n_chains * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (int r = launch_init(ctx, mt_state ? nullptr : d_seeds.p, sched_idx ? d_idx.p : nullptr, mt_state ? d_pos.p : nullptr, space_len, 0, 0))
        return r;
    return commit_chains(ctx);     // (also keeps the host staging vectors alive until the copies have landed)
}

extern "C" int tr_set_kernel(tr_ctx *ctx, int32_t kernel, int32_t chains_per_block, int32_t blocks_per_sm)
{
    if (!ctx) return TR_EINVAL;
    NEED(ctx, kernel == TR_KERNEL_AUTO || kernel == TR_KERNEL_TEAM || kernel == TR_KERNEL_CREW || kernel == TR_KERNEL_CREW2 || kernel == TR_KERNEL_TEAM_PC,
         "Error: kernel must be TR_KERNEL_AUTO, TR_KERNEL_TEAM, TR_KERNEL_CREW, TR_KERNEL_CREW2 or TR_KERNEL_TEAM_PC");
    NEED(ctx, chains_per_block >= 0 && chains_per_block <= 32 && blocks_per_sm >= 0 && blocks_per_sm <= 8,
         "Error: chains_per_block must be 0..32 and blocks_per_sm 0..8");
    ctx->kernel_choice = kernel;
    ctx->crew_chains = chains_per_block;
    ctx->crew_bps = blocks_per_sm;
    return TR_OK;
}

extern "C" const char *tr_kernel_name(const tr_ctx *ctx) { return ctx ? ctx->kernel_name.c_str() : ""; }

extern "C" int tr_abi_sizes(int32_t *out, int32_t cap)
{
    // sizeof and the offsets a binding cannot derive from natural alignment alone: checked at load time by
    // transrot_b200/engine.py and java/TransRotB200.java instead of by comment
    const int32_t v[] = {
        (int32_t)sizeof(tr_system), (int32_t)offsetof(tr_system, mol_site_off), (int32_t)offsetof(tr_system, pair_abcd),
        (int32_t)sizeof(tr_schedule), (int32_t)offsetof(tr_schedule, moves_per_point), (int32_t)offsetof(tr_schedule, write_conf_heat_capacities),
        (int32_t)sizeof(tr_point_rec), (int32_t)offsetof(tr_point_rec, temp), (int32_t)offsetof(tr_point_rec, movie_energy),
        (int32_t)sizeof(tr_trace_rec), (int32_t)offsetof(tr_trace_rec, e_start), (int32_t)offsetof(tr_trace_rec, running),
        (int32_t)sizeof(tr_chain_summary), (int32_t)offsetof(tr_chain_summary, moves), (int32_t)offsetof(tr_chain_summary, min_tooth),
        (int32_t)sizeof(tr_minimum), (int32_t)offsetof(tr_minimum, chain), (int32_t)offsetof(tr_minimum, tooth),
    };
    const int32_t n = (int32_t)(sizeof v / sizeof v[0]);
    if (!out || cap < n) return n;
    for (int32_t i = 0; i < n; i++) out[i] = v[i];
    return n;
}

// ===================================================================================== hot path

extern "C" int tr_run(tr_ctx *ctx, int64_t max_moves)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    const Variant v = ctx->variant;
    KParams P{};
    P.sys = ctx->run_sys;
    P.pair_cache = ctx->d_pair_cache.p;
    P.sched = ctx->d_sched.p;
    P.ctrl = ctx->d_ctrl.p;
    P.sites = ctx->d_sites.p;
    P.com = ctx->d_com.p;
    P.mt = ctx->d_mt.p;
    P.points = ctx->d_points.p;
    P.snap_energy = ctx->d_snap_energy.p;
    P.snap_tooth = ctx->d_snap_tooth.p;
    P.snap_sites = ctx->d_snap_sites.p;
    P.movie_sites = ctx->d_movie_sites.p;
    P.trace = ctx->d_trace.p;
    P.n_chains = ctx->n_chains;
    P.max_moves = max_moves;
    P.trace_moves = ctx->live_trace_moves;
    P.max_points = ctx->max_points;
    P.max_snaps = ctx->max_snaps;
    CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->use_crew2) {
        if (int r = launch_crew2(ctx, P)) return r;
    } else if (ctx->use_team_pc) {
        if (int r = launch_team_pc(ctx, P)) return r;
    } else if (ctx->use_crew) {
        if (int r = launch_crew(ctx, P)) return r;
    } else {
        cudaError_t e;
#define CALL_ANNEAL(L, W, SPL, HE) launch_anneal_t<L, W, SPL, HE>(ctx, P)
        TR_DISPATCH(v, ctx->sys.has_exp, CALL_ANNEAL);
#undef CALL_ANNEAL
        if (e != cudaSuccess) return fail(ctx, TR_ECUDA, "Error: anneal kernel launch failed: %s", cudaGetErrorString(e));
    }
    ctx->launches++;
    CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->have_run_time = true;
    return TR_OK;
}

extern "C" int tr_last_run_ms(tr_ctx *ctx, float *ms)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, ms && ctx->have_run_time, "Error: tr_last_run_ms: no tr_run has been issued");
    CK(ctx, cudaEventSynchronize(ctx->ev1));
    CK(ctx, cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return TR_OK;
}

extern "C" int tr_launch_count(tr_ctx *ctx, int64_t *n)
{
    if (!ctx || !n) return TR_EINVAL;
    *n = ctx->launches;
    return TR_OK;
}

// ====================================================================================== results

extern "C" int tr_get_summary(tr_ctx *ctx, tr_chain_summary *out)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, out != nullptr, "Error: tr_get_summary: out is NULL");
    DevBuf<tr_chain_summary> &d = ctx->d_summary;
    CK(ctx, d.alloc((size_t)ctx->n_chains));
    summary_kernel<<<(unsigned)((ctx->n_chains + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_ctrl.p, ctx->n_chains, d.p);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(out, d.p, (size_t)ctx->n_chains * sizeof(tr_chain_summary), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    // a chain the kernel had to abandon is an error of the run, not a result (the records are still returned)
    for (long long i = 0; i < ctx->n_chains; i++)
        if (out[i].done == 2)
            return fail(ctx, TR_ERNG, "Error: chain %lld was abandoned after %lld moves: its random stream ring could not be refilled",
                        (long long)(ctx->first_chain + i), (long long)out[i].moves);
    return TR_OK;
}

extern "C" int tr_get_sites(tr_ctx *ctx, double *sites, double *com)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    if (sites)
        CK(ctx, cudaMemcpyAsync(sites, ctx->d_sites.p, (size_t)ctx->n_chains * ctx->sys.n_sites * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (com)
        CK(ctx, cudaMemcpyAsync(com, ctx->d_com.p, (size_t)ctx->n_chains * ctx->sys.n_mol * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_max_points(tr_ctx *ctx, int64_t *points_per_chain, int64_t *snapshots_per_chain)
{
    if (!ctx) return TR_EINVAL;
    if (int r = need_chains(ctx)) return r;
    if (points_per_chain) *points_per_chain = ctx->max_points;
    if (snapshots_per_chain) *snapshots_per_chain = ctx->max_snaps;
    return TR_OK;
}

extern "C" int tr_get_points(tr_ctx *ctx, tr_point_rec *out, int64_t cap_per_chain)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, ctx->d_points.p != nullptr, "Error: point records were disabled with tr_set_options");
    NEED(ctx, out && cap_per_chain >= ctx->max_points, "Error: tr_get_points: buffer holds %lld records per chain, %d needed",
         (long long)cap_per_chain, ctx->max_points);
    CK(ctx, cudaMemcpy2DAsync(out, (size_t)cap_per_chain * sizeof(tr_point_rec), ctx->d_points.p,
                              (size_t)ctx->max_points * sizeof(tr_point_rec), (size_t)ctx->max_points * sizeof(tr_point_rec),
                              (size_t)ctx->n_chains, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_get_snapshots(tr_ctx *ctx, double *energy, int32_t *tooth, double *sites, int64_t cap_per_chain)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, cap_per_chain >= ctx->max_snaps, "Error: tr_get_snapshots: buffer holds %lld snapshots per chain, %d needed",
         (long long)cap_per_chain, ctx->max_snaps);
    const size_t C = (size_t)ctx->n_chains, K = (size_t)ctx->max_snaps, cap = (size_t)cap_per_chain;
    if (energy)
        CK(ctx, cudaMemcpy2DAsync(energy, cap * 8, ctx->d_snap_energy.p, K * 8, K * 8, C, cudaMemcpyDeviceToHost, ctx->stream));
    if (tooth)
        CK(ctx, cudaMemcpy2DAsync(tooth, cap * 4, ctx->d_snap_tooth.p, K * 4, K * 4, C, cudaMemcpyDeviceToHost, ctx->stream));
    if (sites) {
        NEED(ctx, ctx->d_snap_sites.p != nullptr, "Error: snapshot structures were disabled with tr_set_options");
        const size_t row = (size_t)ctx->sys.n_sites * 24;
        CK(ctx, cudaMemcpy2DAsync(sites, cap * row, ctx->d_snap_sites.p, K * row, K * row, C, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_get_movie(tr_ctx *ctx, double *sites, int64_t cap_per_chain)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, sites != nullptr, "Error: tr_get_movie: sites is NULL");
    NEED(ctx, ctx->d_movie_sites.p != nullptr, "Error: movie frames were not recorded (tr_set_record_movie before tr_init_chains_*)");
    NEED(ctx, cap_per_chain >= ctx->max_points, "Error: tr_get_movie: buffer holds %lld frames per chain, %d needed",
         (long long)cap_per_chain, ctx->max_points);
    const size_t C = (size_t)ctx->n_chains, K = (size_t)ctx->max_points, cap = (size_t)cap_per_chain;
    const size_t row = (size_t)ctx->sys.n_sites * 3 * sizeof(double);
    CK(ctx, cudaMemcpy2DAsync(sites, cap * row, ctx->d_movie_sites.p, K * row, K * row, C, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_get_trace(tr_ctx *ctx, tr_trace_rec *out)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, ctx->d_trace.p && out, "Error: tracing was not enabled with tr_set_options");
    CK(ctx, cudaMemcpyAsync(out, ctx->d_trace.p, (size_t)ctx->n_chains * (size_t)ctx->live_trace_moves * sizeof(tr_trace_rec),
                            cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_get_rng_state(tr_ctx *ctx, uint32_t *out)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, out != nullptr, "Error: tr_get_rng_state: out is NULL");
    DevBuf<uint32_t> &d_out = ctx->d_rng_out;
    CK(ctx, d_out.alloc((size_t)ctx->n_chains * 3 * MT_WORDS));
    const long long threads = ctx->n_chains * 32;
    rng_export_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_ctrl.p, ctx->d_mt.p, ctx->n_chains, d_out.p);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)ctx->n_chains * 3 * MT_WORDS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_pack_minima(tr_ctx *ctx, tr_minimum *host_out, void *device_out)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, host_out || device_out, "Error: tr_pack_minima: no output buffer");
    tr_minimum *dst = (tr_minimum *)device_out;
    if (!dst) { CK(ctx, ctx->d_minima.alloc((size_t)ctx->n_chains)); dst = ctx->d_minima.p; }
    pack_minima_kernel<<<(unsigned)((ctx->n_chains + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_ctrl.p, ctx->n_chains, ctx->first_chain, dst);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    if (host_out)
        CK(ctx, cudaMemcpyAsync(host_out, dst, (size_t)ctx->n_chains * sizeof(tr_minimum), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_get_min_structure(tr_ctx *ctx, int64_t chain, double *sites, double *energy, int32_t *tooth)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, chain >= 0 && chain < ctx->n_chains, "Error: chain %lld out of range", (long long)chain);
    const size_t K = (size_t)ctx->max_snaps;
    std::vector<double> e(K);
    std::vector<int> t(K);
    CK(ctx, cudaMemcpyAsync(e.data(), ctx->d_snap_energy.p + (size_t)chain * K, K * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(t.data(), ctx->d_snap_tooth.p + (size_t)chain * K, K * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    int best = -1;
    for (size_t k = 0; k < K; k++)
        if (t[k] >= 0 && (best < 0 || e[k] < e[(size_t)best])) best = (int)k;   // strict <: first minimum wins (Space.java:614)
    NEED(ctx, best >= 0, "Error: chain %lld has no snapshot yet", (long long)chain);
    if (energy) *energy = e[(size_t)best];
    if (tooth) *tooth = t[(size_t)best];
    if (sites) {
        NEED(ctx, ctx->d_snap_sites.p != nullptr, "Error: snapshot structures were disabled with tr_set_options");
        const size_t row = (size_t)ctx->sys.n_sites * 3;
        CK(ctx, cudaMemcpyAsync(sites, ctx->d_snap_sites.p + ((size_t)chain * K + (size_t)best) * row, row * 8, cudaMemcpyDeviceToHost,
                                ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return TR_OK;
}

extern "C" int tr_pack_min_structures(tr_ctx *ctx, int64_t first, int64_t count, double *host_out, void *device_out)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    if (int r = need_chains(ctx)) return r;
    NEED(ctx, host_out || device_out, "Error: tr_pack_min_structures: no output buffer");
    NEED(ctx, first >= 0 && count > 0 && first + count <= ctx->n_chains, "Error: tr_pack_min_structures: chains %lld..%lld out of range",
         (long long)first, (long long)(first + count - 1));
    NEED(ctx, ctx->d_snap_sites.p != nullptr, "Error: snapshot structures were disabled with tr_set_options");
    const size_t row = (size_t)ctx->sys.n_sites * 3;
    double *dst = (double *)device_out;
    if (!dst) { CK(ctx, ctx->d_min_sites.alloc((size_t)count * row)); dst = ctx->d_min_sites.p; }
    const long long total = count * (long long)row;
    const size_t K = (size_t)ctx->max_snaps;
    pack_min_structures_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 65535LL * 16), 256, 0, ctx->stream>>>(
        ctx->d_snap_energy.p + (size_t)first * K, ctx->d_snap_tooth.p + (size_t)first * K, ctx->d_snap_sites.p + (size_t)first * K * row, count,
        ctx->max_snaps, (int)row, dst);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    if (host_out) CK(ctx, cudaMemcpyAsync(host_out, dst, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

// ======================================================================================= probes

extern "C" int tr_total_energy(tr_ctx *ctx, const double *sites, int64_t n, double *energy)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, ctx->have_system, "Error: tr_set_system must be called first");
    NEED(ctx, sites && energy && n > 0, "Error: tr_total_energy: bad arguments");
    Variant v;
    if (!pick_variant(ctx, ctx->sys.n_sites, ctx->threads_per_chain, &v)) return fail(ctx, TR_ELIMIT, "Error: system too large");
    DevSystem dsys;
    if (int r = prepare_variant(ctx, ctx->probe_tabs, v, &dsys)) return r;
    DevBuf<double> d_sites, d_out;
    const size_t S3 = (size_t)ctx->sys.n_sites * 3;
    CK(ctx, d_sites.alloc((size_t)n * S3));
    CK(ctx, d_out.alloc((size_t)n));
    CK(ctx, cudaMemcpyAsync(d_sites.p, sites, (size_t)n * S3 * 8, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e;
#define CALL_TOTAL(L, W, SPL, HE) launch_total_t<L, W, SPL, HE>(ctx, dsys, d_sites.p, n, d_out.p)
    TR_DISPATCH(v, ctx->sys.has_exp, CALL_TOTAL);
#undef CALL_TOTAL
    if (e != cudaSuccess) return fail(ctx, TR_ECUDA, "Error: total-energy kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    CK(ctx, cudaMemcpyAsync(energy, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_delta_energy(tr_ctx *ctx, const double *sites, const int32_t *mol, const double *trial, int64_t n, double *e_start,
                               double *e_end, double *d_e)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, ctx->have_system, "Error: tr_set_system must be called first");
    NEED(ctx, sites && mol && trial && e_start && e_end && n > 0, "Error: tr_delta_energy: bad arguments");
    for (int64_t i = 0; i < n; i++) NEED(ctx, mol[i] >= 0 && mol[i] < ctx->sys.n_mol, "Error: mol[%lld] out of range", (long long)i);
    Variant v;
    if (!pick_variant(ctx, ctx->sys.n_sites, ctx->threads_per_chain, &v)) return fail(ctx, TR_ELIMIT, "Error: system too large");
    DevSystem dsys;
    if (int r = prepare_variant(ctx, ctx->probe_tabs, v, &dsys)) return r;
    DevBuf<double> d_sites, d_trial, d_es, d_ee, d_de;
    DevBuf<int> d_mol;
    const size_t S3 = (size_t)ctx->sys.n_sites * 3, T3 = (size_t)TR_MAX_MOL_SITES * 3;
    CK(ctx, d_sites.alloc((size_t)n * S3));
    CK(ctx, d_trial.alloc((size_t)n * T3));
    CK(ctx, d_mol.alloc((size_t)n));
    CK(ctx, d_es.alloc((size_t)n));
    CK(ctx, d_ee.alloc((size_t)n));
    CK(ctx, d_de.alloc((size_t)n));
    CK(ctx, cudaMemcpyAsync(d_sites.p, sites, (size_t)n * S3 * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d_trial.p, trial, (size_t)n * T3 * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d_mol.p, mol, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e;
#define CALL_DELTA(L, W, SPL, HE) launch_delta_t<L, W, SPL, HE>(ctx, dsys, d_sites.p, d_mol.p, d_trial.p, n, d_es.p, d_ee.p, d_de.p)
    TR_DISPATCH(v, ctx->sys.has_exp, CALL_DELTA);
#undef CALL_DELTA
    if (e != cudaSuccess) return fail(ctx, TR_ECUDA, "Error: delta-energy kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    CK(ctx, cudaMemcpyAsync(e_start, d_es.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(e_end, d_ee.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (d_e) CK(ctx, cudaMemcpyAsync(d_e, d_de.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

extern "C" int tr_rng_words(tr_ctx *ctx, const int64_t *seeds, int64_t n, int32_t k, uint32_t *out)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, seeds && out && n > 0 && k > 0, "Error: tr_rng_words: bad arguments");
    DevBuf<long long> d_seeds;
    DevBuf<uint32_t> d_mt, d_out;
    CK(ctx, d_seeds.alloc((size_t)n));
    CK(ctx, d_mt.alloc((size_t)n * 3 * 2 * MT_N + 64));
    CK(ctx, d_out.alloc((size_t)n * 3 * (size_t)k));
    CK(ctx, cudaMemcpyAsync(d_seeds.p, seeds, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->threads_per_chain == 16) rng_words_kernel<16><<<(unsigned)((n + 3) / 4), 64, 0, ctx->stream>>>(d_seeds.p, n, k, d_mt.p, d_out.p);
    else rng_words_kernel<32><<<(unsigned)((n + 3) / 4), 128, 0, ctx->stream>>>(d_seeds.p, n, k, d_mt.p, d_out.p);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)n * 3 * (size_t)k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return TR_OK;
}

// =================================================================================== measurement

extern "C" int tr_measure_fp64_peak(tr_ctx *ctx, double *tflops, double *sm_clock_mhz_est)
{
    if (!ctx) return TR_EINVAL;
    if (int r = bind(ctx)) return r;
    NEED(ctx, tflops != nullptr, "Error: tr_measure_fp64_peak: tflops is NULL");
    const int threads = 256, blocks = ctx->sm_count * 8, iters = 4096;
    DevBuf<double> d_out;
    CK(ctx, d_out.alloc((size_t)threads * blocks));
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; rep++) {
        CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out.p, iters, 1.0 + rep);
        ctx->launches++;
        CK(ctx, cudaGetLastError());
        CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CK(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    const double fmas = (double)threads * blocks * (double)iters * 64.0;
    *tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
    if (sm_clock_mhz_est) *sm_clock_mhz_est = fmas / (best_ms * 1e-3) / ((double)ctx->sm_count * 64.0) / 1e6;
    ctx->have_run_time = false;
    return TR_OK;
}
