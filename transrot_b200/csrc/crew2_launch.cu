// Variant choice, launch geometry and dispatch of the crew2 kernel (crew2.cuh).
#include <cuda_runtime.h>

#include <algorithm>

#include "context.cuh"
#include "crew2.cuh"

using namespace tr;

namespace {

#ifndef TR_CREW2_EW
#define TR_CREW2_EW 7          // energy warps per block (+ 3 control warps), 8-lane teams: four chains each
#endif
#ifndef TR_CREW2_EW32
#define TR_CREW2_EW32 13       // ... whole-warp teams: one chain each (16 warps: 128 registers); fewer if shared memory holds fewer chains
#endif

// Energy warps of a variant's block (compile time: it sets the block size the kernel is compiled for).
template <int L, int M, int SS>
constexpr int energy_warps()
{
    if (L < 32) return TR_CREW2_EW * (L / 8);
    const int stride = C2Tables<L, M, SS>::ctrl_bytes() + C2Tables<L, M, SS>::team_bytes();
    const int fit = (227 * 1024 - 8 * 1024) / stride;          // (8 KB: the chain-independent tables)
    return fit < TR_CREW2_EW32 ? fit : TR_CREW2_EW32;
}

// Instantiated (L lanes per team, M particle slots per lane, SS site slots per particle); capacity L * M particles of up to
// SS sites.  Order = preference: 8-lane teams (four chains per energy warp) while the particles fit.
const Crew2Variant kCrew2[] = {
#ifdef TR_CREW2_L16
    {16, 2, 3},
#endif
    {8, 1, 1}, {8, 1, 3}, {8, 1, 5}, {8, 2, 3}, {8, 3, 3}, {8, 2, 5}, {8, 4, 3}, {8, 3, 5}, {8, 4, 5},
    // whole-warp teams with the pair cache in global memory: 33 and more particles
    {32, 2, 5}, {32, 4, 5}, {32, 8, 3} };

#ifdef TR_CREW2_DEV      // development builds (tools/devbuild.sh): the two variants of the shipped sample and of (H2O)20 only
#define TR_CREW2_CASES(CALL)                                                       \
    switch (v.L * 10000 + v.M * 100 + v.SS) {                                      \
    case 80105: CALL(8, 1, 5); break;                                              \
    case 80303: CALL(8, 3, 3); break;                                              \
    case 160203: CALL(16, 2, 3); break;                                            \
    case 320205: CALL(32, 2, 5); break;                                            \
    case 320803: CALL(32, 8, 3); break;                                            \
    default: return fail(ctx, TR_ELIMIT, "Error: no crew2 kernel for %d lanes x %d particles x %d sites (development build)", v.L, v.M, v.SS); \
    }
#else
#define TR_CREW2_CASES(CALL)                                                       \
    switch (v.L * 10000 + v.M * 100 + v.SS) {                                      \
    case 80101: CALL(8, 1, 1); break;                                              \
    case 80103: CALL(8, 1, 3); break;                                              \
    case 80105: CALL(8, 1, 5); break;                                              \
    case 80203: CALL(8, 2, 3); break;                                              \
    case 80303: CALL(8, 3, 3); break;                                              \
    case 80205: CALL(8, 2, 5); break;                                              \
    case 80403: CALL(8, 4, 3); break;                                              \
    case 80305: CALL(8, 3, 5); break;                                              \
    case 80405: CALL(8, 4, 5); break;                                              \
    case 320205: CALL(32, 2, 5); break;                                            \
    case 320405: CALL(32, 4, 5); break;                                            \
    case 320803: CALL(32, 8, 3); break;                                            \
    default: return fail(ctx, TR_ELIMIT, "Error: no crew2 kernel for %d lanes x %d particles x %d sites", v.L, v.M, v.SS); \
    }
#endif

template <int L, int M, int SS>
void sizes_t(const tr_ctx *ctx, int *tables, int *stride, int *tri, int *chains)
{
    *chains = energy_warps<L, M, SS>() * (32 / L);     // every chain of a block has its own energy team and control lane
    *tables = C2Tables<L, M, SS>::table_bytes(ctx->sys.n_types, ctx->sys.has_exp != 0);
    *stride = C2Tables<L, M, SS>::ctrl_bytes() + C2Tables<L, M, SS>::team_bytes();
    *tri = C2Tables<L, M, SS>::tri_count(ctx->sys.n_mol);
}

int geometry(tr_ctx *ctx, long long n_chains, int *C, int *stride, int *smem, int *tri)
{
    const Crew2Variant v = ctx->crew2;
    int tables = 0, c = 0;
#define CALL_SIZES(L, M, SS) sizes_t<L, M, SS>(ctx, &tables, stride, tri, &c)
    TR_CREW2_CASES(CALL_SIZES)
#undef CALL_SIZES
    if (ctx->crew_chains > 0) c = std::min(c, ctx->crew_chains);
    const int c_fit = (ctx->max_smem_optin - tables) / *stride;
    if (c_fit < 1)
        return fail(ctx, TR_ELIMIT, "Error: one chain needs %d bytes of shared memory, device offers %d", tables + *stride, ctx->max_smem_optin);
    c = std::min(c, c_fit);
    // the same number of waves with evenly filled blocks: 888 chains on 148 SMs run as 148 blocks of 6, not 127 of 7
    if (n_chains > 0 && ctx->crew_chains <= 0) {
        const long long sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        const long long waves = (n_chains + c * sms - 1) / (c * sms);
        c = (int)((n_chains + waves * sms - 1) / (waves * sms));
    }
    *C = c;
    *smem = tables + c * *stride;
    return TR_OK;
}

template <int L, int M, int SS>
cudaError_t launch_t(tr_ctx *ctx, const KParams &P, int C, int smem)
{
    constexpr int EW = energy_warps<L, M, SS>();
    const unsigned blocks = (unsigned)((P.n_chains + C - 1) / C);
    cudaError_t e;
#define TR_C2_LAUNCH(HE, TRC)                                                                                              \
    {                                                                                                                      \
        auto kern = crew2_kernel<L, M, SS, HE, TRC, EW>;                                                                   \
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e; \
        kern<<<blocks, 32 * (C2_CTRL_WARPS + EW), smem, ctx->stream>>>(P, C);                                              \
    }
    if (P.sys.has_exp) { if (P.trace) TR_C2_LAUNCH(true, true) else TR_C2_LAUNCH(true, false) }
    else { if (P.trace) TR_C2_LAUNCH(false, true) else TR_C2_LAUNCH(false, false) }
#undef TR_C2_LAUNCH
    return cudaGetLastError();
}

}  // namespace

namespace tr {

bool crew2_pick(const tr_ctx *ctx, int max_mol_sites, Crew2Variant *out)
{
    for (const Crew2Variant &v : kCrew2) {
        if ((v.L < 32) != (ctx->sys.n_mol <= TR_CREW2_SMALL_MOL)) continue;      // 8-lane teams up to 32 particles, whole warps above
        if (v.L * v.M >= ctx->sys.n_mol && v.SS >= max_mol_sites) { *out = v; return true; }
    }
    return false;
}

int crew2_smem_bytes(tr_ctx *ctx, long long n_chains, int *smem, int *tri_count)
{
    int C = 0, stride = 0;
    return geometry(ctx, n_chains, &C, &stride, smem, tri_count);
}

int launch_crew2(tr_ctx *ctx, const KParams &P)
{
    int C = 0, stride = 0, smem = 0, tri = 0;
    if (int r = geometry(ctx, P.n_chains, &C, &stride, &smem, &tri)) return r;
    const Crew2Variant v = ctx->crew2;
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_LAUNCH(L, M, SS) e = launch_t<L, M, SS>(ctx, P, C, smem)
    TR_CREW2_CASES(CALL_LAUNCH)
#undef CALL_LAUNCH
    if (e != cudaSuccess) return fail(ctx, TR_ECUDA, "Error: crew2 kernel launch failed: %s", cudaGetErrorString(e));
    return TR_OK;
}

}  // namespace tr
