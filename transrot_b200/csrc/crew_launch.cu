// Launch geometry and dispatch of the warp-specialised kernel family (crew.cuh).  Its own translation unit: the
// kernels are the ones the headline workload runs and the ones that change most often.
#include <cuda_runtime.h>

#include <algorithm>

#include "context.cuh"
#include "crew.cuh"

using namespace tr;

namespace {


#ifndef TR_CREW_EW
#define TR_CREW_EW 7          // energy warps per block
#endif
#ifndef TR_CREW_MINB
#define TR_CREW_MINB 2        // blocks per SM the register budget is squeezed for
#endif
#ifndef TR_CREW_BPS
#define TR_CREW_BPS 2         // co-resident blocks per SM the chains-per-block choice aims at
#endif

template <int L, int SPL>
void crew_sizes_t(const tr_ctx *ctx, int max_mol_sites, int *tables, int *stride)
{
    *tables = CrewTables<L, SPL>::table_bytes(ctx->sys.n_types, ctx->sys.has_exp != 0);
    *stride = CrewTables<L, SPL>::chain_bytes(ctx->sys.n_mol, max_mol_sites);
}

int max_mol_sites_of(const tr_ctx *ctx)
{
    int m = 1;
    for (size_t i = 0; i + 1 < ctx->h_mol_off.size(); i++) m = std::max(m, ctx->h_mol_off[i + 1] - ctx->h_mol_off[i]);
    return m;
}

#define TR_CREW_CASES(CALL)                          \
    switch (ctx->variant.L * 100 + ctx->variant.SPL) { \
    case 1601: CALL(16, 1); break;                   \
    case 1602: CALL(16, 2); break;                   \
    case 1604: CALL(16, 4); break;                   \
    case 3204: CALL(32, 4); break;                   \
    case 3206: CALL(32, 6); break;                   \
    default: return fail(ctx, TR_ELIMIT, "Error: no crew kernel for %d lanes x %d sites per lane", ctx->variant.L, ctx->variant.SPL); \
    }

// Chains per block C, bytes per chain region and dynamic shared memory of the crew kernel.  C is chosen so
// that the grid fills every SM with TR_CREW_BPS co-resident blocks in ONE wave when the chains fit; with more
// chains than that the grid simply runs in several waves.
int crew_geometry(tr_ctx *ctx, long long n_chains, int *C, int *stride, int *smem)
{
    int tables = 0;
    const int maxs = max_mol_sites_of(ctx);
#define CALL_SIZES(L, SPL) crew_sizes_t<L, SPL>(ctx, maxs, &tables, stride)
    TR_CREW_CASES(CALL_SIZES)
#undef CALL_SIZES
    const int bps = std::max(1, ctx->crew_bps > 0 ? ctx->crew_bps : TR_CREW_BPS);
    const int per_block = std::min(ctx->max_smem_optin, (ctx->smem_per_sm - 1024 * bps) / bps);
    int c_fit = (per_block - tables) / *stride;
    if (c_fit < 1) c_fit = (ctx->max_smem_optin - tables) / *stride;     // fewer co-resident blocks, but it runs
    if (c_fit < 1)
        return fail(ctx, TR_ELIMIT, "Error: one chain needs %d bytes of shared memory, device offers %d", tables + *stride,
                    ctx->max_smem_optin);
    const long long slots = (long long)ctx->sm_count * bps;
    int c = (int)std::min<long long>((n_chains + slots - 1) / slots, 32);
    c = std::max(1, std::min(c, c_fit));
    const int forced = ctx->crew_chains;
    if (forced > 0) c = std::max(1, std::min(std::min(forced, 32), (ctx->max_smem_optin - tables) / *stride));
    c = std::min(c, TR_CREW_EW * (32 / ctx->variant.L));          // every chain of a block has its own energy team
    *C = c;
    *smem = tables + c * *stride;
    return TR_OK;
}

template <int L, int SPL, int LJS, bool HAS_EXP>
cudaError_t launch_crew_lj(tr_ctx *ctx, const KParams &P, int C, int stride, int smem, int maxs)
{
    constexpr int EW = TR_CREW_EW;
    const unsigned blocks = (unsigned)((P.n_chains + C - 1) / C);
    cudaError_t e;
    if (P.trace) {
        auto kern = crew_kernel<L, SPL, SPL, HAS_EXP, true, EW, 1>;        // tracing: one generic variant
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
        kern<<<blocks, 32 * (1 + EW), smem, ctx->stream>>>(P, C, stride, maxs);
    } else {
        auto kern = crew_kernel<L, SPL, LJS, HAS_EXP, false, EW, TR_CREW_MINB>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
        kern<<<blocks, 32 * (1 + EW), smem, ctx->stream>>>(P, C, stride, maxs);
    }
    return cudaGetLastError();
}

template <int L, int SPL>
cudaError_t launch_crew_t(tr_ctx *ctx, const KParams &P, int C, int stride, int smem, int maxs)
{
    constexpr int THIRD = (SPL + 2) / 3 < 1 ? 1 : (SPL + 2) / 3;
    const int lj_slots = (P.sys.n_lj_pos + L - 1) / L;
    if (P.sys.has_exp) return launch_crew_lj<L, SPL, SPL, true>(ctx, P, C, stride, smem, maxs);
    if (lj_slots == 0) return launch_crew_lj<L, SPL, 0, false>(ctx, P, C, stride, smem, maxs);
    if (lj_slots <= THIRD) return launch_crew_lj<L, SPL, THIRD, false>(ctx, P, C, stride, smem, maxs);
    return launch_crew_lj<L, SPL, SPL, false>(ctx, P, C, stride, smem, maxs);
}

}  // namespace

namespace tr {

int crew_smem_bytes(tr_ctx *ctx, long long n_chains, int *smem)
{
    int C = 0, stride = 0;
    return crew_geometry(ctx, n_chains, &C, &stride, smem);
}

int launch_crew(tr_ctx *ctx, const KParams &P)
{
    int C = 0, stride = 0, smem = 0;
    if (int r = crew_geometry(ctx, P.n_chains, &C, &stride, &smem)) return r;
    const int maxs = max_mol_sites_of(ctx);
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_CREW(L, SPL) e = launch_crew_t<L, SPL>(ctx, P, C, stride, smem, maxs)
    TR_CREW_CASES(CALL_CREW)
#undef CALL_CREW
    if (e != cudaSuccess) return fail(ctx, TR_ECUDA, "Error: crew kernel launch failed: %s", cudaGetErrorString(e));
    return TR_OK;
}

}  // namespace tr
