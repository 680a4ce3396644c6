// Variant choice and dispatch of the team kernel's pair-energy-cache form (anneal.cuh, PCM > 0): particle-major positions,
// eStart from a per-chain cache of particle-pair energies in global memory.  33 and more particles (the crew2 kernel keeps
// the same cache in shared memory for smaller clusters).
#include <cuda_runtime.h>

#include <algorithm>

#include "anneal.cuh"
#include "context.cuh"

using namespace tr;

namespace {

// (lanes, warps per chain, site slots of a thread's first particle, of its second); capacity = 2 x threads particles, particle
// n < threads with up to N0 sites, the others with up to N1.  Order = preference (fewest site slots first).
const TeamPcVariant kTeamPc[] = { {32, 1, 3, 3}, {32, 1, 5, 1}, {32, 1, 5, 3}, {32, 1, 5, 5}, {32, 4, 3, 3}, {32, 4, 5, 5} };

// co-resident blocks the register budget is squeezed for (4 chains per block when a chain is one warp)
template <int L, int W, int N0, int N1> struct PcMinBlocks { static constexpr int value = 1; };
#ifndef TR_PC_MINB_W1
#define TR_PC_MINB_W1 4
#endif
#ifndef TR_PC_MINB_W4
#define TR_PC_MINB_W4 3
#endif
template <int N0, int N1> struct PcMinBlocks<32, 1, N0, N1> { static constexpr int value = TR_PC_MINB_W1; };
template <int N0, int N1> struct PcMinBlocks<32, 4, N0, N1> { static constexpr int value = TR_PC_MINB_W4; };

#define TR_PC_CASES(CALL)                                                         \
    switch (v.L * 1000000 + v.W * 10000 + v.N0 * 100 + v.N1) {                    \
    case 32010303: CALL(32, 1, 3, 3); break;                                      \
    case 32010501: CALL(32, 1, 5, 1); break;                                      \
    case 32010503: CALL(32, 1, 5, 3); break;                                      \
    case 32010505: CALL(32, 1, 5, 5); break;                                      \
    case 32040303: CALL(32, 4, 3, 3); break;                                      \
    case 32040505: CALL(32, 4, 5, 5); break;                                      \
    default: return fail(ctx, TR_ELIMIT, "Error: no pair-cache team kernel for %d x %d threads, %d + %d site slots", v.L, v.W, v.N0, v.N1); \
    }

template <int L, int W, int N0, int N1>
cudaError_t launch_t(tr_ctx *ctx, const KParams &P)
{
    constexpr int SPL = N0 + N1, CPB = Team<L, W>::CPB, MINB = PcMinBlocks<L, W, N0, N1>::value;
    const int smem = Layout<L, W, SPL, N0>::total_bytes(P.sys.n_types, P.sys.has_exp != 0);
    const unsigned blocks = (unsigned)((P.n_chains + CPB - 1) / CPB);
    cudaError_t e;
#define TR_PC_LAUNCH(HE, TRC, MB)                                                                                          \
    {                                                                                                                      \
        auto kern = anneal_kernel<L, W, SPL, SPL, HE, TRC, MB, N0>;                                                         \
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e; \
        kern<<<blocks, Team<L, W>::BLOCK, smem, ctx->stream>>>(P);                                                         \
    }
    if (P.sys.has_exp) { if (P.trace) TR_PC_LAUNCH(true, true, 1) else TR_PC_LAUNCH(true, false, MINB) }
    else { if (P.trace) TR_PC_LAUNCH(false, true, 1) else TR_PC_LAUNCH(false, false, MINB) }
#undef TR_PC_LAUNCH
    return cudaGetLastError();
}

template <int L, int W, int N0, int N1>
void smem_t(const tr_ctx *ctx, int *smem) { *smem = Layout<L, W, N0 + N1, N0>::total_bytes(ctx->sys.n_types, ctx->sys.has_exp != 0); }

}  // namespace

namespace tr {

bool team_pc_pick(const tr_ctx *ctx, TeamPcVariant *out)
{
    for (const TeamPcVariant &v : kTeamPc) {
        if (v.threads() * 2 < ctx->sys.n_mol) continue;
        int most0 = 0, most1 = 0;                      // the fullest first / second particle of any thread
        for (int n = 0; n < ctx->sys.n_mol; n++) {
            const int s = ctx->h_mol_off[(size_t)n + 1] - ctx->h_mol_off[(size_t)n];
            if (n < v.threads()) most0 = std::max(most0, s); else most1 = std::max(most1, s);
        }
        if (most0 <= v.N0 && most1 <= v.N1) { *out = v; return true; }
    }
    return false;
}

int team_pc_smem_bytes(tr_ctx *ctx, int *smem, int *pair_count)
{
    const TeamPcVariant v = ctx->team_pc;
#define CALL_SMEM(L, W, N0, N1) smem_t<L, W, N0, N1>(ctx, smem)
    TR_PC_CASES(CALL_SMEM)
#undef CALL_SMEM
    *pair_count = pair_cache_count(ctx->sys.n_mol);
    return TR_OK;
}

int launch_team_pc(tr_ctx *ctx, const KParams &P)
{
    const TeamPcVariant v = ctx->team_pc;
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_LAUNCH(L, W, N0, N1) e = launch_t<L, W, N0, N1>(ctx, P)
    TR_PC_CASES(CALL_LAUNCH)
#undef CALL_LAUNCH
    if (e != cudaSuccess) return fail(ctx, TR_ECUDA, "Error: pair-cache team kernel launch failed: %s", cudaGetErrorString(e));
    return TR_OK;
}

}  // namespace tr
