// Shared by the translation units of the library: the context behind the opaque tr_ctx, device buffers, error helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/transrot_b200.h"
#include "device_types.cuh"

namespace tr {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count)
    {
        if (p && n == count) return cudaSuccess;      // same batch shape as last time: keep the allocation
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
};

struct Variant { int L, W, SPL; int tpc() const { return W > 1 ? 32 * W : L; } };

// Site <-> position permutation of one kernel variant (which lane and slot holds which site), on the device.
struct PosTables {
    DevBuf<int> pos_site, site_pos, pos_info;
    Variant prepared{0, 0, 0};
    int layout = -1;            // 0: Lennard-Jones sites first (anneal.cuh, crew.cuh); 1: particle-major (crew2.cuh)
    int n_pos = 0, n_lj_pos = 0;
};

// crew2 kernel variant: L lanes per team, M particle slots per lane, SS site slots per particle
struct Crew2Variant { int L, M, SS; };
// pair-cache team kernel variant: W warps of L lanes per chain, two particles per thread with N0 and N1 site slots
struct TeamPcVariant {
    int L, W, N0, N1;
    int threads() const { return W > 1 ? 32 * W : L; }
};

}  // namespace tr

struct tr_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    long long launches = 0;
    bool have_run_time = false;
    int sm_count = 0;
    int max_smem_optin = 0;
    int smem_per_sm = 0;

    // system
    bool have_system = false;
    bool pair_table_symmetric = true;   // {A,B,C,D}[t1][t2] == [t2][t1] bit for bit (the reference always builds it so: Space.java:346-366)
    tr::DevSystem sys{};
    tr::DevBuf<int> d_mol_off, d_site_info, d_moveable, d_site_ghost;
    tr::DevBuf<double> d_site_q, d_mol_radius, d_site_def, d_site_mass;
    tr::DevBuf<double2> d_pair_cd, d_pair_ab;
    tr::PosTables run_tabs, probe_tabs;   // of the live chain batch / of the parity probes
    tr::DevSystem run_sys{};              // sys + the live batch's position tables
    std::vector<int> h_site_info;
    bool have_propagate_data = false;
    std::vector<int> h_mol_off, h_site_lj;

    // schedules
    tr::DevBuf<tr_schedule> d_sched;
    std::vector<tr_schedule> h_sched;

    // options
    long long trace_moves = 0;
    int record_points = 1, record_snapshots = 1, threads_per_chain = 0, record_movie = 0;

    // chains
    long long n_chains = 0, first_chain = 0, pending_chains = 0;
    tr::DevBuf<tr::ChainCtrl> d_ctrl;
    tr::DevBuf<double> d_sites, d_com, d_snap_energy, d_snap_sites, d_movie_sites;
    tr::DevBuf<uint32_t> d_mt;
    tr::DevBuf<int> d_snap_tooth;
    tr::DevBuf<tr_point_rec> d_points;
    tr::DevBuf<tr_trace_rec> d_trace;
    int max_points = 0, max_snaps = 0;
    tr::Variant variant{16, 1, 1};
    bool use_crew = false;        // warp-specialised kernel (crew.cuh): automatic choice while a chain fits one warp
    bool use_crew2 = false;       // crew2.cuh: particle-major teams + pair-energy cache, see crew2_pick
    tr::Crew2Variant crew2{0, 0, 0};
    bool use_team_pc = false;     // anneal.cuh with PCM > 0: the team kernel on particle-major positions with the pair-energy cache
    tr::TeamPcVariant team_pc{0, 0, 0, 0};
    tr::DevBuf<double> d_pair_cache;     // [n_chains][tri_count] particle-pair energies (crew2), part of the chain state
    // tr_set_kernel: 0 = automatic, TR_KERNEL_TEAM / TR_KERNEL_CREW force one family; chains per block / co-resident
    // blocks per SM of the crew kernel (0 = automatic)
    int kernel_choice = 0, crew_chains = 0, crew_bps = 0;
    std::string kernel_name;      // of the live chain batch (tr_kernel_name)
    // what the live chain batch was allocated with (tr_set_options may change the pending values at any time)
    long long live_trace_moves = 0;
    // getter scratch, kept between calls
    tr::DevBuf<tr_chain_summary> d_summary;
    tr::DevBuf<tr_minimum> d_minima;
    tr::DevBuf<double> d_min_sites;
    tr::DevBuf<uint32_t> d_rng_out;
    tr::DevBuf<long long> d_seeds;
    tr::DevBuf<int> d_sched_idx, d_mt_pos;
};

namespace tr {

int fail(tr_ctx *ctx, int code, const char *fmt, ...);

#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return tr::fail(ctx, TR_ECUDA, "Error: CUDA failure in %s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

#define NEED(ctx, cond, ...)                                \
    do {                                                    \
        if (!(cond)) return tr::fail(ctx, TR_EINVAL, __VA_ARGS__); \
    } while (0)

// crew_launch.cu: the warp-specialised kernel family
int crew_smem_bytes(tr_ctx *ctx, long long n_chains, int *smem);
int launch_crew(tr_ctx *ctx, const KParams &P);
// team_pc_launch.cu: the team kernel with the pair-energy cache (TR_KERNEL_TEAM_PC)
constexpr int TR_TEAM_PC_AUTO_SLOTS = 6;   // ... the automatic choice above TR_CREW2_SMALL_MOL particles up to this many site slots per thread
bool team_pc_pick(const tr_ctx *ctx, TeamPcVariant *out);
int team_pc_smem_bytes(tr_ctx *ctx, int *smem, int *pair_count);
int launch_team_pc(tr_ctx *ctx, const KParams &P);
// crew2_launch.cu
constexpr bool TR_CREW2_AUTO_BIG = false;   // whole-warp crew2 variants as the automatic choice above 32 particles
constexpr int TR_CREW2_SMALL_MOL = 32;     // up to here: 8-lane teams, pair cache in shared memory; above: whole-warp teams, cache in global memory
bool crew2_pick(const tr_ctx *ctx, int max_mol_sites, Crew2Variant *out);
int crew2_smem_bytes(tr_ctx *ctx, long long n_chains, int *smem, int *tri_count);
int launch_crew2(tr_ctx *ctx, const KParams &P);

}  // namespace tr
