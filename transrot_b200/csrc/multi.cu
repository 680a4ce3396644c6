// Multi-GPU entry points of the C ABI (include/transrot_b200.h, "multi-GPU" section): one host thread, one tr_ctx per
// device, chains sharded in contiguous global ranges, and the only exchange the path has - the final gather of per-chain
// minima and of the winning (or all) minimum structures - as NCCL collectives between the devices' HBM over
// NVLink / NVSwitch.  The reference is one JVM, one thread (Main.java:17); a Java host therefore
// drives all GPUs of a box from a single process, which is what ncclCommInitAll is for.
//
// Written against the library's own public ABI plus the CUDA runtime.  NCCL is bound at run time (dlopen of
// libnccl.so.2), so a process that already carries an NCCL - torch's, in bench.py - shares it, and a single-GPU host
// never needs the library at all.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/transrot_b200.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

bool load_nccl(NcclApi &api, std::string &why)
{
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { why = dlerror() ? dlerror() : "libnccl.so.2 not found"; return false; }
#define TR_SYM(field, name)                                             \
    *(void **)(&api.field) = dlsym(api.handle, name);                   \
    if (!api.field) { why = std::string("missing symbol ") + name; return false; }
    TR_SYM(CommInitAll, "ncclCommInitAll")
    TR_SYM(CommDestroy, "ncclCommDestroy")
    TR_SYM(GroupStart, "ncclGroupStart")
    TR_SYM(GroupEnd, "ncclGroupEnd")
    TR_SYM(Broadcast, "ncclBroadcast")
    TR_SYM(GetErrorString, "ncclGetErrorString")
    TR_SYM(GetVersion, "ncclGetVersion")
#undef TR_SYM
    return true;
}

thread_local std::string g_multi_create_error;

// (energy, global chain id) minimum over the gathered records; ties go to the lowest chain id (sharding.global_minimum).
__global__ void argmin_minima_kernel(const tr_minimum *all, long long n, tr_minimum *best)
{
    __shared__ double s_e[256];
    __shared__ long long s_i[256];
    double e = 1.7976931348623157e308;
    long long idx = -1;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = all[i].energy;
        if (idx < 0 || v < e) { e = v; idx = i; }      // i grows within a thread: the first minimum is kept
    }
    s_e[threadIdx.x] = e; s_i[threadIdx.x] = idx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double e2 = s_e[threadIdx.x + o];
            const long long i2 = s_i[threadIdx.x + o];
            const long long i1 = s_i[threadIdx.x];
            if (i2 >= 0 && (i1 < 0 || e2 < s_e[threadIdx.x] || (e2 == s_e[threadIdx.x] && i2 < i1))) { s_e[threadIdx.x] = e2; s_i[threadIdx.x] = i2; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *best = all[s_i[0] < 0 ? 0 : s_i[0]];
}

}  // namespace

struct tr_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<tr_ctx *> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<ncclComm_t> comm;
    NcclApi nccl;
    std::string err;
    // the live batch
    long long n_total = 0, first_chain = 0;
    std::vector<long long> lo, hi;         // global chain range of each device, relative to first_chain
    int n_sites = 0;
    // device scratch: gathered minima (every device holds all of them), the winner record and structure buffers
    std::vector<tr_minimum *> d_all;
    std::vector<tr_minimum *> d_best;
    std::vector<double *> d_struct;
    std::vector<size_t> cap_all, cap_struct;
};

namespace {

int mfail(tr_multi *m, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->err = buf; else g_multi_create_error = buf;
    return code;
}

#define MCK(m, call)                                                                                     \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return mfail(m, TR_ECUDA, "Error: CUDA failure in %s: %s", #call, cudaGetErrorString(e__));   \
    } while (0)

#define MNCCL(m, call)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess)                                                                          \
            return mfail(m, TR_ENCCL, "Error: NCCL failure in %s: %s", #call, (m)->nccl.GetErrorString(r__)); \
    } while (0)

// every per-device call goes through here so that the first failure carries the device's own message
#define MEACH(m, i, call)                                                                                \
    do {                                                                                                 \
        int r__ = (call);                                                                                \
        if (r__ != TR_OK) return mfail(m, r__, "device %d: %s", (m)->devices[(size_t)(i)], tr_last_error((m)->ctx[(size_t)(i)])); \
    } while (0)

int grow(tr_multi *m, int i, size_t n_all, size_t n_struct)
{
    MCK(m, cudaSetDevice(m->devices[(size_t)i]));
    if (n_all > m->cap_all[(size_t)i]) {
        if (m->d_all[(size_t)i]) cudaFree(m->d_all[(size_t)i]);
        m->d_all[(size_t)i] = nullptr; m->cap_all[(size_t)i] = 0;
        MCK(m, cudaMalloc((void **)&m->d_all[(size_t)i], n_all * sizeof(tr_minimum)));
        m->cap_all[(size_t)i] = n_all;
    }
    if (n_struct > m->cap_struct[(size_t)i]) {
        if (m->d_struct[(size_t)i]) cudaFree(m->d_struct[(size_t)i]);
        m->d_struct[(size_t)i] = nullptr; m->cap_struct[(size_t)i] = 0;
        MCK(m, cudaMalloc((void **)&m->d_struct[(size_t)i], n_struct * sizeof(double)));
        m->cap_struct[(size_t)i] = n_struct;
    }
    return TR_OK;
}

// all-gather with per-device counts: one broadcast per root inside a single NCCL group (ragged shards need no padding)
int allgather_v(tr_multi *m, std::vector<char *> &bufs, size_t unit_bytes)
{
    MNCCL(m, m->nccl.GroupStart());
    for (int root = 0; root < m->n; root++) {
        const size_t off = (size_t)m->lo[(size_t)root] * unit_bytes, bytes = (size_t)(m->hi[(size_t)root] - m->lo[(size_t)root]) * unit_bytes;
        if (!bytes) continue;
        for (int i = 0; i < m->n; i++)
            MNCCL(m, m->nccl.Broadcast(bufs[(size_t)i] + off, bufs[(size_t)i] + off, bytes, ncclChar, root, m->comm[(size_t)i], m->stream[(size_t)i]));
    }
    MNCCL(m, m->nccl.GroupEnd());
    return TR_OK;
}

}  // namespace

extern "C" void tr_multi_chain_range(int32_t rank, int32_t world, int64_t n_total, int64_t *lo, int64_t *hi)
{
    // contiguous block partition; the first n_total % world devices get one chain more (sharding.chain_range)
    const int64_t base = n_total / world, extra = n_total % world;
    const int64_t l = (int64_t)rank * base + (rank < extra ? rank : extra);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < extra ? 1 : 0);
}

extern "C" int tr_create_multi(const int32_t *devices, int32_t n, tr_multi **out)
{
    if (!out) return mfail(nullptr, TR_EINVAL, "Error: tr_create_multi: out is NULL");
    *out = nullptr;
    if (!devices || n < 1 || n > 64) return mfail(nullptr, TR_EINVAL, "Error: tr_create_multi: give 1..64 devices");
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++)
            if (devices[i] == devices[j]) return mfail(nullptr, TR_EINVAL, "Error: tr_create_multi: device %d listed twice", devices[i]);
    tr_multi *m = new (std::nothrow) tr_multi();
    if (!m) return mfail(nullptr, TR_ENOMEM, "Error: out of host memory");
    m->n = n;
    m->devices.assign(devices, devices + n);
    m->ctx.assign((size_t)n, nullptr);
    m->stream.assign((size_t)n, nullptr);
    m->comm.assign((size_t)n, nullptr);
    m->lo.assign((size_t)n, 0); m->hi.assign((size_t)n, 0);
    m->d_all.assign((size_t)n, nullptr); m->d_best.assign((size_t)n, nullptr); m->d_struct.assign((size_t)n, nullptr);
    m->cap_all.assign((size_t)n, 0); m->cap_struct.assign((size_t)n, 0);
    int rc = TR_OK;
    std::string why;
    for (int i = 0; i < n && rc == TR_OK; i++) {
        rc = tr_create(devices[i], &m->ctx[(size_t)i]);
        if (rc != TR_OK) { why = tr_last_error(nullptr); break; }
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream[(size_t)i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_best[(size_t)i], sizeof(tr_minimum));
        if (e != cudaSuccess) { rc = TR_ECUDA; why = std::string("Error: device ") + std::to_string(devices[i]) + ": " + cudaGetErrorString(e); break; }
        rc = tr_set_stream(m->ctx[(size_t)i], (void *)m->stream[(size_t)i]);
        if (rc != TR_OK) why = tr_last_error(m->ctx[(size_t)i]);
    }
    if (rc == TR_OK && n > 1) {
        if (!load_nccl(m->nccl, why)) { rc = TR_ENCCL; why = "Error: cannot load NCCL: " + why; }
        else {
            ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), n, m->devices.data());
            if (r != ncclSuccess) { rc = TR_ENCCL; why = std::string("Error: ncclCommInitAll: ") + m->nccl.GetErrorString(r); for (auto &c : m->comm) c = nullptr; }
        }
    }
    if (rc != TR_OK) {
        tr_destroy_multi(m);
        return mfail(nullptr, rc, "%s", why.c_str());
    }
    *out = m;
    return TR_OK;
}

extern "C" void tr_destroy_multi(tr_multi *m)
{
    if (!m) return;
    for (int i = 0; i < m->n; i++) {
        cudaSetDevice(m->devices[(size_t)i]);
        if (m->stream[(size_t)i]) cudaStreamSynchronize(m->stream[(size_t)i]);
        if (m->comm[(size_t)i] && m->nccl.CommDestroy) m->nccl.CommDestroy(m->comm[(size_t)i]);
        if (m->ctx[(size_t)i]) tr_destroy(m->ctx[(size_t)i]);
        if (m->d_all[(size_t)i]) cudaFree(m->d_all[(size_t)i]);
        if (m->d_best[(size_t)i]) cudaFree(m->d_best[(size_t)i]);
        if (m->d_struct[(size_t)i]) cudaFree(m->d_struct[(size_t)i]);
        if (m->stream[(size_t)i]) cudaStreamDestroy(m->stream[(size_t)i]);
    }
    delete m;
}

extern "C" const char *tr_multi_last_error(const tr_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

extern "C" int32_t tr_multi_size(const tr_multi *m) { return m ? m->n : 0; }

extern "C" tr_ctx *tr_multi_ctx(tr_multi *m, int32_t i) { return (m && i >= 0 && i < m->n) ? m->ctx[(size_t)i] : nullptr; }

extern "C" int tr_multi_set_system(tr_multi *m, const tr_system *sys)
{
    if (!m) return TR_EINVAL;
    if (!sys) return mfail(m, TR_EINVAL, "Error: tr_multi_set_system: sys is NULL");
    for (int i = 0; i < m->n; i++) MEACH(m, i, tr_set_system(m->ctx[(size_t)i], sys));
    m->n_sites = sys->n_sites;
    m->n_total = 0;
    return TR_OK;
}

extern "C" int tr_multi_set_schedules(tr_multi *m, const tr_schedule *sched, int32_t n_sched)
{
    if (!m) return TR_EINVAL;
    for (int i = 0; i < m->n; i++) MEACH(m, i, tr_set_schedules(m->ctx[(size_t)i], sched, n_sched));
    m->n_total = 0;
    return TR_OK;
}

extern "C" int tr_multi_set_options(tr_multi *m, int64_t trace_moves, int32_t record_points, int32_t record_snapshots, int32_t threads_per_chain)
{
    if (!m) return TR_EINVAL;
    for (int i = 0; i < m->n; i++) MEACH(m, i, tr_set_options(m->ctx[(size_t)i], trace_moves, record_points, record_snapshots, threads_per_chain));
    return TR_OK;
}

extern "C" int tr_multi_init_chains_seeded(tr_multi *m, int64_t n_chains, int64_t first_chain, const int64_t *seeds,
                                           const int32_t *sched_idx, double space_len, int32_t max_prop_failures)
{
    if (!m) return TR_EINVAL;
    if (!seeds || n_chains < m->n) return mfail(m, TR_EINVAL, "Error: tr_multi_init_chains_seeded: need seeds and at least one chain per device");
    m->n_total = 0;
    for (int i = 0; i < m->n; i++) {
        int64_t lo, hi;
        tr_multi_chain_range(i, m->n, n_chains, &lo, &hi);
        m->lo[(size_t)i] = lo; m->hi[(size_t)i] = hi;
        // seeds and parameter sets are indexed by the GLOBAL chain id: results do not depend on the number of devices
        MEACH(m, i, tr_init_chains_seeded(m->ctx[(size_t)i], hi - lo, first_chain + lo, seeds + lo, sched_idx ? sched_idx + lo : nullptr, space_len,
                                          max_prop_failures));
    }
    m->n_total = n_chains;
    m->first_chain = first_chain;
    return TR_OK;
}

extern "C" int tr_multi_run(tr_multi *m, int64_t max_moves)
{
    if (!m) return TR_EINVAL;
    if (m->n_total <= 0) return mfail(m, TR_EINVAL, "Error: no chains initialised (call tr_multi_init_chains_seeded first)");
    for (int i = 0; i < m->n; i++) MEACH(m, i, tr_run(m->ctx[(size_t)i], max_moves));      // asynchronous: all devices anneal concurrently
    return TR_OK;
}

extern "C" int tr_multi_synchronize(tr_multi *m)
{
    if (!m) return TR_EINVAL;
    for (int i = 0; i < m->n; i++) MEACH(m, i, tr_synchronize(m->ctx[(size_t)i]));
    return TR_OK;
}

extern "C" int tr_multi_last_run_ms(tr_multi *m, float *ms_max)
{
    if (!m || !ms_max) return TR_EINVAL;
    float mx = 0.f;
    for (int i = 0; i < m->n; i++) {
        float ms = 0.f;
        MEACH(m, i, tr_last_run_ms(m->ctx[(size_t)i], &ms));
        if (ms > mx) mx = ms;
    }
    *ms_max = mx;
    return TR_OK;
}

extern "C" int tr_gather_minima(tr_multi *m, tr_minimum *all, tr_minimum *best, double *best_sites)
{
    if (!m) return TR_EINVAL;
    if (m->n_total <= 0) return mfail(m, TR_EINVAL, "Error: no chains initialised");
    const size_t row = (size_t)m->n_sites * 3;
    // 1. every device packs its chains' (minimum energy, global chain id, tooth) into its slice of its own copy
    for (int i = 0; i < m->n; i++) {
        if (int r = grow(m, i, (size_t)m->n_total, row)) return r;
        MEACH(m, i, tr_pack_minima(m->ctx[(size_t)i], nullptr, m->d_all[(size_t)i] + m->lo[(size_t)i]));
    }
    // 2. all-gather over NVLink: afterwards every device holds every record
    if (m->n > 1) {
        std::vector<char *> bufs((size_t)m->n);
        for (int i = 0; i < m->n; i++) bufs[(size_t)i] = (char *)m->d_all[(size_t)i];
        if (int r = allgather_v(m, bufs, sizeof(tr_minimum))) return r;
    }
    // 3. arg-min on every device (each "rank" knows the winner without the host)
    for (int i = 0; i < m->n; i++) {
        MCK(m, cudaSetDevice(m->devices[(size_t)i]));
        argmin_minima_kernel<<<1, 256, 0, m->stream[(size_t)i]>>>(m->d_all[(size_t)i], m->n_total, m->d_best[(size_t)i]);
        MCK(m, cudaGetLastError());
    }
    tr_minimum b;
    MCK(m, cudaSetDevice(m->devices[0]));
    MCK(m, cudaMemcpyAsync(&b, m->d_best[0], sizeof b, cudaMemcpyDeviceToHost, m->stream[0]));
    if (all) MCK(m, cudaMemcpyAsync(all, m->d_all[0], (size_t)m->n_total * sizeof(tr_minimum), cudaMemcpyDeviceToHost, m->stream[0]));
    for (int i = 0; i < m->n; i++) { MCK(m, cudaSetDevice(m->devices[(size_t)i])); MCK(m, cudaStreamSynchronize(m->stream[(size_t)i])); }
    if (best) *best = b;
    if (!best_sites) return TR_OK;
    // 4. the owner packs the winning structure and broadcasts it; device 0 hands it to the host
    const long long g = b.chain - m->first_chain;
    int owner = 0;
    for (int i = 0; i < m->n; i++) if (g >= m->lo[(size_t)i] && g < m->hi[(size_t)i]) owner = i;
    MEACH(m, owner, tr_pack_min_structures(m->ctx[(size_t)owner], g - m->lo[(size_t)owner], 1, nullptr, m->d_struct[(size_t)owner]));
    if (m->n > 1) {
        MNCCL(m, m->nccl.GroupStart());
        for (int i = 0; i < m->n; i++)
            MNCCL(m, m->nccl.Broadcast(m->d_struct[(size_t)i], m->d_struct[(size_t)i], row * sizeof(double), ncclChar, owner, m->comm[(size_t)i], m->stream[(size_t)i]));
        MNCCL(m, m->nccl.GroupEnd());
    }
    MCK(m, cudaSetDevice(m->devices[0]));
    MCK(m, cudaMemcpyAsync(best_sites, m->d_struct[0], row * sizeof(double), cudaMemcpyDeviceToHost, m->stream[0]));
    for (int i = 0; i < m->n; i++) { MCK(m, cudaSetDevice(m->devices[(size_t)i])); MCK(m, cudaStreamSynchronize(m->stream[(size_t)i])); }
    return TR_OK;
}

extern "C" int tr_gather_min_structures(tr_multi *m, double *sites)
{
    if (!m) return TR_EINVAL;
    if (m->n_total <= 0) return mfail(m, TR_EINVAL, "Error: no chains initialised");
    if (!sites) return mfail(m, TR_EINVAL, "Error: tr_gather_min_structures: sites is NULL");
    const size_t row = (size_t)m->n_sites * 3;
    for (int i = 0; i < m->n; i++) {
        if (int r = grow(m, i, 0, (size_t)m->n_total * row)) return r;
        MEACH(m, i, tr_pack_min_structures(m->ctx[(size_t)i], 0, m->hi[(size_t)i] - m->lo[(size_t)i], nullptr,
                                           m->d_struct[(size_t)i] + (size_t)m->lo[(size_t)i] * row));
    }
    if (m->n > 1) {
        std::vector<char *> bufs((size_t)m->n);
        for (int i = 0; i < m->n; i++) bufs[(size_t)i] = (char *)m->d_struct[(size_t)i];
        if (int r = allgather_v(m, bufs, row * sizeof(double))) return r;
    }
    MCK(m, cudaSetDevice(m->devices[0]));
    MCK(m, cudaMemcpyAsync(sites, m->d_struct[0], (size_t)m->n_total * row * sizeof(double), cudaMemcpyDeviceToHost, m->stream[0]));
    for (int i = 0; i < m->n; i++) { MCK(m, cudaSetDevice(m->devices[(size_t)i])); MCK(m, cudaStreamSynchronize(m->stream[(size_t)i])); }
    return TR_OK;
}
