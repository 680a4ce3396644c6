// Latency / issue-rate probes for the instructions the anneal kernels are built from (sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu && tools/microbench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_dep(double *out, long long *cyc, int n)
{
    double a = threadIdx.x * 1e-3 + 1.0;
    const double m = 0.999999, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) a = fma(a, m, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
__global__ void k_dfma_ilp(double *out, long long *cyc, int n)
{
    double a[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) a[j] = threadIdx.x * 1e-3 + j;
    const double m = 0.999999, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int j = 0; j < ILP; j++) a[j] = fma(a[j], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_rsq_dep(double *out, long long *cyc, int n)
{
    double a = threadIdx.x * 1e-3 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            double y;
            asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
            a = y + 1.5;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_shfl_dep(double *out, long long *cyc, int n)
{
    double a = threadIdx.x * 1e-3 + 1.5;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lds_dep(double *out, long long *cyc, int n)
{
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 17 + 5) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) p = idx[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_bar(double *out, long long *cyc, int n)
{
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    }
    long long t1 = clock64();
    out[threadIdx.x] = (double)(t1 - t0);
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dconv(double *out, long long *cyc, int n)
{
    unsigned long long v = threadIdx.x + 12345;
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) { acc = (double)(long long)(v + (unsigned long long)acc); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_sincos(double *out, long long *cyc, int n)
{
    double a = threadIdx.x * 1e-3 + 0.1, s, c;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { sincos(a, &s, &c); a = s * 0.4 + c * 0.1; }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_exp(double *out, long long *cyc, int n)
{
    double a = -(threadIdx.x * 1e-3 + 0.1);
    long long t0 = clock64();
    for (int i = 0; i < n; i++) { a = -exp(a / 1.7); }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char *name, F launch, double ops)
{
    double *out; long long *cyc;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 16);
    launch(out, cyc);
    cudaDeviceSynchronize();
    launch(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %10.2f cycles/op  (%s)\n", name, (double)h / ops, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int n = 4096;
    run("DFMA dependent chain, 1 warp", [&](double *o, long long *c) { k_dfma_dep<<<1, 32>>>(o, c, n); }, n * 16.0);
    run("DFMA ILP2, 1 warp (per fma)", [&](double *o, long long *c) { k_dfma_ilp<2><<<1, 32>>>(o, c, n); }, n * 4.0 * 2);
    run("DFMA ILP4, 1 warp (per fma)", [&](double *o, long long *c) { k_dfma_ilp<4><<<1, 32>>>(o, c, n); }, n * 4.0 * 4);
    run("DFMA ILP8, 1 warp (per fma)", [&](double *o, long long *c) { k_dfma_ilp<8><<<1, 32>>>(o, c, n); }, n * 4.0 * 8);
    run("DFMA ILP8, 4 warps/SM (per fma per warp)", [&](double *o, long long *c) { k_dfma_ilp<8><<<1, 128>>>(o, c, n); }, n * 4.0 * 8);
    run("DFMA ILP8, 8 warps/SM (per fma per warp)", [&](double *o, long long *c) { k_dfma_ilp<8><<<1, 256>>>(o, c, n); }, n * 4.0 * 8);
    run("DFMA ILP8, 16 warps/SM (per fma per warp)", [&](double *o, long long *c) { k_dfma_ilp<8><<<1, 512>>>(o, c, n); }, n * 4.0 * 8);
    run("DFMA ILP1, 16 warps/SM (per fma per warp)", [&](double *o, long long *c) { k_dfma_ilp<1><<<1, 512>>>(o, c, n); }, n * 4.0 * 1);
    run("DFMA ILP1, 32 warps/SM (per fma per warp)", [&](double *o, long long *c) { k_dfma_ilp<1><<<1, 1024>>>(o, c, n); }, n * 4.0 * 1);
    run("MUFU.RSQ64H + DADD dependent", [&](double *o, long long *c) { k_rsq_dep<<<1, 32>>>(o, c, n); }, n * 8.0);
    run("shfl.xor f64 + DADD (per level)", [&](double *o, long long *c) { k_shfl_dep<<<1, 32>>>(o, c, n); }, n * 5.0);
    run("LDS dependent (pointer chase)", [&](double *o, long long *c) { k_lds_dep<<<1, 32>>>(o, c, n); }, n * 8.0);
    run("bar.sync, 160 threads", [&](double *o, long long *c) { k_bar<<<1, 160>>>(o, c, n); }, n * 8.0);
    run("bar.sync, 32 threads", [&](double *o, long long *c) { k_bar<<<1, 32>>>(o, c, n); }, n * 8.0);
    run("I2F.F64.S64 + IADD dependent", [&](double *o, long long *c) { k_dconv<<<1, 32>>>(o, c, n); }, n * 8.0);
    run("sincos(double) dependent", [&](double *o, long long *c) { k_sincos<<<1, 32>>>(o, c, n); }, n * 1.0);
    run("exp(double)+div dependent", [&](double *o, long long *c) { k_exp<<<1, 32>>>(o, c, n); }, n * 1.0);
    return 0;
}
