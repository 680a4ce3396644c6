// evalbench.cu - how fast can ONE warp (or two on the same SM sub-partition) run the FP64 evaluation group of crew2.cuh?
// NE independent pair evaluations per staged site, written operation by operation (as c2_evaluate_wide does).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/evalbench tools/evalbench.cu && tools/evalbench
#include <cstdio>
#include <cuda_runtime.h>

template <int NE, int VARIANT>
__global__ void __launch_bounds__(256, 1) bench(const double *in, double *out, long long *cyc, int iters, int warps_mask)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (!((warps_mask >> warp) & 1)) return;
    double x[NE], y[NE], z[NE], q[NE], e[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < NE; j++) { x[j] = in[lane + 32 * j]; y[j] = in[lane + 32 * (j + NE)]; z[j] = in[lane + 32 * (j + 2 * NE)]; q[j] = in[lane + 32 * (j + 3 * NE)]; }
    double nx = in[7], ny = in[8], nz = in[9], kq = in[10];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        double r2[NE], y0[NE], t[NE], u[NE], p[NE], ri[NE];
#pragma unroll
        for (int j = 0; j < NE; j++) { const double d = x[j] - nx; r2[j] = d * d; }
#pragma unroll
        for (int j = 0; j < NE; j++) { const double d = y[j] - ny; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
        for (int j = 0; j < NE; j++) { const double d = z[j] - nz; r2[j] = fma(d, d, r2[j]); }
#pragma unroll
        for (int j = 0; j < NE; j++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0[j]) : "d"(r2[j]));
#pragma unroll
        for (int j = 0; j < NE; j++) t[j] = r2[j] * y0[j];
#pragma unroll
        for (int j = 0; j < NE; j++) u[j] = fma(-t[j], y0[j], 1.0);
#pragma unroll
        for (int j = 0; j < NE; j++) p[j] = fma(u[j], 0.375, 0.5);
#pragma unroll
        for (int j = 0; j < NE; j++) u[j] = y0[j] * u[j];
#pragma unroll
        for (int j = 0; j < NE; j++) ri[j] = fma(p[j], u[j], y0[j]);
        if (VARIANT == 0) {
#pragma unroll
            for (int j = 0; j < NE; j++) e[j % 3] = fma(kq * q[j], ri[j], e[j % 3]);
        } else {
            double s[3] = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < NE; j++) s[j % 3] = fma(q[j], ri[j], s[j % 3]);
#pragma unroll
            for (int k = 0; k < 3; k++) e[k] = fma(kq, s[k], e[k]);
        }
        nx += 1e-9; ny -= 1e-9; nz += 2e-9;        // next staged site
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    out[threadIdx.x] = e[0] + e[1] + e[2];
}

template <int NE, int VARIANT>
void run(const char *name, const double *d_in, double *d_out, long long *d_cyc, int mask)
{
    const int iters = 3000;
    bench<NE, VARIANT><<<1, 256>>>(d_in, d_out, d_cyc, iters, mask);
    bench<NE, VARIANT><<<1, 256>>>(d_in, d_out, d_cyc, iters, mask);
    long long c[8] = {0};
    cudaMemcpy(c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    int w = 0;
    while (!((mask >> w) & 1)) w++;
    const int fp64 = NE * (VARIANT == 0 ? 13 : 12) + (VARIANT ? 3 : 0) + 3;
    printf("%-34s NE=%d warps mask 0x%02x: %7.1f cycles per staged site, %5.2f cycles per FP64 instruction (%d per site)\n", name, NE, mask,
           (double)c[w] / iters, (double)c[w] / iters / fp64, fp64);
}

int main()
{
    double h[32 * 40];
    for (int i = 0; i < 32 * 40; i++) h[i] = 1.0 + 0.37 * (i % 17) + 0.011 * i;
    double *d_in, *d_out; long long *d_cyc;
    cudaMalloc(&d_in, sizeof h); cudaMalloc(&d_out, 256 * 8); cudaMalloc(&d_cyc, 64);
    cudaMemcpy(d_in, h, sizeof h, cudaMemcpyHostToDevice);
    run<3, 0>("3 in flight", d_in, d_out, d_cyc, 0x01);
    run<6, 0>("6 in flight", d_in, d_out, d_cyc, 0x01);
    run<9, 0>("9 in flight", d_in, d_out, d_cyc, 0x01);
    run<9, 1>("9 in flight, kq folded", d_in, d_out, d_cyc, 0x01);
    run<9, 0>("9 in flight, 2 warps same SMSP", d_in, d_out, d_cyc, 0x11);
    run<9, 0>("9 in flight, 2 warps other SMSPs", d_in, d_out, d_cyc, 0x03);
    run<3, 0>("3 in flight, 2 warps same SMSP", d_in, d_out, d_cyc, 0x11);
    run<9, 0>("9 in flight, 4 warps (one per SMSP)", d_in, d_out, d_cyc, 0x0f);
    run<9, 0>("9 in flight, 8 warps (two per SMSP)", d_in, d_out, d_cyc, 0xff);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
