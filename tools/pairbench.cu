// Cycles of the pair phase of one (H2O)20 move for ONE warp alone on an SM (no contention): 3 staged sites x 4 slots.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I transrot_b200/csrc -o tools/pairbench tools/pairbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "anneal.cuh"
using namespace tr;

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_pair(double *out, long long *cyc, int n, const double2 *cdtab)
{
    __shared__ __align__(16) StageSite stage[WARPS][3];
    __shared__ __align__(16) double2 cd[16];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 16) cd[threadIdx.x] = cdtab[threadIdx.x];
    if (lane < 3) {
        StageSite st;
        st.ox = 0.1 * lane; st.oy = 0.2; st.oz = 0.3; st.nx = 0.15 * lane; st.ny = 0.25; st.nz = 0.35; st.kq = 100.0; st.row = 0; st.lj = lane == 0;
        stage[w][lane] = st;
    }
    __syncthreads();
    double x[4], y[4], z[4], qm[4];
    int col[4];
    for (int k = 0; k < 4; k++) { x[k] = 3.0 + lane + 0.1 * k; y[k] = 2.0 + k; z[k] = 1.0 + 0.3 * lane; qm[k] = 0.4; col[k] = (k & 1) * 16; }
    double acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        double eS = 0.0, eE = 0.0;
        for (int a = 0; a < 3; a++) {
            const StageSite st = stage[w][a];
            if (st.lj) stage_site_energy<4, 2, 1>(st, (const unsigned char *)cd, (const unsigned char *)cd, x, y, z, qm, col, eS, eE);
            else stage_site_energy<4, 0, 0>(st, (const unsigned char *)cd, (const unsigned char *)cd, x, y, z, qm, col, eS, eE);
        }
        acc += eE - eS;
        x[0] += 1e-9 * acc;      // loop-carried dependency like the real chain (next move needs this one's result)
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    double *out; long long *cyc; double2 *cd;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 12); cudaMalloc(&cd, 16 * sizeof(double2));
    double2 h[16]; for (int i = 0; i < 16; i++) h[i] = make_double2(600.0 * (i == 0), 580000.0 * (i == 0));
    cudaMemcpy(cd, h, sizeof h, cudaMemcpyHostToDevice);
    const int n = 2000;
    long long c;
    k_pair<1><<<1, 32>>>(out, cyc, n, cd); cudaDeviceSynchronize();
    k_pair<1><<<1, 32>>>(out, cyc, n, cd); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 warp / SM : %.0f cycles per pair phase (24 evaluations per lane)\n", (double)c / n);
    k_pair<4><<<1, 128>>>(out, cyc, n, cd); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 warps / SM (1 per SMSP): %.0f\n", (double)c / n);
    k_pair<8><<<1, 256>>>(out, cyc, n, cd); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("8 warps / SM (2 per SMSP): %.0f\n", (double)c / n);
    k_pair<16><<<1, 512>>>(out, cyc, n, cd); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("16 warps / SM (4 per SMSP): %.0f   (%s)\n", (double)c / n, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
