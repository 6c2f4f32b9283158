// Microbenchmark: how close do the register-resident fp64 codelets of the spectrogram
// kernel get to the fp64 pipe's peak, as a function of warps per SM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_codelets fp64_codelets.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../asr_dfcnn_transformer_b200/csrc/asrk_fft.cuh"
using namespace asrk;

__constant__ double c_tab[1200];

template <int WARPS, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) k(double* out, int iters, int role_in) {
    // role made warp-uniform the CUTLASS way
    const int role = __shfl_sync(0xffffffffu, (threadIdx.x >> 5) % 10 + role_in, 0);
    cplx z[20], y[20];
#pragma unroll
    for (int i = 0; i < 20; ++i) z[i] = cplx{(double)(threadIdx.x + i) * 1e-3, (double)(blockIdx.x - i) * 1e-3};
    const cplx* tw = reinterpret_cast<const cplx*>(c_tab + 400) + role * 20;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // pass 1: DFT20 + twiddle (tables from the constant bank, uniform)
            dft20(z, y);
#pragma unroll
            for (int k1 = 1; k1 < 20; ++k1) z[k1] = cmul(y[k1], tw[k1]);
            z[0] = y[0];
        } else {                    // pass 2: 2 x DFT10 + 10 split pairs
            cplx a[10], b[10], za[10], zb[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) { a[i] = z[i]; b[i] = z[10 + i]; }
            dft10(a, za);
            dft10(b, zb);
            const cplx* P = reinterpret_cast<const cplx*>(c_tab + 800) + role;
#pragma unroll
            for (int s = 0; s < 10; ++s) {
                double pk, pm;
                split_pair(za[s], zb[9 - s], P[20 * s], pk, pm);
                z[s] = cplx{pk * 1e-3, za[s].y * 0.5};
                z[10 + s] = cplx{pm * 1e-3, zb[s].x * 0.5};
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 20; ++i) acc += z[i].x + z[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WARPS, int MODE>
void run(double* d, int sms, double fp64_per_iter) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WARPS, MODE><<<sms, WARPS * 32>>>(d, 10, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<WARPS, MODE><<<sms, WARPS * 32>>>(d, iters, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double inst = fp64_per_iter * iters * WARPS * 32.0 * sms;     // lane-instructions
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("mode %d warps %2d: %.3f ms  %.2f fp64 lane-instr/clk/SM (at %d MHz nominal; 64 = peak)  err=%s\n", MODE, WARPS, ms,
           inst / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double h[1200]; for (int i = 0; i < 1200; ++i) h[i] = 0.5 + 1e-3 * i;
    cudaMemcpyToSymbol(c_tab, h, sizeof(h));
    double* d; cudaMalloc(&d, sizeof(double) * sms * 1024);
    // pass 1: 224 (DFT20) + 76 (twiddles) = 300 fp64 instructions per iteration
    run<4, 0>(d, sms, 300); run<8, 0>(d, sms, 300); run<12, 0>(d, sms, 300); run<16, 0>(d, sms, 300);
    // pass 2: 184 + 160 + 40 = 384
    run<4, 1>(d, sms, 384); run<8, 1>(d, sms, 384); run<12, 1>(d, sms, 384); run<16, 1>(d, sms, 384);
    return 0;
}
