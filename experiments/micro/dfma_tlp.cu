// DFMA issue rate against resident warps per SM and independent chains per thread (does 16 warps x 10 chains reach the FP64 peak?)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = __fma_rn(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}
template <int ILP>
void run(double *d, int sms) {
    for (int thr : {128, 256, 384, 512, 768, 1024}) {
        int iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<ILP><<<sms, thr>>>(d, 100, 1.0000001, 0.5);
        cudaEventRecord(e0);
        k<ILP><<<sms, thr>>>(d, iters, 1.0000001, 0.5);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)sms * thr * iters * ILP;
        printf("ILP %2d warps/SM %2d: %.2f T DFMA/s\n", ILP, thr / 32, inst / ms / 1e9);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *d; cudaMalloc(&d, 8);
    run<1>(d, p.multiProcessorCount); run<2>(d, p.multiProcessorCount); run<4>(d, p.multiProcessorCount);
    run<8>(d, p.multiProcessorCount); run<16>(d, p.multiProcessorCount);
    return 0;
}
