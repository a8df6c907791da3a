// Micro-benchmarks for the two pipes that bound the RDF pair kernel on B200:
// FP64 vector throughput (no FMA contraction, as the bin-deciding arithmetic requires)
// and shared-memory / L2 atomic-increment throughput on a histogram-like address stream.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o tools/microbench tools/microbench.cu
// Prints one JSON object.  SURVEY.md 8(d): "Builder must measure both with micro-benchmarks".
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---- FP64: 8 independent chains of (sub, mul, add) per thread, no FMA ----
__global__ void __launch_bounds__(512) k_fp64(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double d = x[i] - b;      // DADD
            double s = d * d;         // DMUL
            x[i] = s + a;             // DADD
        }
    }
    double s = 0; 
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(512) k_dfma(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = __fma_rn(x[i], b, a);
            x[i] = __fma_rn(x[i], b, a);
            x[i] = __fma_rn(x[i], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

// ---- shared atomics on a histogram of `words` u32 counters ----
// mode 0: ATOMS on random addresses, all lanes active
// mode 1: ATOMS, ~25% of lanes active (predicated)
// mode 2: global RED (atomicAdd without return) on random addresses within `words`
// mode 3: warp-private non-atomic RMW with __match_any_sync conflict resolution
__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(512) k_atom(uint32_t* gh, int words, int iters, unsigned long long* sink) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t mask = (uint32_t)words - 1u;  // words is a power of two
    for (int it = 0; it < iters; ++it) {
        uint32_t r = lcg(s);
        uint32_t idx = r & mask;
        if (MODE == 0) {
            atomicAdd(&sh[idx], 1u);
        } else if (MODE == 1) {
            if ((r >> 20) & 3u) continue;   // 25 % active
            atomicAdd(&sh[idx], 1u);
        } else if (MODE == 2) {
            atomicAdd(&gh[idx], 1u);
        } else if (MODE == 3) {
            // warp-private slice: each warp owns words/nwarps counters
            int nw = blockDim.x >> 5, w = threadIdx.x >> 5;
            uint32_t per = (uint32_t)words / nw;          // power of two if nw is
            uint32_t id = w * per + (idx & (per - 1u));
            unsigned m = __match_any_sync(0xffffffffu, id);
            int leader = __ffs(m) - 1;
            if ((threadIdx.x & 31) == leader) sh[id] += __popc(m);
            __syncwarp();
        }
    }
    __syncthreads();
    unsigned long long t = 0;
    for (int i = threadIdx.x; i < words; i += blockDim.x) t += sh[i];
    if (t == 0xdeadbeefULL) sink[0] = t;
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    double* dout; CK(cudaMalloc(&dout, 64));
    uint32_t* gh; CK(cudaMalloc(&gh, 1 << 22)); CK(cudaMemset(gh, 0, 1 << 22));
    unsigned long long* sink; CK(cudaMalloc(&sink, 64));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", p.name, sms, p.clockRate);

    {   // FP64 non-FMA
        int iters = 4096, blocks = sms * 4, thr = 512;
        float ms = time_ms([&] { k_fp64<<<blocks, thr>>>(dout, iters, 1.000001, 0.5); }, 5);
        double ops = (double)blocks * thr * iters * 8 * 3;
        printf(", \"fp64_nofma_gops\": %.1f", ops / ms * 1e-6);
        ms = time_ms([&] { k_dfma<<<blocks, thr>>>(dout, iters, 1.000001, 0.5); }, 5);
        printf(", \"fp64_dfma_ginst\": %.1f", ops / ms * 1e-6);
    }
    int words = 16384; size_t shb = words * 4;
    int iters = 8192, thr = 512;
    CK(cudaFuncSetAttribute(k_atom<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_atom<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_atom<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    CK(cudaFuncSetAttribute(k_atom<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int bps = 1; bps <= 2; ++bps) {
        int blocks = sms * bps;
        double n = (double)blocks * thr * iters;
        float ms = time_ms([&] { k_atom<0><<<blocks, thr, shb>>>(gh, words, iters, sink); }, 5);
        printf(", \"atoms_rand_full_gops_bps%d\": %.1f", bps, n / ms * 1e-6);
        ms = time_ms([&] { k_atom<1><<<blocks, thr, shb>>>(gh, words, iters, sink); }, 5);
        printf(", \"atoms_rand_quarter_gops_bps%d\": %.1f", bps, n * 0.25 / ms * 1e-6);
        ms = time_ms([&] { k_atom<3><<<blocks, thr, shb>>>(gh, words, iters, sink); }, 5);
        printf(", \"warp_private_rmw_gops_bps%d\": %.1f", bps, n / ms * 1e-6);
    }
    {
        int blocks = sms * 2; double n = (double)blocks * thr * iters;
        for (int w = 16384; w <= 262144; w *= 16) {
            float ms = time_ms([&] { k_atom<2><<<blocks, thr, shb>>>(gh, w, iters, sink); }, 5);
            printf(", \"redg_rand_gops_words%d\": %.1f", w, n / ms * 1e-6);
        }
    }
    printf("}\n");
    return 0;
}
