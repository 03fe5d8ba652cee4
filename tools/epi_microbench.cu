// Microbenchmark for the packed 16-bit top-2 "insert" the tensor-engine epilogue is made of (sm_100a).
// One insert updates the running (best, second best) of TWO independent 16-bit lanes packed in one register:
//     m1 = min(m1, max(m0, v));  m0 = min(m0, v)
// Variants (all exact on the value range the ORB engine uses: non-negative multiples of 1/2 below 1024 as f16 bit patterns):
//   u16   3 x VIMNMX.U16x2            (min.u16x2 / max.u16x2)
//   f16   3 x HMNMX2                  (min.f16x2 / max.f16x2)
//   relu  5 FMA-pipe ops              d = relu(m0 - v); m0 -= d; x = v + d; e = relu(m1 - x); m1 -= e   (HFMA2.RELU + HADD2)
//   mixes of the above on independent chains, to see which of them issue to different pipes.
// Output: packed inserts / clk / SM (each covers two scores).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/epi_microbench tools/epi_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int CHAINS = 8;
constexpr int NIN = 8;

__device__ __forceinline__ uint32_t minu2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t maxu2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t minh2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t maxh2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// relu(c - a) as one HFMA2.RELU: a * (-1) + c
__device__ __forceinline__ uint32_t relu_sub(uint32_t c, uint32_t a, uint32_t neg1) { uint32_t r; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(neg1), "r"(c)); return r; }
__device__ __forceinline__ uint32_t subh2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addh2(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int KIND>   // 0 = u16, 1 = f16, 2 = relu, 3 = u16 with the 3-input min (tree step: two values at once)
__device__ __forceinline__ void insert(uint32_t& m0, uint32_t& m1, uint32_t v, uint32_t neg1) {
    if (KIND == 0) { m1 = minu2(m1, maxu2(m0, v)); m0 = minu2(m0, v); }
    else if (KIND == 1) { m1 = minh2(m1, maxh2(m0, v)); m0 = minh2(m0, v); }
    else {
        const uint32_t d = relu_sub(m0, v, neg1);
        m0 = subh2(m0, d);
        const uint32_t x = addh2(v, d);
        const uint32_t e = relu_sub(m1, x, neg1);
        m1 = subh2(m1, e);
    }
}

// NA chains of kind KA followed by CHAINS-NA chains of kind KB, interleaved in the instruction stream
template <int KA, int KB, int NA>
__global__ void __launch_bounds__(1024, 1) bench(const uint32_t* __restrict__ in, uint32_t* out, long long* cycles) {
    uint32_t m0[CHAINS], m1[CHAINS], v[NIN];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { m0[i] = 0x63D063D0u; m1[i] = 0x63D063D0u; }     // 1000.0 | 1000.0
#pragma unroll
    for (int i = 0; i < NIN; ++i) v[i] = in[threadIdx.x * NIN + i];
    const uint32_t neg1 = in[1024 * NIN];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < NIN; ++k) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (i < NA) insert<KA>(m0[i], m1[i], v[(k + i) % NIN], neg1);
                else insert<KB>(m0[i], m1[i], v[(k + i) % NIN], neg1);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += m0[i] ^ m1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// correctness of the relu form against the integer form on the value range in use
__global__ void check_relu(const uint32_t* __restrict__ in, uint32_t* bad) {
    const uint32_t neg1 = in[1024 * NIN];
    uint32_t a0 = 0x63D063D0u, a1 = 0x63D063D0u, b0 = a0, b1 = a1;
    for (int i = 0; i < 1024 * NIN; ++i) {
        const uint32_t v = in[(i * 37 + threadIdx.x * 101) % (1024 * NIN)];
        insert<0>(a0, a1, v, neg1);
        insert<2>(b0, b1, v, neg1);
        if (a0 != b0 || a1 != b1) atomicAdd(bad, 1u);
    }
}

template <int KA, int KB, int NA>
int run(const char* name, const uint32_t* in, uint32_t* out, long long* cyc, int nsm) {
    bench<KA, KB, NA><<<nsm, 1024>>>(in, out, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KA, KB, NA><<<nsm, 1024>>>(in, out, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
    const double inserts = 1024.0 * ITERS * NIN * CHAINS;
    printf("{\"variant\": \"%s\", \"packed_inserts_per_clk_per_sm\": %.2f, \"scores_per_clk_per_sm\": %.2f, \"cycles\": %lld, \"ms\": %.4f, \"implied_mhz\": %.0f}\n",
           name, inserts / (double)mx, 2 * inserts / (double)mx, mx, ms, (double)mx / (ms * 1e3));
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, nsm);
    // inputs: f16 bit patterns of multiples of 1/2 in [0, 600]
    const int n = 1024 * NIN + 1;
    uint32_t* h = new uint32_t[n];
    uint32_t lcg = 12345u;
    auto half_bits = [](float x) -> uint32_t {           // x = k/2, 0 <= x < 1024: exact conversion by hand
        if (x == 0.f) return 0u;
        int e = 0; float m = x;
        while (m >= 2.f) { m *= 0.5f; ++e; }
        while (m < 1.f) { m *= 2.f; --e; }
        return (uint32_t)(((e + 15) << 10) | (int)((m - 1.f) * 1024.f + 0.5f));
    };
    for (int i = 0; i < n - 1; ++i) {
        lcg = lcg * 1664525u + 1013904223u; const uint32_t a = (lcg >> 8) % 1201u;
        lcg = lcg * 1664525u + 1013904223u; const uint32_t b = (lcg >> 8) % 1201u;
        h[i] = half_bits(a * 0.5f) | (half_bits(b * 0.5f) << 16);
    }
    h[n - 1] = 0xBC00BC00u;     // -1.0 | -1.0
    uint32_t *in, *out, *bad; long long* cyc;
    CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * 1024)); CK(cudaMalloc(&cyc, sizeof(long long) * 256));
    CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
    CK(cudaMemcpy(in, h, n * 4, cudaMemcpyHostToDevice));
    check_relu<<<1, 128>>>(in, bad);
    CK(cudaDeviceSynchronize());
    uint32_t hb = 0; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"relu insert == u16 insert on multiples of 1/2 in [0,600]\", \"mismatches\": %u}\n", hb);
    run<0, 0, 8>("u16 (3 x VIMNMX.U16x2)", in, out, cyc, nsm);
    run<1, 1, 8>("f16 (3 x HMNMX2)", in, out, cyc, nsm);
    run<2, 2, 8>("relu (2 x HFMA2.RELU + 3 x HADD2)", in, out, cyc, nsm);
    run<0, 1, 4>("mix u16:f16 4:4", in, out, cyc, nsm);
    run<0, 2, 4>("mix u16:relu 4:4", in, out, cyc, nsm);
    run<0, 2, 5>("mix u16:relu 5:3", in, out, cyc, nsm);
    run<0, 2, 6>("mix u16:relu 6:2", in, out, cyc, nsm);
    run<1, 2, 4>("mix f16:relu 4:4", in, out, cyc, nsm);
    run<1, 2, 5>("mix f16:relu 5:3", in, out, cyc, nsm);
    return 0;
}
