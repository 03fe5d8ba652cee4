// Microbenchmark: 256-bit Hamming distance with carry-save compression before POPC (fewer POPC, more LOP3),
// in the same register-tiled loop as orb_match_pairs_kernel (8 rows/thread, top-2 both directions, REDUX merge).
// NPOPC = 8: plain; 6: two CSAs; 4: four CSAs (Harley-Seal step).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <int NPOPC>
__device__ __forceinline__ uint32_t hamming(const uint32_t (&a)[8], const uint4& b0, const uint4& b1) {
    const uint32_t x0 = a[0] ^ b0.x, x1 = a[1] ^ b0.y, x2 = a[2] ^ b0.z, x3 = a[3] ^ b0.w, x4 = a[4] ^ b1.x, x5 = a[5] ^ b1.y,
                   x6 = a[6] ^ b1.z, x7 = a[7] ^ b1.w;
    if (NPOPC == 8) return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
    const uint32_t s0 = xor3(x0, x1, x2), c0 = maj3(x0, x1, x2), s1 = xor3(x3, x4, x5), c1 = maj3(x3, x4, x5);
    if (NPOPC == 6) return __popc(s0) + __popc(s1) + __popc(x6) + __popc(x7) + 2 * (__popc(c0) + __popc(c1));
    const uint32_t s2 = xor3(s0, s1, x6), c2 = maj3(s0, s1, x6);
    if (NPOPC == 5) return __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
    const uint32_t s3 = xor3(c0, c1, c2), c3 = maj3(c0, c1, c2);
    return __popc(s2) + __popc(x7) + 2 * __popc(s3) + 4 * __popc(c3);
}

__device__ __forceinline__ void sort2(uint32_t& lo, uint32_t& hi) { const uint32_t a = lo, b = hi; lo = min(a, b); hi = max(a, b); }
// top-2 of two sorted pairs with 3-input min
__device__ __forceinline__ void merge_pairs(uint32_t& lo, uint32_t& hi, uint32_t lo2, uint32_t hi2) {
    const uint32_t mx = max(lo, lo2);
    lo = min(lo, lo2);
    hi = __vimin3_u32(mx, hi, hi2);
}

// top-2 insert with 2 compares (ALU) + 3 predicated moves (can issue on the FMA pipe as IMAD.MOV)
__device__ __forceinline__ void insert_pred(uint32_t k, uint32_t& m0, uint32_t& m1) {
    asm volatile(
        "{\n\t.reg .pred p0, p1;\n\t"
        "setp.lt.u32 p0, %2, %0;\n\t"
        "setp.lt.u32 p1, %2, %1;\n\t"
        "@p1 mov.u32 %1, %2;\n\t"
        "@p0 mov.u32 %1, %0;\n\t"
        "@p0 mov.u32 %0, %2;\n\t}"
        : "+r"(m0), "+r"(m1) : "r"(k));
}

template <int NPOPC, int COLS_TOO>
__global__ void __launch_bounds__(512, 1) mixp(uint32_t* out, long long* cycles, uint32_t seed, int ncols) {
    __shared__ uint4 cols[2 * 256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { uint32_t v = seed * (i + 7); cols[i] = make_uint4(v, v * 3, v * 5, v * 7); }
    uint32_t a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int w = 0; w < 8; ++w) a[r][w] = seed * (threadIdx.x * 64 + r * 8 + w + 1);
    uint32_t m0[8], m1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { m0[r] = 0xffffffffu; m1[r] = 0xffffffffu; }
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < ncols; ++j) {
        uint4 b0 = cols[(j & 255) * 2], b1 = cols[(j & 255) * 2 + 1];
        uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t d = hamming<NPOPC>(a[r], b0, b1);
            uint32_t kr = (d << 16) + j, kc = (d << 16) + r * 512;
            insert_pred(kr, m0[r], m1[r]);
            if (COLS_TOO) insert_pred(kc, c0, c1);
            else { c1 = min(c1, max(c0, kc)); c0 = min(c0, kc); }
        }
        uint32_t g0 = __reduce_min_sync(0xffffffffu, c0);
        uint32_t x = (c0 == g0) ? c1 : c0;
        uint32_t g1 = __reduce_min_sync(0xffffffffu, x);
        acc += g0 ^ g1;
    }
    long long t1 = clock64();
    uint32_t s = acc;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += m0[r] ^ m1[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NPOPC>
__global__ void __launch_bounds__(512, 1) mix3(uint32_t* out, long long* cycles, uint32_t seed, int ncols) {
    __shared__ uint4 cols[2 * 256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { uint32_t v = seed * (i + 7); cols[i] = make_uint4(v, v * 3, v * 5, v * 7); }
    uint32_t a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int w = 0; w < 8; ++w) a[r][w] = seed * (threadIdx.x * 64 + r * 8 + w + 1);
    uint32_t m0[8], m1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { m0[r] = 0xffffffffu; m1[r] = 0xffffffffu; }
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < ncols; ++j) {
        uint4 b0 = cols[(j & 255) * 2], b1 = cols[(j & 255) * 2 + 1];
        uint32_t kc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t d = hamming<NPOPC>(a[r], b0, b1);
            uint32_t kr = (d << 16) + j;
            kc[r] = (d << 16) + r * 512;
            m1[r] = min(m1[r], max(m0[r], kr)); m0[r] = min(m0[r], kr);
        }
        // column direction: tournament top-2 of the 8 fresh keys (17 ops instead of 20)
        sort2(kc[0], kc[1]); sort2(kc[2], kc[3]); sort2(kc[4], kc[5]); sort2(kc[6], kc[7]);
        merge_pairs(kc[0], kc[1], kc[2], kc[3]); merge_pairs(kc[4], kc[5], kc[6], kc[7]);
        merge_pairs(kc[0], kc[1], kc[4], kc[5]);
        uint32_t c0 = kc[0], c1 = kc[1];
        uint32_t g0 = __reduce_min_sync(0xffffffffu, c0);
        uint32_t x = (c0 == g0) ? c1 : c0;
        uint32_t g1 = __reduce_min_sync(0xffffffffu, x);
        acc += g0 ^ g1;
    }
    long long t1 = clock64();
    uint32_t s = acc;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += m0[r] ^ m1[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NPOPC>
__global__ void __launch_bounds__(512, 1) mix(uint32_t* out, long long* cycles, uint32_t seed, int ncols) {
    __shared__ uint4 cols[2 * 256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { uint32_t v = seed * (i + 7); cols[i] = make_uint4(v, v * 3, v * 5, v * 7); }
    uint32_t a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int w = 0; w < 8; ++w) a[r][w] = seed * (threadIdx.x * 64 + r * 8 + w + 1);
    uint32_t m0[8], m1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { m0[r] = 0xffffffffu; m1[r] = 0xffffffffu; }
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < ncols; ++j) {
        uint4 b0 = cols[(j & 255) * 2], b1 = cols[(j & 255) * 2 + 1];
        uint32_t c0 = 0xffffffffu, c1 = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t d = hamming<NPOPC>(a[r], b0, b1);
            uint32_t kr = (d << 16) + j, kc = (d << 16) + r * 512;
            m1[r] = min(m1[r], max(m0[r], kr)); m0[r] = min(m0[r], kr);
            c1 = min(c1, max(c0, kc)); c0 = min(c0, kc);
        }
        uint32_t g0 = __reduce_min_sync(0xffffffffu, c0);
        uint32_t x = (c0 == g0) ? c1 : c0;
        uint32_t g1 = __reduce_min_sync(0xffffffffu, x);
        acc += g0 ^ g1;
    }
    long long t1 = clock64();
    uint32_t s = acc;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += m0[r] ^ m1[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NPOPC, int V>
void run(uint32_t* out, long long* cyc, int nsm) {
    for (int rep = 0; rep < 2; ++rep) {
        if (V == 1) mix3<NPOPC><<<nsm, 512>>>(out, cyc, 977u, 4096); else if (V == 2) mixp<NPOPC, 0><<<nsm, 512>>>(out, cyc, 977u, 4096); else if (V == 3) mixp<NPOPC, 1><<<nsm, 512>>>(out, cyc, 977u, 4096); else mix<NPOPC><<<nsm, 512>>>(out, cyc, 977u, 4096);
        cudaDeviceSynchronize();
        long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
        double dists = 512.0 * 8 * 4096;
        if (rep) printf("{\"op\": \"orb_inner_loop_csa\", \"variant\": %d, \"popc_per_distance\": %d, \"distances_per_clk_per_sm\": %.3f, \"equiv_popc8_per_clk_per_sm\": %.2f, \"cycles\": %lld}\n",
                        V, NPOPC, dists / mx, 8 * dists / mx, mx);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * nsm * 512); cudaMalloc(&cyc, sizeof(long long) * 256);
    run<8, 0>(out, cyc, nsm); run<6, 0>(out, cyc, nsm); run<5, 0>(out, cyc, nsm); run<4, 0>(out, cyc, nsm); run<5, 1>(out, cyc, nsm); run<4, 1>(out, cyc, nsm); run<4, 2>(out, cyc, nsm); run<4, 3>(out, cyc, nsm); run<5, 2>(out, cyc, nsm);
    return 0;
}
