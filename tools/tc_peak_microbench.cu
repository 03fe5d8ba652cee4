// Tensor-core / TMEM probes for the matching engine (sm_100a). Three parts, one JSON line each:
//   1. F16-accumulator exactness: one 128 x 128 ORB tile (bits as e4m3 0/1, negate-A, augmentation) through
//      tcgen05.mma kind::f8f6f4 with D format F16, read back with tcgen05.ld (plain and .pack::16b), compared with the
//      CPU Hamming distance. Establishes the register layout of .pack::16b (reg i = column 2i | column 2i+1 << 16).
//   2. MMA peak: one persistent CTA per SM issuing back-to-back M=128 MMAs (9 K-steps per tile, 2 TMEM stages) for
//      kind::f16 (bf16) and kind::f8f6f4 (e4m3) with F32 / F16 accumulators, N = 128 / 256. The measured UTCQMMA rate is the
//      denominator of the ORB engine's tensor roofline (bench.py reads profiles/r02_tc_peak.jsonl).
//   3. TMEM read rate: 16 warps looping tcgen05.ld (x16, x32, with and without .pack::16b).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o tools/tc_peak_microbench tools/tc_peak_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../eacham_b200/csrc/tc_match_kernels.cuh"

using namespace eacham;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, addr = tc::smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 24) && !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
    return done != 0;
}

__host__ __device__ constexpr uint32_t idesc_any(bool fp8, bool d_f32, uint32_t m, uint32_t n, bool negate_a) {
    return ((d_f32 ? 1u : 0u) << 4) | (fp8 ? 0u : ((1u << 7) | (1u << 10))) | ((negate_a ? 1u : 0u) << 13) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16_pack(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
#define R32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
    "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), \
    "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), \
    "=r"(v[30]), "=r"(v[31])
#define L32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " L32 : R32(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32_pack(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " L32 : R32(v) : "r"(taddr) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// 1. F16 accumulator exactness + .pack::16b layout
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) f16_tile_kernel(const uint8_t* a_blk, const uint8_t* b_blk, uint32_t* out_plain, uint32_t* out_pack,
                                                          uint32_t* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + tc::kAOperandBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * tc::kAOperandBytes);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { tc::mbar_init(&bars[0], 1); tc::mbar_init(&bars[1], 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(tslot, 128);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        tc::mbar_expect_tx(&bars[0], 2 * tc::kAOperandBytes);
        tc::bulk_g2s(sA, a_blk, tc::kAOperandBytes, &bars[0]);
        tc::bulk_g2s(sB, b_blk, tc::kDataBytes, &bars[0]);
        tc::bulk_g2s(sB + tc::kDataBytes, b_blk + tc::kDataBytes + tc::kAugBytes, tc::kAugBytes, &bars[0]);
        if (!mbar_wait_bounded(&bars[0], 0)) atomicAdd(err, 1u);
        tc::tc_fence_after();
        const uint64_t base = tc::make_smem_desc_base(tc::kLBO, tc::kSBO);
        const uint32_t idesc = idesc_any(true, false, 128, 128, true);          // e4m3 x e4m3 -> F16, negate A
        for (int ks = 0; ks < tc::kKSteps; ++ks) {
            const uint64_t da = tc::smem_desc(base, tc::smem_u32(sA) + ks * 2 * tc::kChunkStride);
            const uint64_t db = tc::smem_desc(base, tc::smem_u32(sB) + ks * 2 * tc::kChunkStride);
            tc::mma_f8(tmem, da, db, idesc, ks > 0);
        }
        tc::mma_commit(&bars[1]);
    }
    if (!mbar_wait_bounded(&bars[1], 0)) atomicAdd(err, 1u);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
        for (int k = 0; k < 16; ++k) out_plain[(size_t)tid * 128 + c0 + k] = v[k];
    }
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[16];
        tmem_ld16_pack(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
        for (int k = 0; k < 16; ++k) out_pack[(size_t)tid * 64 + c0 / 2 + k] = v[k];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------------------
// 2. MMA peak
// ------------------------------------------------------------------------------------------------------------
template <bool FP8, bool DF32, int N>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(const uint8_t* src, int tiles, long long* cycles, uint32_t* err) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                   // 128 rows x 18 chunks
    uint8_t* sB = smem + tc::kAOperandBytes;              // N rows x 18 chunks (chunk stride N * 16)
    constexpr uint32_t kBBytes = (N / 128) * tc::kAOperandBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tc::kAOperandBytes + kBBytes);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 3; ++i) tc::mbar_init(&bars[i], 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(tslot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tslot;
    if (tid == 0) {
        tc::mbar_expect_tx(&bars[2], tc::kAOperandBytes + kBBytes);
        tc::bulk_g2s(sA, src, tc::kAOperandBytes, &bars[2]);
        for (int b = 0; b < N / 128; ++b) tc::bulk_g2s(sB + b * tc::kAOperandBytes, src + (b + 1) * tc::kAOperandBytes, tc::kAOperandBytes, &bars[2]);
        bool ok = mbar_wait_bounded(&bars[2], 0);
        tc::tc_fence_after();
        const uint64_t abase = tc::make_smem_desc_base(tc::kLBO, tc::kSBO);
        const uint64_t bbase = tc::make_smem_desc_base(N * 16, tc::kSBO);
        const uint32_t idesc = idesc_any(FP8, DF32, 128, N, true);
        const long long t0 = clock64();
        for (int t = 0; t < tiles && ok; ++t) {
            const int st = t & 1, use = t >> 1;
            if (use > 0) ok = mbar_wait_bounded(&bars[st], (use - 1) & 1);
            tc::tc_fence_after();
            const uint32_t d = tmem + st * 256;
#pragma unroll
            for (int ks = 0; ks < tc::kKSteps; ++ks) {
                const uint64_t da = tc::smem_desc(abase, tc::smem_u32(sA) + ks * 2 * tc::kChunkStride);
                const uint64_t db = tc::smem_desc(bbase, tc::smem_u32(sB) + ks * 2 * (N * 16));
                if (FP8) tc::mma_f8(d, da, db, idesc, ks > 0); else tc::mma_bf16(d, da, db, idesc, ks > 0);
            }
            tc::mma_commit(&bars[st]);
        }
        for (int t = tiles - 2; t < tiles && ok; ++t) if (t >= 0) ok = mbar_wait_bounded(&bars[t & 1], (t >> 1) & 1);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
        if (!ok) atomicAdd(err, 1u);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------
// 3. TMEM read rate: 16 warps (4 per lane quadrant), each reading `cols` columns per iteration
// ------------------------------------------------------------------------------------------------------------
template <int MODE>   // 0: 2 x x16 plain (32 cols), 1: x32 plain (32 cols), 2: 2 x x16.pack (64 cols), 3: x32.pack (64 cols)
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tslot, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tslot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (uint32_t)((it * 64 + (warp >> 2) * 128) & 511);
        if (MODE == 0) {
            uint32_t a[16], b[16];
            tc::tmem_ld16(tmem + (col & 511), a); tc::tmem_ld16(tmem + ((col + 16) & 511), b);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += a[k] ^ b[k];
        } else if (MODE == 1) {
            uint32_t a[32];
            tmem_ld32(tmem + (col & 511), a);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += a[k];
        } else if (MODE == 2) {
            uint32_t a[16], b[16];
            tmem_ld16_pack(tmem + (col & 511), a); tmem_ld16_pack(tmem + ((col + 32) & 511), b);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) acc += a[k] ^ b[k];
        } else {
            uint32_t a[32];
            tmem_ld32_pack(tmem + (col & 511), a);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += a[k];
        }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tslot, 512);
}

static uint16_t half_bits_of_halfint(int twice) {      // value = twice / 2, exact
    if (twice == 0) return 0;
    int e = 0; double m = twice * 0.5;
    while (m >= 2.0) { m *= 0.5; ++e; }
    while (m < 1.0) { m *= 2.0; --e; }
    return (uint16_t)(((e + 15) << 10) | (int)((m - 1.0) * 1024.0 + 0.5));
}

template <bool FP8, bool DF32, int N>
int run_peak(const char* name, const uint8_t* src, long long* cyc, uint32_t* err, int nsm, double clock_hz) {
    const size_t smem = tc::kAOperandBytes * (1 + N / 128) + 128;
    CK(cudaFuncSetAttribute(mma_peak_kernel<FP8, DF32, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = 20000;
    mma_peak_kernel<FP8, DF32, N><<<nsm, 128, smem>>>(src, 2000, cyc, err);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_peak_kernel<FP8, DF32, N><<<nsm, 128, smem>>>(src, tiles, cyc, err);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
    uint32_t he = 0; CK(cudaMemcpy(&he, err, 4, cudaMemcpyDeviceToHost));
    const double k_per = FP8 ? 32.0 : 16.0;
    const double flops = 2.0 * 128 * N * k_per * tc::kKSteps * (double)tiles * nsm;
    printf("{\"mma\": \"%s\", \"M\": 128, \"N\": %d, \"k_steps\": %d, \"tiles_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.1f, \"cycles_per_mma\": %.1f, "
           "\"implied_mhz\": %.0f, \"timeouts\": %u}\n",
           name, N, tc::kKSteps, tiles, ms, flops / (ms * 1e-3) / 1e12, (double)mx / ((double)tiles * tc::kKSteps), (double)mx / (ms * 1e3), he);
    (void)clock_hz;
    return 0;
}

template <int MODE>
int run_ld(const char* name, int cols, long long* cyc, uint32_t* sink, int nsm) {
    const int iters = 20000;
    tmem_ld_kernel<MODE><<<nsm, 512>>>(200, cyc, sink);
    CK(cudaDeviceSynchronize());
    tmem_ld_kernel<MODE><<<nsm, 512>>>(iters, cyc, sink);
    CK(cudaDeviceSynchronize());
    long long h[256]; CK(cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < nsm; ++i) mx = h[i] > mx ? h[i] : mx;
    // per iteration all 16 warps read `cols` columns x 32 lanes
    printf("{\"tmem_ld\": \"%s\", \"columns_per_warp_iter\": %d, \"cycles_per_iter\": %.1f, \"tmem_cells_per_clk_per_sm\": %.1f}\n", name, cols,
           (double)mx / iters, 16.0 * 32 * cols * iters / (double)mx);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
    uint32_t* err; CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));

    // ---- 1. exactness of the F16 accumulator on an ORB tile ----
    {
        const int rows_a = 121, rows_b = 128;
        std::vector<uint8_t> a(128 * 32), b(128 * 32);
        srand(7);
        for (auto& x : a) x = (uint8_t)(rand() & 0xFF);
        for (auto& x : b) x = (uint8_t)(rand() & 0xFF);
        memcpy(&b[5 * 32], &a[7 * 32], 32);                                 // duplicate: distance 0
        memset(&a[9 * 32], 0x00, 32); memset(&b[11 * 32], 0xFF, 32);       // distance 256
        for (int k = 0; k < 32; ++k) b[13 * 32 + k] = a[3 * 32 + k] ^ (k == 0 ? 1 : 0);   // distance 1 (D = 0.5)
        uint8_t *da, *db, *ta, *tb; uint32_t *o1, *o2;
        CK(cudaMalloc(&da, a.size())); CK(cudaMalloc(&db, b.size())); CK(cudaMalloc(&ta, tc::kBlockBytes)); CK(cudaMalloc(&tb, tc::kBlockBytes));
        CK(cudaMalloc(&o1, 128 * 128 * 4)); CK(cudaMalloc(&o2, 128 * 64 * 4));
        CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
        tcm::orb_tc_prep_kernel<<<16, 256>>>(da, rows_a, ta, 1);
        tcm::orb_tc_prep_kernel<<<16, 256>>>(db, rows_b, tb, 1);
        const size_t smem = 2 * tc::kAOperandBytes + 64;
        CK(cudaFuncSetAttribute(f16_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        f16_tile_kernel<<<1, 128, smem>>>(ta, tb, o1, o2, err);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h1(128 * 128), h2(128 * 64);
        CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost));
        int bad_plain = 0, bad_pack = 0, bad_hi = 0, bad_pad = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 128; ++j) {
                int ham = 0;
                for (int k = 0; k < 32; ++k) ham += __builtin_popcount((unsigned)(a[i * 32 + k] ^ b[j * 32 + k]));
                const uint32_t pl = h1[i * 128 + j];
                const uint32_t pk = (h2[i * 64 + j / 2] >> (16 * (j & 1))) & 0xFFFFu;
                if (i >= rows_a) {            // padding row: n/2 = 300, so D = 300 + popcount(b_j)/2 (a is zero)
                    int m = 0; for (int k = 0; k < 32; ++k) m += __builtin_popcount((unsigned)b[j * 32 + k]);
                    const uint16_t want = half_bits_of_halfint(600 + m);
                    bad_pad += ((pl & 0xFFFFu) != want) || (pk != want);
                    continue;
                }
                const uint16_t want = half_bits_of_halfint(ham);
                bad_plain += (pl & 0xFFFFu) != want;
                bad_hi += (pl >> 16) != 0;
                bad_pack += pk != want;
            }
        uint32_t he = 0; CK(cudaMemcpy(&he, err, 4, cudaMemcpyDeviceToHost));
        printf("{\"test\": \"f16_accumulator_orb_tile\", \"bad_plain_low16\": %d, \"plain_high16_nonzero\": %d, \"bad_pack16b\": %d, \"bad_padding\": %d, "
               "\"timeouts\": %u, \"samples_plain\": [\"%08x\", \"%08x\", \"%08x\"], \"samples_pack\": [\"%08x\", \"%08x\"], \"dup\": \"%08x\", \"d256\": \"%08x\", \"d1\": \"%08x\"}\n",
               bad_plain, bad_hi, bad_pack, bad_pad, he, h1[0], h1[1], h1[128], h2[0], h2[64], h1[7 * 128 + 5], h1[9 * 128 + 11], h1[3 * 128 + 13]);
    }

    // ---- 2. MMA peaks ----
    {
        const size_t bytes = 3 * (size_t)tc::kAOperandBytes;
        std::vector<uint8_t> h(bytes);
        srand(3);
        for (auto& x : h) x = (rand() & 1) ? 0x38 : 0x00;        // e4m3 0/1; as bf16 pairs 0x3838 / 0x0038 ...: finite small values
        uint8_t* src; CK(cudaMalloc(&src, bytes)); CK(cudaMemcpy(src, h.data(), bytes, cudaMemcpyHostToDevice));
        long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * 256));
        run_peak<false, true, 128>("bf16->f32 (UTCHMMA)", src, cyc, err, nsm, p.clockRate * 1e3);
        run_peak<false, true, 256>("bf16->f32 (UTCHMMA)", src, cyc, err, nsm, p.clockRate * 1e3);
        run_peak<true, true, 128>("e4m3->f32 (UTCQMMA)", src, cyc, err, nsm, p.clockRate * 1e3);
        run_peak<true, true, 256>("e4m3->f32 (UTCQMMA)", src, cyc, err, nsm, p.clockRate * 1e3);
        run_peak<true, false, 128>("e4m3->f16 (UTCQMMA)", src, cyc, err, nsm, p.clockRate * 1e3);
        run_peak<true, false, 256>("e4m3->f16 (UTCQMMA)", src, cyc, err, nsm, p.clockRate * 1e3);
    }

    // ---- 3. TMEM read rate ----
    {
        long long* cyc; uint32_t* sink;
        CK(cudaMalloc(&cyc, sizeof(long long) * 256)); CK(cudaMalloc(&sink, sizeof(uint32_t) * nsm * 512));
        run_ld<0>("2 x 32x32b.x16", 32, cyc, sink, nsm);
        run_ld<1>("32x32b.x32", 32, cyc, sink, nsm);
        run_ld<2>("2 x 32x32b.x16.pack::16b", 64, cyc, sink, nsm);
        run_ld<3>("32x32b.x32.pack::16b", 64, cyc, sink, nsm);
    }
    return 0;
}
