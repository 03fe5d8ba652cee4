// Unit test for the tcgen05 plumbing (descriptors, layout, negate-A, TMEM load) used by the SIFT scorer:
// one CTA computes D = |a_i - b_j|^2 / 2 for a 128 x 128 tile from pre-tiled bf16 blocks and compares with the CPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I. -o tools/tc_unit_test tools/tc_unit_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../eacham_b200/csrc/tc_match_kernels.cuh"

using namespace eacham;

__global__ void __launch_bounds__(128, 1) tile_kernel(const uint8_t* a_blk, const uint8_t* b_blk, float* out, int negate) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;                                   // 36,864 (chunks 0..17)
    uint8_t* sB = smem + tc::kAOperandBytes;              // data 32,768 + B-aug 4,096
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
        tc::mbar_wait(&bars[0], 0);
        tc::tc_fence_after();
        const uint64_t base = tc::make_smem_desc_base(tc::kLBO, tc::kSBO);
        const uint32_t idesc = tc::make_idesc_bf16_f32(128, 128, negate != 0);
        for (int ks = 0; ks < tc::kKSteps; ++ks) {
            const uint64_t da = tc::smem_desc(base, tc::smem_u32(sA) + ks * 2 * tc::kChunkStride);
            const uint64_t db = tc::smem_desc(base, tc::smem_u32(sB) + ks * 2 * tc::kChunkStride);
            tc::mma_bf16(tmem, da, db, idesc, ks > 0);
        }
        tc::mma_commit(&bars[1]);
    }
    tc::mbar_wait(&bars[1], 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
        for (int k = 0; k < 16; ++k) out[(size_t)tid * 128 + c0 + k] = __uint_as_float(v[k]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

static float bf16_round(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
    r &= 0xFFFF0000u; float y; memcpy(&y, &r, 4); return y;
}

int main() {
    const int R = 128;
    std::vector<float> a(R * 128), b(R * 128);
    srand(1);
    for (int mode = 0; mode < 2; ++mode) {
        for (auto& x : a) x = mode == 0 ? (float)(rand() % 256) : (float)rand() / RAND_MAX * 0.2f;
        for (auto& x : b) x = mode == 0 ? (float)(rand() % 256) : (float)rand() / RAND_MAX * 0.2f;
        for (int k = 0; k < 128; ++k) b[5 * 128 + k] = a[7 * 128 + k];     // an exact duplicate: D must be 0
        float *da, *db, *dout; uint8_t *ta, *tb;
        cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, R * R * 4);
        cudaMalloc(&ta, tc::kBlockBytes); cudaMalloc(&tb, tc::kBlockBytes);
        cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
        const int rows_a = 120, rows_b = 128;      // 8 padding rows in A
        tcm::sift_prep_kernel<<<16, 256>>>(da, rows_a, ta, 1);
        tcm::sift_prep_kernel<<<16, 256>>>(db, rows_b, tb, 1);
        const size_t smem = 2 * tc::kAOperandBytes + 64;
        cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tile_kernel<<<1, 128, smem>>>(ta, tb, dout, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> out(R * R);
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double max_rel = 0, max_abs = 0; int bad_pad = 0;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < R; ++j) {
                if (i >= rows_a) { bad_pad += !(out[i * R + j] > 1e29f); continue; }
                double d = 0;
                for (int k = 0; k < 128; ++k) { double df = (double)bf16_round(a[i * 128 + k]) - (double)bf16_round(b[j * 128 + k]); d += df * df; }
                d *= 0.5;
                double err = fabs(out[i * R + j] - d);
                max_abs = fmax(max_abs, err); max_rel = fmax(max_rel, err / fmax(d, 1e-3));
            }
        printf("{\"test\": \"tc_tile\", \"mode\": \"%s\", \"max_abs_err\": %.6g, \"max_rel_err\": %.3g, \"dup_value\": %g, \"pad_rows_bad\": %d, \"sample\": [%g, %g, %g]}\n",
               mode == 0 ? "integer-valued" : "float", max_abs, max_rel, out[7 * R + 5], bad_pad, out[0], out[1], out[R]);
    }
    return 0;
}
