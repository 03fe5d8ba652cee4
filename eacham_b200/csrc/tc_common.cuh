// tc_common.cuh -- raw PTX wrappers for the Blackwell (sm_100a) pieces the SIFT scorer uses:
// mbarrier, 1-D bulk copy (TMA engine, SASS UBLKCP), tcgen05 alloc / mma / commit / ld / fences, and the
// shared-memory + instruction descriptors for tcgen05.mma kind::f16 with K-major, un-swizzled operands.
//
// Operand layout ("UMMA canonical K-major INTERLEAVE", in 16-byte units ((8,n),2):((1,SBO),LBO)):
//   a core matrix is 8 rows x 16 bytes (8 bf16 along K), stored as 128 contiguous bytes (row r at r*16);
//   SBO = byte distance between consecutive 8-row groups, LBO = byte distance between the two K-adjacent
//   core matrices one K=16 MMA step consumes.
// The descriptor tcgen05.mma takes is 64 bits:
//   [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) swizzle = 0
// The 32-bit instruction descriptor (kind::f16):
//   [4,6) D format (1 = F32) | [7,10) A format (1 = BF16) | [10,13) B format | 13 negate A | 14 negate B |
//   15 A major (0 = K) | 16 B major | [17,23) N >> 3 | [24,29) M >> 4
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace eacham {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done, addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One lane of a CONVERGED warp. Code that issues uniform-datapath instructions (tcgen05.mma / commit, bulk copies) must sit under
// this predicate inside warp-uniform control flow: under a plain `if (lane == 0)` the compiler cannot prove uniformity and wraps
// every such instruction in a waterfall loop (ELECT + BRA.U.ANY in SASS), which made each MMA cost ~80-110 issue cycles.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- tensor memory ----------------------------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) TMEM columns; the base address lands in *slot (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__host__ __device__ constexpr uint64_t make_smem_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t smem_desc(uint64_t base, uint32_t smem_addr) {
    return base | (uint64_t)((smem_addr >> 4) & 0x3FFF);
}
// a descriptor from its two words (low word: address field + LBO, high word: SBO + version; see above)
__device__ __forceinline__ uint64_t desc_from(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(uint32_t m, uint32_t n, bool negate_a) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((negate_a ? 1u : 0u) << 13) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 step, issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// FP8 (e4m3 x e4m3 -> f32) variant: one K=32 step (32 bytes per row, i.e. the same two 16-byte K-chunks as a bf16 K=16 step)
__host__ __device__ constexpr uint32_t make_idesc_e4m3(uint32_t m, uint32_t n, bool negate_a, bool d_f32) {
    return ((d_f32 ? 1u : 0u) << 4) | ((negate_a ? 1u : 0u) << 13) | ((n >> 3) << 17) | ((m >> 4) << 24);   // A/B format 0 = E4M3; D format 0 = F16, 1 = F32
}
__host__ __device__ constexpr uint32_t make_idesc_e4m3_f32(uint32_t m, uint32_t n, bool negate_a) { return make_idesc_e4m3(m, n, negate_a, true); }
__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 16 consecutive 32-bit columns: thread l of the warp receives TMEM lane (base lane + l).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- pre-tiled bf16 block layout shared by the prep kernel and the scorer ------------------------------------
// One block = 128 descriptor rows. 20 K-chunks of 8 bf16 (16 bytes) per row; chunk c of row r lives at
// c * 2048 + r * 16. Chunks 0..15: the 128 descriptor values. Chunks 16,17: augmentation used when the image is
// the A (first) operand: [-n/2 split in 3 bf16, 1, 1, 1, 0 x 10]. Chunks 18,19: augmentation used as the B (second)
// operand: [1, 1, 1, -m/2 split in 3 bf16, 0 x 10]. With negate-A the MMA then yields
//   D = -a.b + n/2 + m/2 = |a - b|^2 / 2      (n = |a|^2, m = |b|^2 of the bf16-rounded rows).
constexpr int kBlockRows = 128;
constexpr int kChunkStride = kBlockRows * 16;            // 2048 bytes between K-adjacent core matrices (LBO)
constexpr int kDataChunks = 16, kAugChunks = 2;
constexpr int kBlockBytes = (kDataChunks + 2 * kAugChunks) * kChunkStride;   // 40,960
constexpr int kDataBytes = kDataChunks * kChunkStride;                       // 32,768
constexpr int kAugBytes = kAugChunks * kChunkStride;                         // 4,096
constexpr int kAOperandBytes = kDataBytes + kAugBytes;                       // chunks 0..17 contiguous: 36,864
constexpr int kKSteps = 9;                                                   // 8 data + 1 augmentation, K = 16 each
constexpr uint32_t kLBO = kChunkStride, kSBO = 128;

}  // namespace tc
}  // namespace eacham
