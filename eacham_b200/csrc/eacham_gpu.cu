// eacham_gpu.cu -- C ABI (include/eacham_gpu.h) over the sm_100a matching kernels.
//
// Host-side runtime for the one hot path of fatlipp/eacham this library replaces: descriptor arena + upload
// (Node::GetDescriptors feeding Match, /root/reference/apps/sfm/main.cpp:107-108), the reference-shaped
// single-direction Match (/root/reference/modules/base/features/FeatureMatcherFlann.cpp:14-30) and the batched
// pair loop (/root/reference/apps/sfm/main.cpp:84-147). No torch types, no exceptions across the boundary,
// no CPU fallback.
#include <algorithm>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <exception>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/eacham_gpu.h"
#include "l2_kernels.cuh"
#include "orb_kernels.cuh"
#include "tc_match_kernels.cuh"
#include "tc_orb_kernels.cuh"
#include "tc_sift_kernels.cuh"
#include "verify_kernels.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            return fail(_e == cudaErrorMemoryAllocation ? EACHAM_ERR_OUT_OF_MEMORY : EACHAM_ERR_CUDA,       \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);        \
    } while (0)

constexpr size_t kAlign = 128;
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline size_t row_bytes(int kind) { return kind == EACHAM_KIND_ORB256 ? 32 : 512; }

struct ImageHost {
    int kind = -1;
    uint32_t rows = 0;
    size_t offset = 0;     // byte offset in staging == arena
    bool present = false;
    bool has_data = false;   // bytes staged by set_descriptors
    bool reserved = false;   // shape declared by reserve: bytes arrive in the device arena by broadcast
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    int ensure(size_t n) {
        if (n <= cap) return EACHAM_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)16);
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return fail(EACHAM_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e)); }
        cap = want;
        return EACHAM_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct eacham_gpu_handle {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;

    std::vector<ImageHost> images;
    uint8_t* staging = nullptr;   // pinned
    size_t staging_cap = 0, staging_used = 0, staging_waste = 0;
    bool committed = false;
    bool any_data = false;

    DevBuf<uint8_t> arena;
    size_t arena_bytes = 0;
    DevBuf<eacham::orb::ImageDesc> d_images;
    uint32_t max_rows[2] = {0, 0};
    // SIFT tensor-core path: pre-tiled bf16 copy of every F32X128 image, built lazily after commit (+ broadcast)
    DevBuf<uint8_t> tc_arena;
    DevBuf<eacham::tcm::ImageDescTc> d_images_tc;
    DevBuf<uint32_t> d_block_start;
    std::vector<size_t> tc_offsets;
    size_t tc_bytes = 0;
    bool tc_dirty = true;
    DevBuf<uint8_t> tc_scratch;
    uint32_t cfg_flags = 0;

    DevBuf<eacham_pair_t> d_pairs;
    DevBuf<uint32_t> d_order;
    DevBuf<int32_t> d_dbg_idx;                // debug kNN of one pair: [rows12 * 2 | rows21 * 2]
    DevBuf<float> d_dbg_dist;
    bool dbg_on = false;
    std::vector<uint32_t> order_host, order_count;
    DevBuf<eacham_pair_result_t> d_results;
    DevBuf<eacham_match_t> d_matches;
    DevBuf<uint32_t> d_counter;               // [0] work counter
    DevBuf<unsigned long long> d_cursor;      // [0] match cursor
    size_t last_n_pairs = 0;
    unsigned long long last_total = 0;
    uint64_t cfg_match_entries = 0;

    // single-pair scratch
    DevBuf<uint8_t> d_q, d_t;
    DevBuf<unsigned long long> d_partial;
    DevBuf<int32_t> d_idx;
    DevBuf<float> d_dist;
    DevBuf<uint32_t> d_match, d_match2;
    DevBuf<uint8_t> d_flush;

    cudaEvent_t ev[8] = {};
    eacham_gpu_timing timing = {};

    // ---- geometric verification (verify_kernels.cuh) ----
    std::vector<std::vector<float>> kp_host;  // per image: x, y interleaved
    bool kp_dirty = true;
    DevBuf<float2> d_kp;
    DevBuf<unsigned long long> d_kp_offset;
    DevBuf<double> d_hyps;
    DevBuf<eacham_verify_result> d_vres;
    DevBuf<float> d_vmed;
    DevBuf<uint8_t> d_vmask;
    uint32_t last_max_first = 0, last_max_second = 0;

    // ---- the per-call Match() route on the tensor-core engines (match_single.cuh) ----
    struct CacheEntry {                       // one image's descriptors, resident on the device across calls
        const void* host = nullptr; uint32_t rows = 0; size_t stride = 0; int kind = -1; uint64_t hash = 0;
        DevBuf<uint8_t> raw, tc;
        int refs = 0; bool ready = false, loading = false; uint64_t tick = 0;
    };
    struct CallSlot {                         // what one in-flight Match() call needs privately
        cudaStream_t stream = nullptr;
        DevBuf<uint32_t> d_out, d_counter; DevBuf<eacham_pair_t> d_pair;
        uint32_t* h_out = nullptr; size_t h_out_cap = 0; eacham_pair_t* h_pair = nullptr;
        bool busy = false;
    };
    static constexpr int kCacheEntries = 128, kCallSlots = 8;
    std::mutex cache_mu;
    std::condition_variable cache_cv;
    std::vector<CacheEntry> cache;
    DevBuf<eacham::tcm::ImageDescTc> d_cache_descs;      // entry i's image record (absolute device addresses)
    std::vector<CallSlot> slots;
    uint64_t cache_tick = 0, cache_hits = 0, cache_misses = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ensure_staging(eacham_gpu_handle* h, size_t need) {
    if (need <= h->staging_cap) return EACHAM_OK;
    size_t cap = std::max(need, std::max(h->staging_cap * 2, (size_t)1 << 20));
    uint8_t* n = nullptr;
    CUDA_TRY(cudaMallocHost(&n, cap));
    if (h->staging) { memcpy(n, h->staging, std::min(h->staging_used, h->staging_cap)); cudaFreeHost(h->staging); }
    h->staging = n; h->staging_cap = cap;
    return EACHAM_OK;
}

// Gives image_id a region of the arena layout. Only set_descriptors backs the layout with pinned staging memory
// (`backed`); reserve just computes offsets. A region whose image changes size is abandoned and reclaimed by the
// compaction in commit.
int place_image(eacham_gpu_handle* h, uint32_t image_id, int kind, uint32_t rows, bool backed) {
    if (kind != EACHAM_KIND_ORB256 && kind != EACHAM_KIND_F32X128) return fail(EACHAM_ERR_INVALID_ARG, "unknown descriptor kind %d", kind);
    if (rows > 65535u) return fail(EACHAM_ERR_TOO_LARGE, "image %u has %u descriptors; at most 65535 are supported", image_id, rows);
    if (image_id > (1u << 26)) return fail(EACHAM_ERR_INVALID_ARG, "image id %u out of range", image_id);
    const bool known = image_id < h->images.size() && h->images[image_id].present;
    const bool reuse = known && h->images[image_id].kind == kind && h->images[image_id].rows == rows;
    size_t offset = reuse ? h->images[image_id].offset : align_up(h->staging_used, kAlign);
    const size_t end = offset + (size_t)rows * row_bytes(kind);
    if (backed) {
        int rc = ensure_staging(h, align_up(std::max(end, h->staging_used), kAlign));
        if (rc) return rc;
    }
    if (image_id >= h->images.size()) h->images.resize((size_t)image_id + 1);
    ImageHost& im = h->images[image_id];
    if (!reuse) {
        if (known) h->staging_waste += align_up((size_t)im.rows * row_bytes(im.kind), kAlign);
        im.offset = offset;
        h->staging_used = end;
        im.kind = kind; im.rows = rows; im.present = true; im.has_data = false; im.reserved = false;
    }
    h->committed = false;
    return EACHAM_OK;
}

// Re-packs the staged images in id order so that abandoned regions (images re-set with another row count) do not
// accumulate; only possible while every image's bytes are still in staging (no reserved images).
void compact_staging(eacham_gpu_handle* h) {
    if (h->staging_waste == 0) return;
    for (const ImageHost& im : h->images) if (im.present && im.reserved) return;
    std::vector<uint32_t> order;
    for (uint32_t i = 0; i < h->images.size(); ++i) if (h->images[i].present) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h->images[a].offset < h->images[b].offset; });
    size_t used = 0;
    for (uint32_t i : order) {
        ImageHost& im = h->images[i];
        const size_t off = align_up(used, kAlign), bytes = (size_t)im.rows * row_bytes(im.kind);
        if (off != im.offset && im.has_data && bytes) memmove(h->staging + off, h->staging + im.offset, bytes);
        im.offset = off;
        used = off + bytes;
    }
    h->staging_used = used;
    h->staging_waste = 0;
}

}  // namespace

extern "C" {

int eacham_gpu_abi_version(void) { return EACHAM_GPU_ABI_VERSION; }

const char* eacham_gpu_last_error(void) { return g_last_error.c_str(); }

int eacham_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void eacham_gpu_default_opts(eacham_match_opts* o) {
    if (!o) return;
    o->ratio = 0.8; o->min_dir = 30; o->min_mutual = 30; o->cross_check = 1; o->emit_all = 0;
}

int eacham_gpu_create(const eacham_gpu_config* cfg, eacham_gpu_handle** out) {
    if (!out) return fail(EACHAM_ERR_INVALID_ARG, "out handle pointer is null");
    *out = nullptr;
    int dev = cfg ? cfg->device : 0;
    int n = eacham_gpu_device_count();
    if (n <= 0) return fail(EACHAM_ERR_NO_DEVICE, "no CUDA device visible; libeacham_gpu has no CPU fallback");
    if (dev < 0 || dev >= n) return fail(EACHAM_ERR_INVALID_ARG, "device %d out of range (0..%d)", dev, n - 1);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(EACHAM_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    eacham_gpu_handle* h = new (std::nothrow) eacham_gpu_handle();
    if (!h) return fail(EACHAM_ERR_OUT_OF_MEMORY, "host allocation failed");
    h->device = dev;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->cfg_match_entries = cfg ? cfg->match_buffer_entries : 0;
    h->cfg_flags = cfg ? cfg->flags : 0;
    DeviceGuard g(dev);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
    if (e != cudaSuccess) { eacham_gpu_destroy(h); return fail(EACHAM_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e)); }
    if (h->d_counter.ensure(16) || h->d_cursor.ensure(16)) { eacham_gpu_destroy(h); return EACHAM_ERR_OUT_OF_MEMORY; }
    *out = h;
    return EACHAM_OK;
}

void eacham_gpu_destroy(eacham_gpu_handle* h) {
    if (!h) return;
    {
        DeviceGuard g(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        h->tc_arena.release(); h->d_images_tc.release(); h->d_block_start.release(); h->tc_scratch.release();
        h->arena.release(); h->d_images.release(); h->d_pairs.release(); h->d_order.release(); h->d_dbg_idx.release(); h->d_dbg_dist.release();
        h->d_kp.release(); h->d_kp_offset.release(); h->d_hyps.release(); h->d_vres.release(); h->d_vmed.release(); h->d_vmask.release(); h->d_results.release(); h->d_matches.release();
        h->d_counter.release(); h->d_cursor.release(); h->d_q.release(); h->d_t.release(); h->d_partial.release();
        h->d_idx.release(); h->d_dist.release(); h->d_match.release(); h->d_match2.release(); h->d_flush.release();
        if (h->staging) cudaFreeHost(h->staging);
        for (auto& c : h->cache) { c.raw.release(); c.tc.release(); }
        h->d_cache_descs.release();
        for (auto& sl : h->slots) {
            if (sl.stream) { cudaStreamSynchronize(sl.stream); cudaStreamDestroy(sl.stream); }
            sl.d_out.release(); sl.d_counter.release(); sl.d_pair.release();
            if (sl.h_out) cudaFreeHost(sl.h_out);
            if (sl.h_pair) cudaFreeHost(sl.h_pair);
        }
        for (auto& e : h->ev) if (e) cudaEventDestroy(e);
        if (h->stream) cudaStreamDestroy(h->stream);
    }
    delete h;
}

int eacham_gpu_set_descriptors(eacham_gpu_handle* h, uint32_t image_id, int kind, const void* data, uint32_t rows,
                               size_t row_stride_bytes) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    if (kind != EACHAM_KIND_ORB256 && kind != EACHAM_KIND_F32X128) return fail(EACHAM_ERR_INVALID_ARG, "unknown descriptor kind %d", kind);
    if (rows > 0 && !data) return fail(EACHAM_ERR_INVALID_ARG, "null descriptor pointer for image %u", image_id);
    const size_t rb = row_bytes(kind);
    if (rows > 0 && row_stride_bytes < rb) return fail(EACHAM_ERR_INVALID_ARG, "row stride %zu < row size %zu", row_stride_bytes, rb);
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = place_image(h, image_id, kind, rows, /*backed=*/true);       // nothing is modified when this fails
    if (rc) return rc;
    ImageHost& im = h->images[image_id];
    uint8_t* dst = h->staging + im.offset;
    if (row_stride_bytes == rb) memcpy(dst, data, (size_t)rows * rb);
    else for (uint32_t r = 0; r < rows; ++r) memcpy(dst + (size_t)r * rb, (const uint8_t*)data + (size_t)r * row_stride_bytes, rb);
    im.has_data = true; im.reserved = false;
    h->any_data = true;
    return EACHAM_OK;
}

// Many images in one call (image ids first_id .. first_id + n - 1): arguments are validated up front, the staging buffer grows once, and
// the copies into it run on a few host threads (a single memcpy stream fills pinned memory at ~10 GB/s, far below the PCIe link it feeds).
int eacham_gpu_set_descriptors_batch(eacham_gpu_handle* h, uint32_t first_id, uint32_t n, int kind, const void* const* data, const uint32_t* rows,
                                     const size_t* row_stride_bytes) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    if (kind != EACHAM_KIND_ORB256 && kind != EACHAM_KIND_F32X128) return fail(EACHAM_ERR_INVALID_ARG, "unknown descriptor kind %d", kind);
    if (n == 0) return EACHAM_OK;
    if (!data || !rows) return fail(EACHAM_ERR_INVALID_ARG, "null array");
    if ((uint64_t)first_id + n > (1u << 26)) return fail(EACHAM_ERR_INVALID_ARG, "image ids %u .. +%u out of range", first_id, n);
    const size_t rb = row_bytes(kind);
    size_t total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (rows[i] > 65535u) return fail(EACHAM_ERR_TOO_LARGE, "image %u has %u descriptors; at most 65535 are supported", first_id + i, rows[i]);
        if (rows[i] > 0 && !data[i]) return fail(EACHAM_ERR_INVALID_ARG, "null descriptor pointer for image %u", first_id + i);
        if (rows[i] > 0 && row_stride_bytes && row_stride_bytes[i] < rb)
            return fail(EACHAM_ERR_INVALID_ARG, "image %u: row stride %zu < row size %zu", first_id + i, row_stride_bytes[i], rb);
        total += align_up((size_t)rows[i] * rb, kAlign);
    }
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = ensure_staging(h, align_up(h->staging_used, kAlign) + total + kAlign);      // one growth instead of one per image
    if (rc) return rc;
    for (uint32_t i = 0; i < n; ++i)
        if ((rc = place_image(h, first_id + i, kind, rows[i], /*backed=*/true))) return rc;
    auto copy_range = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t i = lo; i < hi; ++i) {
            const ImageHost& im = h->images[first_id + i];
            uint8_t* dst = h->staging + im.offset;
            const size_t stride = row_stride_bytes ? row_stride_bytes[i] : rb;
            if (stride == rb) memcpy(dst, data[i], (size_t)rows[i] * rb);
            else for (uint32_t r = 0; r < rows[i]; ++r) memcpy(dst + (size_t)r * rb, (const uint8_t*)data[i] + (size_t)r * stride, rb);
        }
    };
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned workers = (unsigned)std::min<size_t>(std::min(hw, 8u), std::max<size_t>(1, total >> 22));      // one per 4 MiB, at most 8
    if (workers <= 1) copy_range(0, n);
    else {
        // contiguous ranges of about equal bytes; a range whose thread cannot be started is copied by the caller
        std::vector<std::thread> pool;
        uint32_t lo = 0;
        size_t acc = 0, share = (total + workers - 1) / workers;
        for (uint32_t i = 0; i < n; ++i) {
            acc += align_up((size_t)rows[i] * rb, kAlign);
            if (acc >= share || i + 1 == n) {
                try { pool.emplace_back(copy_range, lo, i + 1); }
                catch (const std::exception&) { copy_range(lo, i + 1); }
                lo = i + 1; acc = 0;
            }
        }
        for (std::thread& t : pool) t.join();
    }
    for (uint32_t i = 0; i < n; ++i) { ImageHost& im = h->images[first_id + i]; im.has_data = true; im.reserved = false; }
    h->any_data = true;
    return EACHAM_OK;
}

int eacham_gpu_reserve(eacham_gpu_handle* h, uint32_t image_id, int kind, uint32_t rows) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = place_image(h, image_id, kind, rows, /*backed=*/false);
    if (rc) return rc;
    h->images[image_id].reserved = true; h->images[image_id].has_data = false;
    return EACHAM_OK;
}

int eacham_gpu_clear(eacham_gpu_handle* h) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->images.clear();
    h->staging_used = 0; h->staging_waste = 0; h->committed = false; h->any_data = false; h->arena_bytes = 0;
    h->max_rows[0] = h->max_rows[1] = 0;
    h->kp_host.clear(); h->kp_dirty = true;
    return EACHAM_OK;
}

int eacham_gpu_commit(eacham_gpu_handle* h) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    compact_staging(h);
    const size_t bytes = align_up(std::max(h->staging_used, (size_t)1), kAlign);
    int rc = h->arena.ensure(bytes + kAlign);
    if (rc) return rc;
    if ((rc = h->d_images.ensure(std::max(h->images.size(), (size_t)1)))) return rc;
    std::vector<eacham::orb::ImageDesc> table(h->images.size());
    h->max_rows[0] = h->max_rows[1] = 0;
    for (size_t i = 0; i < h->images.size(); ++i) {
        const ImageHost& im = h->images[i];
        table[i].offset = im.offset;
        table[i].rows = im.present ? im.rows : 0;
        table[i].kind = im.present ? (uint32_t)im.kind : 0xffffffffu;
        if (im.present) h->max_rows[im.kind] = std::max(h->max_rows[im.kind], im.rows);
    }
    h->tc_offsets.assign(h->images.size(), 0);
    h->tc_bytes = 0;
    for (size_t i = 0; i < h->images.size(); ++i) {
        const ImageHost& im = h->images[i];
        if (im.present && (im.kind == EACHAM_KIND_F32X128 || !(h->cfg_flags & EACHAM_CFG_ORB_POPC))) {
            h->tc_offsets[i] = h->tc_bytes;
            h->tc_bytes += (size_t)((im.rows + 127) / 128) * eacham::tc::kBlockBytes;
        }
    }
    h->tc_dirty = true;
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    // one H2D copy of everything staged (ranges of reserved images carry no host bytes: a broadcast fills them)
    const size_t staged = std::min(h->staging_used, h->staging_cap);
    if (h->any_data && staged > 0)
        CUDA_TRY(cudaMemcpyAsync(h->arena.p, h->staging, staged, cudaMemcpyHostToDevice, h->stream));
    if (!table.empty())
        CUDA_TRY(cudaMemcpyAsync(h->d_images.p, table.data(), table.size() * sizeof(table[0]), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaEventElapsedTime(&h->timing.upload_ms, h->ev[0], h->ev[1]));
    h->arena_bytes = bytes;
    h->committed = true;
    return EACHAM_OK;
}

int eacham_gpu_arena(eacham_gpu_handle* h, void** device_ptr, size_t* bytes) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->committed) return fail(EACHAM_ERR_NOT_COMMITTED, "arena queried before eacham_gpu_commit");
    if (device_ptr) *device_ptr = h->arena.p;
    if (bytes) *bytes = h->arena_bytes;
    return EACHAM_OK;
}

int eacham_gpu_image_info(eacham_gpu_handle* h, uint32_t image_id, int* kind, uint32_t* rows, size_t* arena_offset) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (image_id >= h->images.size() || !h->images[image_id].present) return fail(EACHAM_ERR_INVALID_ARG, "image %u has no descriptors", image_id);
    const ImageHost& im = h->images[image_id];
    if (kind) *kind = im.kind;
    if (rows) *rows = im.rows;
    if (arena_offset) *arena_offset = im.offset;
    return EACHAM_OK;
}

int eacham_gpu_last_timing(eacham_gpu_handle* h, eacham_gpu_timing* t) {
    if (!h || !t) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    *t = h->timing;
    return EACHAM_OK;
}

int eacham_gpu_flush_l2(eacham_gpu_handle* h, size_t bytes) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = h->d_flush.ensure(bytes);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_flush.p, 0x5a, bytes, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return EACHAM_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------------------
// single pair, single direction
// -------------------------------------------------------------------------------------------------------------
namespace {

bool match_single_eligible(eacham_gpu_handle* h, int kind, double ratio, size_t q_stride, size_t t_stride);
int match_single_tc(eacham_gpu_handle* h, int kind, const void* query, uint32_t q_rows, size_t q_stride, const void* train, uint32_t t_rows,
                    size_t t_stride, double ratio, eacham_match_t* out, size_t cap, size_t* n_out);

// kNN-2 of q (device, nq rows) in t (device, nt rows); outputs on device. Any of idx/dist, match may be null.
int knn2_device(eacham_gpu_handle* h, int kind, const void* dq, uint32_t nq, const void* dt, uint32_t nt, double ratio,
                int32_t* d_idx, float* d_dist, uint32_t* d_match, uint32_t* launches) {
    if (nq == 0) return EACHAM_OK;
    using namespace eacham;
    if (kind == EACHAM_KIND_ORB256) {
        const uint32_t row_blocks = (nq + orb::kKnnRowBlock - 1) / orb::kKnnRowBlock;
        uint32_t want = std::max(1u, (uint32_t)(2 * h->sm_count) / row_blocks);
        uint32_t tiles = std::max(1u, (nt + orb::kKnnTile - 1) / orb::kKnnTile);
        uint32_t nsplit = std::min(want, tiles);
        uint32_t cols_per_split = ((tiles + nsplit - 1) / nsplit) * orb::kKnnTile;
        nsplit = std::max(1u, (nt + cols_per_split - 1) / cols_per_split);
        int rc = h->d_partial.ensure((size_t)nsplit * nq);
        if (rc) return rc;
        orb::orb_knn2_partial_kernel<<<dim3(row_blocks, nsplit), orb::kKnnThreads, 0, h->stream>>>(
            (const uint8_t*)dq, nq, (const uint8_t*)dt, nt, cols_per_split, (uint2*)h->d_partial.p);
        orb::orb_knn2_finalize_kernel<<<(nq + 255) / 256, 256, 0, h->stream>>>((const uint2*)h->d_partial.p, nq, nsplit, ratio,
                                                                                d_idx, d_dist, d_match);
    } else {
        const uint32_t row_blocks = (nq + l2::kBM - 1) / l2::kBM;
        uint32_t want = std::max(1u, (uint32_t)(2 * h->sm_count) / row_blocks);
        uint32_t tiles = std::max(1u, (nt + l2::kBN - 1) / l2::kBN);
        uint32_t nsplit = std::min(want, tiles);
        uint32_t cols_per_split = ((tiles + nsplit - 1) / nsplit) * l2::kBN;
        nsplit = std::max(1u, (nt + cols_per_split - 1) / cols_per_split);
        int rc = h->d_partial.ensure((size_t)nsplit * nq * 2);
        if (rc) return rc;
        l2::l2_knn2_partial_kernel<<<dim3(row_blocks, nsplit), l2::kThreads, 0, h->stream>>>(
            (const float*)dq, nq, (const float*)dt, nt, cols_per_split, h->d_partial.p);
        l2::l2_knn2_finalize_kernel<<<(nq + 255) / 256, 256, 0, h->stream>>>(h->d_partial.p, nq, nsplit, ratio, d_idx, d_dist,
                                                                              d_match);
    }
    if (launches) *launches += 2;
    CUDA_TRY(cudaGetLastError());
    return EACHAM_OK;
}

int upload_rows(eacham_gpu_handle* h, DevBuf<uint8_t>& dst, int kind, const void* src, uint32_t rows, size_t stride) {
    const size_t rb = row_bytes(kind);
    int rc = dst.ensure(std::max((size_t)rows * rb, (size_t)rb));
    if (rc) return rc;
    if (rows == 0) return EACHAM_OK;
    if (stride < rb) return fail(EACHAM_ERR_INVALID_ARG, "row stride %zu < row size %zu", stride, rb);
    CUDA_TRY(cudaMemcpy2DAsync(dst.p, rb, src, stride, rb, rows, cudaMemcpyHostToDevice, h->stream));
    return EACHAM_OK;
}

}  // namespace

extern "C" {

int eacham_gpu_knn2(eacham_gpu_handle* h, int kind, const void* query, uint32_t q_rows, size_t q_stride,
                    const void* train, uint32_t t_rows, size_t t_stride, int32_t* idx, float* dist) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    if (kind != EACHAM_KIND_ORB256 && kind != EACHAM_KIND_F32X128) return fail(EACHAM_ERR_INVALID_ARG, "unknown descriptor kind %d", kind);
    if ((q_rows && !query) || (t_rows && !train) || (q_rows && (!idx || !dist))) return fail(EACHAM_ERR_INVALID_ARG, "null pointer argument");
    if (q_rows > 65535u || t_rows > 65535u) return fail(EACHAM_ERR_TOO_LARGE, "at most 65535 descriptors per image");
    if (q_rows == 0) return EACHAM_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc;
    if ((rc = upload_rows(h, h->d_q, kind, query, q_rows, q_stride))) return rc;
    if ((rc = upload_rows(h, h->d_t, kind, train, t_rows, t_stride))) return rc;
    if ((rc = h->d_idx.ensure((size_t)q_rows * 2))) return rc;
    if ((rc = h->d_dist.ensure((size_t)q_rows * 2))) return rc;
    if ((rc = knn2_device(h, kind, h->d_q.p, q_rows, h->d_t.p, t_rows, 0.8, h->d_idx.p, h->d_dist.p, nullptr, nullptr))) return rc;
    CUDA_TRY(cudaMemcpyAsync(idx, h->d_idx.p, (size_t)q_rows * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dist, h->d_dist.p, (size_t)q_rows * 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return EACHAM_OK;
}

int eacham_gpu_match(eacham_gpu_handle* h, int kind, const void* query, uint32_t q_rows, size_t q_stride,
                     const void* train, uint32_t t_rows, size_t t_stride, double ratio, eacham_match_t* out, size_t cap,
                     size_t* n_out) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    if (kind != EACHAM_KIND_ORB256 && kind != EACHAM_KIND_F32X128) return fail(EACHAM_ERR_INVALID_ARG, "unknown descriptor kind %d", kind);
    if ((q_rows && !query) || (t_rows && !train) || !n_out || (cap && !out)) return fail(EACHAM_ERR_INVALID_ARG, "null pointer argument");
    if (q_rows > 65535u || t_rows > 65535u) return fail(EACHAM_ERR_TOO_LARGE, "at most 65535 descriptors per image");
    *n_out = 0;
    if (q_rows == 0 || t_rows == 0) return EACHAM_OK;
    if (match_single_eligible(h, kind, ratio, q_stride, t_stride))
        return match_single_tc(h, kind, query, q_rows, q_stride, train, t_rows, t_stride, ratio, out, cap, n_out);
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc;
    if ((rc = upload_rows(h, h->d_q, kind, query, q_rows, q_stride))) return rc;
    if ((rc = upload_rows(h, h->d_t, kind, train, t_rows, t_stride))) return rc;
    if ((rc = h->d_match.ensure(q_rows))) return rc;
    if ((rc = knn2_device(h, kind, h->d_q.p, q_rows, h->d_t.p, t_rows, ratio, nullptr, nullptr, h->d_match.p, nullptr))) return rc;
    std::vector<uint32_t> m(q_rows);
    CUDA_TRY(cudaMemcpyAsync(m.data(), h->d_match.p, (size_t)q_rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    size_t n = 0;
    for (uint32_t i = 0; i < q_rows; ++i)
        if (m[i] != EACHAM_NONE) {
            if (n < cap) { out[n].query = i; out[n].train = m[i]; }
            ++n;
        }
    *n_out = n;
    if (n > cap) return fail(EACHAM_ERR_BUFFER_TOO_SMALL, "match buffer holds %zu entries, %zu needed", cap, n);
    return EACHAM_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------------------
// batched pairs
// -------------------------------------------------------------------------------------------------------------
namespace {

// Builds the tensor-core copy (tc_common.cuh block layout) of every image that has one -- bf16 for F32X128, one e4m3 element
// per bit for ORB256 -- from the raw arena, in ONE launch over the image table. Runs lazily at the first match_pairs after a
// commit, i.e. after a possible NCCL broadcast has filled the arena on this device. Timed into timing.prep_ms.
int prepare_tc(eacham_gpu_handle* h) {
    using namespace eacham;
    if (!h->tc_dirty) return EACHAM_OK;
    int rc;
    if ((rc = h->tc_arena.ensure(std::max(h->tc_bytes, (size_t)tc::kBlockBytes)))) return rc;
    if ((rc = h->d_images_tc.ensure(std::max(h->images.size(), (size_t)1)))) return rc;
    if ((rc = h->d_block_start.ensure(h->images.size() + 1))) return rc;
    std::vector<tcm::ImageDescTc> table(h->images.size());
    std::vector<uint32_t> block_start(h->images.size() + 1, 0);
    uint32_t blocks = 0;
    for (size_t i = 0; i < h->images.size(); ++i) {
        const ImageHost& im = h->images[i];
        table[i].offset = im.offset; table[i].tc_offset = h->tc_offsets[i];
        table[i].rows = im.present ? im.rows : 0; table[i].kind = im.present ? (uint32_t)im.kind : 0xffffffffu;
        table[i].max_norm_bits = 0u; table[i].bf16_exact = 1u;            // refined by the prep kernel (F32X128)
        block_start[i] = blocks;
        const bool has_tc = im.present && im.rows > 0 && (im.kind == EACHAM_KIND_F32X128 || !(h->cfg_flags & EACHAM_CFG_ORB_POPC));
        if (has_tc) blocks += (im.rows + 127) / 128;
    }
    block_start[h->images.size()] = blocks;
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    if (!table.empty()) {
        CUDA_TRY(cudaMemcpyAsync(h->d_images_tc.p, table.data(), table.size() * sizeof(table[0]), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_block_start.p, block_start.data(), block_start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    }
    if (blocks > 0) {
        tcm::tc_prep_all_kernel<<<blocks * 16, 256, 0, h->stream>>>(h->arena.p, h->tc_arena.p, h->d_images_tc.p, h->d_block_start.p,
                                                                     (uint32_t)h->images.size());
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));               // the host tables above go out of scope
    CUDA_TRY(cudaEventElapsedTime(&h->timing.prep_ms, h->ev[0], h->ev[1]));
    h->tc_dirty = false;
    return EACHAM_OK;
}

int launch_pairs(eacham_gpu_handle* h, const eacham_pair_t* pairs, size_t n_pairs, const eacham_match_opts* opts_in) {
    using namespace eacham;
    if (!h->committed) return fail(EACHAM_ERR_NOT_COMMITTED, "eacham_gpu_match_pairs called before eacham_gpu_commit");
    if (n_pairs > 0xfffffff0ull) return fail(EACHAM_ERR_INVALID_ARG, "too many pairs");
    eacham_match_opts o;
    if (opts_in) o = *opts_in; else eacham_gpu_default_opts(&o);
    h->timing.kernel_launches = 0;
    h->timing.kernel_ms = h->timing.pairs_h2d_ms = h->timing.d2h_ms = 0.f;
    h->last_n_pairs = 0;                 // a failed call leaves nothing to fetch
    h->last_total = 0;
    if (n_pairs == 0) return EACHAM_OK;

    // validate the pair list against the image table; all pairs of a call share one descriptor kind
    int kind = -1;
    uint32_t max_first = 0, max_second = 0;
    for (size_t i = 0; i < n_pairs; ++i) {
        const uint32_t a = pairs[i].first, b = pairs[i].second;
        if (a >= h->images.size() || b >= h->images.size() || !h->images[a].present || !h->images[b].present)
            return fail(EACHAM_ERR_NOT_COMMITTED, "pair %zu = (%u, %u) names an image without descriptors", i, a, b);
        if (!(h->images[a].has_data || h->images[a].reserved) || !(h->images[b].has_data || h->images[b].reserved))
            return fail(EACHAM_ERR_NOT_COMMITTED, "pair %zu = (%u, %u) names an image whose descriptors were never set", i, a, b);
        if (kind < 0) kind = h->images[a].kind;
        if (h->images[a].kind != kind || h->images[b].kind != kind)
            return fail(EACHAM_ERR_KIND_MISMATCH, "pair %zu = (%u, %u) mixes descriptor kinds", i, a, b);
        max_first = std::max(max_first, h->images[a].rows);
        max_second = std::max(max_second, h->images[b].rows);
    }

    int rc;
    if ((rc = h->d_pairs.ensure(n_pairs))) return rc;
    if ((rc = h->d_order.ensure(n_pairs))) return rc;
    if ((rc = h->d_results.ensure(n_pairs))) return rc;
    // Processing order (results are written at the pair's input index).
    // ORB: the image x image grid is cut into kBlock x kBlock blocks and the pairs of one block are handed out together (counting sort by
    // block, stable), so that the ~148 pairs in flight share a few dozen images that stay L2-resident instead of streaming 148 different
    // second images from HBM.
    // SIFT: by SECOND image (counting sort, stable). A CTA keeps 256 rows of `first` resident and streams all of `second` past them, 32
    // times per pair at 8k rows; an 8k image is 2.6 MB of operand blocks + 4 MB of fp32 rows for the re-rank. With 16 x 16 blocks the 16
    // second images in flight, the first images' fp32 rows and the per-CTA column state (58 MB) oversubscribed L2 (ncu at 500 images:
    // 34 % hit rate, 2.1 TB/s from HBM, SM clock 1.65 GHz at the power cap). Ordered by `second`, all CTAs stream the SAME image at the
    // same time against 148 different resident ones.
    {
        constexpr uint32_t kBlock = 16;
        const bool by_second = kind == EACHAM_KIND_F32X128;
        const uint32_t nb = by_second ? 0u : (uint32_t)((h->images.size() + kBlock - 1) / kBlock);
        const size_t n_keys = by_second ? h->images.size() : (size_t)nb * nb;
        const bool sparse = n_keys > 4 * n_pairs + 1024;          // sparse lists over huge id ranges: keep the input order
        auto key_of = [&](size_t i) -> size_t {
            return by_second ? (size_t)pairs[i].second : (size_t)(pairs[i].first / kBlock) * nb + pairs[i].second / kBlock;
        };
        h->order_host.resize(n_pairs);
        if (sparse) {
            for (size_t i = 0; i < n_pairs; ++i) h->order_host[i] = (uint32_t)i;
        } else {
            h->order_count.assign(n_keys + 1, 0u);
            for (size_t i = 0; i < n_pairs; ++i) ++h->order_count[key_of(i) + 1];
            for (size_t b = 1; b < h->order_count.size(); ++b) h->order_count[b] += h->order_count[b - 1];
            for (size_t i = 0; i < n_pairs; ++i) h->order_host[h->order_count[key_of(i)]++] = (uint32_t)i;
        }
    }
    size_t want_entries = h->cfg_match_entries ? (size_t)h->cfg_match_entries
                                               : std::max((size_t)1 << 20, n_pairs * (size_t)192);
    if (h->d_matches.cap < want_entries && (rc = h->d_matches.ensure(want_entries))) return rc;

    // The tensor-core ORB engines keep no indices in their hot loop; they rely on a match being a strict unique minimum, which
    // holds for ratio <= 1. A ratio above 1 (ties can pass) goes to the XOR+POPC kernels, whose packed keys carry OpenCV's tie order.
    const bool use_tc = (kind == EACHAM_KIND_F32X128 && !(h->cfg_flags & EACHAM_CFG_SIFT_EXACT_FP32)) ||
                        (kind == EACHAM_KIND_ORB256 && !(h->cfg_flags & EACHAM_CFG_ORB_POPC) && o.ratio <= 1.0);
    if (use_tc && (rc = prepare_tc(h))) return rc;
    CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_pairs.p, pairs, n_pairs * sizeof(eacham_pair_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_order.p, h->order_host.data(), n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));

    for (int attempt = 0; attempt < 2; ++attempt) {
        CUDA_TRY(cudaMemsetAsync(h->d_counter.p, 0, 2 * sizeof(uint32_t), h->stream));       // [0] pair queue, [1] exact-scan fallbacks
        CUDA_TRY(cudaMemsetAsync(h->d_cursor.p, 0, sizeof(unsigned long long), h->stream));
        CUDA_TRY(cudaEventRecord(h->ev[4], h->stream));
        if (kind == EACHAM_KIND_ORB256 && !use_tc && max_first <= orb::kMaxRowsFused && max_second <= orb::kMaxRowsFused) {
            orb::PairParams p;
            p.arena = h->arena.p; p.images = h->d_images.p; p.pairs = h->d_pairs.p; p.n_pairs = (uint32_t)n_pairs;
            p.work_counter = h->d_counter.p; p.order = h->d_order.p;
            p.ratio = o.ratio; p.min_dir = o.min_dir; p.min_mutual = o.min_mutual; p.cross_check = o.cross_check; p.emit_all = o.emit_all;
            p.results = h->d_results.p; p.matches = h->d_matches.p; p.matches_cap = h->d_matches.cap; p.cursor = h->d_cursor.p;
            p.smem_cols = (uint32_t)align_up(std::max(max_second, 4u), 4);
            p.smem_rows = (uint32_t)align_up(std::max(max_first, 8u), 8);
            const size_t smem = orb::pair_smem_bytes(p.smem_cols, p.smem_rows);
            if (smem > h->smem_optin) return fail(EACHAM_ERR_TOO_LARGE, "pair kernel needs %zu bytes of shared memory (> %zu)", smem, h->smem_optin);
            CUDA_TRY(cudaFuncSetAttribute(orb::orb_match_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned grid = (unsigned)std::min<size_t>(n_pairs, (size_t)h->sm_count);
            orb::orb_match_pairs_kernel<<<grid, orb::kThreads, smem, h->stream>>>(p);
            CUDA_TRY(cudaGetLastError());
            h->timing.kernel_launches += 1;
        } else if (use_tc) {
            // tensor-core engine, one persistent CTA per pair: SIFT = bf16 scoring + exact FP32 re-rank; ORB (default engine) = FP8 {0,1}
            // operands, distances exact in the accumulator
            tcm::PairParamsTc p;
            p.arena = h->arena.p; p.tc_arena = h->tc_arena.p; p.images = h->d_images_tc.p; p.pairs = h->d_pairs.p; p.n_pairs = (uint32_t)n_pairs;
            p.ratio = o.ratio; p.min_dir = o.min_dir; p.min_mutual = o.min_mutual; p.cross_check = o.cross_check; p.emit_all = o.emit_all;
            p.results = h->d_results.p; p.matches = h->d_matches.p; p.matches_cap = h->d_matches.cap; p.cursor = h->d_cursor.p;
            p.rows_cap = (uint32_t)align_up(std::max(max_first, 1u), 128); p.cols_cap = (uint32_t)align_up(std::max(max_second, 1u), 128);
            const unsigned grid = (unsigned)std::min<size_t>(n_pairs, (size_t)h->sm_count);
            if ((rc = h->tc_scratch.ensure(tcm::tc_scratch_bytes_per_cta(p.rows_cap, p.cols_cap) * grid))) return rc;
            p.scratch = h->tc_scratch.p;
            p.work_counter = h->d_counter.p; p.order = h->d_order.p; p.exact_fallbacks = h->d_counter.p + 1;
            p.dbg_idx12 = p.dbg_idx21 = nullptr; p.dbg_dist12 = p.dbg_dist21 = nullptr;
            p.single_dir = 0; p.single_out = nullptr;
            if (h->dbg_on && n_pairs == 1) {
                p.dbg_idx12 = h->d_dbg_idx.p; p.dbg_dist12 = h->d_dbg_dist.p;
                p.dbg_idx21 = h->d_dbg_idx.p + 2 * (size_t)max_first; p.dbg_dist21 = h->d_dbg_dist.p + 2 * (size_t)max_first;
            }
            const size_t smem = sizeof(tcm::SmemTc) + 128;
            if (kind == EACHAM_KIND_ORB256 && !(h->cfg_flags & EACHAM_CFG_ORB_TC_V1)) {
                // default: F16 accumulators, packed epilogue (tc_orb_kernels.cuh)
                const size_t smem2 = sizeof(tco::SmemOrb) + 128;
                if (h->cfg_flags & EACHAM_CFG_ORB_TC_ALU_SORT) {
                    CUDA_TRY(cudaFuncSetAttribute(tco::orb_tc_match_pairs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                    tco::orb_tc_match_pairs_kernel<false><<<grid, tco::kThreads, smem2, h->stream>>>(p);
                } else {
                    CUDA_TRY(cudaFuncSetAttribute(tco::orb_tc_match_pairs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                    tco::orb_tc_match_pairs_kernel<true><<<grid, tco::kThreads, smem2, h->stream>>>(p);
                }
            } else if (kind == EACHAM_KIND_ORB256) {
                CUDA_TRY(cudaFuncSetAttribute(tcm::tc_match_pairs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                tcm::tc_match_pairs_kernel<true><<<grid, tcm::kThreadsTc, smem, h->stream>>>(p);
            } else if (h->cfg_flags & EACHAM_CFG_SIFT_TC_V1) {
                CUDA_TRY(cudaFuncSetAttribute(tcm::tc_match_pairs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                tcm::tc_match_pairs_kernel<false><<<grid, tcm::kThreadsTc, smem, h->stream>>>(p);
            } else {
                // default: both directions as pruned thread-local scans over D and D^T (tc_sift_kernels.cuh)
                const size_t smem3 = sizeof(tcs::SmemSift) + 128;
                CUDA_TRY(cudaFuncSetAttribute(tcs::sift_tc_match_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
                tcs::sift_tc_match_pairs_kernel<<<grid, tcs::kThreads, smem3, h->stream>>>(p);
            }
            CUDA_TRY(cudaGetLastError());
            h->timing.kernel_launches += 1;
        } else {
            // exact per-pair path: two kNN directions + finalisation (SIFT all-FP32 when cfg.flags & 1; ORB beyond the fused limit)
            const uint32_t max_rows = std::max(max_first, max_second);
            if ((rc = h->d_match.ensure(max_rows)) || (rc = h->d_match2.ensure(max_rows))) return rc;
            for (size_t i = 0; i < n_pairs; ++i) {
                const ImageHost& A = h->images[pairs[i].first];
                const ImageHost& B = h->images[pairs[i].second];
                if (A.rows && (rc = knn2_device(h, kind, h->arena.p + A.offset, A.rows, h->arena.p + B.offset, B.rows, o.ratio, nullptr,
                                                nullptr, h->d_match.p, &h->timing.kernel_launches))) return rc;
                if (B.rows && (rc = knn2_device(h, kind, h->arena.p + B.offset, B.rows, h->arena.p + A.offset, A.rows, o.ratio, nullptr,
                                                nullptr, h->d_match2.p, &h->timing.kernel_launches))) return rc;
                FinalizeParams f;
                f.m12 = h->d_match.p; f.n1 = A.rows; f.m21 = h->d_match2.p; f.n2 = B.rows;
                f.min_dir = o.min_dir; f.min_mutual = o.min_mutual; f.cross_check = o.cross_check; f.emit_all = o.emit_all;
                f.result = h->d_results.p + i; f.matches = h->d_matches.p; f.matches_cap = h->d_matches.cap; f.cursor = h->d_cursor.p;
                pair_finalize_kernel<<<1, 256, 0, h->stream>>>(f);
                h->timing.kernel_launches += 1;
            }
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaEventRecord(h->ev[5], h->stream));
        unsigned long long used = 0;
        uint32_t counters[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(&used, h->d_cursor.p, sizeof(used), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaMemcpyAsync(counters, h->d_counter.p, sizeof(counters), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->timing.exact_fallbacks = counters[1];
        if (used <= h->d_matches.cap) { h->last_total = used; break; }
        if (attempt == 1) return fail(EACHAM_ERR_OUT_OF_MEMORY, "match buffer overflow persisted after regrow");
        if ((rc = h->d_matches.ensure((size_t)used + 1024))) return rc;   // deterministic need: rerun once
    }
    CUDA_TRY(cudaEventElapsedTime(&h->timing.pairs_h2d_ms, h->ev[2], h->ev[3]));
    CUDA_TRY(cudaEventElapsedTime(&h->timing.kernel_ms, h->ev[4], h->ev[5]));
    h->last_n_pairs = n_pairs;           // only a completed batch can be fetched
    h->last_max_first = max_first; h->last_max_second = max_second;
    return EACHAM_OK;
}

int fetch(eacham_gpu_handle* h, eacham_pair_result_t* res, size_t n_pairs, eacham_match_t* buf, size_t buf_cap, size_t* buf_used) {
    if (n_pairs != h->last_n_pairs) return fail(EACHAM_ERR_INVALID_ARG, "fetch of %zu pairs but the last batch had %zu", n_pairs, h->last_n_pairs);
    if (buf_used) *buf_used = (size_t)h->last_total;
    if (n_pairs == 0) return EACHAM_OK;
    CUDA_TRY(cudaEventRecord(h->ev[6], h->stream));
    if (res) CUDA_TRY(cudaMemcpyAsync(res, h->d_results.p, n_pairs * sizeof(eacham_pair_result_t), cudaMemcpyDeviceToHost, h->stream));
    const size_t n_copy = std::min(std::min((size_t)h->last_total, buf_cap), h->d_matches.cap);
    if (buf && n_copy) CUDA_TRY(cudaMemcpyAsync(buf, h->d_matches.p, n_copy * sizeof(eacham_match_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev[7], h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaEventElapsedTime(&h->timing.d2h_ms, h->ev[6], h->ev[7]));
    if (h->last_total > buf_cap) return fail(EACHAM_ERR_BUFFER_TOO_SMALL, "match buffer holds %zu entries, %llu needed", buf_cap, h->last_total);
    return EACHAM_OK;
}

}  // namespace

extern "C" {

int eacham_gpu_match_pairs_device(eacham_gpu_handle* h, const eacham_pair_t* pairs, size_t n_pairs,
                                  const eacham_match_opts* opts, size_t* total_matches) {
    if (!h || (n_pairs && !pairs)) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = launch_pairs(h, pairs, n_pairs, opts);
    if (total_matches) *total_matches = (size_t)h->last_total;
    return rc;
}

int eacham_gpu_fetch_results(eacham_gpu_handle* h, eacham_pair_result_t* res, size_t n_pairs, eacham_match_t* buf,
                             size_t buf_cap, size_t* buf_used) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    return fetch(h, res, n_pairs, buf, buf_cap, buf_used);
}

int eacham_gpu_device_results(eacham_gpu_handle* h, void** results, void** matches, size_t* n_pairs, size_t* n_matches) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (results) *results = h->d_results.p;
    if (matches) *matches = h->d_matches.p;
    if (n_pairs) *n_pairs = h->last_n_pairs;
    if (n_matches) *n_matches = (size_t)h->last_total;
    return EACHAM_OK;
}

int eacham_gpu_match_pairs(eacham_gpu_handle* h, const eacham_pair_t* pairs, size_t n_pairs, const eacham_match_opts* opts,
                           eacham_pair_result_t* res, eacham_match_t* buf, size_t buf_cap, size_t* buf_used) {
    if (!h || (n_pairs && (!pairs || !res)) || (buf_cap && !buf)) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    int rc = launch_pairs(h, pairs, n_pairs, opts);
    if (rc) return rc;
    return fetch(h, res, n_pairs, buf, buf_cap, buf_used);
}

}  // extern "C"

extern "C" int eacham_gpu_debug_pair_knn2(eacham_gpu_handle* h, uint32_t first, uint32_t second, const eacham_match_opts* opts, int32_t* idx12,
                                          float* dist12, int32_t* idx21, float* dist21) {
    if (!h || !idx12 || !dist12 || !idx21 || !dist21) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    if (first >= h->images.size() || second >= h->images.size() || !h->images[first].present || !h->images[second].present)
        return fail(EACHAM_ERR_NOT_COMMITTED, "pair (%u, %u) names an image without descriptors", first, second);
    if (h->images[first].kind != EACHAM_KIND_F32X128 || (h->cfg_flags & EACHAM_CFG_SIFT_EXACT_FP32))
        return fail(EACHAM_ERR_INVALID_ARG, "the debug kNN view exists for F32X128 pairs on the tensor-core engine only");
    const size_t n1 = h->images[first].rows, n2 = h->images[second].rows;
    int rc;
    if ((rc = h->d_dbg_idx.ensure(2 * (n1 + n2) + 4)) || (rc = h->d_dbg_dist.ensure(2 * (n1 + n2) + 4))) return rc;
    CUDA_TRY(cudaMemsetAsync(h->d_dbg_idx.p, 0xFF, (2 * (n1 + n2) + 4) * sizeof(int32_t), h->stream));
    eacham_pair_t pr; pr.first = first; pr.second = second;
    h->dbg_on = true;
    rc = launch_pairs(h, &pr, 1, opts);
    h->dbg_on = false;
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(idx12, h->d_dbg_idx.p, 2 * n1 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dist12, h->d_dbg_dist.p, 2 * n1 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(idx21, h->d_dbg_idx.p + 2 * n1, 2 * n2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dist21, h->d_dbg_dist.p + 2 * n1, 2 * n2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return EACHAM_OK;
}

extern "C" int eacham_gpu_set_keypoints(eacham_gpu_handle* h, uint32_t image_id, const float* xy, uint32_t rows, size_t row_stride_bytes) {
    if (!h) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    if (rows > 0 && !xy) return fail(EACHAM_ERR_INVALID_ARG, "null keypoint pointer for image %u", image_id);
    if (rows > 0 && row_stride_bytes < 2 * sizeof(float)) return fail(EACHAM_ERR_INVALID_ARG, "keypoint row stride %zu < 8", row_stride_bytes);
    if (image_id > (1u << 26)) return fail(EACHAM_ERR_INVALID_ARG, "image id %u out of range", image_id);
    std::lock_guard<std::mutex> lk(h->mu);
    if (image_id >= h->kp_host.size()) h->kp_host.resize((size_t)image_id + 1);
    std::vector<float>& v = h->kp_host[image_id];
    v.resize((size_t)rows * 2);
    for (uint32_t r = 0; r < rows; ++r) memcpy(&v[2 * (size_t)r], reinterpret_cast<const uint8_t*>(xy) + (size_t)r * row_stride_bytes, 2 * sizeof(float));
    h->kp_dirty = true;
    return EACHAM_OK;
}

extern "C" int eacham_gpu_verify_pairs(eacham_gpu_handle* h, int model, const double* hyps, const eacham_verify_opts* opts, eacham_verify_result* res,
                                       float* medians, uint8_t* mask) {
    using namespace eacham;
    if (!h || !hyps || !opts || !res) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    if (model != EACHAM_MODEL_ESSENTIAL && model != EACHAM_MODEL_HOMOGRAPHY) return fail(EACHAM_ERR_INVALID_ARG, "unknown model %d", model);
    if (opts->n_hyp == 0) return fail(EACHAM_ERR_INVALID_ARG, "need at least one hypothesis");
    if (model == EACHAM_MODEL_ESSENTIAL && !(opts->focal > 0.0)) return fail(EACHAM_ERR_INVALID_ARG, "focal length must be positive");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    const size_t n_pairs = h->last_n_pairs;
    if (n_pairs == 0) return fail(EACHAM_ERR_NOT_COMMITTED, "no completed match_pairs batch on this handle to verify");
    const uint32_t max_matches = std::min(h->last_max_first, h->last_max_second);       // a pair has at most min(rows) mutual matches
    if (max_matches > verify::kMaxMatches) return fail(EACHAM_ERR_TOO_LARGE, "up to %u matches per pair are supported for verification", verify::kMaxMatches);
    int rc;
    if (h->kp_dirty) {                                   // keypoint table: every image of the arena needs one row per descriptor
        std::vector<unsigned long long> off(h->images.size() + 1, 0);
        size_t total = 0;
        for (size_t i = 0; i < h->images.size(); ++i) {
            off[i] = total;
            const size_t have = i < h->kp_host.size() ? h->kp_host[i].size() / 2 : 0;
            if (h->images[i].present && have != h->images[i].rows)
                return fail(EACHAM_ERR_INVALID_ARG, "image %zu has %u descriptors but %zu keypoints", i, h->images[i].rows, have);
            total += have;
        }
        off[h->images.size()] = total;
        if ((rc = h->d_kp.ensure(std::max(total, (size_t)1))) || (rc = h->d_kp_offset.ensure(off.size()))) return rc;
        std::vector<float> flat(2 * std::max(total, (size_t)1));
        for (size_t i = 0; i < h->images.size() && i < h->kp_host.size(); ++i)
            if (!h->kp_host[i].empty()) memcpy(&flat[2 * off[i]], h->kp_host[i].data(), h->kp_host[i].size() * sizeof(float));
        CUDA_TRY(cudaMemcpyAsync(h->d_kp.p, flat.data(), 2 * total * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_kp_offset.p, off.data(), off.size() * sizeof(off[0]), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->kp_dirty = false;
    }
    const size_t n_h = (size_t)opts->n_hyp * (opts->shared ? 1 : n_pairs) * 9;
    if ((rc = h->d_hyps.ensure(n_h)) || (rc = h->d_vres.ensure(n_pairs))) return rc;
    if (medians && (rc = h->d_vmed.ensure(n_pairs * opts->n_hyp))) return rc;
    if (mask && (rc = h->d_vmask.ensure(std::max((size_t)h->last_total, (size_t)1)))) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->d_hyps.p, hyps, n_h * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    verify::Params p;
    p.pairs = h->d_pairs.p; p.results = h->d_results.p; p.matches = h->d_matches.p; p.keypoints = h->d_kp.p; p.kp_offset = h->d_kp_offset.p;
    p.hyps = h->d_hyps.p; p.n_pairs = (uint32_t)n_pairs; p.n_hyp = opts->n_hyp; p.shared = opts->shared ? 1u : 0u; p.model = (uint32_t)model;
    p.focal = opts->focal; p.cx = opts->cx; p.cy = opts->cy;
    p.out = h->d_vres.p; p.medians = medians ? h->d_vmed.p : nullptr; p.mask = mask ? h->d_vmask.p : nullptr;
    uint32_t cap = 64;
    while (cap < max_matches) cap <<= 1;
    p.cap = cap;
    const size_t smem = (size_t)cap * (sizeof(float4) + 2 * sizeof(float));
    CUDA_TRY(cudaFuncSetAttribute(verify::verify_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max(smem, (size_t)1024)));
    const unsigned grid = (unsigned)std::min<size_t>(n_pairs, (size_t)h->sm_count * per_sm);
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    verify::verify_pairs_kernel<<<grid, verify::kThreads, smem, h->stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    CUDA_TRY(cudaMemcpyAsync(res, h->d_vres.p, n_pairs * sizeof(eacham_verify_result), cudaMemcpyDeviceToHost, h->stream));
    if (medians) CUDA_TRY(cudaMemcpyAsync(medians, h->d_vmed.p, n_pairs * opts->n_hyp * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (mask && h->last_total) CUDA_TRY(cudaMemcpyAsync(mask, h->d_vmask.p, (size_t)h->last_total, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaEventElapsedTime(&h->timing.verify_ms, h->ev[0], h->ev[1]));
    return EACHAM_OK;
}

#include "match_single.cuh"
#include "multi.cuh"
