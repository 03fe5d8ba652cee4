// match_single.cuh -- the reference-shaped per-call route, eacham_gpu_match == FeatureMatcherFlann::Match
// (/root/reference/modules/base/features/FeatureMatcherFlann.cpp:14-30), on the tensor-core engines.
//
// What the reference does with it: TBB workers call Match concurrently on ONE matcher object, once per ORDERED image pair, always
// with the descriptor matrices owned by the graph nodes (/root/reference/apps/sfm/main.cpp:98-109). So
//   * descriptors are cached on the device, keyed by (host pointer, rows, stride, kind) and validated by a hash of sampled rows:
//     an image is uploaded and converted to tensor-core operands once, not 2 (n - 1) times;
//   * every call runs on its own stream with its own small buffers (8 call slots), so concurrent callers overlap on the GPU;
//   * one call = the single-direction mode of the pair kernels (PairParamsTc::single_dir): one CTA per 128 query rows sweeping all
//     train tiles, row direction only, ratio test and index recovery on chip; the host only compacts the per-row result.
// Included at the end of eacham_gpu.cu (one translation unit).
#pragma once

namespace {

uint64_t sample_hash(const void* data, uint32_t rows, size_t stride, size_t rb) {
    // FNV-1a over 16 rows spread across the matrix (first and last included)
    uint64_t hsh = 1469598103934665603ull ^ ((uint64_t)rows * 0x9E3779B97F4A7C15ull);
    const uint32_t n = std::min<uint32_t>(rows, 16u);
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t r = n > 1 ? (uint32_t)(((uint64_t)k * (rows - 1)) / (n - 1)) : 0u;
        const uint64_t* w = reinterpret_cast<const uint64_t*>(static_cast<const uint8_t*>(data) + (size_t)r * stride);
        uint64_t acc = 0;
        for (size_t i = 0; i < rb / 8; ++i) { uint64_t v; memcpy(&v, w + i, 8); acc = (acc ^ v) * 0x100000001B3ull; }
        hsh = (hsh ^ acc) * 0x100000001B3ull;
    }
    return hsh;
}

bool match_single_eligible(eacham_gpu_handle* h, int kind, double ratio, size_t q_stride, size_t t_stride) {
    if (h->cfg_flags & EACHAM_CFG_MATCH_LEGACY) return false;
    if (kind == EACHAM_KIND_ORB256) {
        // the value-only ORB engine needs a strict unique minimum (ratio <= 1); the round-1 tensor kernel and the XOR+POPC engine
        // selections keep the packed-key kernels for Match
        if (!(ratio <= 1.0) || (h->cfg_flags & (EACHAM_CFG_ORB_POPC | EACHAM_CFG_ORB_TC_V1))) return false;
        return q_stride >= 32 && t_stride >= 32;
    }
    if (h->cfg_flags & EACHAM_CFG_SIFT_EXACT_FP32) return false;
    return q_stride >= 512 && t_stride >= 512;
}

int ensure_single_state(eacham_gpu_handle* h) {          // cache_mu held
    if (!h->cache.empty()) return EACHAM_OK;
    int rc = h->d_cache_descs.ensure(eacham_gpu_handle::kCacheEntries);
    if (rc) return rc;
    h->slots.resize(eacham_gpu_handle::kCallSlots);
    for (auto& sl : h->slots) {
        CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        if ((rc = sl.d_counter.ensure(16)) || (rc = sl.d_pair.ensure(16))) return rc;
        CUDA_TRY(cudaHostAlloc((void**)&sl.h_pair, sizeof(eacham_pair_t), cudaHostAllocDefault));
    }
    h->cache.resize(eacham_gpu_handle::kCacheEntries);
    return EACHAM_OK;
}

// Finds or loads the device copy of one descriptor matrix and pins it (refs + 1). `slot` provides the stream for a load.
int cache_acquire(eacham_gpu_handle* h, eacham_gpu_handle::CallSlot& slot, int kind, const void* data, uint32_t rows, size_t stride, int* entry_out) {
    using namespace eacham;
    const size_t rb = row_bytes(kind);
    const bool use_cache = !(h->cfg_flags & EACHAM_CFG_MATCH_NO_CACHE);
    const uint64_t hsh = use_cache ? sample_hash(data, rows, stride, rb) : 0;
    std::unique_lock<std::mutex> lk(h->cache_mu);
    int idx = -1;
    for (;;) {
        idx = -1;
        if (use_cache)
            for (int i = 0; i < (int)h->cache.size(); ++i) {
                auto& c = h->cache[i];
                if (c.host == data && c.rows == rows && c.stride == stride && c.kind == kind && c.hash == hsh && (c.ready || c.loading)) { idx = i; break; }
            }
        if (idx >= 0 && h->cache[idx].loading) { h->cache_cv.wait(lk); continue; }     // another caller is uploading this very image
        break;
    }
    if (idx >= 0) {
        auto& c = h->cache[idx];
        ++c.refs; c.tick = ++h->cache_tick; ++h->cache_hits;
        *entry_out = idx;
        return EACHAM_OK;
    }
    // miss: take the least recently used entry nobody holds
    for (;;) {
        uint64_t best = ~0ull;
        for (int i = 0; i < (int)h->cache.size(); ++i) {
            auto& c = h->cache[i];
            if (c.refs == 0 && !c.loading && (!c.ready ? 0ull : c.tick) < best) { best = !c.ready ? 0ull : c.tick; idx = i; }
        }
        if (idx >= 0) break;
        h->cache_cv.wait(lk);                             // every entry is pinned by an in-flight call: wait for one to finish
    }
    auto& c = h->cache[idx];
    c.host = data; c.rows = rows; c.stride = stride; c.kind = kind; c.hash = hsh;
    c.ready = false; c.loading = true; c.refs = 1; c.tick = ++h->cache_tick; ++h->cache_misses;
    lk.unlock();

    // upload + tensor-core operand prep on the caller's slot stream (no lock held: other callers keep matching)
    int rc = EACHAM_OK;
    const uint32_t nblk = (rows + 127) / 128;
    tcm::ImageDescTc desc;
    do {
        if ((rc = c.raw.ensure(std::max((size_t)rows * rb, rb))) || (rc = c.tc.ensure((size_t)std::max(nblk, 1u) * tc::kBlockBytes))) break;
        desc.offset = reinterpret_cast<unsigned long long>(c.raw.p);
        desc.tc_offset = reinterpret_cast<unsigned long long>(c.tc.p);
        desc.rows = rows; desc.kind = (uint32_t)kind; desc.max_norm_bits = 0u; desc.bf16_exact = 1u;
        cudaError_t e = cudaMemcpyAsync(h->d_cache_descs.p + idx, &desc, sizeof(desc), cudaMemcpyHostToDevice, slot.stream);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(c.raw.p, rb, data, stride, rb, rows, cudaMemcpyHostToDevice, slot.stream);
        if (e == cudaSuccess) {
            if (kind == EACHAM_KIND_F32X128)
                tcm::sift_prep_kernel<<<nblk * 16, 256, 0, slot.stream>>>(reinterpret_cast<const float*>(c.raw.p), rows, c.tc.p, nblk,
                                                                          &h->d_cache_descs.p[idx].max_norm_bits, &h->d_cache_descs.p[idx].bf16_exact);
            else
                tcm::orb_tc_prep_kernel<<<nblk * 16, 256, 0, slot.stream>>>(c.raw.p, rows, c.tc.p, nblk);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(slot.stream);          // `data` and `desc` may go away after this call
        if (e != cudaSuccess) rc = fail(EACHAM_ERR_CUDA, "descriptor upload failed: %s", cudaGetErrorString(e));
    } while (0);
    lk.lock();
    c.loading = false;
    c.ready = rc == EACHAM_OK;
    if (rc) { c.refs = 0; c.host = nullptr; }
    h->cache_cv.notify_all();
    *entry_out = idx;
    return rc;
}

void cache_release(eacham_gpu_handle* h, int idx) {
    std::lock_guard<std::mutex> lk(h->cache_mu);
    auto& c = h->cache[idx];
    if (--c.refs == 0 && (h->cfg_flags & EACHAM_CFG_MATCH_NO_CACHE)) { c.ready = false; c.host = nullptr; }
    h->cache_cv.notify_all();
}

int match_single_tc(eacham_gpu_handle* h, int kind, const void* query, uint32_t q_rows, size_t q_stride, const void* train, uint32_t t_rows,
                    size_t t_stride, double ratio, eacham_match_t* out, size_t cap, size_t* n_out) {
    using namespace eacham;
    DeviceGuard g(h->device);
    int rc;
    int si = -1;
    {
        std::unique_lock<std::mutex> lk(h->cache_mu);
        if ((rc = ensure_single_state(h))) return rc;
        for (;;) {
            for (int i = 0; i < (int)h->slots.size(); ++i) if (!h->slots[i].busy) { si = i; break; }
            if (si >= 0) break;
            h->cache_cv.wait(lk);
        }
        h->slots[si].busy = true;
    }
    auto& slot = h->slots[si];
    int qe = -1, te = -1;
    auto finish = [&](int code) {
        if (qe >= 0) cache_release(h, qe);
        if (te >= 0) cache_release(h, te);
        std::lock_guard<std::mutex> lk(h->cache_mu);
        slot.busy = false;
        h->cache_cv.notify_all();
        return code;
    };
    if ((rc = cache_acquire(h, slot, kind, query, q_rows, q_stride, &qe))) { qe = -1; return finish(rc); }
    if ((rc = cache_acquire(h, slot, kind, train, t_rows, t_stride, &te))) { te = -1; return finish(rc); }
    if ((rc = slot.d_out.ensure(q_rows))) return finish(rc);
    if (slot.h_out_cap < q_rows) {
        if (slot.h_out) cudaFreeHost(slot.h_out);
        slot.h_out = nullptr; slot.h_out_cap = 0;
        const size_t want = std::max<size_t>(q_rows, 8192);
        if (cudaHostAlloc((void**)&slot.h_out, want * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return finish(fail(EACHAM_ERR_OUT_OF_MEMORY, "pinned result buffer allocation failed"));
        }
        slot.h_out_cap = want;
    }
    slot.h_pair->first = (uint32_t)qe; slot.h_pair->second = (uint32_t)te;

    tcm::PairParamsTc p{};
    p.arena = nullptr; p.tc_arena = nullptr; p.images = h->d_cache_descs.p; p.pairs = slot.d_pair.p;
    p.n_pairs = (q_rows + 127) / 128;                     // work items = 128-row blocks of the query image, one CTA each
    p.ratio = ratio; p.min_dir = 0; p.min_mutual = 0; p.cross_check = 0; p.emit_all = 1;
    p.results = nullptr; p.matches = nullptr; p.matches_cap = 0; p.cursor = nullptr;
    p.scratch = nullptr; p.rows_cap = 0; p.cols_cap = 0;
    p.work_counter = slot.d_counter.p; p.order = nullptr; p.exact_fallbacks = slot.d_counter.p + 1;
    p.single_dir = 1; p.single_out = slot.d_out.p;
    const unsigned grid = std::min<unsigned>(p.n_pairs, (unsigned)h->sm_count);
    cudaError_t e = cudaMemcpyAsync(slot.d_pair.p, slot.h_pair, sizeof(eacham_pair_t), cudaMemcpyHostToDevice, slot.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(slot.d_counter.p, 0, 2 * sizeof(uint32_t), slot.stream);
    if (e == cudaSuccess) {
        if (kind == EACHAM_KIND_ORB256) {
            const size_t smem = sizeof(tco::SmemOrb) + 128;
            if (h->cfg_flags & EACHAM_CFG_ORB_TC_ALU_SORT) {
                e = cudaFuncSetAttribute(tco::orb_tc_match_pairs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess) tco::orb_tc_match_pairs_kernel<false><<<grid, tco::kThreads, smem, slot.stream>>>(p);
            } else {
                e = cudaFuncSetAttribute(tco::orb_tc_match_pairs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e == cudaSuccess) tco::orb_tc_match_pairs_kernel<true><<<grid, tco::kThreads, smem, slot.stream>>>(p);
            }
        } else if (h->cfg_flags & EACHAM_CFG_SIFT_TC_V1) {
            const size_t smem = sizeof(tcm::SmemTc) + 128;
            e = cudaFuncSetAttribute(tcm::tc_match_pairs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) tcm::tc_match_pairs_kernel<false><<<grid, tcm::kThreadsTc, smem, slot.stream>>>(p);
        } else {
            const size_t smem = sizeof(tcs::SmemSift) + 128;
            e = cudaFuncSetAttribute(tcs::sift_tc_match_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) tcs::sift_tc_match_pairs_kernel<<<grid, tcs::kThreads, smem, slot.stream>>>(p);
        }
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(slot.h_out, slot.d_out.p, (size_t)q_rows * sizeof(uint32_t), cudaMemcpyDeviceToHost, slot.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(slot.stream);
    if (e != cudaSuccess) return finish(fail(EACHAM_ERR_CUDA, "single-direction match failed: %s", cudaGetErrorString(e)));
    size_t n = 0;
    for (uint32_t i = 0; i < q_rows; ++i)
        if (slot.h_out[i] != EACHAM_NONE) {
            if (n < cap) { out[n].query = i; out[n].train = slot.h_out[i]; }
            ++n;
        }
    *n_out = n;
    if (n > cap) return finish(fail(EACHAM_ERR_BUFFER_TOO_SMALL, "match buffer holds %zu entries, %zu needed", cap, n));
    return finish(EACHAM_OK);
}

}  // namespace
