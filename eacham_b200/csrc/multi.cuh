// multi.cuh -- one process, several B200s of one NVSwitch box, behind the C ABI (include/eacham_gpu.h, eacham_gpu_multi_*).
//
// The pair list of /root/reference/apps/sfm/main.cpp:84-92 is a set of independent units, so the path shards with no data-path
// exchange: every device holds the whole descriptor arena, every device matches its share of the pair list (whole blocks of the image x image grid) and copies ITS shard of
// the results over ITS PCIe link into its slice of the caller's buffers. The only collective is one broadcast of the arena:
// devices[0] receives the staged bytes (one H2D) and ncclBroadcast (ncclCommInitAll, one stream per device, NVLink through
// NVSwitch) replicates them. NCCL is bound at run time (dlopen of libnccl.so.2) so the single-device library has no NCCL
// dependency; EACHAM_CFG_MULTI_PARALLEL_H2D replaces the broadcast by one H2D copy per device from the same pinned staging
// buffer. Included at the end of eacham_gpu.cu (one translation unit: it uses the handle's internals).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <thread>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && Broadcast && GroupStart && GroupEnd && GetErrorString; }
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(api.lib, "ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(dlsym(api.lib, "ncclBroadcast"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(dlsym(api.lib, "ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(dlsym(api.lib, "ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    });
    return api;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct eacham_gpu_multi {
    std::vector<eacham_gpu_handle*> dev;          // dev[0] owns the pinned staging buffer
    std::vector<ncclComm_t> comms;                // empty when EACHAM_CFG_MULTI_PARALLEL_H2D
    uint32_t flags = 0;
    std::mutex mu;
    // per device, reused across calls
    std::vector<std::vector<eacham_pair_t>> shard;
    std::vector<std::vector<uint32_t>> shard_index;          // input index of every shard element
    std::vector<uint32_t> block_rank;                        // scratch of the block-level sharding
    std::vector<eacham_pair_result_t*> res_pinned;
    std::vector<size_t> res_pinned_cap;
    eacham_gpu_multi_timing timing = {};
};

namespace {

// Runs fn(g) on one host thread per device and returns the first non-zero status (message kept for the caller's thread).
template <class Fn>
int for_each_device(eacham_gpu_multi* m, Fn fn) {
    const size_t n = m->dev.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    if (n == 1) {
        rc[0] = fn(0);
        return rc[0];
    }
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t g = 0; g < n; ++g)
        th.emplace_back([&, g] {
            rc[g] = fn(g);
            if (rc[g]) msg[g] = g_last_error;            // thread-local: carry it over to the caller
        });
    for (auto& t : th) t.join();
    for (size_t g = 0; g < n; ++g)
        if (rc[g]) { g_last_error = "device " + std::to_string(m->dev[g]->device) + ": " + msg[g]; return rc[g]; }
    return EACHAM_OK;
}

}  // namespace

extern "C" {

int eacham_gpu_create_multi(const int32_t* devices, uint32_t n_devices, const eacham_gpu_config* cfg, eacham_gpu_multi** out) {
    if (!out) return fail(EACHAM_ERR_INVALID_ARG, "out handle pointer is null");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > 64) return fail(EACHAM_ERR_INVALID_ARG, "need 1..64 devices");
    eacham_gpu_multi* m = new (std::nothrow) eacham_gpu_multi();
    if (!m) return fail(EACHAM_ERR_OUT_OF_MEMORY, "host allocation failed");
    m->flags = cfg ? cfg->flags : 0;
    for (uint32_t g = 0; g < n_devices; ++g) {
        eacham_gpu_config c = cfg ? *cfg : eacham_gpu_config{};
        c.device = devices[g];
        eacham_gpu_handle* h = nullptr;
        int rc = eacham_gpu_create(&c, &h);
        if (rc) { eacham_gpu_destroy_multi(m); return rc; }
        m->dev.push_back(h);
    }
    m->shard.resize(n_devices);
    m->shard_index.resize(n_devices);
    m->res_pinned.assign(n_devices, nullptr);
    m->res_pinned_cap.assign(n_devices, 0);
    if (!(m->flags & EACHAM_CFG_MULTI_PARALLEL_H2D)) {
        NcclApi& api = nccl_api();
        if (!api.ok()) { eacham_gpu_destroy_multi(m); return fail(EACHAM_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols"); }
        std::vector<int> devlist(devices, devices + n_devices);
        m->comms.resize(n_devices);
        ncclResult_t r = api.CommInitAll(m->comms.data(), (int)n_devices, devlist.data());
        if (r != ncclSuccess) {
            m->comms.clear();
            eacham_gpu_destroy_multi(m);
            return fail(EACHAM_ERR_NCCL, "ncclCommInitAll over %u devices failed: %s", n_devices, api.GetErrorString(r));
        }
    }
    *out = m;
    return EACHAM_OK;
}

void eacham_gpu_destroy_multi(eacham_gpu_multi* m) {
    if (!m) return;
    for (size_t g = 0; g < m->comms.size(); ++g)
        if (m->comms[g]) { DeviceGuard dg(m->dev[g]->device); nccl_api().CommDestroy(m->comms[g]); }
    for (size_t g = 0; g < m->res_pinned.size(); ++g) if (m->res_pinned[g]) cudaFreeHost(m->res_pinned[g]);
    for (eacham_gpu_handle* h : m->dev) eacham_gpu_destroy(h);
    delete m;
}

uint32_t eacham_gpu_multi_device_count(eacham_gpu_multi* m) { return m ? (uint32_t)m->dev.size() : 0u; }

int eacham_gpu_multi_set_descriptors(eacham_gpu_multi* m, uint32_t image_id, int kind, const void* data, uint32_t rows, size_t row_stride_bytes) {
    if (!m) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(m->mu);
    return eacham_gpu_set_descriptors(m->dev[0], image_id, kind, data, rows, row_stride_bytes);
}

int eacham_gpu_multi_set_descriptors_batch(eacham_gpu_multi* m, uint32_t first_id, uint32_t n, int kind, const void* const* data, const uint32_t* rows,
                                           const size_t* row_stride_bytes) {
    if (!m) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(m->mu);
    return eacham_gpu_set_descriptors_batch(m->dev[0], first_id, n, kind, data, rows, row_stride_bytes);
}

int eacham_gpu_multi_clear(eacham_gpu_multi* m) {
    if (!m) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(m->mu);
    for (eacham_gpu_handle* h : m->dev) { int rc = eacham_gpu_clear(h); if (rc) return rc; }
    return EACHAM_OK;
}

int eacham_gpu_multi_commit(eacham_gpu_multi* m) {
    if (!m) return fail(EACHAM_ERR_INVALID_ARG, "null handle");
    std::lock_guard<std::mutex> lk(m->mu);
    const double t0 = now_ms();
    eacham_gpu_handle* h0 = m->dev[0];
    int rc = eacham_gpu_commit(h0);                              // layout + one H2D copy on devices[0]
    if (rc) return rc;
    const double t1 = now_ms();
    const size_t n = m->dev.size();
    // the other devices lay out the same arena (offsets only: no pinned memory behind it)
    rc = for_each_device(m, [&](size_t g) -> int {
        if (g == 0) return EACHAM_OK;
        eacham_gpu_handle* h = m->dev[g];
        int r = eacham_gpu_clear(h);
        for (size_t i = 0; i < h0->images.size() && !r; ++i)
            if (h0->images[i].present) r = eacham_gpu_reserve(h, (uint32_t)i, h0->images[i].kind, h0->images[i].rows);
        if (!r) r = eacham_gpu_commit(h);
        if (!r && h->staging_used != h0->staging_used) r = fail(EACHAM_ERR_CUDA, "arena layouts differ between devices");
        return r;
    });
    if (rc) return rc;
    const size_t bytes = std::min(h0->staging_used, h0->staging_cap);
    if (n > 1 && bytes > 0) {
        if (m->comms.empty()) {
            rc = for_each_device(m, [&](size_t g) -> int {       // one H2D per device, each over its own PCIe link
                if (g == 0) return EACHAM_OK;
                eacham_gpu_handle* h = m->dev[g];
                DeviceGuard dg(h->device);
                CUDA_TRY(cudaMemcpyAsync(h->arena.p, h0->staging, bytes, cudaMemcpyHostToDevice, h->stream));
                CUDA_TRY(cudaStreamSynchronize(h->stream));
                return EACHAM_OK;
            });
            if (rc) return rc;
        } else {
            NcclApi& api = nccl_api();
            ncclResult_t r = api.GroupStart();
            for (size_t g = 0; g < n && r == ncclSuccess; ++g) {
                DeviceGuard dg(m->dev[g]->device);
                r = api.Broadcast(h0->arena.p, m->dev[g]->arena.p, bytes, ncclUint8, 0, m->comms[g], m->dev[g]->stream);
            }
            const ncclResult_t r2 = api.GroupEnd();
            if (r == ncclSuccess) r = r2;
            if (r != ncclSuccess) return fail(EACHAM_ERR_NCCL, "ncclBroadcast of the descriptor arena failed: %s", api.GetErrorString(r));
            for (size_t g = 0; g < n; ++g) {
                DeviceGuard dg(m->dev[g]->device);
                CUDA_TRY(cudaStreamSynchronize(m->dev[g]->stream));
            }
        }
    }
    const double t2 = now_ms();
    m->timing.upload_ms = (float)(t1 - t0);
    m->timing.broadcast_ms = (float)(t2 - t1);
    return EACHAM_OK;
}

int eacham_gpu_multi_match_pairs(eacham_gpu_multi* m, const eacham_pair_t* pairs, size_t n_pairs, const eacham_match_opts* opts,
                                 eacham_pair_result_t* res, eacham_match_t* buf, size_t buf_cap, size_t* buf_used) {
    if (!m || (n_pairs && (!pairs || !res)) || (buf_cap && !buf)) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    const size_t n = m->dev.size();
    const double t0 = now_ms();
    // shard: whole 16 x 16 blocks of the image x image grid go to one device (largest block first, each to the device with the least pairs so far), so that every
    // device works through complete blocks whose ~32 images stay resident in its L2 (pair k -> device k % n would leave each device
    // 1/n of every block and n times as many images in flight). Sparse id ranges fall back to k % n.
    for (size_t g = 0; g < n; ++g) {
        m->shard[g].clear(); m->shard_index[g].clear();
        m->shard[g].reserve(n_pairs / n + 256); m->shard_index[g].reserve(n_pairs / n + 256);
    }
    {
        uint32_t kBlock = 16;
        uint32_t max_id = 0;
        for (size_t k = 0; k < n_pairs; ++k) max_id = std::max(max_id, std::max(pairs[k].first, pairs[k].second));
        size_t nb = (size_t)max_id / kBlock + 1;
        bool blocked = n > 1 && nb * nb <= 4 * n_pairs + 1024;
        while (blocked) {                                  // small sets: finer blocks until there are enough of them to balance
            m->block_rank.assign(nb * nb, 0u);
            for (size_t k = 0; k < n_pairs; ++k) ++m->block_rank[(size_t)(pairs[k].first / kBlock) * nb + pairs[k].second / kBlock];
            size_t occupied = 0;
            for (uint32_t c : m->block_rank) occupied += c != 0;
            if (kBlock == 1 || occupied >= 8 * n) break;
            kBlock /= 2;
            nb = (size_t)max_id / kBlock + 1;
            blocked = nb * nb <= 4 * n_pairs + 1024;
        }
        if (blocked) {
            // blocks differ in size (diagonal blocks, window lists): largest first, each to the device with the least pairs so far
            std::vector<uint32_t> ids;
            for (uint32_t b = 0; b < m->block_rank.size(); ++b) if (m->block_rank[b]) ids.push_back(b);
            std::stable_sort(ids.begin(), ids.end(), [&](uint32_t a, uint32_t b) { return m->block_rank[a] > m->block_rank[b]; });
            std::vector<size_t> load(n, 0);
            for (uint32_t b : ids) {
                const size_t g = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
                load[g] += m->block_rank[b];
                m->block_rank[b] = 1u + (uint32_t)g;       // from here on: 0 = empty, else device + 1
            }
        }
        for (size_t k = 0; k < n_pairs; ++k) {
            const size_t g = blocked ? m->block_rank[(size_t)(pairs[k].first / kBlock) * nb + pairs[k].second / kBlock] - 1u : k % n;
            m->shard[g].push_back(pairs[k]);
            m->shard_index[g].push_back((uint32_t)k);
        }
    }
    std::vector<size_t> total(n, 0);
    int rc = for_each_device(m, [&](size_t g) -> int {
        return eacham_gpu_match_pairs_device(m->dev[g], m->shard[g].data(), m->shard[g].size(), opts, &total[g]);
    });
    if (rc) return rc;
    const double t1 = now_ms();
    std::vector<size_t> base(n + 1, 0);
    for (size_t g = 0; g < n; ++g) base[g + 1] = base[g] + total[g];
    if (buf_used) *buf_used = base[n];
    // every device copies its own shard into its slice of the caller's buffers, in parallel (pin `buf` with
    // eacham_gpu_host_alloc for full PCIe rate); per-pair records are re-interleaved into input order with offsets rebased
    rc = for_each_device(m, [&](size_t g) -> int {
        eacham_gpu_handle* h = m->dev[g];
        const size_t np = m->shard[g].size();
        if (np == 0) return EACHAM_OK;
        DeviceGuard dg(h->device);
        if (m->res_pinned_cap[g] < np) {
            if (m->res_pinned[g]) cudaFreeHost(m->res_pinned[g]);
            m->res_pinned[g] = nullptr; m->res_pinned_cap[g] = 0;
            const size_t cap = np + np / 4 + 64;
            CUDA_TRY(cudaHostAlloc((void**)&m->res_pinned[g], cap * sizeof(eacham_pair_result_t), cudaHostAllocPortable));
            m->res_pinned_cap[g] = cap;
        }
        const size_t room = buf_cap > base[g] ? buf_cap - base[g] : 0;
        size_t used = 0;
        int r = eacham_gpu_fetch_results(h, m->res_pinned[g], np, room ? buf + base[g] : nullptr, room, &used);
        if (r && r != EACHAM_ERR_BUFFER_TOO_SMALL) return r;
        const eacham_pair_result_t* src = m->res_pinned[g];
        for (size_t i = 0; i < np; ++i) {
            eacham_pair_result_t v = src[i];
            v.offset += base[g];
            res[m->shard_index[g][i]] = v;
        }
        return EACHAM_OK;
    });
    if (rc) return rc;
    const double t2 = now_ms();
    m->timing.match_ms = (float)(t1 - t0);
    m->timing.d2h_ms = (float)(t2 - t1);
    m->timing.kernel_ms_max = 0.f; m->timing.prep_ms_max = 0.f; m->timing.kernel_launches = 0;
    for (eacham_gpu_handle* h : m->dev) {
        m->timing.kernel_ms_max = std::max(m->timing.kernel_ms_max, h->timing.kernel_ms);
        m->timing.prep_ms_max = std::max(m->timing.prep_ms_max, h->timing.prep_ms);
        m->timing.kernel_launches += h->timing.kernel_launches;
    }
    if (base[n] > buf_cap) return fail(EACHAM_ERR_BUFFER_TOO_SMALL, "match buffer holds %zu entries, %zu needed", buf_cap, base[n]);
    return EACHAM_OK;
}

int eacham_gpu_multi_last_timing(eacham_gpu_multi* m, eacham_gpu_multi_timing* t) {
    if (!m || !t) return fail(EACHAM_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lk(m->mu);
    *t = m->timing;
    return EACHAM_OK;
}

void* eacham_gpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, std::max(bytes, (size_t)1), cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); fail(EACHAM_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}

void eacham_gpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
