// group.cu -- one process, N GPUs of one box (include/rtb200.h, "multi-GPU group").
//
// The reference drives exactly one device (RayTracer.cpp:2128-2131 takes devices[0]; raytrace_gpgpu :330-344 enqueues on
// one queue). A group is N contexts, one per GPU, behind the same seam:
//   * the scene is packed and uploaded once (GPU 0) and its blob -- ONE contiguous buffer -- is broadcast to the other GPUs
//     with ncclBroadcast over NVLink (ncclCommInitAll: all ranks in this process), then adopted there without a host
//     round trip;
//   * a frame is split in interleaved bands of `band_rows` rows; every GPU traces its bands and stores the pixels straight
//     into the caller's frame: page-locked host memory is written by all GPUs at once, each over its own PCIe link
//     (zero copy, rows assembled to 128/512-byte writes); a pageable frame goes through a frame on GPU 0 that the other
//     GPUs write over NVLink (cudaDeviceEnablePeerAccess; the gather is fused into the kernels' stores) + one copy.
//   * launches are issued by one worker thread per GPU so that no GPU waits for the host to get round to it.
// Only the public C ABI of the contexts is used here; NCCL is bound at run time (dlopen) so that librtb200.so loads
// on hosts without it -- rt_group_upload_scene then fails with a clear message unless option "broadcast" = 1
// (peer copies) is set.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtb200.h"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string problem;

    bool load() {
        if (handle) return true;
        if (!problem.empty()) return false;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            problem = std::string("NCCL not found (dlopen libnccl.so.2: ") + dlerror() + ")";
            return false;
        }
        bool ok = true;
        auto sym = [&](const char* n) {
            void* p = dlsym(handle, n);
            if (!p) ok = false;
            return p;
        };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        GetVersion = (decltype(GetVersion))sym("ncclGetVersion");
        if (!ok) {
            problem = "libnccl.so.2 lacks a required symbol";
            handle = nullptr;
        }
        return ok;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;
thread_local std::string g_group_create_error;

}  // namespace

struct rt_group {
    int n = 0;
    std::vector<int> dev;
    std::vector<rt_context*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev0, ev1;
    std::vector<void*> d_blob;          // ranks > 0: this GPU's copy of the scene blob (owned)
    std::vector<ncclComm_t> comms;
    void* d_frame0 = nullptr;           // frame on GPU 0 that the peers write (pageable destinations)
    size_t frame0_bytes = 0;
    bool peers_enabled = false;
    int band_rows = 16;
    int broadcast_mode = 0;             // 0 = ncclBroadcast, 1 = cudaMemcpyPeerAsync from GPU 0 (diagnostic)
    double stats[RT_GROUP_STATS] = {0};
    std::string err;
    // one worker thread per GPU
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::atomic<uint64_t> generation{0};
    std::function<int(int)> job;
    std::vector<int> rc;
    int pending = 0;
    bool stop = false;
};

static int gerr(rt_group* g, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (g) g->err = buf;
    else g_group_create_error = buf;
    return code;
}

static void worker_main(rt_group* g, int rank) {
    cudaSetDevice(g->dev[rank]);
    uint64_t seen = 0;
    for (;;) {
        // a frame takes a fraction of a millisecond: spin briefly for the next job before sleeping on the condition variable
        const auto t0 = std::chrono::steady_clock::now();
        while (g->generation.load(std::memory_order_acquire) == seen &&
               std::chrono::steady_clock::now() - t0 < std::chrono::microseconds(200)) {
        }
        {
            std::unique_lock<std::mutex> lk(g->m);
            g->cv_job.wait(lk, [&] { return g->stop || g->generation.load(std::memory_order_relaxed) != seen; });
            if (g->stop) return;
            seen = g->generation.load(std::memory_order_relaxed);
        }
        const int r = g->job(rank);
        {
            std::lock_guard<std::mutex> lk(g->m);
            g->rc[rank] = r;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

// Run fn(rank) on every GPU's worker thread; returns the first non-zero result (the failing context holds the message).
static int run_all(rt_group* g, std::function<int(int)> fn) {
    {
        std::lock_guard<std::mutex> lk(g->m);
        g->job = std::move(fn);
        g->pending = g->n;
        g->generation.fetch_add(1, std::memory_order_release);
    }
    g->cv_job.notify_all();
    std::unique_lock<std::mutex> lk(g->m);
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
    for (int r = 0; r < g->n; r++)
        if (g->rc[r]) {
            g->err = std::string("GPU ") + std::to_string(g->dev[r]) + ": " + rt_last_error(g->ctx[r]);
            return g->rc[r];
        }
    return RT_OK;
}

extern "C" const char* rt_group_last_error(const rt_group* g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }
extern "C" int rt_group_size(const rt_group* g) { return g ? g->n : 0; }
extern "C" rt_context* rt_group_context(rt_group* g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }

extern "C" int rt_destroy_group(rt_group* g) {
    if (!g) return RT_OK;
    if (!g->workers.empty()) {
        {
            std::lock_guard<std::mutex> lk(g->m);
            g->stop = true;
        }
        g->cv_job.notify_all();
        for (auto& t : g->workers) t.join();
    }
    int prev = -1;
    cudaGetDevice(&prev);
    for (int r = 0; r < (int)g->ctx.size(); r++) {
        cudaSetDevice(g->dev[r]);
        if (g->ctx[r]) rt_destroy(g->ctx[r]);  // before the blob it borrows
        if (r < (int)g->d_blob.size() && g->d_blob[r]) cudaFree(g->d_blob[r]);
        if (r < (int)g->ev0.size() && g->ev0[r]) cudaEventDestroy(g->ev0[r]);
        if (r < (int)g->ev1.size() && g->ev1[r]) cudaEventDestroy(g->ev1[r]);
        if (r < (int)g->stream.size() && g->stream[r]) cudaStreamDestroy(g->stream[r]);
    }
    if (g->d_frame0) {
        cudaSetDevice(g->dev[0]);
        cudaFree(g->d_frame0);
    }
    if (!g->comms.empty() && g_nccl.CommDestroy)
        for (ncclComm_t c : g->comms)
            if (c) g_nccl.CommDestroy(c);
    if (prev >= 0) cudaSetDevice(prev);
    delete g;
    return RT_OK;
}

extern "C" int rt_create_group(int n_gpus, const int* device_ordinals, rt_group** out_group) {
    if (!out_group) return gerr(nullptr, RT_E_INVALID, "rt_create_group: out_group is NULL");
    *out_group = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return gerr(nullptr, RT_E_NO_DEVICE, "rt_create_group: no CUDA device; this library has no CPU fallback");
    if (n_gpus < 1 || n_gpus > count) return gerr(nullptr, RT_E_INVALID, "rt_create_group: %d GPUs requested, %d present", n_gpus, count);
    int prev = -1;
    cudaGetDevice(&prev);
    rt_group* g = new rt_group();
    g->n = n_gpus;
    g->dev.resize(n_gpus);
    g->ctx.assign(n_gpus, nullptr);
    g->stream.assign(n_gpus, nullptr);
    g->ev0.assign(n_gpus, nullptr);
    g->ev1.assign(n_gpus, nullptr);
    g->d_blob.assign(n_gpus, nullptr);
    g->rc.assign(n_gpus, 0);
    int rc = RT_OK;
    for (int r = 0; r < n_gpus && rc == RT_OK; r++) {
        g->dev[r] = device_ordinals ? device_ordinals[r] : r;
        for (int q = 0; q < r; q++)
            if (g->dev[q] == g->dev[r]) rc = gerr(nullptr, RT_E_INVALID, "rt_create_group: device %d listed twice", g->dev[r]);
        if (rc) break;
        if ((rc = rt_create(g->dev[r], &g->ctx[r]))) {
            gerr(nullptr, rc, "rt_create_group: %s", rt_last_error(nullptr));
            break;
        }
        cudaSetDevice(g->dev[r]);
        if (cudaStreamCreateWithFlags(&g->stream[r], cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&g->ev0[r]) != cudaSuccess ||
            cudaEventCreate(&g->ev1[r]) != cudaSuccess) {
            rc = gerr(nullptr, RT_E_CUDA, "rt_create_group: stream/event creation failed on device %d", g->dev[r]);
            break;
        }
        rt_set_stream(g->ctx[r], g->stream[r]);
    }
    if (prev >= 0) cudaSetDevice(prev);
    if (rc) {
        rt_destroy_group(g);
        return rc;
    }
    for (int r = 0; r < n_gpus; r++) g->workers.emplace_back(worker_main, g, r);
    *out_group = g;
    return RT_OK;
}

extern "C" int rt_group_set_option(rt_group* g, const char* name, int value) {
    if (!g || !name) return RT_E_INVALID;
    if (!strcmp(name, "band_rows")) {
        if (value < 4 || (value % 4)) return gerr(g, RT_E_INVALID, "band_rows must be a positive multiple of 4 (got %d)", value);
        g->band_rows = value;
        return RT_OK;
    }
    if (!strcmp(name, "broadcast")) {
        g->broadcast_mode = value ? 1 : 0;
        return RT_OK;
    }
    for (int r = 0; r < g->n; r++) {
        const int rc = rt_set_option(g->ctx[r], name, value);
        if (rc) return gerr(g, rc, "%s", rt_last_error(g->ctx[r]));
    }
    return RT_OK;
}

static int enable_peers(rt_group* g) {
    if (g->peers_enabled) return RT_OK;
    for (int r = 1; r < g->n; r++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, g->dev[r], g->dev[0]);
        if (!can) return gerr(g, RT_E_CUDA, "GPU %d cannot access GPU %d's memory (no peer access)", g->dev[r], g->dev[0]);
    }
    const int rc = run_all(g, [g](int r) {
        if (r == 0) return (int)RT_OK;
        const cudaError_t e = cudaDeviceEnablePeerAccess(g->dev[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int)RT_E_CUDA;
        cudaGetLastError();
        return (int)RT_OK;
    });
    if (rc) return gerr(g, rc, "cudaDeviceEnablePeerAccess failed");
    g->peers_enabled = true;
    return RT_OK;
}

extern "C" int rt_group_upload_scene(rt_group* g, const float* verts, int V, const int32_t* indices, int T, const void* nodes, int N,
                                     const int32_t* tri_indices, int R, const float* normals, int Vn, const int32_t* normal_indices,
                                     const void* materials, int M, const int32_t* tri_to_material) {
    if (!g) return RT_E_INVALID;
    int rc = rt_upload_scene(g->ctx[0], verts, V, indices, T, nodes, N, tri_indices, R, normals, Vn, normal_indices, materials, M, tri_to_material);
    if (rc) return gerr(g, rc, "%s", rt_last_error(g->ctx[0]));
    g->stats[RT_GROUP_STAT_BROADCAST_MS] = 0.0;
    g->stats[RT_GROUP_STAT_BLOB_BYTES] = 0.0;
    void* blob0 = nullptr;
    size_t bytes = 0;
    if ((rc = rt_scene_blob(g->ctx[0], &blob0, &bytes))) return gerr(g, rc, "%s", rt_last_error(g->ctx[0]));
    g->stats[RT_GROUP_STAT_BLOB_BYTES] = (double)bytes;
    if (g->n == 1) return RT_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    // receive buffers; a context keeps borrowing its previous blob until the new one is adopted, so the old ones are
    // released only at the end
    std::vector<void*> old_blob(g->n, nullptr);
    auto release_old = [&]() {
        int p = -1;
        cudaGetDevice(&p);
        for (int r = 1; r < g->n; r++)
            if (old_blob[r]) {
                cudaSetDevice(g->dev[r]);
                cudaFree(old_blob[r]);
            }
        if (p >= 0) cudaSetDevice(p);
    };
    // on a failure before the new blobs are adopted the contexts still borrow the old ones: put those back and drop the new
    std::vector<bool> swapped(g->n, false);
    auto abandon_new = [&]() {
        int p = -1;
        cudaGetDevice(&p);
        for (int r = 1; r < g->n; r++)
            if (swapped[r]) {
                cudaSetDevice(g->dev[r]);
                cudaStreamSynchronize(g->stream[r]);  // a collective may still be writing into it
                cudaFree(g->d_blob[r]);
                g->d_blob[r] = old_blob[r];
                old_blob[r] = nullptr;
                swapped[r] = false;
            }
        if (p >= 0) cudaSetDevice(p);
    };
    for (int r = 1; r < g->n; r++) {
        cudaSetDevice(g->dev[r]);
        rt_synchronize(g->ctx[r]);
        void* fresh = nullptr;
        if (cudaMalloc(&fresh, bytes) != cudaSuccess) {
            if (prev >= 0) cudaSetDevice(prev);
            abandon_new();
            return gerr(g, RT_E_CUDA, "cudaMalloc(%zu) for the scene blob failed on GPU %d", bytes, g->dev[r]);
        }
        old_blob[r] = g->d_blob[r];
        g->d_blob[r] = fresh;
        swapped[r] = true;
    }
    if (prev >= 0) cudaSetDevice(prev);
    rt_synchronize(g->ctx[0]);
    const auto t0 = std::chrono::steady_clock::now();
    if (g->broadcast_mode == 0) {
        std::lock_guard<std::mutex> lk(g_nccl_mutex);
        if (!g_nccl.load()) {
            abandon_new();
            return gerr(g, RT_E_CUDA, "%s; set group option \"broadcast\" = 1 for peer copies", g_nccl.problem.c_str());
        }
        if (g->comms.empty()) {
            g->comms.assign(g->n, nullptr);
            const ncclResult_t e = g_nccl.CommInitAll(g->comms.data(), g->n, g->dev.data());
            if (e != ncclSuccess) {
                g->comms.clear();
                abandon_new();
                return gerr(g, RT_E_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(e));
            }
        }
        const auto t1 = std::chrono::steady_clock::now();  // communicator setup is one-off and not part of the broadcast time
        g->stats[RT_GROUP_STAT_COMM_INIT_MS] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        cudaGetDevice(&prev);
        cudaSetDevice(g->dev[0]);
        cudaEventRecord(g->ev0[0], g->stream[0]);
        ncclResult_t e = g_nccl.GroupStart();
        for (int r = 0; r < g->n && e == ncclSuccess; r++)
            e = g_nccl.Broadcast(blob0, r == 0 ? blob0 : g->d_blob[r], bytes, ncclUint8, 0, g->comms[r], g->stream[r]);
        const ncclResult_t e2 = g_nccl.GroupEnd();
        if (e == ncclSuccess) e = e2;
        cudaEventRecord(g->ev1[0], g->stream[0]);
        if (prev >= 0) cudaSetDevice(prev);
        if (e != ncclSuccess) {
            abandon_new();
            return gerr(g, RT_E_CUDA, "ncclBroadcast failed: %s", g_nccl.GetErrorString(e));
        }
    } else {
        if ((rc = enable_peers(g))) {
            abandon_new();
            return rc;
        }
        cudaGetDevice(&prev);
        cudaSetDevice(g->dev[0]);
        cudaEventRecord(g->ev0[0], g->stream[0]);
        for (int r = 1; r < g->n; r++)
            if (cudaMemcpyPeerAsync(g->d_blob[r], g->dev[r], blob0, g->dev[0], bytes, g->stream[0]) != cudaSuccess) {
                cudaStreamSynchronize(g->stream[0]);
                if (prev >= 0) cudaSetDevice(prev);
                abandon_new();
                return gerr(g, RT_E_CUDA, "cudaMemcpyPeerAsync to GPU %d failed", g->dev[r]);
            }
        cudaEventRecord(g->ev1[0], g->stream[0]);
        if (prev >= 0) cudaSetDevice(prev);
    }
    // every rank: wait for its part of the collective (peer copies: all on GPU 0's stream), then adopt the blob
    cudaGetDevice(&prev);
    cudaSetDevice(g->dev[0]);
    const cudaError_t sync0 = cudaStreamSynchronize(g->stream[0]);
    if (prev >= 0) cudaSetDevice(prev);
    if (sync0 != cudaSuccess) {
        abandon_new();
        return gerr(g, RT_E_CUDA, "scene broadcast failed on GPU %d: %s", g->dev[0], cudaGetErrorString(sync0));
    }
    rc = run_all(g, [g, bytes](int r) {
        if (cudaStreamSynchronize(g->stream[r]) != cudaSuccess) return (int)RT_E_CUDA;
        return r == 0 ? (int)RT_OK : rt_adopt_scene_blob(g->ctx[r], g->d_blob[r], bytes);
    });
    if (rc) return rc;
    release_old();
    float ms = 0.0f;
    cudaGetDevice(&prev);
    cudaSetDevice(g->dev[0]);
    if (cudaEventElapsedTime(&ms, g->ev0[0], g->ev1[0]) == cudaSuccess) g->stats[RT_GROUP_STAT_BROADCAST_MS] = ms;
    if (prev >= 0) cudaSetDevice(prev);
    return RT_OK;
}

extern "C" int rt_group_set_params(rt_group* g, const float params[32]) {
    if (!g || !params) return RT_E_INVALID;
    for (int r = 0; r < g->n; r++) {
        const int rc = rt_set_params(g->ctx[r], params);
        if (rc) return gerr(g, rc, "%s", rt_last_error(g->ctx[r]));
    }
    return RT_OK;
}

// Is `host` page-locked and visible to every GPU of the group? (then the kernels store straight into it)
static bool zero_copy_target(rt_group* g, void* host, void** alias) {
    cudaPointerAttributes attr;
    memset(&attr, 0, sizeof attr);
    if (((uintptr_t)host & 3u) ||  // frame words are stored as 4-byte words: an odd destination goes through the copy instead
        cudaPointerGetAttributes(&attr, host) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
        cudaGetLastError();
        return false;
    }
    int unified = 1;
    for (int r = 0; r < g->n; r++) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrUnifiedAddressing, g->dev[r]);
        unified &= v;
    }
    if (!unified) return false;
    *alias = attr.devicePointer;
    return true;
}

enum TiledPass { PASS_FRAME = 0, PASS_PRIMARY = 1, PASS_PRIMARY_SHADOW = 2 };

static int tiled_pass(rt_group* g, int pass, int w, int h, void* out_host) {
    if (!g || !out_host || w <= 0 || h <= 0) return g ? gerr(g, RT_E_INVALID, "tiled pass: bad arguments") : RT_E_INVALID;
    const size_t bytes = (size_t)w * h * 4;
    void* dst = nullptr;
    const bool direct = zero_copy_target(g, out_host, &dst);
    int rc;
    if (!direct) {
        if (g->n > 1 && (rc = enable_peers(g))) return rc;
        if (g->frame0_bytes < bytes) {
            int prev = -1;
            cudaGetDevice(&prev);
            cudaSetDevice(g->dev[0]);
            if (g->d_frame0) cudaFree(g->d_frame0);
            g->d_frame0 = nullptr;
            g->frame0_bytes = 0;
            const cudaError_t e = cudaMalloc(&g->d_frame0, bytes);
            if (prev >= 0) cudaSetDevice(prev);
            if (e != cudaSuccess) return gerr(g, RT_E_CUDA, "cudaMalloc(%zu) for the gather frame failed", bytes);
            g->frame0_bytes = bytes;
        }
        dst = g->d_frame0;
    }
    const int band_rows = g->band_rows;
    const auto t0 = std::chrono::steady_clock::now();
    rc = run_all(g, [g, pass, w, h, dst, band_rows](int r) {
        rt_context* c = g->ctx[r];
        cudaEventRecord(g->ev0[r], g->stream[r]);
        int e;
        if (pass == PASS_FRAME) e = rt_render_frame_device(c, w, h, r, g->n, band_rows, (uint32_t*)dst);
        else if (pass == PASS_PRIMARY) e = rt_primary_gather_device(c, w, h, r, g->n, band_rows, nullptr, (int32_t*)dst);
        else e = rt_primary_shadow_device(c, w, h, r, g->n, band_rows, nullptr, nullptr, (int32_t*)dst);
        if (e) return e;
        cudaEventRecord(g->ev1[r], g->stream[r]);
        return rt_synchronize(c);
    });
    if (rc) return rc;
    if (!direct) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(g->dev[0]);
        const cudaError_t e = cudaMemcpy(out_host, g->d_frame0, bytes, cudaMemcpyDeviceToHost);
        if (prev >= 0) cudaSetDevice(prev);
        if (e != cudaSuccess) return gerr(g, RT_E_CUDA, "frame copy to the host failed: %s", cudaGetErrorString(e));
    }
    g->stats[RT_GROUP_STAT_LAST_CALL_MS] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    g->stats[RT_GROUP_STAT_ZERO_COPY] = direct ? 1.0 : 0.0;
    int prev = -1;
    cudaGetDevice(&prev);
    for (int r = 0; r < g->n && r < RT_GROUP_MAX_GPUS; r++) {
        float ms = 0.0f;
        cudaSetDevice(g->dev[r]);
        if (cudaEventElapsedTime(&ms, g->ev0[r], g->ev1[r]) != cudaSuccess) cudaGetLastError();
        g->stats[RT_GROUP_STAT_RANK_KERNEL_MS + r] = ms;
    }
    if (prev >= 0) cudaSetDevice(prev);
    return RT_OK;
}

extern "C" int rt_render_frame_tiled(rt_group* g, int w, int h, uint32_t* out_host) { return tiled_pass(g, PASS_FRAME, w, h, out_host); }
extern "C" int rt_primary_tiled(rt_group* g, int w, int h, int with_shadow, int32_t* frame_host) {
    return tiled_pass(g, with_shadow ? PASS_PRIMARY_SHADOW : PASS_PRIMARY, w, h, frame_host);
}

extern "C" int rt_group_stats(rt_group* g, double out[RT_GROUP_STATS]) {
    if (!g || !out) return RT_E_INVALID;
    memcpy(out, g->stats, sizeof g->stats);
    return RT_OK;
}
