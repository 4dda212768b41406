// rtb200.cu -- the extern "C" boundary of include/rtb200.h over the sm_100a kernels.
// Replaces the OpenCL context/queue/kernel plumbing of the reference (RayTracer.cpp:93-109,
// 858-1005, 1222-1287, 2050-2433). No CPU fallback: every entry point needs a CUDA device.
#include "../../include/rtb200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "kernels.cuh"
#include "kernels_lanes.cuh"
#include "scene_blob.h"

using namespace rtb;

static_assert(sizeof(rt_ray) == 32 && sizeof(rt_hit) == 16, "ABI record sizes");

// rt_context::d_counter is 32 x u64: [0] work-queue head of the context stream, [1..RT_FRAME_SLOTS] heads of the frame
// slots, [8] self-test mismatches, [16] traversals started (RT_CNT_RAYS_TRACED, accumulated by the kernels)
static const int kRayCounterSlot = 16;

struct rt_context {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // device->host copies that overlap tracing (rt_primary)
    cudaEvent_t chunk_events[16] = {nullptr};
    cudaStream_t out_stream = nullptr;    // device->host copies of rt_trace chunks
    cudaEvent_t out_events[16] = {nullptr};
    // frames in flight (rt_render_frame_begin / _end)
    // Row assembly scratch of one in-flight pass (TraceArgs::stage / group_count): a frame-shaped stage in local memory and
    // [work-queue head, 256 B][one arrival counter per tile group], zeroed by one memset per launch.
    struct RowAsm { void* stage = nullptr; size_t stage_bytes = 0; void* counts = nullptr; size_t counts_bytes = 0; };
    RowAsm rowasm;
    struct FrameSlot { cudaEvent_t done = nullptr; cudaEvent_t start = nullptr; cudaStream_t stream = nullptr; bool pending = false; void* d_out = nullptr; size_t d_bytes = 0; uint32_t* host = nullptr; size_t bytes = 0; RowAsm rowasm; } slots[RT_FRAME_SLOTS];
    void* d_sort = nullptr;               // ray-sorting scratch (keys, permutation, sorted rays, sorted hits, cub temp)
    size_t sort_bytes = 0;
    // scene
    uint8_t* d_blob = nullptr;
    size_t blob_bytes = 0;
    bool blob_owned = false;
    BlobHeader hdr;
    SceneView view;
    bool have_scene = false;
    // params
    ParamsBlock params;
    bool have_params = false;
    // scratch
    unsigned long long* d_counter = nullptr;
    void* d_stage_in = nullptr;
    size_t stage_in_bytes = 0;
    void* d_stage_out = nullptr;
    size_t stage_out_bytes = 0;
    void* d_scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int* d_flags = nullptr;
    int* d_offsets = nullptr;
    size_t flags_count = 0;
    // options
    int opt_smem_top = 0;     // number of top pairs staged in shared memory (0 = off)
    int opt_blocks_per_sm = 0;  // 0 = occupancy-derived
    int opt_top_pairs = 2047;   // BFS-ordered prefix chosen at pack time
    int opt_scheduler = -1;     // 0 = one 32-ray batch per warp (trace_kernel), 1 = persistent lanes (trace_lanes_kernel),
                                // -1 = auto: batch for in-kernel camera/shadow rays (coherent), lanes for ray buffers
    int opt_refill = 16;         // persistent lanes: refill when this many lanes are empty
    int opt_inner_exit = 8;     // persistent lanes: leave the inner phase when fewer lanes than this still descend
    int opt_fast_box = 0;       // 1 = approximate reciprocal/FMA box test for CLOSEST-hit batch kernels (NOT bit-exact; experiment)
    int opt_tile_order = 0;     // primary-ray tile order (see TraceArgs::tile_order)
    int opt_zero_copy = 1;      // rt_primary: store hits directly into pinned host memory when the destination is pinned
    int opt_exact_div = 0;      // 1 = always use the compiler's full division in the box test
    int opt_overlap_frames = 1; // rt_render_frame_begin: one-kernel frames on per-slot streams (frames in flight overlap)
    int opt_store_group = -1;   // row assembly of 4-byte/pixel frames: -1 = auto (on when the frame is host or peer memory),
                                // 0 = off, 2 = groups of 4 tiles (128-byte rows), 4 = groups of 16 tiles (512-byte rows)
    int opt_l2_warm = 0;        // bulk L2 prefetch of node pairs + packed triangles at the start of every traversal launch:
                                // 0 off, 1 pairs + triangles, 2 pairs only, 3 triangles only (kernels.cuh warm_l2)
    int opt_l2_warm_chunk_kb = 16;
    int opt_smem_carveout = -1;  // experiment: cudaFuncAttributePreferredSharedMemoryCarveout of the traversal kernels (percent of the
    bool carveout_touched = false;  // L1/shared array given to shared memory; -1 = the driver's choice, 0 = largest L1)
    int opt_l2_persist_kb = 0;  // experiment: L2 persisting access window over the first N KB of the node pairs (the BFS-ordered top)
    int opt_batch_inner_exit = -1;  // batch kernels: early exit from the inner loop (traverse.cuh INNER_EXIT): -1 = when the scene's
                                // traversal data (node pairs + triangles) does not fit L2, 0 = never, 1 = always
    size_t l2_bytes = 0;
    int opt_gate_cull = 1;      // camera-ray kernels: tiles outside the scene box's screen rectangle leave the queue (cull_setup)
    int opt_tile_hints = 1;     // temporal tile scheduling of the camera-ray kernels (kernels.cuh "tile scheduler")
    int opt_hint_heavy_pct = 12;                        // the slowest N percent of the tiles start first
    int opt_hint_light_pct = -1;                        // the quickest N percent run last; -1 = 35 (65 for the shaded-frame kernel,
                                                        // whose normal tiles last long enough to need more cover behind them)
    int opt_hint_split_pct = 80, opt_hint_keep_pct = 30;  // percent of the previous launch's span: split a tile / keep a row split
    // What a launch learnt about its tiles, kept per frame geometry for the next launch of the same geometry. Two buffers
    // per slot: launch k reads the one launch k-1 wrote and writes the other.
    struct HintSlot {
        int kind = -1, w = 0, h = 0, part = 0, n_parts = 0, band_tile_rows = 0, frame_slot = -2;
        long long num_batches = 0;
        unsigned int* buf[2] = {nullptr, nullptr};
        size_t bytes = 0;
        int cur = 0;          // buf[cur] = written by the last launch
        bool valid = false;   // buf[cur] holds hints
        uint64_t last_use = 0;
    } hint_slots[16];
    uint64_t hint_clock = 0;
    uint64_t counters[RT_CNT_COUNT] = {0};
#ifdef RTB_TIMELINE
    unsigned long long* d_timeline = nullptr;  // tools build only (see kernels.cuh)
#endif
    // resident blocks per SM of each kernel (occupancy query is a slow host call: done once per kernel and smem size)
    struct OccEntry { const void* fn; size_t smem; int per_sm; } occ_cache[32];
    int occ_count = 0;
    std::string err;
};

static thread_local std::string g_create_error;  // rt_create may run concurrently on several threads (one per GPU)

static int set_err(rt_context* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_error = buf;
    return code;
}

#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return set_err(ctx, RT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// Every entry point runs on its context's device and leaves the calling thread's current device as it found it: a
// process that also uses another CUDA framework, or drives several contexts from one thread, keeps its own device.
struct DeviceScope {
    int prev = -1;
    bool ok = false;
    explicit DeviceScope(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (prev == device) || cudaSetDevice(device) == cudaSuccess;
        if (prev == device) prev = -1;  // nothing to restore
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(ctx)                         \
    DeviceScope device_scope__((ctx)->device); \
    if (!device_scope__.ok) return set_err(ctx, RT_E_CUDA, "cudaSetDevice(%d) failed", (ctx)->device)

extern "C" const char* rt_version(void) { return "rtb200 0.2 (sm_100a)"; }

extern "C" const char* rt_last_error(const rt_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static int init_context(rt_context* ctx, int device_ordinal) {
    ctx->device = device_ordinal;
    DeviceScope scope(device_ordinal);
    if (!scope.ok) return set_err(nullptr, RT_E_CUDA, "cudaSetDevice(%d) failed", device_ordinal);
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, device_ordinal));
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->num_sms = prop.multiProcessorCount;
    CK(nullptr, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CK(nullptr, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(nullptr, cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
    for (auto& ev : ctx->chunk_events) CK(nullptr, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->out_events) CK(nullptr, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& sl : ctx->slots) {
        CK(nullptr, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        CK(nullptr, cudaEventCreateWithFlags(&sl.start, cudaEventDisableTiming));
        CK(nullptr, cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    }
    CK(nullptr, cudaMalloc(&ctx->d_counter, 256));
    CK(nullptr, cudaMemset(ctx->d_counter, 0, 256));
    memset(&ctx->hdr, 0, sizeof ctx->hdr);
    memset(&ctx->view, 0, sizeof ctx->view);
    return RT_OK;
}

extern "C" int rt_create(int device_ordinal, rt_context** out_ctx) {
    if (!out_ctx) return set_err(nullptr, RT_E_INVALID, "rt_create: out_ctx is NULL");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_err(nullptr, RT_E_NO_DEVICE, "rt_create: no CUDA device (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_ordinal < 0 || device_ordinal >= count)
        return set_err(nullptr, RT_E_INVALID, "rt_create: device %d out of range [0,%d)", device_ordinal, count);
    rt_context* ctx = new rt_context();
    const int rc = init_context(ctx, device_ordinal);
    if (rc) {  // the message is in g_create_error; release whatever was created
        rt_destroy(ctx);
        return rc;
    }
    *out_ctx = ctx;
    return RT_OK;
}

// Hints describe a frame of the scene that was current when they were recorded: a new scene or new thresholds start over.
static void forget_hints(rt_context* ctx) {
    for (auto& hs : ctx->hint_slots) hs.valid = false;
}

static void free_scene(rt_context* ctx) {
    forget_hints(ctx);
    if (ctx->d_blob && ctx->blob_owned) cudaFree(ctx->d_blob);
    ctx->d_blob = nullptr;
    ctx->blob_bytes = 0;
    ctx->blob_owned = false;
    ctx->have_scene = false;
}

extern "C" int rt_destroy(rt_context* ctx) {
    if (!ctx) return RT_OK;
    DeviceScope scope(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_scene(ctx);
    cudaFree(ctx->d_counter);
    cudaFree(ctx->d_stage_in);
    cudaFree(ctx->d_stage_out);
    cudaFree(ctx->d_scan_tmp);
    cudaFree(ctx->d_flags);
    cudaFree(ctx->d_offsets);
    for (auto& ev : ctx->chunk_events)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->out_events)
        if (ev) cudaEventDestroy(ev);
    for (auto& sl : ctx->slots) {
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.start) cudaEventDestroy(sl.start);
        if (sl.stream) cudaStreamDestroy(sl.stream);
        cudaFree(sl.d_out);
        cudaFree(sl.rowasm.stage);
        cudaFree(sl.rowasm.counts);
    }
    cudaFree(ctx->rowasm.stage);
    cudaFree(ctx->rowasm.counts);
    for (auto& hs : ctx->hint_slots) {
        cudaFree(hs.buf[0]);
        cudaFree(hs.buf[1]);
    }
    cudaFree(ctx->d_sort);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return RT_OK;
}

extern "C" int rt_set_stream(rt_context* ctx, void* cuda_stream) {
    if (!ctx) return RT_E_INVALID;
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    // The tile hints of one launch are read by the next launch of the same geometry; stream order is what makes that safe.
    // On a new stream nothing orders the two, and a half-written hint buffer could hide rows from the queue: start over.
    if (s != ctx->stream) forget_hints(ctx);
    ctx->stream = s;
    return RT_OK;
}

extern "C" int rt_synchronize(rt_context* ctx) {
    if (!ctx) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

// A blob may arrive from another process (rt_adopt_scene_blob after a broadcast): every section the kernels will index
// must lie inside the allocation before any pointer is derived from the header.
static const char* header_problem(const BlobHeader& h, size_t bytes) {
    if (h.num_pairs < 0 || h.num_tris < 0 || h.V < 0 || h.T < 0 || h.Vn < 0 || h.M < 0 || h.top_pairs < 0) return "negative count";
    struct { uint64_t off, size; } sec[] = {
        {h.off_pairs, 64ull * (uint64_t)h.num_pairs},       {h.off_tris, 48ull * ((uint64_t)h.num_tris + 1)},
        {h.off_verts, 16ull * (uint64_t)h.V},                {h.off_indices, 12ull * (uint64_t)h.T},
        {h.off_normals, 16ull * (uint64_t)h.Vn},             {h.off_normal_indices, h.Vn ? 12ull * (uint64_t)h.T : 0ull},
        {h.off_mat_diffuse, 16ull * (uint64_t)h.M},          {h.off_tri_to_material, h.M ? 4ull * (uint64_t)h.T : 0ull}};
    for (const auto& s : sec)
        if (s.off < sizeof(BlobHeader) || (s.off & 255u) || s.off > bytes || s.size > bytes - s.off) return "section outside the allocation";
    if (h.top_pairs > h.num_pairs) return "top_pairs exceeds num_pairs";
    if (h.root_ref >= 0 ? h.root_ref >= h.num_pairs : (h.root_ref != kRefPoison && ~h.root_ref > h.num_tris)) return "root reference out of range";
    return nullptr;
}

static int apply_l2_window(rt_context* ctx);
static int bind_blob(rt_context* ctx, const BlobHeader& h, uint8_t* d_blob, size_t bytes, bool owned) {
    if (h.magic != kBlobMagic || h.version != kBlobVersion || h.total_bytes != bytes)
        return set_err(ctx, RT_E_INVALID, "scene blob header mismatch (magic %08x version %u bytes %llu vs %zu)", h.magic,
                       h.version, (unsigned long long)h.total_bytes, bytes);
    if (const char* why = header_problem(h, bytes)) return set_err(ctx, RT_E_INVALID, "scene blob header rejected: %s", why);
    free_scene(ctx);
    ctx->d_blob = d_blob;
    ctx->blob_bytes = bytes;
    ctx->blob_owned = owned;
    ctx->hdr = h;
    SceneView& v = ctx->view;
    v.pairs = (const float4*)(d_blob + h.off_pairs);
    v.tris = (const float4*)(d_blob + h.off_tris);
    v.verts = (const float4*)(d_blob + h.off_verts);
    v.indices = (const int*)(d_blob + h.off_indices);
    v.normals = (const float4*)(d_blob + h.off_normals);
    v.normal_indices = (const int*)(d_blob + h.off_normal_indices);
    v.mat_diffuse = (const float4*)(d_blob + h.off_mat_diffuse);
    v.tri_to_material = (const int*)(d_blob + h.off_tri_to_material);
    v.root_ref = h.root_ref;
    v.num_pairs = h.num_pairs;
    v.top_pairs = h.top_pairs;
    v.coords_in_window = ctx->opt_exact_div ? 0 : h.coords_in_window;
    ctx->have_scene = true;
    if (ctx->opt_l2_persist_kb) return apply_l2_window(ctx);
    return RT_OK;
}

extern "C" int rt_upload_scene(rt_context* ctx, const float* verts, int V, const int32_t* indices, int T, const void* nodes,
                               int N, const int32_t* tri_indices, int R, const float* normals, int Vn,
                               const int32_t* normal_indices, const void* materials, int M, const int32_t* tri_to_material) {
    if (!ctx) return RT_E_INVALID;
    ON_DEVICE(ctx);
    SceneInputs in = {verts, V, indices, T, nodes, N, tri_indices, R, normals, Vn, normal_indices, materials, M, tri_to_material};
    uint8_t* blob = nullptr;
    uint64_t bytes = 0;
    char err[256] = {0};
    if (pack_scene(in, ctx->opt_top_pairs, &blob, &bytes, err, sizeof err)) return set_err(ctx, RT_E_INVALID, "%s", err);
    uint8_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d, blob, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    BlobHeader h;
    memcpy(&h, blob, sizeof h);
    free(blob);
    if (e != cudaSuccess) {
        if (d) cudaFree(d);
        return set_err(ctx, RT_E_CUDA, "rt_upload_scene: uploading %llu bytes failed: %s", (unsigned long long)bytes,
                       cudaGetErrorString(e));
    }
    ctx->counters[RT_CNT_H2D_BYTES] += bytes;
    return bind_blob(ctx, h, d, bytes, true);
}

extern "C" int rt_scene_blob(rt_context* ctx, void** out_device_ptr, size_t* out_bytes) {
    if (!ctx || !out_device_ptr || !out_bytes) return RT_E_INVALID;
    if (!ctx->have_scene) return set_err(ctx, RT_E_NO_SCENE, "rt_scene_blob: no scene uploaded");
    *out_device_ptr = ctx->d_blob;
    *out_bytes = ctx->blob_bytes;
    return RT_OK;
}

extern "C" int rt_copy_scene_blob(rt_context* ctx, void* dst_device_ptr, size_t bytes) {
    if (!ctx || !dst_device_ptr) return RT_E_INVALID;
    if (!ctx->have_scene) return set_err(ctx, RT_E_NO_SCENE, "rt_copy_scene_blob: no scene uploaded");
    if (bytes != ctx->blob_bytes) return set_err(ctx, RT_E_INVALID, "rt_copy_scene_blob: %zu bytes given, blob has %zu", bytes, ctx->blob_bytes);
    ON_DEVICE(ctx);
    CK(ctx, cudaMemcpyAsync(dst_device_ptr, ctx->d_blob, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

extern "C" int rt_adopt_scene_blob(rt_context* ctx, void* device_ptr, size_t bytes) {
    if (!ctx || !device_ptr || bytes < sizeof(BlobHeader)) return RT_E_INVALID;
    if ((uintptr_t)device_ptr & 255u)  // sections are 256-byte aligned relative to the base; the kernels use 32-byte loads
        return set_err(ctx, RT_E_INVALID, "rt_adopt_scene_blob: the blob must be 256-byte aligned (got %p)", device_ptr);
    ON_DEVICE(ctx);
    BlobHeader h;
    CK(ctx, cudaMemcpyAsync(&h, device_ptr, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = bind_blob(ctx, h, (uint8_t*)device_ptr, bytes, false);
    if (rc) return rc;
    // the header is consistent; now the contents it describes (see validate_blob_kernel)
    unsigned int* d_bad = (unsigned int*)(ctx->d_counter + 9);
    CK(ctx, cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), ctx->stream));
    validate_blob_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->view.pairs, h.num_pairs, ctx->view.tris, h.num_tris, h.Vn && h.M ? h.T : 0,
                                                                    h.root_ref, d_bad);
    CK(ctx, cudaGetLastError());
    unsigned int bad = 0;
    CK(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) {
        free_scene(ctx);
        return set_err(ctx, RT_E_INVALID, "scene blob rejected: %u child references / triangle records point outside their sections", bad);
    }
    return RT_OK;
}

extern "C" int rt_set_params(rt_context* ctx, const float params[32]) {
    if (!ctx || !params) return RT_E_INVALID;
    memcpy(&ctx->params, params, sizeof(ParamsBlock));
    ctx->have_params = true;
    return RT_OK;
}

// Experiment (VERDICT r1 task 8): mark the first opt_l2_persist_kb KB of the node pairs -- the breadth-first tree top when
// the scene was packed with a matching "top_pairs" -- as persisting in L2 for the work of the context stream.
static int apply_l2_window(rt_context* ctx) {
    if (!ctx->have_scene) return RT_OK;
    DeviceScope scope(ctx->device);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    size_t bytes = (size_t)ctx->opt_l2_persist_kb * 1024;
    const size_t pair_bytes = (size_t)ctx->hdr.num_pairs * 64;
    if (bytes > pair_bytes) bytes = pair_bytes;
    if (bytes) {
        int max_window = 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        if ((size_t)max_window < bytes) bytes = (size_t)max_window;
        size_t carve = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
        CK(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
        attr.accessPolicyWindow.base_ptr = (void*)ctx->view.pairs;
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = carve >= bytes ? 1.0f : (float)carve / (float)bytes;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    CK(ctx, cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (!bytes) cudaCtxResetPersistingL2Cache();
    return RT_OK;
}

extern "C" int rt_set_option(rt_context* ctx, const char* name, int value) {
    if (!ctx || !name) return RT_E_INVALID;
    if (!strcmp(name, "smem_top")) ctx->opt_smem_top = value < 0 ? 0 : value;
    else if (!strcmp(name, "blocks_per_sm")) ctx->opt_blocks_per_sm = value < 0 ? 0 : value;
    else if (!strcmp(name, "scheduler")) ctx->opt_scheduler = value < 0 ? -1 : (value ? 1 : 0);
    else if (!strcmp(name, "refill")) ctx->opt_refill = value < 1 ? 1 : (value > 32 ? 32 : value);
    else if (!strcmp(name, "inner_exit")) ctx->opt_inner_exit = value < 0 ? 0 : (value > 32 ? 32 : value);
    else if (!strcmp(name, "fast_box")) ctx->opt_fast_box = value ? 1 : 0;
    else if (!strcmp(name, "tile_order")) ctx->opt_tile_order = value < 0 ? 0 : value;
    else if (!strcmp(name, "zero_copy")) ctx->opt_zero_copy = value ? 1 : 0;
    else if (!strcmp(name, "exact_div")) {
        ctx->opt_exact_div = value ? 1 : 0;
        if (ctx->have_scene) ctx->view.coords_in_window = ctx->opt_exact_div ? 0 : ctx->hdr.coords_in_window;
    } else if (!strcmp(name, "overlap_frames")) ctx->opt_overlap_frames = value ? 1 : 0;
    else if (!strcmp(name, "store_group")) ctx->opt_store_group = (value == 0 || value == 2 || value == 4) ? value : -1;
    else if (!strcmp(name, "smem_carveout")) {
        ctx->opt_smem_carveout = value < 0 ? -1 : (value > 100 ? 100 : value);
        ctx->carveout_touched = true;
        ctx->occ_count = 0;  // re-applied at each kernel's next launch
    }
    else if (!strcmp(name, "l2_warm")) ctx->opt_l2_warm = (value >= 0 && value <= 3) ? value : 0;
    else if (!strcmp(name, "l2_warm_chunk_kb")) ctx->opt_l2_warm_chunk_kb = value < 1 ? 1 : (value > 1024 ? 1024 : value);
    else if (!strcmp(name, "l2_persist_kb")) { ctx->opt_l2_persist_kb = value < 0 ? 0 : value; return apply_l2_window(ctx); }
    else if (!strcmp(name, "gate_cull")) ctx->opt_gate_cull = value ? 1 : 0;
    else if (!strcmp(name, "inner_exit_batch")) ctx->opt_batch_inner_exit = value < 0 ? -1 : (value ? 1 : 0);
    else if (!strcmp(name, "tile_hints")) { ctx->opt_tile_hints = value ? 1 : 0; forget_hints(ctx); }
    else if (!strcmp(name, "hint_heavy_pct")) { ctx->opt_hint_heavy_pct = value < 1 ? 1 : value; forget_hints(ctx); }
    else if (!strcmp(name, "hint_light_pct")) { ctx->opt_hint_light_pct = value < 0 ? -1 : (value > 90 ? 90 : value); forget_hints(ctx); }
    else if (!strcmp(name, "hint_split_pct")) { ctx->opt_hint_split_pct = value < 1 ? 1 : value; forget_hints(ctx); }
    else if (!strcmp(name, "hint_keep_pct")) { ctx->opt_hint_keep_pct = value < 1 ? 1 : value; forget_hints(ctx); }
    else if (!strcmp(name, "top_pairs")) ctx->opt_top_pairs = value < 0 ? 0 : value;  // takes effect at the next upload
    else return set_err(ctx, RT_E_INVALID, "rt_set_option: unknown option '%s'", name);
    return RT_OK;
}

// What the most recent camera-ray launch recorded for its successor: [0] one-row entries on the split list, [1] tiles on the
// heavy list, [2] tiles timed, [3] the launch span in SM clock cycles. All zero when tile hints are off / nothing ran yet.
extern "C" int rt_tile_hint_stats(rt_context* ctx, uint64_t out[4]) {
    if (!ctx || !out) return RT_E_INVALID;
    out[0] = out[1] = out[2] = out[3] = 0;
    const rt_context::HintSlot* last = nullptr;
    for (const auto& hs : ctx->hint_slots)
        if (hs.valid && (!last || hs.last_use > last->last_use)) last = &hs;
    if (!last) return RT_OK;
    ON_DEVICE(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& sl : ctx->slots)
        if (sl.pending) CK(ctx, cudaEventSynchronize(sl.done));
    unsigned int hdr[16];
    CK(ctx, cudaMemcpy(hdr, last->buf[last->cur], sizeof hdr, cudaMemcpyDeviceToHost));
    out[0] = hdr[0];
    out[1] = hdr[1];
    out[2] = hdr[2];
    out[3] = hdr[3];
    return RT_OK;
}

extern "C" int rt_scene_info(rt_context* ctx, int64_t out[6]) {
    if (!ctx || !out) return RT_E_INVALID;
    if (!ctx->have_scene) return set_err(ctx, RT_E_NO_SCENE, "rt_scene_info: no scene uploaded");
    out[0] = ctx->hdr.num_pairs;
    out[1] = ctx->hdr.num_tris;
    out[2] = (int64_t)ctx->blob_bytes;
    out[3] = ctx->hdr.max_depth;
    out[4] = ctx->hdr.coords_in_window;
    out[5] = ctx->hdr.top_pairs;
    return RT_OK;
}

extern "C" int rt_get_counters(rt_context* ctx, uint64_t out[RT_CNT_COUNT]) {
    if (!ctx || !out) return RT_E_INVALID;
    ON_DEVICE(ctx);
    // RT_CNT_RAYS_TRACED lives on the device (the kernels count the traversals they start): wait for the work issued so far
    unsigned long long traced = 0;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& sl : ctx->slots)
        if (sl.pending) CK(ctx, cudaEventSynchronize(sl.done));
    CK(ctx, cudaMemcpy(&traced, ctx->d_counter + kRayCounterSlot, sizeof traced, cudaMemcpyDeviceToHost));
    ctx->counters[RT_CNT_RAYS_TRACED] = traced;
    memcpy(out, ctx->counters, sizeof ctx->counters);
    return RT_OK;
}
extern "C" int rt_reset_counters(rt_context* ctx) {
    if (!ctx) return RT_E_INVALID;
    ON_DEVICE(ctx);
    memset(ctx->counters, 0, sizeof ctx->counters);
    CK(ctx, cudaMemsetAsync(ctx->d_counter + kRayCounterSlot, 0, sizeof(unsigned long long), ctx->stream));
    return RT_OK;
}

// ---- launch helpers ------------------------------------------------------------------------------

static int ensure(rt_context* ctx, void** p, size_t* have, size_t need) {
    if (*have >= need) return RT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    CK(ctx, cudaMalloc(p, need));
    *have = need;
    return RT_OK;
}


template <typename K>
static int blocks_per_sm(rt_context* ctx, K kernel, size_t smem, int* out) {
    for (int i = 0; i < ctx->occ_count; i++)
        if (ctx->occ_cache[i].fn == (const void*)kernel && ctx->occ_cache[i].smem == smem) {
            *out = ctx->occ_cache[i].per_sm;
            return RT_OK;
        }
    if (smem > 48 * 1024) CK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (ctx->carveout_touched) CK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->opt_smem_carveout));
    int per_sm = 1;
    CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlockThreads, smem));
    if (per_sm < 1) per_sm = 1;
    if (ctx->occ_count < 32) ctx->occ_cache[ctx->occ_count++] = {(const void*)kernel, smem, per_sm};
    *out = per_sm;
    return RT_OK;
}

static void warm_setup(const rt_context* ctx, TraceArgs& a) {
    a.warm_base = nullptr;
    a.warm_bytes = 0;
    a.warm_chunk = (unsigned int)ctx->opt_l2_warm_chunk_kb * 1024u;
    if (!ctx->opt_l2_warm) return;
    const unsigned long long pairs_bytes = ctx->hdr.off_tris - ctx->hdr.off_pairs, tris_bytes = ctx->hdr.off_verts - ctx->hdr.off_tris;
    const char* pairs = (const char*)ctx->view.pairs;
    if (ctx->opt_l2_warm == 3) {
        a.warm_base = pairs + pairs_bytes;
        a.warm_bytes = tris_bytes;
    } else {
        a.warm_base = pairs;
        a.warm_bytes = ctx->opt_l2_warm == 2 ? pairs_bytes : pairs_bytes + tris_bytes;
    }
}

template <typename K, typename... Extra>
static int launch_persistent(rt_context* ctx, K kernel, TraceArgs& a, size_t smem, cudaStream_t stream, unsigned long long* counter,
                             size_t zero_bytes, Extra... extra) {
    if (!stream) stream = ctx->stream;
    if (!counter) {
        counter = ctx->d_counter;
        zero_bytes = sizeof(unsigned long long);
    }
    int per_sm = ctx->opt_blocks_per_sm;
    {
        int occ = 1, rc = blocks_per_sm(ctx, kernel, smem, &occ);  // also raises the dynamic-smem limit once
        if (rc) return rc;
        if (per_sm <= 0) per_sm = occ;
    }
    long long blocks = (long long)per_sm * ctx->num_sms;
    const long long needed = (a.num_batches + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (blocks > needed) blocks = needed;
    if (blocks < 1) return RT_OK;  // nothing to do
    a.work_counter = counter;
    a.ray_counter = ctx->d_counter + kRayCounterSlot;
    warm_setup(ctx, a);
#ifdef RTB_TIMELINE
    a.timeline = ctx->d_timeline;
#endif
    if (zero_bytes) CK(ctx, cudaMemsetAsync(counter, 0, zero_bytes, stream));  // queue head (+ the row-assembly arrival counters behind it)
    kernel<<<(unsigned)blocks, kBlockThreads, smem, stream>>>(a, extra...);
    CK(ctx, cudaGetLastError());
    ctx->counters[RT_CNT_KERNEL_LAUNCHES]++;
    return RT_OK;
}
// trace_kernel / render_kernel take (args, smem_count)
template <typename K>
static int launch_persistent(rt_context* ctx, K kernel, TraceArgs& a, int smem_count, cudaStream_t stream = nullptr,
                             unsigned long long* counter = nullptr, size_t zero_bytes = sizeof(unsigned long long)) {
    return launch_persistent(ctx, kernel, a, (size_t)smem_count * 64, stream, counter, zero_bytes, smem_count);
}

// ---- scene-box culling of the tile queue ---------------------------------------------------------------------------------
// A pixel traverses only if its ray passes the scene-AABB gate (vR.cl:1196); on the bench frame 64 % of the 8x4 tiles hold no
// such pixel, and fetching + testing them one by one was ~5 % of a launch's warp time (40 us of fixed cost for 1/8 of a 4K
// frame). The rays start ON the image plane and leave the eye, so every point they can reach is campos + lambda * (image_pos
// - campos) with lambda >= 1: the pixels whose rays can touch the box lie inside the bounding rectangle of the box corners
// projected through the eye onto the image plane -- provided every corner is in front of the plane. That rectangle, grown by
// two pixels (the device decides each pixel with fp32 arithmetic; its error is ~1e-6 of the frame) and to tile / row-assembly
// group boundaries, is what the queue enumerates; everything outside is written by store-only fill items. A camera inside or
// beside the box, a degenerate basis or non-finite numbers leave the rectangle at the whole frame.
static bool stores_to_host_memory(const TraceArgs& a) {
    auto in_host_memory = [](const void* p) {
        if (!p) return false;
        cudaPointerAttributes attr;
        memset(&attr, 0, sizeof attr);
        if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return attr.type == cudaMemoryTypeHost;
    };
    return in_host_memory(a.frame_out) || in_host_memory(a.hits_out) || in_host_memory(a.shadow_hits_out);
}

// Inclusive pixel bounds [x0, x1] x [y0, y1] (clamped to the frame; x0 > x1 or y0 > y1: nothing on screen) of every pixel
// whose ray can pass the scene gate, or false when no such bound is known (a box corner is not beyond the image plane).
static bool cull_pixel_rect(const ParamsBlock& P, int w, int h, long long out[4]) {
    const double A[3] = {P.a.x, P.a.y, P.a.z}, B[3] = {P.b.x, P.b.y, P.b.z}, Cc[3] = {P.c.x, P.c.y, P.c.z};
    const double E[3] = {P.campos.x, P.campos.y, P.campos.z};
    const double lo[3] = {P.aabb_min.x, P.aabb_min.y, P.aabb_min.z}, hi[3] = {P.aabb_max.x, P.aabb_max.y, P.aabb_max.z};
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int corner = 0; corner < 8; corner++) {
        double D[3], R[3];  // D = corner - eye, R = eye - c;  solve  A xf + B yf - D s = R
        for (int i = 0; i < 3; i++) {
            D[i] = ((corner >> i) & 1 ? hi[i] : lo[i]) - E[i];
            R[i] = E[i] - Cc[i];
        }
        auto det3 = [](const double* u, const double* v, const double* w3) {
            return u[0] * (v[1] * w3[2] - v[2] * w3[1]) - v[0] * (u[1] * w3[2] - u[2] * w3[1]) + w3[0] * (u[1] * v[2] - u[2] * v[1]);
        };
        const double nD[3] = {-D[0], -D[1], -D[2]};
        const double det = det3(A, B, nD);
        if (!(std::fabs(det) > 1e-300)) return false;
        const double xf = det3(R, B, nD) / det, yf = det3(A, R, nD) / det, sc = det3(A, B, R) / det;
        if (!(sc > 1e-9 && sc <= 1.0) || !std::isfinite(xf) || !std::isfinite(yf)) return false;  // corner not beyond the image plane
        const double px = xf * w + 0.5, py = yf * h + 0.5;  // xf = (x - 0.5) / w
        xmin = px < xmin ? px : xmin; xmax = px > xmax ? px : xmax;
        ymin = py < ymin ? py : ymin; ymax = py > ymax ? py : ymax;
    }
    if (!(xmax - xmin < 1e15) || !(ymax - ymin < 1e15)) return false;
    const double margin = 2.0;
    long long x0 = (long long)std::floor(xmin - margin), x1 = (long long)std::ceil(xmax + margin);
    long long y0 = (long long)std::floor(ymin - margin), y1 = (long long)std::ceil(ymax + margin);
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > w - 1) x1 = w - 1;
    if (y1 > h - 1) y1 = h - 1;
    out[0] = x0; out[1] = x1; out[2] = y0; out[3] = y1;
    return true;
}
// Host-only probe of the rectangle (tests/test_cull_rect.py checks it against the oracle's gate pixel by pixel).
extern "C" int rt_cull_rect_host(const float params[32], int w, int h, int64_t out_x0_x1_y0_y1[4]) {
    if (!params || !out_x0_x1_y0_y1 || w <= 0 || h <= 0) return RT_E_INVALID;
    ParamsBlock P;
    memcpy(&P, params, sizeof P);
    long long r[4];
    if (!cull_pixel_rect(P, w, h, r)) {  // unknown: the whole frame
        r[0] = 0; r[1] = w - 1; r[2] = 0; r[3] = h - 1;
    }
    for (int i = 0; i < 4; i++) out_x0_x1_y0_y1[i] = r[i];
    return RT_OK;
}

static void cull_setup(const rt_context* ctx, TraceArgs& a, bool frame_kernel = false) {
    const int tile_rows_total = (a.h + 3) / 4;
    const long long K = a.tiles_x ? a.num_batches / a.tiles_x : 0;  // this rank's tile rows
    a.in_tx0 = 0;
    a.in_tx1 = a.tiles_x;
    a.in_k0 = 0;
    a.in_k1 = (int)K;
    a.n_fill = 0;
    // (a pass that stores straight into host memory is paced by PCIe and wants its stores -- two thirds of them miss records
    // -- spread over the launch, not bunched in fill items at the end: it keeps the plain enumeration, like it keeps the plain queue)
    // The shaded frame is the exception: 4 bytes per pixel behind 0.7 ms of tracing are far from the link's limit, so it is
    // culled too, with the fill items FIRST -- their burst of (black) pixels crosses PCIe while the slow tiles trace.
    const bool host_sink = stores_to_host_memory(a);
    a.fill_first = (host_sink && frame_kernel) ? 1 : 0;
    if (!ctx->opt_gate_cull || a.rays_out || a.tile_order != 0 || K < 1 || (host_sink && !frame_kernel)) return;
    long long rect[4];
    if (!cull_pixel_rect(a.params, a.w, a.h, rect)) return;
    const long long x0 = rect[0], x1 = rect[1], y0 = rect[2], y1 = rect[3];
    int tx0 = 0, tx1 = 0, ty0 = 0, ty1 = 0;  // empty when the box is off screen
    if (x0 <= x1 && y0 <= y1) {
        tx0 = (int)(x0 / 8); tx1 = (int)(x1 / 8) + 1;
        ty0 = (int)(y0 / 4); ty1 = (int)(y1 / 4) + 1;
        if (a.group_log2) {  // whole row-assembly groups: a group is either traced (and assembled) or filled
            const int g = 1 << a.group_log2;
            tx0 = tx0 / g * g;
            tx1 = (tx1 + g - 1) / g * g;
            if (tx1 > a.tiles_x) tx1 = a.tiles_x;
        }
        if (ty1 > tile_rows_total) ty1 = tile_rows_total;
    }
    // this rank's tile rows are ascending in the frame: the ones inside [ty0, ty1) form one run [k0, k1)
    int k0 = (int)K, k1 = (int)K;
    for (long long k = 0; k < K; k++) {
        const long long bk = k / a.band_tile_rows;
        const long long row = ((long long)a.part + bk * a.n_parts) * a.band_tile_rows + (k - bk * a.band_tile_rows);
        if (row >= ty0 && row < ty1) {
            if (k0 == (int)K) k0 = (int)k;
            k1 = (int)k + 1;
        }
    }
    if (k0 == (int)K) k0 = k1 = 0;
    if (tx0 >= tx1) { tx0 = tx1 = 0; k0 = k1 = 0; }
    const long long inside = (long long)(k1 - k0) * (tx1 - tx0);
    if (inside * 10 > a.num_batches * 9) return;  // nothing worth culling: keep the plain enumeration
    a.in_tx0 = tx0; a.in_tx1 = tx1; a.in_k0 = k0; a.in_k1 = k1;
    a.n_fill = (unsigned int)(K * ((a.tiles_x + 31) / 32));
}

// ---- temporal tile scheduling: hint slots (device side: kernels.cuh "tile scheduler") -----------------------------------
enum { HINT_KIND_PRIMARY = 0, HINT_KIND_PRIMARY_SHADOW = 1, HINT_KIND_FRAME = 2 };
// Bind the hint buffers of this launch's frame geometry to `a`: hint_in = what the previous launch of the same geometry
// (same kernel kind, frame size, band partition and frame slot) recorded, hint_out = where this launch records. Frames in
// flight on different frame slots have their own slots, so two launches never share a buffer.
static int attach_hints(rt_context* ctx, int kind, TraceArgs& a, int frame_slot, cudaStream_t stream, rt_context::HintSlot** out_slot,
                        unsigned long long** counter, size_t* zero_bytes) {
    *out_slot = nullptr;
    a.hint_in = nullptr;
    a.hint_out = nullptr;
    a.hint_heavy_pct = ctx->opt_hint_heavy_pct;
    a.hint_light_pct = ctx->opt_hint_light_pct >= 0 ? ctx->opt_hint_light_pct : (kind == HINT_KIND_FRAME ? 65 : 35);
    a.hint_split_pct = ctx->opt_hint_split_pct;
    a.hint_keep_pct = ctx->opt_hint_keep_pct;
    if (!ctx->opt_tile_hints || a.tile_order != 0 || a.num_batches < 1 || a.num_batches > (1ll << 29)) return RT_OK;
    // A pass that stores straight into host memory is paced by the PCIe link, which wants the stores spread evenly over
    // the launch; starting the slow tiles first and ending on the quick ones bunches the stores at the end (measured:
    // rt_primary into pinned memory 0.71 -> 0.78 ms). Such passes keep the plain row-major queue.
    if (kind != HINT_KIND_FRAME && stores_to_host_memory(a)) return RT_OK;
    rt_context::HintSlot* hs = nullptr;
    rt_context::HintSlot* lru = &ctx->hint_slots[0];
    for (auto& c : ctx->hint_slots) {
        if (c.kind == kind && c.w == a.w && c.h == a.h && c.part == a.part && c.n_parts == a.n_parts && c.band_tile_rows == a.band_tile_rows &&
            c.frame_slot == frame_slot && c.num_batches == a.num_batches) {
            hs = &c;
            break;
        }
        if (c.last_use < lru->last_use) lru = &c;
    }
    const size_t bytes = sizeof(unsigned int) * ((size_t)kHintHeader + 7 * (size_t)a.num_batches);
    if (!hs) {  // a geometry not seen lately: take over the least recently used slot
        hs = lru;
        if (hs->bytes < bytes) {
            cudaFree(hs->buf[0]);  // (synchronises with launches that may still read them)
            cudaFree(hs->buf[1]);
            hs->buf[0] = hs->buf[1] = nullptr;
            hs->bytes = 0;
            CK(ctx, cudaMalloc((void**)&hs->buf[0], bytes));
            CK(ctx, cudaMalloc((void**)&hs->buf[1], bytes));
            hs->bytes = bytes;
        }
        hs->kind = kind; hs->w = a.w; hs->h = a.h; hs->part = a.part; hs->n_parts = a.n_parts; hs->band_tile_rows = a.band_tile_rows;
        hs->frame_slot = frame_slot; hs->num_batches = a.num_batches;
        hs->cur = 0;
        hs->valid = false;
    }
    hs->last_use = ++ctx->hint_clock;
    a.hint_in = hs->valid ? hs->buf[hs->cur] : nullptr;
    a.hint_out = hs->buf[hs->cur ^ 1];
    CK(ctx, cudaMemsetAsync(a.hint_out, 0, kHintHeader * sizeof(unsigned int), stream ? stream : ctx->stream));  // list counts, histogram
    // a pass without row-assembly counters keeps its work-queue head in the header just zeroed: one memset per launch, not two
    if (!a.group_log2) {
        *counter = reinterpret_cast<unsigned long long*>(a.hint_out + 8);
        *zero_bytes = 0;
    }
    *out_slot = hs;
    return RT_OK;
}
static void commit_hints(rt_context::HintSlot* hs, int rc) {
    if (!hs) return;
    if (rc == RT_OK) {
        hs->cur ^= 1;
        hs->valid = true;
    } else {
        hs->valid = false;
    }
}

// Destination of a pass's 4-byte/pixel frame. Decides whether rows are assembled (TraceArgs::stage): always when the
// `store_group` option says so; in auto mode when `dst` is page-locked host memory or memory of another GPU, where a
// tile's 32-byte row pieces would travel as quarter-filled PCIe / NVLink writes. Local frames are stored directly.
// `long_kernel`: the pass is the shaded frame (0.7 ms of tracing behind 4 bytes per pixel) -- alone on its link it gains nothing
// from assembled rows and pays for the group hand-shake (measured 0.757 vs 0.788 ms into pinned memory), so auto mode assembles
// its host frames only when several GPUs share the host's write bandwidth (n_parts > 1).
static int frame_sink(rt_context* ctx, TraceArgs& a, void* dst, rt_context::RowAsm& ra, unsigned long long** counter, size_t* zero_bytes,
                      bool long_kernel = false) {
    a.frame_out = (unsigned int*)dst;
    a.group_log2 = 0;
    if (!dst) return RT_OK;
    int mode = ctx->opt_store_group;
    if (mode < 0) {
        cudaPointerAttributes attr;
        memset(&attr, 0, sizeof attr);
        mode = 0;
        if (cudaPointerGetAttributes(&attr, dst) == cudaSuccess) {
            if (attr.type == cudaMemoryTypeHost) mode = (long_kernel && a.n_parts == 1) ? 0 : 4;
            else if (attr.type == cudaMemoryTypeDevice && attr.device != ctx->device) mode = 2;
        } else {
            cudaGetLastError();
        }
    }
    if (mode == 4 && (((uintptr_t)dst & 15u) || (a.w & 3))) mode = 2;  // 512-byte rows are stored as 16 bytes per lane
    if (!mode) return RT_OK;
    const int group_tiles = 1 << mode;
    const long long groups_x = (a.tiles_x + group_tiles - 1) / group_tiles;
    const long long tile_rows = a.tiles_x ? a.num_batches / a.tiles_x : 0;
    const size_t counts_bytes = 256 + (size_t)(tile_rows * groups_x) * sizeof(unsigned int);
    int rc;
    if ((rc = ensure(ctx, &ra.stage, &ra.stage_bytes, (size_t)a.w * a.h * 4))) return rc;
    if ((rc = ensure(ctx, &ra.counts, &ra.counts_bytes, counts_bytes))) return rc;
    a.stage = (unsigned int*)ra.stage;
    a.group_count = (unsigned int*)ra.counts + 64;
    a.group_log2 = mode;
    a.groups_x = (int)groups_x;
    *counter = (unsigned long long*)ra.counts;
    *zero_bytes = counts_bytes;
    return RT_OK;
}

template <typename K>
static int launch_lanes(rt_context* ctx, K kernel, TraceArgs& a, long long total_items) {
    int per_sm = ctx->opt_blocks_per_sm;
    if (per_sm <= 0) {
        int rc = blocks_per_sm(ctx, kernel, 0, &per_sm);
        if (rc) return rc;
    }
    long long blocks = (long long)per_sm * ctx->num_sms;
    const long long needed = (total_items + kBlockThreads - 1) / kBlockThreads;
    if (blocks > needed) blocks = needed;
    if (blocks < 1) return RT_OK;
    a.work_counter = ctx->d_counter;
    a.ray_counter = ctx->d_counter + kRayCounterSlot;
    warm_setup(ctx, a);
    CK(ctx, cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), ctx->stream));
    kernel<<<(unsigned)blocks, kBlockThreads, 0, ctx->stream>>>(a, ctx->opt_refill, ctx->opt_inner_exit);
    CK(ctx, cudaGetLastError());
    ctx->counters[RT_CNT_KERNEL_LAUNCHES]++;
    return RT_OK;
}

// Device alias of a pinned (page-locked) host buffer, or nullptr for pageable memory.
static void* pinned_alias(const void* host_ptr) {
    cudaPointerAttributes attr;
    memset(&attr, 0, sizeof attr);
    if (cudaPointerGetAttributes(&attr, host_ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        return attr.devicePointer;
    cudaGetLastError();  // pageable memory is not an error
    return nullptr;
}

// rt_ray / rt_hit buffers are read and written as 16-byte vectors and frames as 4-byte words, although the C structs only
// promise 4-byte alignment: a buffer carved at an odd offset would fault inside the kernel (a sticky CUDA error that takes
// the whole process' CUDA context with it), so it is refused on the host. nullptr (an optional buffer) passes.
static bool misaligned(const void* p, unsigned bytes) { return ((uintptr_t)p & (uintptr_t)(bytes - 1)) != 0; }
#define NEED_ALIGNED(ctx, who, ptr, bytes)                                                                                  \
    do {                                                                                                                    \
        if (misaligned((ptr), (bytes)))                                                                                     \
            return set_err((ctx), RT_E_INVALID, "%s: %s = %p is not %d-byte aligned", (who), #ptr, (const void*)(ptr), (int)(bytes)); \
    } while (0)

static int smem_top_count(const rt_context* ctx) {
    int c = ctx->opt_smem_top;
    if (c > ctx->hdr.top_pairs) c = ctx->hdr.top_pairs;
    if (c > 3400) c = 3400;  // 3400 * 64 B = 217.6 KB < 227 KB
    return c;
}

// Which instances of the traversal kernels run: the ones for scenes whose traversal data does not fit L2 (batch kernels:
// early exit from the inner loop, traverse.cuh INNER_EXIT; all of them: 256-byte L2 fills on node loads) or the plain ones.
static bool big_scene_instances(const rt_context* ctx) {
    if (ctx->opt_batch_inner_exit >= 0) return ctx->opt_batch_inner_exit != 0;
    const size_t traversal_bytes = 64 * (size_t)ctx->hdr.num_pairs + 48 * ((size_t)ctx->hdr.num_tris + 1);
    return ctx->l2_bytes && traversal_bytes > ctx->l2_bytes;
}

static int require(rt_context* ctx, bool params, bool shading) {
    if (!ctx) return RT_E_INVALID;
    if (!ctx->have_scene) return set_err(ctx, RT_E_NO_SCENE, "no scene uploaded (rt_upload_scene / rt_adopt_scene_blob)");
    if (params && !ctx->have_params) return set_err(ctx, RT_E_NO_PARAMS, "no params set (rt_set_params)");
    if (shading && (ctx->hdr.Vn == 0 || ctx->hdr.M == 0))
        return set_err(ctx, RT_E_INVALID, "scene was uploaded without normals/materials: shading is unavailable");
    return RT_OK;
}

static int band_setup(rt_context* ctx, TraceArgs& a, int w, int h, int part, int n_parts, int band_rows) {
    if (w <= 0 || h <= 0 || n_parts < 1 || part < 0 || part >= n_parts || band_rows < 4 || (band_rows % 4))
        return set_err(ctx, RT_E_INVALID, "bad frame/band arguments (w=%d h=%d part=%d/%d band_rows=%d; band_rows must be a multiple of 4)",
                       w, h, part, n_parts, band_rows);
    a.w = w;
    a.h = h;
    a.tiles_x = (w + 7) / 8;
    a.part = part;
    a.n_parts = n_parts;
    a.band_tile_rows = band_rows / 4;
    const long long bands = ((long long)h + band_rows - 1) / band_rows;
    const long long owned = bands > part ? (bands - part + n_parts - 1) / n_parts : 0;
    a.num_batches = owned * a.band_tile_rows * a.tiles_x;
    if (a.num_batches >= (1ll << 31) || (long long)a.tiles_x * ((h + 3) / 4) >= (1ll << 31))
        return set_err(ctx, RT_E_INVALID, "frame of %d x %d pixels has too many 8x4 tiles (the kernels index them in 32 bits)", w, h);
    a.tile_order = ctx->opt_tile_order;
    {   // an odd multiplier near 0.618 * count that is coprime to the count being permuted
        const unsigned long long cnt = a.tile_order == 3 ? (unsigned long long)(a.num_batches >> 6) : (unsigned long long)a.num_batches;
        unsigned long long m = (unsigned long long)(0.6180339887 * (double)cnt) | 1ull;
        auto gcd = [](unsigned long long x, unsigned long long y) { while (y) { unsigned long long t = x % y; x = y; y = t; } return x; };
        while (cnt > 1 && gcd(m, cnt) != 1) m += 2;
        a.order_mul = (unsigned int)(m ? m : 1);
    }
    return RT_OK;
}

static int do_trace_device(rt_context* ctx, int mode, long long n, const rt_ray* d_rays, rt_hit* d_hits) {
    NEED_ALIGNED(ctx, "trace", d_rays, 16);
    NEED_ALIGNED(ctx, "trace", d_hits, 16);
    TraceArgs a;
    memset(&a, 0, sizeof a);
    a.scene = ctx->view;
    a.n = n;
    a.num_batches = (n + 31) / 32;
    a.rays_in = (const float4*)d_rays;
    a.hits_out = (float4*)d_hits;
    const int st = smem_top_count(ctx);
    int rc;
    if (ctx->opt_scheduler != 0 && !st)
        rc = mode == RT_CLOSEST ? (big_scene_instances(ctx) ? launch_lanes(ctx, trace_lanes_kernel<SRC_BUFFER, false, true>, a, n) : launch_lanes(ctx, trace_lanes_kernel<SRC_BUFFER, false>, a, n))
                                : (big_scene_instances(ctx) ? launch_lanes(ctx, trace_lanes_kernel<SRC_BUFFER, true, true>, a, n) : launch_lanes(ctx, trace_lanes_kernel<SRC_BUFFER, true>, a, n));
    else if (mode == RT_CLOSEST && ctx->opt_fast_box && !st)
        rc = launch_persistent(ctx, trace_kernel<SRC_BUFFER, false, false, true>, a, 0);
    else if (mode == RT_CLOSEST)
        rc = st ? launch_persistent(ctx, trace_kernel<SRC_BUFFER, false, true>, a, st)
                : big_scene_instances(ctx) ? launch_persistent(ctx, trace_kernel<SRC_BUFFER, false, false, false, true>, a, 0)
                                      : launch_persistent(ctx, trace_kernel<SRC_BUFFER, false, false>, a, 0);
    else
        rc = st ? launch_persistent(ctx, trace_kernel<SRC_BUFFER, true, true>, a, st)
                : big_scene_instances(ctx) ? launch_persistent(ctx, trace_kernel<SRC_BUFFER, true, false, false, true>, a, 0)
                                      : launch_persistent(ctx, trace_kernel<SRC_BUFFER, true, false>, a, 0);
    return rc;
}

// Trace a ray buffer in coherence order: key generation -> radix sort -> gather -> trace -> scatter the hits back to
// the caller's order. Per-ray results are those of rt_trace_device, bit for bit; only the order of execution changes.
extern "C" int rt_trace_sorted_device(rt_context* ctx, int mode, int64_t n, const rt_ray* d_rays, rt_hit* d_hits) {
    int rc = require(ctx, false, false);
    if (rc) return rc;
    if ((mode != RT_CLOSEST && mode != RT_ANY) || n < 0 || n > 0x7fffffff || (n > 0 && (!d_rays || !d_hits)))
        return set_err(ctx, RT_E_INVALID, "rt_trace_sorted_device: bad arguments");
    NEED_ALIGNED(ctx, "rt_trace_sorted_device", d_rays, 16);
    NEED_ALIGNED(ctx, "rt_trace_sorted_device", d_hits, 16);
    ON_DEVICE(ctx);
    if (n == 0) return RT_OK;
    const size_t N = (size_t)n, npad = (N + 63) & ~(size_t)63;
    size_t tmp = 0;
    CK(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned int*)nullptr, (unsigned int*)nullptr, (unsigned int*)nullptr,
                                            (unsigned int*)nullptr, (int)n, 0, 30, ctx->stream));
    tmp = (tmp + 255) & ~(size_t)255;
    const size_t need = 16 * npad + 32 * npad + 16 * npad + tmp;  // 4 x u32 arrays, sorted rays, sorted hits, cub temp
    if ((rc = ensure(ctx, &ctx->d_sort, &ctx->sort_bytes, need))) return rc;
    uint8_t* base = (uint8_t*)ctx->d_sort;
    unsigned int *keys_in = (unsigned int*)base, *keys_out = keys_in + npad, *idx_in = keys_out + npad, *perm = idx_in + npad;
    float4* sorted_rays = (float4*)(base + 16 * npad);
    float4* sorted_hits = (float4*)(base + 48 * npad);
    void* cub_tmp = base + 64 * npad;
    const float3 lo = make_float3(ctx->hdr.root_min[0], ctx->hdr.root_min[1], ctx->hdr.root_min[2]);
    float3 inv;
    inv.x = 1.0f / fmaxf(ctx->hdr.root_max[0] - lo.x, 1e-20f);
    inv.y = 1.0f / fmaxf(ctx->hdr.root_max[1] - lo.y, 1e-20f);
    inv.z = 1.0f / fmaxf(ctx->hdr.root_max[2] - lo.z, 1e-20f);
    const int threads = 256;
    const unsigned blocks = (unsigned)((N + threads - 1) / threads);
    ray_sort_keys_kernel<<<blocks, threads, 0, ctx->stream>>>(n, (const float4*)d_rays, lo, inv, keys_in, idx_in);
    CK(ctx, cudaGetLastError());
    CK(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp, tmp, keys_in, keys_out, idx_in, perm, (int)n, 0, 30, ctx->stream));
    ray_gather_kernel<<<blocks, threads, 0, ctx->stream>>>(n, (const float4*)d_rays, perm, sorted_rays);
    CK(ctx, cudaGetLastError());
    if ((rc = do_trace_device(ctx, mode, n, (const rt_ray*)sorted_rays, (rt_hit*)sorted_hits))) return rc;
    hit_scatter_kernel<<<blocks, threads, 0, ctx->stream>>>(n, sorted_hits, perm, (float4*)d_hits);
    CK(ctx, cudaGetLastError());
    ctx->counters[RT_CNT_KERNEL_LAUNCHES] += 3;
    return RT_OK;
}

extern "C" int rt_trace_device(rt_context* ctx, int mode, int64_t n, const rt_ray* d_rays, rt_hit* d_hits) {
    int rc = require(ctx, false, false);
    if (rc) return rc;
    if ((mode != RT_CLOSEST && mode != RT_ANY) || n < 0 || (n > 0 && (!d_rays || !d_hits)))
        return set_err(ctx, RT_E_INVALID, "rt_trace_device: bad arguments");
    ON_DEVICE(ctx);
    if (n == 0) return RT_OK;
    return do_trace_device(ctx, mode, n, d_rays, d_hits);
}


// Host-buffer batch operator. The batch is cut into chunks: the upload of chunk c+1 (copy stream) overlaps the
// tracing of chunk c (context stream), and results go either straight into pinned host memory (zero copy) or
// through a third stream's device->host copies.
extern "C" int rt_trace(rt_context* ctx, int mode, int64_t n, const rt_ray* rays_host, rt_hit* hits_host) {
    int rc = require(ctx, false, false);
    if (rc) return rc;
    if ((mode != RT_CLOSEST && mode != RT_ANY) || n < 0 || (n > 0 && (!rays_host || !hits_host)))
        return set_err(ctx, RT_E_INVALID, "rt_trace: bad arguments");
    if (n == 0) return RT_OK;
    ON_DEVICE(ctx);
    if ((rc = ensure(ctx, &ctx->d_stage_in, &ctx->stage_in_bytes, (size_t)n * sizeof(rt_ray)))) return rc;
    // host buffers are only copied from / to (any alignment will do) unless the kernel stores into them directly
    rt_hit* direct_out = ctx->opt_zero_copy && !misaligned(hits_host, 16) ? (rt_hit*)pinned_alias(hits_host) : nullptr;
    if (!direct_out && (rc = ensure(ctx, &ctx->d_stage_out, &ctx->stage_out_bytes, (size_t)n * sizeof(rt_hit)))) return rc;
    const int64_t chunk_rays = 1 << 18;  // 8 MB of rays per chunk
    int chunks = (int)((n + chunk_rays - 1) / chunk_rays);
    if (chunks > 16) chunks = 16;
    const int64_t per = (((n + chunks - 1) / chunks) + 31) & ~(int64_t)31;
    // rays_host / hits_host are borrowed for the duration of this call only: on a failure part-way through, the copies
    // already enqueued must have finished before returning
    auto drain = [&]() {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->out_stream);
    };
#define CK_DRAIN(call)                                                                                                      \
    do {                                                                                                                    \
        cudaError_t e__ = (call);                                                                                           \
        if (e__ != cudaSuccess) {                                                                                           \
            drain();                                                                                                        \
            return set_err(ctx, RT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);  \
        }                                                                                                                   \
    } while (0)
    for (int c = 0; c < chunks; c++) {
        const int64_t lo = (int64_t)c * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        const rt_ray* d_in = (const rt_ray*)ctx->d_stage_in + lo;
        CK_DRAIN(cudaMemcpyAsync((void*)d_in, rays_host + lo, (size_t)(hi - lo) * sizeof(rt_ray), cudaMemcpyHostToDevice, ctx->copy_stream));
        CK_DRAIN(cudaEventRecord(ctx->chunk_events[c], ctx->copy_stream));
        CK_DRAIN(cudaStreamWaitEvent(ctx->stream, ctx->chunk_events[c], 0));
        rt_hit* d_out = direct_out ? direct_out + lo : (rt_hit*)ctx->d_stage_out + lo;
        if ((rc = do_trace_device(ctx, mode, hi - lo, d_in, d_out))) {
            drain();
            return rc;
        }
        if (!direct_out) {
            CK_DRAIN(cudaEventRecord(ctx->out_events[c], ctx->stream));
            CK_DRAIN(cudaStreamWaitEvent(ctx->out_stream, ctx->out_events[c], 0));
            CK_DRAIN(cudaMemcpyAsync(hits_host + lo, d_out, (size_t)(hi - lo) * sizeof(rt_hit), cudaMemcpyDeviceToHost, ctx->out_stream));
        }
    }
#undef CK_DRAIN
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (!direct_out) CK(ctx, cudaStreamSynchronize(ctx->out_stream));
    ctx->counters[RT_CNT_H2D_BYTES] += (uint64_t)n * sizeof(rt_ray);
    ctx->counters[RT_CNT_D2H_BYTES] += (uint64_t)n * sizeof(rt_hit);
    return RT_OK;
}

static int primary_impl(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits, rt_ray* d_rays_out,
                        int32_t* d_idx_frame) {
    int rc = require(ctx, true, false);
    if (rc) return rc;
    if (!d_hits && !d_idx_frame) return set_err(ctx, RT_E_INVALID, "primary pass: no output buffer given");
    NEED_ALIGNED(ctx, "primary pass", d_hits, 16);
    NEED_ALIGNED(ctx, "primary pass", d_rays_out, 16);
    NEED_ALIGNED(ctx, "primary pass", d_idx_frame, 4);
    ON_DEVICE(ctx);
    TraceArgs a;
    memset(&a, 0, sizeof a);
    a.scene = ctx->view;
    a.params = ctx->params;
    if ((rc = band_setup(ctx, a, w, h, part, n_parts, band_rows))) return rc;
    a.hits_out = (float4*)d_hits;
    a.rays_out = (float4*)d_rays_out;
    unsigned long long* counter = nullptr;
    size_t zero_bytes = 0;
    if ((rc = frame_sink(ctx, a, d_idx_frame, ctx->rowasm, &counter, &zero_bytes))) return rc;
    const int st = smem_top_count(ctx);
    if (ctx->opt_scheduler == 1 && !st && d_hits && !d_idx_frame)
        return (big_scene_instances(ctx) ? launch_lanes(ctx, trace_lanes_kernel<SRC_PRIMARY, false, true>, a, a.num_batches * 32) : launch_lanes(ctx, trace_lanes_kernel<SRC_PRIMARY, false>, a, a.num_batches * 32));
    cull_setup(ctx, a);
    rt_context::HintSlot* hs = nullptr;
    if ((rc = attach_hints(ctx, HINT_KIND_PRIMARY, a, -1, nullptr, &hs, &counter, &zero_bytes))) return rc;
    if (ctx->opt_fast_box && !st)
        rc = launch_persistent(ctx, trace_kernel<SRC_PRIMARY, false, false, true>, a, 0, nullptr, counter, zero_bytes);
    else
        rc = st ? launch_persistent(ctx, trace_kernel<SRC_PRIMARY, false, true>, a, st, nullptr, counter, zero_bytes)
                : big_scene_instances(ctx) ? launch_persistent(ctx, trace_kernel<SRC_PRIMARY, false, false, false, true>, a, 0, nullptr, counter, zero_bytes)
                                      : launch_persistent(ctx, trace_kernel<SRC_PRIMARY, false, false>, a, 0, nullptr, counter, zero_bytes);
    commit_hints(hs, rc);
    return rc;
}

// Primary + shadow in one launch (primary_shadow_kernel).
static int primary_shadow_impl(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                               rt_hit* d_shadow_hits, int32_t* d_vis_frame) {
    int rc = require(ctx, true, false);
    if (rc) return rc;
    if (!d_hits && !d_shadow_hits && !d_vis_frame) return set_err(ctx, RT_E_INVALID, "primary+shadow pass: no output buffer given");
    NEED_ALIGNED(ctx, "primary+shadow pass", d_hits, 16);
    NEED_ALIGNED(ctx, "primary+shadow pass", d_shadow_hits, 16);
    NEED_ALIGNED(ctx, "primary+shadow pass", d_vis_frame, 4);
    ON_DEVICE(ctx);
    TraceArgs a;
    memset(&a, 0, sizeof a);
    a.scene = ctx->view;
    a.params = ctx->params;
    if ((rc = band_setup(ctx, a, w, h, part, n_parts, band_rows))) return rc;
    a.hits_out = (float4*)d_hits;
    a.shadow_hits_out = (float4*)d_shadow_hits;
    unsigned long long* counter = nullptr;
    size_t zero_bytes = 0;
    if ((rc = frame_sink(ctx, a, d_vis_frame, ctx->rowasm, &counter, &zero_bytes))) return rc;
    cull_setup(ctx, a);
    rt_context::HintSlot* hs = nullptr;
    if ((rc = attach_hints(ctx, HINT_KIND_PRIMARY_SHADOW, a, -1, nullptr, &hs, &counter, &zero_bytes))) return rc;
    rc = big_scene_instances(ctx) ? launch_persistent(ctx, primary_shadow_kernel<true>, a, (size_t)0, nullptr, counter, zero_bytes)
                             : launch_persistent(ctx, primary_shadow_kernel<false>, a, (size_t)0, nullptr, counter, zero_bytes);
    commit_hints(hs, rc);
    return rc;
}

extern "C" int rt_primary_shadow_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                                        rt_hit* d_shadow_hits, int32_t* d_vis_frame) {
    return primary_shadow_impl(ctx, w, h, part, n_parts, band_rows, d_hits, d_shadow_hits, d_vis_frame);
}

// Host-buffer form: the 4-byte/pixel visibility frame of the fused primary + shadow pass. Page-locked destination: the
// kernel assembles full rows and stores them straight into host memory; pageable: device frame + one copy.
extern "C" int rt_primary_shadow(rt_context* ctx, int w, int h, int32_t* vis_host) {
    int rc = require(ctx, true, false);
    if (rc) return rc;
    if (!vis_host || w <= 0 || h <= 0) return set_err(ctx, RT_E_INVALID, "rt_primary_shadow: bad arguments");
    ON_DEVICE(ctx);
    const size_t bytes = (size_t)w * h * 4;
    void* alias = ctx->opt_zero_copy && !misaligned(vis_host, 4) ? pinned_alias(vis_host) : nullptr;
    if (alias) {
        if ((rc = primary_shadow_impl(ctx, w, h, 0, 1, 4, nullptr, nullptr, (int32_t*)alias))) return rc;
    } else {
        if ((rc = ensure(ctx, &ctx->d_stage_out, &ctx->stage_out_bytes, bytes))) return rc;
        if ((rc = primary_shadow_impl(ctx, w, h, 0, 1, 4, nullptr, nullptr, (int32_t*)ctx->d_stage_out))) return rc;
        CK(ctx, cudaMemcpyAsync(vis_host, ctx->d_stage_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->counters[RT_CNT_H2D_BYTES] += sizeof(ParamsBlock);
    ctx->counters[RT_CNT_D2H_BYTES] += bytes;
    return RT_OK;
}

extern "C" int rt_primary_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                                 rt_ray* d_rays_out) {
    if (ctx && !d_hits) return set_err(ctx, RT_E_INVALID, "rt_primary_device: d_hits is NULL");
    return primary_impl(ctx, w, h, part, n_parts, band_rows, d_hits, d_rays_out, nullptr);
}

extern "C" int rt_primary_gather_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                                        int32_t* d_idx_frame) {
    return primary_impl(ctx, w, h, part, n_parts, band_rows, d_hits, nullptr, d_idx_frame);
}

// ---- peer-visible framebuffer (CUDA IPC): one process per GPU, rank 0 owns the frame, the others map it ----
extern "C" int rt_ipc_alloc(rt_context* ctx, size_t bytes, void** out_device_ptr, unsigned char out_handle[64]) {
    if (!ctx || !out_device_ptr || !out_handle || !bytes) return RT_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    ON_DEVICE(ctx);
    void* p = nullptr;
    CK(ctx, cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return set_err(ctx, RT_E_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(out_handle, &h, 64);
    *out_device_ptr = p;
    return RT_OK;
}
extern "C" int rt_ipc_open(rt_context* ctx, const unsigned char handle[64], void** out_device_ptr) {
    if (!ctx || !handle || !out_device_ptr) return RT_E_INVALID;
    ON_DEVICE(ctx);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(ctx, cudaIpcOpenMemHandle(out_device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}
extern "C" int rt_ipc_close(rt_context* ctx, void* device_ptr) {
    if (!ctx || !device_ptr) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaIpcCloseMemHandle(device_ptr));
    return RT_OK;
}
extern "C" int rt_ipc_free(rt_context* ctx, void* device_ptr) {
    if (!ctx || !device_ptr) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaFree(device_ptr));
    return RT_OK;
}
// Page-lock caller-owned host memory (e.g. a POSIX shared-memory framebuffer that several ranks map) and return the
// device alias kernels can store into directly (zero copy over this GPU's own PCIe link).
extern "C" int rt_host_register(rt_context* ctx, void* host_ptr, size_t bytes, void** out_device_alias) {
    if (!ctx || !host_ptr || !bytes || !out_device_alias) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaHostRegister(host_ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    cudaError_t e = cudaHostGetDevicePointer(out_device_alias, host_ptr, 0);
    if (e != cudaSuccess) {
        cudaHostUnregister(host_ptr);
        return set_err(ctx, RT_E_CUDA, "cudaHostGetDevicePointer failed: %s", cudaGetErrorString(e));
    }
    return RT_OK;
}
extern "C" int rt_host_unregister(rt_context* ctx, void* host_ptr) {
    if (!ctx || !host_ptr) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaHostUnregister(host_ptr));
    return RT_OK;
}
// Frame-complete signal for a consumer that polls memory instead of joining a barrier: stream-ordered after the work
// enqueued so far, one 32-bit word (page-locked host memory registered with rt_host_register, or device memory).
__global__ void signal_word_kernel(volatile unsigned int* word, unsigned int value) {
    __threadfence_system();
    *word = value;
}
extern "C" int rt_signal(rt_context* ctx, void* d_word, uint32_t value) {
    if (!ctx || !d_word || ((uintptr_t)d_word & 3u)) return RT_E_INVALID;
    ON_DEVICE(ctx);
    signal_word_kernel<<<1, 1, 0, ctx->stream>>>((volatile unsigned int*)d_word, value);
    CK(ctx, cudaGetLastError());
    return RT_OK;
}
extern "C" int rt_memcpy_to_host(rt_context* ctx, void* dst_host, const void* src_device, size_t bytes) {
    if (!ctx || !dst_host || !src_device) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->counters[RT_CNT_D2H_BYTES] += bytes;
    return RT_OK;
}

// Host-buffer primary pass: the frame is cut into up to 8 contiguous row chunks; chunk c is traced on
// the context stream while chunk c-1 is copied to the host on the copy stream.
extern "C" int rt_primary(rt_context* ctx, int w, int h, rt_hit* hits_host) {
    int rc = require(ctx, true, false);
    if (rc) return rc;
    if (!hits_host || w <= 0 || h <= 0) return set_err(ctx, RT_E_INVALID, "rt_primary: bad arguments");
    ON_DEVICE(ctx);
    const size_t bytes = (size_t)w * h * sizeof(rt_hit);
    // Pinned (page-locked) destination: let the kernel store the hit records straight into host memory over
    // PCIe (zero copy) -- the transfer then overlaps the tracing completely and no staging copy exists.
    if (ctx->opt_zero_copy && !misaligned(hits_host, 16)) {
        if (void* alias = pinned_alias(hits_host)) {
            if ((rc = rt_primary_device(ctx, w, h, 0, 1, 4, (rt_hit*)alias, nullptr))) return rc;
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->counters[RT_CNT_H2D_BYTES] += sizeof(ParamsBlock);
            ctx->counters[RT_CNT_D2H_BYTES] += bytes;
            return RT_OK;
        }
    }
    if ((rc = ensure(ctx, &ctx->d_stage_out, &ctx->stage_out_bytes, bytes))) return rc;
    int chunks = 8;
    int rows = (((h + chunks - 1) / chunks) + 3) & ~3;  // rows per chunk, multiple of the 4-row warp tile
    chunks = (h + rows - 1) / rows;
    for (int c = 0; c < chunks; c++) {
        if ((rc = rt_primary_device(ctx, w, h, c, chunks, rows, (rt_hit*)ctx->d_stage_out, nullptr))) return rc;
        CK(ctx, cudaEventRecord(ctx->chunk_events[c], ctx->stream));
        CK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_events[c], 0));
        const int y0 = c * rows, y1 = (y0 + rows < h) ? y0 + rows : h;
        const size_t off = (size_t)y0 * w * sizeof(rt_hit), len = (size_t)(y1 - y0) * w * sizeof(rt_hit);
        CK(ctx, cudaMemcpyAsync((uint8_t*)hits_host + off, (uint8_t*)ctx->d_stage_out + off, len, cudaMemcpyDeviceToHost,
                                ctx->copy_stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->copy_stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->counters[RT_CNT_H2D_BYTES] += sizeof(ParamsBlock);
    ctx->counters[RT_CNT_D2H_BYTES] += bytes;
    return RT_OK;
}

extern "C" int rt_shadow_device(rt_context* ctx, int64_t n, const rt_ray* d_rays, const rt_hit* d_hits, rt_hit* d_shadow_hits,
                                rt_ray* d_shadow_rays_out) {
    int rc = require(ctx, true, false);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!d_rays || !d_hits || !d_shadow_hits))) return set_err(ctx, RT_E_INVALID, "rt_shadow_device: bad arguments");
    if (n == 0) return RT_OK;
    NEED_ALIGNED(ctx, "rt_shadow_device", d_rays, 16);
    NEED_ALIGNED(ctx, "rt_shadow_device", d_hits, 16);
    NEED_ALIGNED(ctx, "rt_shadow_device", d_shadow_hits, 16);
    NEED_ALIGNED(ctx, "rt_shadow_device", d_shadow_rays_out, 16);
    ON_DEVICE(ctx);
    TraceArgs a;
    memset(&a, 0, sizeof a);
    a.scene = ctx->view;
    a.params = ctx->params;
    a.n = n;
    a.num_batches = (n + 31) / 32;
    a.rays_in = (const float4*)d_rays;
    a.hits_in = (const float4*)d_hits;
    a.hits_out = (float4*)d_shadow_hits;
    a.rays_out = (float4*)d_shadow_rays_out;
    const int st = smem_top_count(ctx);
    if (ctx->opt_scheduler == 1 && !st)
        rc = (big_scene_instances(ctx) ? launch_lanes(ctx, trace_lanes_kernel<SRC_SHADOW, true, true>, a, n) : launch_lanes(ctx, trace_lanes_kernel<SRC_SHADOW, true>, a, n));
    else
        rc = st ? launch_persistent(ctx, trace_kernel<SRC_SHADOW, true, true>, a, st)
                : big_scene_instances(ctx) ? launch_persistent(ctx, trace_kernel<SRC_SHADOW, true, false, false, true>, a, 0)
                                      : launch_persistent(ctx, trace_kernel<SRC_SHADOW, true, false>, a, 0);
    return rc;
}

extern "C" int rt_diffuse_rays_device(rt_context* ctx, int64_t n, const rt_ray* d_rays, const rt_hit* d_hits, int spp,
                                      uint32_t seed, rt_ray* d_out_rays, int64_t* d_count) {
    int rc = require(ctx, false, false);
    if (rc) return rc;
    if (n <= 0 || n > 0x7fffffff || !d_rays || !d_hits || spp < 1 || !d_out_rays)
        return set_err(ctx, RT_E_INVALID, "rt_diffuse_rays_device: bad arguments");
    NEED_ALIGNED(ctx, "rt_diffuse_rays_device", d_rays, 16);
    NEED_ALIGNED(ctx, "rt_diffuse_rays_device", d_hits, 16);
    NEED_ALIGNED(ctx, "rt_diffuse_rays_device", d_out_rays, 16);
    NEED_ALIGNED(ctx, "rt_diffuse_rays_device", d_count, 8);
    ON_DEVICE(ctx);
    if (ctx->flags_count < (size_t)n) {
        cudaFree(ctx->d_flags);
        cudaFree(ctx->d_offsets);
        ctx->d_flags = ctx->d_offsets = nullptr;
        ctx->flags_count = 0;
        CK(ctx, cudaMalloc(&ctx->d_flags, (size_t)n * 4));
        CK(ctx, cudaMalloc(&ctx->d_offsets, (size_t)n * 4));
        ctx->flags_count = (size_t)n;
    }
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    diffuse_flags_kernel<<<blocks, threads, 0, ctx->stream>>>(n, (const float4*)d_hits, ctx->d_flags);
    CK(ctx, cudaGetLastError());
    size_t tmp = 0;
    CK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_flags, ctx->d_offsets, (int)n, ctx->stream));
    if ((rc = ensure(ctx, &ctx->d_scan_tmp, &ctx->scan_tmp_bytes, tmp ? tmp : 1))) return rc;
    CK(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_scan_tmp, tmp, ctx->d_flags, ctx->d_offsets, (int)n, ctx->stream));
    diffuse_rays_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->view, n, (const float4*)d_rays, (const float4*)d_hits,
                                                             ctx->d_offsets, spp, seed, (float4*)d_out_rays,
                                                             (long long*)d_count);
    CK(ctx, cudaGetLastError());
    ctx->counters[RT_CNT_KERNEL_LAUNCHES] += 2;
    return RT_OK;
}

// `slot` >= 0: the frame runs on that frame slot's own stream, work-queue head and row-assembly scratch (frames in flight
// overlap on the GPU); -1: on the context stream.
static int render_frame_impl(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, uint32_t* d_out, int slot) {
    int rc = require(ctx, true, true);
    if (rc) return rc;
    if (!d_out) return set_err(ctx, RT_E_INVALID, "rt_render_frame_device: d_out is NULL");
    NEED_ALIGNED(ctx, "rt_render_frame_device", d_out, 4);
    ON_DEVICE(ctx);
    TraceArgs a;
    memset(&a, 0, sizeof a);
    a.scene = ctx->view;
    a.params = ctx->params;
    if ((rc = band_setup(ctx, a, w, h, part, n_parts, band_rows))) return rc;
    cudaStream_t stream = slot >= 0 ? ctx->slots[slot].stream : nullptr;
    unsigned long long* counter = slot >= 0 ? ctx->d_counter + 1 + slot : nullptr;
    size_t zero_bytes = sizeof(unsigned long long);
    if ((rc = frame_sink(ctx, a, d_out, slot >= 0 ? ctx->slots[slot].rowasm : ctx->rowasm, &counter, &zero_bytes, true))) return rc;
    const int st = smem_top_count(ctx);
    cull_setup(ctx, a, true);
    rt_context::HintSlot* hs = nullptr;
    if ((rc = attach_hints(ctx, HINT_KIND_FRAME, a, slot, stream, &hs, &counter, &zero_bytes))) return rc;
    rc = st ? launch_persistent(ctx, render_kernel<true>, a, st, stream, counter, zero_bytes)
            : big_scene_instances(ctx) ? launch_persistent(ctx, render_kernel<false, true>, a, 0, stream, counter, zero_bytes)
                                  : launch_persistent(ctx, render_kernel<false>, a, 0, stream, counter, zero_bytes);
    commit_hints(hs, rc);
    return rc;
}

extern "C" int rt_render_frame_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, uint32_t* d_out) {
    return render_frame_impl(ctx, w, h, part, n_parts, band_rows, d_out, -1);
}

extern "C" int rt_render_frame(rt_context* ctx, int w, int h, uint32_t* out_host) {
    int rc = require(ctx, true, true);
    if (rc) return rc;
    if (!out_host || w <= 0 || h <= 0) return set_err(ctx, RT_E_INVALID, "rt_render_frame: bad arguments");
    ON_DEVICE(ctx);
    const size_t bytes = (size_t)w * h * 4;
    void* alias = ctx->opt_zero_copy && !misaligned(out_host, 4) ? pinned_alias(out_host) : nullptr;
    if (alias) {  // pinned destination: the kernel stores the pixels straight into host memory
        if ((rc = rt_render_frame_device(ctx, w, h, 0, 1, 4, (uint32_t*)alias))) return rc;
    } else {
        if ((rc = ensure(ctx, &ctx->d_stage_out, &ctx->stage_out_bytes, bytes))) return rc;
        if ((rc = rt_render_frame_device(ctx, w, h, 0, 1, 4, (uint32_t*)ctx->d_stage_out))) return rc;
        CK(ctx, cudaMemcpyAsync(out_host, ctx->d_stage_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->counters[RT_CNT_H2D_BYTES] += sizeof(ParamsBlock);
    ctx->counters[RT_CNT_D2H_BYTES] += bytes;
    return RT_OK;
}

// Pipelined frames: begin enqueues, end waits. Each launch carries its Params block by value, so rt_set_params for
// frame k+1 may be called while frame k is still in flight.
extern "C" int rt_render_frame_begin(rt_context* ctx, int w, int h, uint32_t* out_host, int slot) {
    int rc = require(ctx, true, true);
    if (rc) return rc;
    if (!out_host || w <= 0 || h <= 0) return set_err(ctx, RT_E_INVALID, "rt_render_frame_begin: bad arguments");
    if (slot < 0 || slot >= RT_FRAME_SLOTS) return set_err(ctx, RT_E_INVALID, "rt_render_frame_begin: slot %d outside 0..%d", slot, RT_FRAME_SLOTS - 1);
    rt_context::FrameSlot& sl = ctx->slots[slot];
    if (sl.pending) return set_err(ctx, RT_E_INVALID, "rt_render_frame_begin: slot %d still has a frame in flight (call rt_render_frame_end first)", slot);
    ON_DEVICE(ctx);
    const size_t bytes = (size_t)w * h * 4;
    // Frames run on the slot's own stream with the slot's own work counter, ordered after everything already enqueued on
    // the context stream: the ramp-down of frame k (12-19 % of a launch) overlaps the start of frame k+1.
    const bool own = ctx->opt_overlap_frames != 0;
    if (own) {
        CK(ctx, cudaEventRecord(sl.start, ctx->stream));
        CK(ctx, cudaStreamWaitEvent(sl.stream, sl.start, 0));
    }
    void* alias = ctx->opt_zero_copy && !misaligned(out_host, 4) ? pinned_alias(out_host) : nullptr;
    if (alias) {
        if ((rc = render_frame_impl(ctx, w, h, 0, 1, 4, (uint32_t*)alias, own ? slot : -1))) return rc;
        sl.host = nullptr;
    } else {  // pageable destination: per-slot device frame, copied out by rt_render_frame_end
        if ((rc = ensure(ctx, &sl.d_out, &sl.d_bytes, bytes))) return rc;
        if ((rc = render_frame_impl(ctx, w, h, 0, 1, 4, (uint32_t*)sl.d_out, own ? slot : -1))) return rc;
        sl.host = out_host;
    }
    sl.bytes = bytes;
    CK(ctx, cudaEventRecord(sl.done, own ? sl.stream : ctx->stream));
    sl.pending = true;
    return RT_OK;
}

extern "C" int rt_render_frame_end(rt_context* ctx, int slot) {
    if (!ctx) return RT_E_INVALID;
    if (slot < 0 || slot >= RT_FRAME_SLOTS) return set_err(ctx, RT_E_INVALID, "rt_render_frame_end: slot %d outside 0..%d", slot, RT_FRAME_SLOTS - 1);
    rt_context::FrameSlot& sl = ctx->slots[slot];
    if (!sl.pending) return set_err(ctx, RT_E_INVALID, "rt_render_frame_end: no frame in flight in slot %d", slot);
    ON_DEVICE(ctx);
    CK(ctx, cudaEventSynchronize(sl.done));  // on failure the slot stays pending: the caller may retry or destroy the context
    if (sl.host) CK(ctx, cudaMemcpy(sl.host, sl.d_out, sl.bytes, cudaMemcpyDeviceToHost));
    sl.pending = false;
    ctx->counters[RT_CNT_H2D_BYTES] += sizeof(ParamsBlock);
    ctx->counters[RT_CNT_D2H_BYTES] += sl.bytes;
    return RT_OK;
}

// GPU self test of the hoisted division (device_math.cuh): `samples` random operand pairs inside the
// admitted window, bitwise comparison against the compiler's IEEE division.
extern "C" int rt_selftest(rt_context* ctx, int64_t samples, uint32_t seed, uint64_t* out_mismatches) {
    if (!ctx || !out_mismatches || samples < 1) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaMemsetAsync(ctx->d_counter + 8, 0, sizeof(unsigned long long), ctx->stream));
    const int threads = 256, iters = 4096;
    long long blocks = (samples + (long long)threads * iters - 1) / ((long long)threads * iters);
    if (blocks < 1) blocks = 1;
    selftest_division_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(seed, iters, ctx->d_counter + 8);
    CK(ctx, cudaGetLastError());
    unsigned long long bad = 0;
    CK(ctx, cudaMemcpyAsync(&bad, ctx->d_counter + 8, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    *out_mismatches = bad;
    return RT_OK;
}

// Range probe of the hoisted division: numerator exponent fixed to `x_exponent` (unbiased, < -126 = denormal).
extern "C" int rt_selftest_range(rt_context* ctx, int64_t samples, uint32_t seed, int x_exponent, uint64_t* out_mismatches) {
    if (!ctx || !out_mismatches || samples < 1) return RT_E_INVALID;
    ON_DEVICE(ctx);
    CK(ctx, cudaMemsetAsync(ctx->d_counter + 8, 0, sizeof(unsigned long long), ctx->stream));
    const int threads = 256, iters = 1024;
    long long blocks = (samples + (long long)threads * iters - 1) / ((long long)threads * iters);
    if (blocks < 1) blocks = 1;
    selftest_division_range_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(seed, iters, x_exponent, ctx->d_counter + 8);
    CK(ctx, cudaGetLastError());
    unsigned long long bad = 0;
    CK(ctx, cudaMemcpyAsync(&bad, ctx->d_counter + 8, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    *out_mismatches = bad;
    return RT_OK;
}

// Exhaustive self test: all 2^23 numerator mantissas against `md_count` divisor mantissas starting at md_begin, at the
// exponents (ex_x, ex_d) and signs (bit 0: divisor negative, bit 1: numerator negative). The full square is 2^23
// divisor mantissas; callers split it into chunks (tools/div_exhaustive.py) so that a launch stays short.
extern "C" int rt_selftest_exhaustive(rt_context* ctx, uint32_t md_begin, uint32_t md_count, int ex_x, int ex_d, uint32_t signs,
                                      uint64_t* out_mismatches) {
    if (!ctx || !out_mismatches) return RT_E_INVALID;
    if (md_count < 1 || md_count > 65535u || md_begin > (1u << 23) - md_count || ex_x < -100 || ex_x > 60 || ex_d < -20 || ex_d > 19)
        return set_err(ctx, RT_E_INVALID, "rt_selftest_exhaustive: chunk of 1..65535 divisor mantissas inside [0, 2^23); exponents inside the window");
    ON_DEVICE(ctx);
    CK(ctx, cudaMemsetAsync(ctx->d_counter + 8, 0, sizeof(unsigned long long), ctx->stream));
    selftest_division_exhaustive_kernel<<<dim3(16, md_count), 256, 0, ctx->stream>>>(md_begin, ex_x, ex_d, signs, ctx->d_counter + 8);
    CK(ctx, cudaGetLastError());
    unsigned long long bad = 0;
    CK(ctx, cudaMemcpyAsync(&bad, ctx->d_counter + 8, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    *out_mismatches = bad;
    return RT_OK;
}

#ifdef RTB_TIMELINE
// tools build only (make timeline): per-batch start/end timestamps of the persistent-warp kernels go to d_buf
// (2 x uint64 per batch of the next launches; NULL = off)
extern "C" int rt_debug_timeline(rt_context* ctx, void* d_buf) {
    if (!ctx) return RT_E_INVALID;
    ctx->d_timeline = (unsigned long long*)d_buf;
    return RT_OK;
}
#endif

// Host-only: validate + pack the reference arrays into the blob layout WITHOUT touching a device (the packing is
// what rt_upload_scene uploads). For tools and the CPU tests of the layout; free the result with rt_free_host.
extern "C" int rt_pack_scene_host(const float* verts, int V, const int32_t* indices, int T, const void* nodes, int N,
                                  const int32_t* tri_indices, int R, const float* normals, int Vn, const int32_t* normal_indices,
                                  const void* materials, int M, const int32_t* tri_to_material, int top_pairs, void** out_blob,
                                  size_t* out_bytes, char* err, int err_len) {
    if (!out_blob || !out_bytes) return RT_E_INVALID;
    SceneInputs in = {verts, V, indices, T, nodes, N, tri_indices, R, normals, Vn, normal_indices, materials, M, tri_to_material};
    uint8_t* blob = nullptr;
    uint64_t bytes = 0;
    if (pack_scene(in, top_pairs, &blob, &bytes, err, err_len)) return RT_E_INVALID;
    *out_blob = blob;
    *out_bytes = (size_t)bytes;
    return RT_OK;
}
extern "C" void rt_free_host(void* p) { free(p); }
