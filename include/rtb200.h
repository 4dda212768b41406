/* rtb200.h -- C ABI of the B200-native ray-tracing hot path (librtb200.so).
 *
 * This boundary replaces the OpenCL host<->device seam of the reference
 * (itmanager85/real-time-opencl-raytracer); each entry point cites what it stands in for.
 * File names are relative to the reference tree; "vR.cl" = x64/Release/volumeRender.cl.
 *
 *   reference call (RayTracer.cpp)                         replacement
 *   ------------------------------------------------------ ---------------------------
 *   setupCL(): context, queue, program, kernel :2370-2433   rt_create
 *   11 x clCreateBuffer(COPY_HOST_PTR)         :942-984     rt_upload_scene
 *   17 x clSetKernelArg                        :1228-1260   (bound inside the context)
 *   clEnqueueWriteBuffer(params, 128 B)        :671         rt_set_params
 *   clEnqueueNDRangeKernel + clFinish +
 *   clEnqueueReadBuffer(w*h*4)                 :332-343     rt_render_frame
 *   traverse_bvh(Ray*, tHit, ..., needClosestHit) vR.cl:658 rt_trace  (batch operator)
 *   ~RayTraceData, cleanup()                   :208-228,1263-1287  rt_destroy
 *
 * Conventions: every function returns 0 on success (the reference's SDK_SUCCESS) and a non-zero
 * RT_E_* code on failure; nothing throws or exits; rt_last_error() gives the text. Host pointers
 * are borrowed for the duration of the call only. A context is used from one host thread at a
 * time; different contexts (one per GPU) may be driven concurrently. There is NO CPU fallback:
 * without a CUDA device rt_create fails with RT_E_NO_DEVICE.
 *
 * All scene inputs are in the REFERENCE layouts (no re-packing required from the caller):
 *   verts          float4[V], w = 1                         Mesh::vertices            (Mesh.h:73)
 *   indices        int[3T]                                  Mesh::indices             (Mesh.h:72)
 *   nodes          48 B[N] {float4 min, float4 max, int left, right, tri_off, tri_cnt}
 *                                                           BVH_Cuda::bvh_nodes       (BVH_Cuda.h:12-29)
 *   tri_indices    int[R], each = 3 * triangle id           BVH_Cuda::tri_indices     (BVH_Cuda.h:90-93)
 *   normals        float4[Vn]; normal_indices int[3T]       Mesh::normals(_indices)   (Mesh.h:76-77)
 *   materials      176 B[M] {int4 technique, 10 x float4}   Mesh::materials           (Mesh.h:20-33)
 *   tri_to_material int[T]                                  Mesh::triangle_index_to_material_index
 *   params         8 x float4 {a,b,c,campos,light_pos,light_color,aabb_min,aabb_max}  (RayTracer.cpp:115-161)
 *
 * Results are bit-identical to the reference traversal as restated in oracle/oracle.c: same hit
 * index (3 * triangle id, or -1), same t, and the Moller-Trumbore u,v of the accepted triangle.
 */
#ifndef RTB200_H
#define RTB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_context rt_context;

enum {
    RT_OK = 0,
    RT_E_INVALID = 1,    /* bad argument / malformed scene */
    RT_E_NO_DEVICE = 2,  /* no usable CUDA device: there is no CPU fallback */
    RT_E_CUDA = 3,       /* a CUDA runtime call failed */
    RT_E_NO_SCENE = 4,   /* trace/render before rt_upload_scene */
    RT_E_NO_PARAMS = 5   /* render before rt_set_params */
};

enum { RT_CLOSEST = 0, RT_ANY = 1 }; /* needClosestHit = true / false (vR.cl:659) */

/* (float)UINT_MAX: the reference's initial hit distance (HitRecordInit, vR.cl:223-227) */
#define RT_T_INIT 4294967296.0f

/* One ray of the batch operator, 32 bytes. `dir` is used exactly as given: the caller normalises
 * it the way RayInit does (vR.cl:205-211). tmax is the initial *tHit (RT_T_INIT in the reference). */
typedef struct { float ox, oy, oz, tmax, dx, dy, dz, reserved; } rt_ray;
/* One result, 16 bytes: idx = 3 * triangle id or -1; t = final *tHit; u,v = barycentrics of the
 * accepted triangle (weights of its 2nd and 3rd vertex), 0 when idx < 0. */
typedef struct { int32_t idx; float t, u, v; } rt_hit;
/* Alignment: DEVICE buffers of rt_ray / rt_hit (every `d_*` argument, and rt_host_register aliases) must be 16-byte
 * aligned -- the kernels move records as 16-byte vectors -- and frames 4-byte aligned (the int64 count of
 * rt_diffuse_rays_device 8). cudaMalloc, torch and rt_ipc_alloc all give 256 bytes or better; a buffer carved at an odd
 * offset is refused with RT_E_INVALID before anything is launched. HOST buffers may have any alignment: a page-locked
 * one that is suitably aligned is written directly by the kernel, any other goes through a device buffer and a copy. */

/* ---- lifetime ------------------------------------------------------------------------------- */
int rt_create(int device_ordinal, rt_context** out_ctx);   /* replaces setupCL() */
int rt_destroy(rt_context* ctx);                            /* replaces cleanup() + ~RayTraceData */
const char* rt_last_error(const rt_context* ctx);           /* ctx may be NULL: last rt_create error */
const char* rt_version(void);

/* Run all work of this context on an existing CUDA stream (a cudaStream_t passed as void*), e.g.
 * the caller's framework stream so that its events bracket our kernels. NULL = the context's own. */
int rt_set_stream(rt_context* ctx, void* cuda_stream);
int rt_synchronize(rt_context* ctx);

/* ---- scene ---------------------------------------------------------------------------------- */
/* Replaces the clCreateBuffer calls of initRayTrace() (RayTracer.cpp:942-984). normals,
 * normal_indices, materials, tri_to_material may be NULL (with Vn = M = 0) when only rt_trace /
 * rt_primary_* are used; rt_render_frame then fails with RT_E_INVALID. The flat BVH is validated
 * (index ranges, acyclic) and re-packed (once, on the host) into 64-byte node pairs + 48-byte
 * pre-subtracted triangles; see DESIGN.md "Data layout in HBM". */
int rt_upload_scene(rt_context* ctx, const float* verts, int V, const int32_t* indices, int T, const void* nodes, int N,
                    const int32_t* tri_indices, int R, const float* normals, int Vn, const int32_t* normal_indices,
                    const void* materials, int M, const int32_t* tri_to_material);

/* The packed device scene is ONE contiguous, position-independent device allocation so that it can
 * be broadcast to other GPUs as a single buffer (NCCL) and adopted there without a host round trip. */
int rt_scene_blob(rt_context* ctx, void** out_device_ptr, size_t* out_bytes);
/* Device-to-device copy of the blob into caller-owned device memory (e.g. the NCCL broadcast buffer). */
int rt_copy_scene_blob(rt_context* ctx, void* dst_device_ptr, size_t bytes);
/* Adopt a blob that already sits in this device's memory (borrowed: the caller keeps it alive). device_ptr must be
 * 256-byte aligned; the header and every section range are validated before use. */
int rt_adopt_scene_blob(rt_context* ctx, void* device_ptr, size_t bytes);

/* Host-only twin of the packing step of rt_upload_scene: same validation, same blob (see csrc/scene_blob.h), returned
 * in malloc'ed host memory (free with rt_free_host). Needs no device; err (optional) receives the message on failure. */
int rt_pack_scene_host(const float* verts, int V, const int32_t* indices, int T, const void* nodes, int N,
                       const int32_t* tri_indices, int R, const float* normals, int Vn, const int32_t* normal_indices,
                       const void* materials, int M, const int32_t* tri_to_material, int top_pairs, void** out_blob,
                       size_t* out_bytes, char* err, int err_len);
void rt_free_host(void* p);

/* Replaces clEnqueueWriteBuffer(params) in updateCamera() (RayTracer.cpp:671). */
int rt_set_params(rt_context* ctx, const float params[32]);

/* ---- the hot path --------------------------------------------------------------------------- */
/* Whole frame = the raytracer_bvh kernel (vR.cl:1043-1547): primary rays, <= 3 path segments of
 * closest-hit + any-hit shadow + mirror reflection, GGX shading, RGBA8 pack. out_host receives
 * w*h uint32 pixels, row-major, pixel = (b<<16)|(g<<8)|r as rgbToInt (vR.cl:186-195).
 * Replaces raytrace_gpgpu() (RayTracer.cpp:330-344). */
int rt_render_frame(rt_context* ctx, int w, int h, uint32_t* out_host);
/* The same frame without the per-frame stall of displayGL()'s clFinish + blocking read (RayTracer.cpp:284-293,
 * 341-343): _begin enqueues the frame described by the last rt_set_params into out_host and returns; _end blocks
 * until the frame of that slot is complete in out_host. Up to RT_FRAME_SLOTS frames may be in flight, one per slot;
 * rt_set_params for the next frame may be called at once (every launch carries its Params block by value).
 * out_host must stay valid until _end: page-locked memory (rt_host_register, cudaHostAlloc) receives the pixels
 * straight from the kernel while the next frame traces; pageable memory is filled by a copy inside _end.
 * Frames in flight run on per-slot streams (ordered after the work already enqueued on the context stream), so the
 * ramp-down of one frame overlaps the start of the next: 0.81 -> 0.67 ms per 1080p frame with two slots. */
#define RT_FRAME_SLOTS 4
int rt_render_frame_begin(rt_context* ctx, int w, int h, uint32_t* out_host, int slot);
int rt_render_frame_end(rt_context* ctx, int slot);

/* Batch operator = traverse_bvh (vR.cl:658-1010) on n caller-supplied rays; host buffers. */
int rt_trace(rt_context* ctx, int mode, int64_t n, const rt_ray* rays_host, rt_hit* hits_host);

/* Primary-ray pass of a w x h frame with the current Params (vR.cl:1156-1196 + the first
 * traverse_bvh call of vR.cl:1238): hits_host receives w*h records, pixel i = y*w+x. Pixels that
 * fail the scene-AABB gate get idx = -1, t = RT_T_INIT. Host buffer out; the device->host copy of
 * finished row bands overlaps the tracing of the next ones. */
int rt_primary(rt_context* ctx, int w, int h, rt_hit* hits_host);

/* ---- device-resident variants (benchmarks, multi-GPU plumbing; pointers are device memory) ---- */
int rt_trace_device(rt_context* ctx, int mode, int64_t n, const rt_ray* d_rays, rt_hit* d_hits);
/* Same results as rt_trace_device, ray for ray and bit for bit, but the batch is traced in coherence order (key =
 * direction octant | Morton code of the origin | coarse direction; radix sort, gather, trace, scatter back): an optional
 * pre-pass for large incoherent batches. n < 2^31. */
int rt_trace_sorted_device(rt_context* ctx, int mode, int64_t n, const rt_ray* d_rays, rt_hit* d_hits);
/* Primary rays of a w x h frame generated in-kernel from the Params block (vR.cl:1156-1196) and
 * traced closest-hit; pixel i = y*w+x. Pixels failing the scene-AABB gate get idx = -1,
 * t = RT_T_INIT. d_rays_out (optional, may be NULL) receives the generated rays. Only rows
 * y with (y / band_rows) % n_parts == part are produced (part = 0, n_parts = 1: whole frame);
 * untouched rows are left as they are. */
int rt_primary_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                      rt_ray* d_rays_out);
/* Same primary pass with the gather fused into the store: besides (or instead of, d_hits may be NULL) the
 * 16-byte hit records, every pixel's hit index (3 * triangle id or -1) is written to d_idx_frame[y*w+x].
 * d_idx_frame may live on ANOTHER GPU (a peer mapping obtained with rt_ipc_open, or cudaDeviceEnablePeerAccess inside
 * one process): each rank then stores its bands straight into rank 0's framebuffer over NVLink and no separate gather
 * collective is needed. It may also be page-locked HOST memory (rt_host_register). For such remote frames the kernel
 * assembles rows: tiles land in a local stage and the last warp of a group of 4 / 16 horizontally adjacent tiles writes
 * the group's rows as 128- / 512-byte segments (option "store_group": -1 auto, 0 off, 2, 4). */
int rt_primary_gather_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                             int32_t* d_idx_frame);
/* Primary + shadow in ONE launch (BASELINE configs 3 and 5): per pixel the camera ray, the closest hit (first traverse_bvh
 * call, vR.cl:1238) and, for every hit, the any-hit shadow ray towards the light built exactly as vR.cl:1314,1407-1441.
 * One launch pays one ramp-down and one fixed cost where rt_primary_device + rt_shadow_device pay two. Outputs, each
 * optional (NULL): d_hits = the closest-hit records; d_shadow_hits = the shadow rays' any-hit records (idx = -1,
 * t = RT_T_INIT where the pixel has no hit); d_vis_frame = one 4-byte word per pixel,
 *     vis = -1 (no hit)   or   3 * triangle id + (occluded ? 1 : 0),
 * with the reference's rule occluded = shadow idx >= 0 && shadow t > 0.025 (vR.cl:1444-1449). 3 * triangle id is a
 * multiple of 3, so vis / 3 * 3 is the hit index and vis % 3 the shadow bit. Band arguments as rt_primary_device;
 * d_vis_frame may be peer or page-locked host memory like d_idx_frame of rt_primary_gather_device. */
int rt_primary_shadow_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, rt_hit* d_hits,
                             rt_hit* d_shadow_hits, int32_t* d_vis_frame);
/* Host-buffer form of the same pass: vis_host receives w*h visibility words (page-locked memory is written by the kernel
 * directly, in full 128/512-byte rows; pageable memory through a device frame and one copy). */
int rt_primary_shadow(rt_context* ctx, int w, int h, int32_t* vis_host);
/* Peer-visible device buffer across processes of one node (CUDA IPC): the owner allocates and passes the
 * 64-byte handle to its peers (any transport), they map it. */
int rt_ipc_alloc(rt_context* ctx, size_t bytes, void** out_device_ptr, unsigned char out_handle[64]);
int rt_ipc_open(rt_context* ctx, const unsigned char handle[64], void** out_device_ptr);
int rt_ipc_close(rt_context* ctx, void* device_ptr);
int rt_ipc_free(rt_context* ctx, void* device_ptr);
int rt_memcpy_to_host(rt_context* ctx, void* dst_host, const void* src_device, size_t bytes);
/* Page-lock caller-owned host memory and get the device alias the kernels can store into directly. With a POSIX
 * shared-memory framebuffer registered by every rank, each GPU writes its bands straight into the consumer's host
 * memory over its OWN PCIe link (the gather needs neither NVLink nor a device->host copy on rank 0). */
int rt_host_register(rt_context* ctx, void* host_ptr, size_t bytes, void** out_device_alias);
int rt_host_unregister(rt_context* ctx, void* host_ptr);
/* Stream-ordered completion signal: after everything enqueued on the context so far has finished (and its stores are
 * visible system-wide), the 32-bit word at d_word (a device alias from rt_host_register, or device memory) is set to
 * `value`. Lets the consumer of a multi-GPU frame poll one word per rank instead of joining a barrier every frame. */
int rt_signal(rt_context* ctx, void* d_word, uint32_t value);
/* Shadow rays built in-kernel from rays + their closest hits exactly as vR.cl:1314,1407-1441 and
 * traced any-hit. Entries whose hit idx < 0 produce idx = -1, t = RT_T_INIT. d_shadow_rays_out may
 * be NULL. */
int rt_shadow_device(rt_context* ctx, int64_t n, const rt_ray* d_rays, const rt_hit* d_hits, rt_hit* d_shadow_hits,
                     rt_ray* d_shadow_rays_out);
/* Synthetic incoherent workload (not in the reference, BASELINE config 4): `spp` cosine-weighted
 * hemisphere rays about the geometric normal of every hit, origin = hit point + n * 1e-3, hash RNG
 * seeded by (ray index * spp + s + seed). Writes rays for valid hits compacted in index order and
 * the count to *d_count (device int64). */
int rt_diffuse_rays_device(rt_context* ctx, int64_t n, const rt_ray* d_rays, const rt_hit* d_hits, int spp, uint32_t seed,
                           rt_ray* d_out_rays, int64_t* d_count);
/* Frame into a device framebuffer (w*h uint32); band partition as in rt_primary_device. */
int rt_render_frame_device(rt_context* ctx, int w, int h, int part, int n_parts, int band_rows, uint32_t* d_out);

/* ---- multi-GPU group: one process, N GPUs of one box ------------------------------------------------
 * The reference drives one device (RayTracer.cpp:2128-2131 takes devices[0]; raytrace_gpgpu, :330-344, enqueues on one
 * queue). A group is N contexts behind the same seam: the scene is packed and uploaded once and its blob is broadcast
 * to the other GPUs with ncclBroadcast over NVLink (ncclCommInitAll; NCCL is bound at run time); a frame is split in
 * interleaved bands of `band_rows` rows (default 16, group option "band_rows"), every GPU traces its bands -- launches are
 * issued by one worker thread per GPU -- and stores its pixels straight into the caller's frame: page-locked host memory
 * is written by all GPUs at once, each over its own PCIe link; a pageable frame is gathered in a frame on GPU 0 that
 * the other GPUs write over NVLink (peer access) and copied out once. The calls return when the frame is complete,
 * like raytrace_gpgpu(). Frames are bit-identical to the single-context calls. */
typedef struct rt_group rt_group;
#define RT_GROUP_MAX_GPUS 16
enum {
    RT_GROUP_STAT_BROADCAST_MS = 0, /* device time of the last scene broadcast (CUDA events on GPU 0's stream) */
    RT_GROUP_STAT_COMM_INIT_MS = 1, /* one-off ncclCommInitAll wall time */
    RT_GROUP_STAT_BLOB_BYTES = 2,
    RT_GROUP_STAT_LAST_CALL_MS = 3, /* wall time of the last tiled call */
    RT_GROUP_STAT_ZERO_COPY = 4,    /* 1 if the last tiled call stored straight into the caller's (page-locked) frame */
    RT_GROUP_STAT_RANK_KERNEL_MS = 8, /* [8 + rank]: device time of that GPU's part of the last tiled call */
    RT_GROUP_STATS = 8 + RT_GROUP_MAX_GPUS
};
int rt_create_group(int n_gpus, const int* device_ordinals /* NULL = 0..n-1 */, rt_group** out_group);
int rt_destroy_group(rt_group* g);
const char* rt_group_last_error(const rt_group* g); /* g may be NULL: last rt_create_group error */
int rt_group_size(const rt_group* g);
rt_context* rt_group_context(rt_group* g, int rank); /* borrowed; e.g. for rt_get_counters */
/* rt_upload_scene on GPU 0 + broadcast of the packed blob + rt_adopt_scene_blob on the others. */
int rt_group_upload_scene(rt_group* g, const float* verts, int V, const int32_t* indices, int T, const void* nodes, int N,
                          const int32_t* tri_indices, int R, const float* normals, int Vn, const int32_t* normal_indices,
                          const void* materials, int M, const int32_t* tri_to_material);
int rt_group_set_params(rt_group* g, const float params[32]);
/* "band_rows" (multiple of 4), "broadcast" (0 = ncclBroadcast, 1 = peer copies from GPU 0); any other name is forwarded
 * to rt_set_option of every context. */
int rt_group_set_option(rt_group* g, const char* name, int value);
/* rt_render_frame over the group: replaces raytrace_gpgpu() (RayTracer.cpp:330-344) for N GPUs. */
int rt_render_frame_tiled(rt_group* g, int w, int h, uint32_t* out_host);
/* The 4-byte/pixel primary pass over the group: with_shadow = 0 -> hit-index frame (rt_primary_gather_device),
 * with_shadow = 1 -> visibility frame of the fused primary + shadow pass (rt_primary_shadow_device). */
int rt_primary_tiled(rt_group* g, int w, int h, int with_shadow, int32_t* frame_host);
int rt_group_stats(rt_group* g, double out[RT_GROUP_STATS]);

/* ---- introspection -------------------------------------------------------------------------- */
/* RT_CNT_RAYS_TRACED = traversals started (one per traverse_bvh call of the reference: primary rays that pass the scene
 * gate, shadow rays, reflection rays, caller rays), counted by the kernels themselves; rt_get_counters waits for the
 * work issued so far before it reads the count. */
enum { RT_CNT_KERNEL_LAUNCHES = 0, RT_CNT_RAYS_TRACED = 1, RT_CNT_H2D_BYTES = 2, RT_CNT_D2H_BYTES = 3, RT_CNT_COUNT = 8 };
int rt_get_counters(rt_context* ctx, uint64_t out[RT_CNT_COUNT]);
int rt_reset_counters(rt_context* ctx);
/* Traversal variant selector for experiments (see DESIGN.md "Kernels"); 0 = default. */
int rt_set_option(rt_context* ctx, const char* name, int value);
/* Temporal tile scheduling (option "tile_hints", default 1): a camera-ray launch (rt_primary*, rt_primary_shadow*,
 * rt_render_frame*) records how long each 8x4-pixel tile took; the next launch of the same frame geometry starts the slow
 * tiles first and runs the slowest as four one-row items (DESIGN.md "Tile scheduler"). Results never depend on the hints.
 * Options: "hint_heavy_pct" (12) / "hint_light_pct" (-1 = 35, 65 for shaded frames): the slowest N percent of the tiles start first, the quickest N
 * percent run last; "hint_split_pct" (80), "hint_keep_pct" (30): a tile is split / a row stays split when it took more
 * than that percentage of the previous launch's span (measured: tools/hint_sweep.sh). out: [0] rows on the split list, [1] tiles on the heavy list, [2] tiles timed, [3] launch span
 * (SM cycles) recorded by the most recent such launch. */
int rt_tile_hint_stats(rt_context* ctx, uint64_t out[4]);
/* Option "gate_cull" (default 1): camera-ray launches enumerate only the tiles inside the screen-space bounding rectangle
 * of the scene box (a pixel outside cannot pass the gate of vR.cl:1196) and write the rest with store-only items.
 * rt_cull_rect_host (no device needed) returns that rectangle in pixels, inclusive: [x0, x1] x [y0, y1] (x0 > x1 or y0 > y1:
 * the box is off screen; the whole frame when no bound is known, e.g. the eye is inside the box). */
int rt_cull_rect_host(const float params[32], int w, int h, int64_t out_x0_x1_y0_y1[4]);
/* Option "inner_exit_batch" (-1 auto / 0 / 1): which instance of the batch kernels' traversal loop runs -- with 1 a warp
 * leaves the inner-node loop as soon as fewer than 8 of its lanes still descend while others wait on a leaf. Same results;
 * auto picks it when the scene's node pairs + triangles do not fit the L2 cache (measured: +8-10 % there, -2-10 % otherwise). */
/* Option "l2_warm" (default 0; 1 node pairs + triangles, 2 pairs, 3 triangles; "l2_warm_chunk_kb", 16): every traversal
 * launch first asks L2 for the scene's traversal data with bulk prefetches. An experiment kept for measurement: launches that
 * find L2 flushed are only 1-2.5 % slower than warm ones and the warm-up costs more than that (profiles/r2_experiments.md 14). */
/* Option "smem_carveout" (-1 = the driver's choice; 0..100 = cudaFuncAttributePreferredSharedMemoryCarveout of the traversal
 * kernels): an experiment kept for measurement -- the default split is within noise of the best (r2_experiments.md 15). */
/* GPU self test of the box test's hoisted exact division against the compiler's IEEE division on
 * `samples` random operand pairs; *out_mismatches must come back 0. */
int rt_selftest(rt_context* ctx, int64_t samples, uint32_t seed, uint64_t* out_mismatches);
/* Same comparison with the numerator's binary exponent fixed (probe of the admitted operand window). */
int rt_selftest_range(rt_context* ctx, int64_t samples, uint32_t seed, int x_exponent, uint64_t* out_mismatches);
/* Exhaustive form: ALL 2^23 numerator mantissas x md_count divisor mantissas from md_begin (chunks of <= 65535; the full
 * square is 2^46 quotients, about a GPU-minute) at binary exponents ex_x / ex_d and signs (bit 0: divisor < 0, bit 1:
 * numerator < 0). Power-of-two scaling is exact inside the window, so one exponent pair stands for all. */
int rt_selftest_exhaustive(rt_context* ctx, uint32_t md_begin, uint32_t md_count, int ex_x, int ex_d, uint32_t signs,
                           uint64_t* out_mismatches);
/* Scene statistics: [0] node pairs, [1] packed triangles, [2] blob bytes, [3] max tree depth,
 * [4] 1 if every box coordinate admits the hoisted exact division (else the full division is used), [5] BFS-ordered pairs */
int rt_scene_info(rt_context* ctx, int64_t out[6]);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
