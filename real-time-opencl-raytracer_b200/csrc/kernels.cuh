// kernels.cuh -- sm_100a kernels around traverse(): ray sources (buffer / camera / shadow), the
// persistent warp scheduler, the frame megakernel (shading + pack) and the diffuse-ray generator.
#pragma once
#include <cstdint>

#include "traverse.cuh"

namespace rtb {

struct ParamsBlock {  // the reference's 128-byte Params (RayTracer.cpp:115-161): 8 x float4, w = 1
    float4 a, b, c, campos, light_pos, light_color, aabb_min, aabb_max;
};

struct RayIO {  // rt_ray / rt_hit as two / one 16-byte vectors
    float4 o_tmax;  // ox oy oz tmax
    float4 d_pad;   // dx dy dz reserved
};

enum RaySource { SRC_BUFFER = 0, SRC_PRIMARY = 1, SRC_SHADOW = 2 };

struct TraceArgs {
    SceneView scene;
    ParamsBlock params;
    // work description
    long long n;              // SRC_BUFFER / SRC_SHADOW: number of rays
    int w, h;                 // SRC_PRIMARY: frame size
    int tiles_x;              // SRC_PRIMARY: 8-pixel-wide tiles per row
    int part, n_parts;        // SRC_PRIMARY: interleaved row bands, this rank's share
    int band_tile_rows;       // band_rows / 4
    long long num_batches;    // warp-sized work items
    int tile_order;           // SRC_PRIMARY: 0 row-major tiles, 1 reversed, 2 multiplicative permutation, 3 permuted 64-tile chunks
    unsigned int order_mul;   // odd multiplier coprime to the permuted count (host-chosen)
    // buffers
    const float4* rays_in;    // SRC_BUFFER, SRC_SHADOW (2 x float4 per ray)
    const float4* hits_in;    // SRC_SHADOW (closest hits of rays_in)
    float4* hits_out;         // 1 x float4 per ray {bits(idx), t, u, v}
    float4* shadow_hits_out;  // primary_shadow_kernel only: the any-hit record of each pixel's shadow ray (optional)
    float4* rays_out;         // optional: the generated rays (SRC_PRIMARY, SRC_SHADOW)
    // The 4-byte/pixel frame of the pass (optional): the shaded pixel (render_kernel), the hit index (SRC_PRIMARY) or the
    // visibility word (primary_shadow_kernel). May be memory of ANOTHER GPU (peer mapping: the gather of a multi-GPU frame
    // is fused into this store) or page-locked host memory (zero copy).
    unsigned int* frame_out;
    unsigned long long* work_counter;  // persistent-warp work queue head (zeroed before launch)
    unsigned long long* ray_counter;   // += traversals started (traverse() calls) by this launch; one atomicAdd per warp at exit
    // Row assembly for frame_out in remote (host / peer) memory. A warp's 8x4 tile is four 32-byte row pieces; PCIe and
    // NVLink want full 128-byte writes. With group_log2 > 0 the tiles store into `stage` (a frame-shaped scratch in local
    // memory, L2-resident) and the LAST warp to finish a group of 2^group_log2 horizontally adjacent tiles copies the
    // group's four rows to frame_out as 128-byte (4 tiles) or 512-byte (16 tiles) row segments.
    unsigned int* stage;
    unsigned int* group_count;  // one arrival counter per (tile row of this rank, group), zeroed before launch
    int group_log2;             // 0 = off, 2 = 4 tiles (32 px), 4 = 16 tiles (128 px)
    int groups_x;
    // Scene-box culling of the tile queue (camera-ray kernels, host side: cull_setup in rtb200.cu). Tiles outside the
    // screen-space bounding rectangle of the scene AABB cannot pass the gate: the queue's main pass enumerates only tile
    // columns [in_tx0, in_tx1) of this rank's tile rows [in_k0, in_k1), and `n_fill` store-only items (one per 32 tiles of
    // a tile row) write the miss values of everything outside. n_fill == 0: no culling, the rectangle is the whole frame.
    int in_tx0, in_tx1, in_k0, in_k1;
    unsigned int n_fill;
    int fill_first;  // the fill items lead the queue instead of closing it (frames stored straight into host memory: their
                     // burst of stores then overlaps the slow tiles' tracing instead of trailing the launch)
    // Temporal tile scheduling (camera-ray kernels; see "tile scheduler" below): what the previous launch of this frame
    // geometry learnt about its tiles, and where this launch records the same for the next one. Either may be NULL.
    const unsigned int* hint_in;
    unsigned int* hint_out;
    int hint_heavy_pct, hint_light_pct;   // the slowest / quickest N percent of the tiles that traced
    int hint_split_pct, hint_keep_pct;    // percent of the previous launch's span
    // L2 warm-up (option "l2_warm"): the first thread of every block asks L2 for its share of [warm_base, + warm_bytes)
    // -- node pairs and packed triangles, contiguous in the blob -- before it starts tracing. 0 bytes = off.
    const char* warm_base;
    unsigned long long warm_bytes;
    unsigned int warm_chunk;  // bytes per bulk prefetch, a multiple of 16
#ifdef RTB_TIMELINE
    unsigned long long* timeline;  // tools build only: {start ns, end ns | smid << 56} per batch (tools/timeline_probe.py)
#endif
};

#ifdef RTB_TIMELINE
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int tl_smid() {
    unsigned int s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}
#define RTB_TL_BEGIN(a) const unsigned long long tl_t0__ = (a).timeline ? tl_now() : 0ull
#define RTB_TL_END(a, batch, lane)                                                              \
    do {                                                                                        \
        __syncwarp();                                                                           \
        if ((a).timeline && (lane) == 0) {                                                      \
            (a).timeline[2 * (batch)] = tl_t0__;                                                \
            (a).timeline[2 * (batch) + 1] = (tl_now() & 0x00ffffffffffffffull) | ((unsigned long long)tl_smid() << 56); \
        }                                                                                       \
    } while (0)
#else
#define RTB_TL_BEGIN(a)
#define RTB_TL_END(a, batch, lane)
#endif

#ifndef RTB_MINB_BATCH
#define RTB_MINB_BATCH 1
#endif
#ifndef RTB_MINB_PRIMARY
#define RTB_MINB_PRIMARY 8  /* camera-ray batch kernel: 8 blocks x 128 threads = 64 registers, no spills */
#endif
#ifndef RTB_MINB_LANES
#define RTB_MINB_LANES 1
#endif
#ifndef RTB_BLOCK_THREADS
#define RTB_BLOCK_THREADS 128
#endif
static constexpr int kBlockThreads = RTB_BLOCK_THREADS;
static constexpr int kWarpsPerBlock = kBlockThreads / 32;

// ---- ray sources ---------------------------------------------------------------------------------

// Cheap, conservative pre-test of the scene-AABB gate: true only if the pixel's ray CERTAINLY fails the exact gate
// below, so that the exact ray (2 IEEE divisions, normalise, 3 more divisions) need not be built for it. Most pixels of
// a typical frame look past the scene; they cost ~25 instead of ~150 instructions. The slab test is invariant under
// positive scaling of the direction, so the unnormalised direction and approximate reciprocals (error ~1e-6) are
// enough; decisions within a 1e-3 relative margin, and anything non-finite, are left to the exact path.
__device__ __forceinline__ bool certainly_gated_out(const ParamsBlock& P, unsigned x, unsigned y, unsigned w, unsigned h) {
    const float xf = ((float)x - 0.5f) * rcp_approx((float)w), yf = ((float)y - 0.5f) * rcp_approx((float)h);
    const float ox = __fmaf_rn(P.b.x, yf, __fmaf_rn(P.a.x, xf, P.c.x));
    const float oy = __fmaf_rn(P.b.y, yf, __fmaf_rn(P.a.y, xf, P.c.y));
    const float oz = __fmaf_rn(P.b.z, yf, __fmaf_rn(P.a.z, xf, P.c.z));
    const float ix = rcp_approx(ox - P.campos.x), iy = rcp_approx(oy - P.campos.y), iz = rcp_approx(oz - P.campos.z);
    const float ax = (P.aabb_min.x - ox) * ix, bx = (P.aabb_max.x - ox) * ix;
    const float ay = (P.aabb_min.y - oy) * iy, by = (P.aabb_max.y - oy) * iy;
    const float az = (P.aabb_min.z - oz) * iz, bz = (P.aabb_max.z - oz) * iz;
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(ax), fabsf(bx)), fmaxf(fabsf(ay), fabsf(by))), fmaxf(fabsf(az), fabsf(bz)));
    if (!(mag < 1e30f)) return false;  // inf / NaN somewhere (axis-parallel ray, degenerate box): decide exactly
    const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    const float margin = 1e-3f * mag;
    return (tmax < tmin - margin) || (tmax < -margin);
}

// volumeRender.cl:1156-1196: pixel -> primary ray; returns the scene-AABB gate
__device__ __forceinline__ bool primary_ray(const ParamsBlock& P, unsigned x, unsigned y, unsigned w, unsigned h, Ray& r) {
    // (x-0.5)/((float)w): exact in fp32 (see oracle/oracle.c primary_ray)
    const float xf = ((float)x - 0.5f) / ((float)w);
    const float yf = ((float)y - 0.5f) / ((float)h);
    const f3 t1 = add3(ld3(P.c), scale3(ld3(P.a), xf));
    const f3 t2 = scale3(ld3(P.b), yf);
    const f3 image_pos = add3(t1, t2);
    r = ray_init(image_pos, sub3(image_pos, ld3(P.campos)));
    return scene_gate(ld3(P.aabb_min), ld3(P.aabb_max), r);
}

// volumeRender.cl:1314 + 1407-1441: shadow ray from a path vertex
__device__ __forceinline__ Ray shadow_ray(f3 light_pos, const Ray& r, float t, f3& hitpoint) {
    hitpoint = add3(r.ori, scale3(r.dir, (t - 0.001f)));
    const f3 L = normalize3(sub3(light_pos, hitpoint));
    return ray_init(add3(hitpoint, scale3(L, 0.001f)), L);
}

// Map a primary-ray batch (one warp = one 8x4 pixel tile) to pixel coordinates. tx = tile column, k = index of the tile row
// among this rank's tile rows (what the row-assembly counters are indexed by).
// k-th tile row of this rank (interleaved bands of band_tile_rows tile rows) -> tile row of the frame
__device__ __forceinline__ unsigned int global_tile_row(const TraceArgs& a, unsigned int kk) {
    const unsigned int btr = (unsigned int)a.band_tile_rows;
    const unsigned int bk = kk / btr;
    return ((unsigned int)a.part + bk * (unsigned int)a.n_parts) * btr + (kk - bk * btr);
}
__device__ __forceinline__ void tile_pixel(const TraceArgs& a, long long batch, int lane, int& x, int& y, int& tx, long long& k) {
    if (a.tile_order == 1) batch = a.num_batches - 1 - batch;
    else if (a.tile_order == 2) batch = (long long)(((unsigned long long)batch * a.order_mul) % (unsigned long long)a.num_batches);
    else if (a.tile_order == 3) {
        const long long chunks = a.num_batches >> 6;  // whole 64-tile chunks are permuted, the remainder stays in place
        const long long c = batch >> 6;
        if (c < chunks) batch = (long long)(((unsigned long long)c * a.order_mul) % (unsigned long long)chunks) * 64 + (batch & 63);
    }
    // 32-bit index arithmetic: a frame has < 2^31 tiles (band_setup rejects more). The four divisions / remainders below
    // run once per tile in every lane; as 64-bit operations they were ~10 % of the primary kernel's instructions (ncu).
    const unsigned int b = (unsigned int)batch, tiles_x = (unsigned int)a.tiles_x;
    const unsigned int kk = b / tiles_x;  // index among this rank's tile rows
    tx = (int)(b - kk * tiles_x);
    k = (long long)kk;
    x = tx * 8 + (lane & 7);
    y = (int)(global_tile_row(a, kk) * 4u + (unsigned int)(lane >> 3));
}
__device__ __forceinline__ void tile_pixel(const TraceArgs& a, long long batch, int lane, int& x, int& y) {
    int tx;
    long long k;
    tile_pixel(a, batch, lane, x, y, tx, k);
}

// L2 warm-up: cp.async.bulk.prefetch.L2 moves a whole chunk DRAM -> L2 with one instruction and no register or shared-memory
// destination. A launch that finds L2 cold (another working set went through it since the last frame) otherwise pays one
// DRAM round trip per first touch of a node or a triangle, inside the dependent chain of a traversal.
__device__ __forceinline__ void warm_l2(const TraceArgs& a) {
    if (!a.warm_bytes || threadIdx.x != 0) return;
    const unsigned long long chunk = a.warm_chunk;
    for (unsigned long long off = (unsigned long long)blockIdx.x * chunk; off < a.warm_bytes; off += (unsigned long long)gridDim.x * chunk) {
        const unsigned long long left = a.warm_bytes - off;
        const unsigned int n = (unsigned int)(left < chunk ? left : chunk);
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.warm_base + off), "r"(n) : "memory");
    }
}

// RT_CNT_RAYS_TRACED: every lane counts the traversals it starts; one warp reduction + atomicAdd when the warp retires.
__device__ __forceinline__ void retire_ray_count(const TraceArgs& a, unsigned int traced, int lane) {
    traced = __reduce_add_sync(0xffffffffu, traced);
    if (lane == 0 && traced) atomicAdd(a.ray_counter, (unsigned long long)traced);
}

// Where a tile's lanes put their 4-byte pixel: the destination itself, or the local stage when rows are assembled.
__device__ __forceinline__ unsigned int* pixel_sink(const TraceArgs& a) { return a.group_log2 ? a.stage : a.frame_out; }

// Row assembly (see TraceArgs::stage). Called by ALL lanes of a warp after they stored their tile's pixels to the stage.
// Last-arriver pattern: fence, count the tile in, and the warp that completes the group copies it out. The stage was
// written by other SMs, so it is read around L1 (ld.global.cg).
__device__ __forceinline__ void finish_tile(const TraceArgs& a, int tx, long long k, int y0, int lane, unsigned int rows_done = 4u) {
    if (!a.group_log2) return;
    __threadfence();
    __syncwarp();
    const int gx = tx >> a.group_log2;
    const int first_tile = gx << a.group_log2;
    const int tiles_in_group = min(1 << a.group_log2, a.tiles_x - first_tile);
    unsigned int arrived = 0;
    // arrivals are counted in pixel rows (4 per tile): a tile may arrive in pieces (tile scheduler: quarter / partial tiles)
    if (lane == 0) arrived = atomicAdd(a.group_count + k * a.groups_x + gx, rows_done);
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if ((int)(arrived + rows_done) != 4 * tiles_in_group) return;
    __threadfence();
    const int x0 = first_tile * 8;
    const int span = tiles_in_group * 8;  // pixels per row of this group
    if (a.group_log2 == 4 && (a.w & 3) == 0) {  // 16 tiles: one 512-byte row segment per warp store (16 B per lane)
        const int c = lane * 4;
        if (c < span && x0 + c < a.w) {
            for (int r = 0; r < 4 && y0 + r < a.h; r++) {
                const size_t at = (size_t)(y0 + r) * a.w + x0 + c;
                *reinterpret_cast<uint4*>(a.frame_out + at) = __ldcg(reinterpret_cast<const uint4*>(a.stage + at));
            }
        }
    } else {  // 128 bytes per warp store
        for (int r = 0; r < 4 && y0 + r < a.h; r++) {
            const size_t row = (size_t)(y0 + r) * a.w;
            for (int c = lane; c < span; c += 32)
                if (x0 + c < a.w) a.frame_out[row + x0 + c] = __ldcg(a.stage + row + x0 + c);
        }
    }
}


// ---- tile scheduler: persistent warps + temporal hints ------------------------------------------------------------------
// A camera-ray launch is a queue of 8x4-pixel tiles drained by persistent warps. Measured on the bench frame (per-tile
// timestamps, tools/timeline_probe.py): the median tile takes ~18 us but tiles whose rays skim a terrain ridge take 8x
// that (every lane ~100 node visits + ~50 triangle tests, and the lanes drift apart); wherever they sit in the queue they
// finish last, and a launch ends with 30-50 % of its duration spent at a few percent occupancy -- the tail that limits a
// frame split over 8 GPUs. A real-time renderer draws almost the same frame again, so every launch RECORDS what it
// learnt for the next launch of the same frame geometry (host side: rtb200.cu, hint slots):
//   * the launch's span (longest life of a persistent warp), per tile the time it took (SM clock) and a histogram of those
//     times. The slowest `heavy_pct` percent of the tiles go on the HEAVY list; a tile that took more than `split_pct`
//     percent of the previous launch's span -- one that by itself decides when the launch ends -- goes on the SPLIT list
//     as four one-row (8x1 pixel) items; the quickest `light_pct` percent (of the tiles that traced at all) go on the
//     LIGHT list
//   * the next launch's queue is [split rows][heavy tiles][all tiles in row-major order, minus what the lists cover]
//     [light tiles], the lists walked longest-first: the long tiles start first, the longest run as four warps with 8
//     coherent lanes each instead of one warp whose 32 lanes serialise each other, and the launch ends on its shortest
//     tiles, so the idle time at the end is a short tile's duration instead of a long one's
// Hints only change WHO traces a pixel and WHEN; every pixel is still traced exactly once with the same arithmetic, so
// results do not depend on them (tests/test_gpu_sched.py). A stale hint (the camera moved) costs time, never correctness.
//
// Hint buffer (32-bit words): [0] split entries [1] heavy entries [2] tiles timed [3] launch span (cycles) [4] light
// entries; [16, 16 + 128) histogram of tile times in bins of 4096 cycles;  [H ..) split list: (batch << 2) | row;
// [H + 4B ..) heavy list: batch;  [H + 5B ..) light list: batch;  [H + 6B ..) one byte per (batch, row): non-zero = that
// row is covered by a list entry (H = kHintHeader, B = num_batches; the lists can hold every row / tile: no overflow).
enum { kHintHist = 16, kHintBins = 128, kHintBinShift = 12, kHintHeader = kHintHist + kHintBins };
static constexpr unsigned int kFillItem = 0x10u;
struct TileWork {
    long long batch;    // tile index among this rank's tiles (row-major over all tile columns); for a fill item its index
    unsigned int rows;  // bit r: this warp handles pixel row r of the tile (0xF = the whole tile); kFillItem = a fill item
    unsigned int t0;    // SM clock when the item was fetched
};
// s_sched (shared, per block): [0] split entries [1] heavy entries [2] heavy threshold [3] split threshold [4] keep threshold
// [5] SM clock at block start [6] light entries [7] light threshold
__device__ __forceinline__ void sched_init(const TraceArgs& a, unsigned int* s_sched) {
    // The histogram is fetched by the whole block at once (one load per thread) and walked in shared memory, instead of
    // ~110 dependent global loads by one thread (each step decides whether the next is needed) ahead of the first tile.
    __shared__ unsigned int s_hist[kHintBins];
    if (a.hint_in)
        for (int i = threadIdx.x; i < kHintBins; i += blockDim.x) s_hist[i] = __ldg(a.hint_in + kHintHist + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int n_split = 0, n_heavy = 0, n_light = 0, thr_h = 0xffffffffu, thr_s = 0xffffffffu, thr_k = 0xffffffffu, thr_l = 0u;
        if (a.hint_in) {
            n_split = min(__ldg(a.hint_in + 0), (unsigned int)(4 * a.num_batches));
            n_heavy = min(__ldg(a.hint_in + 1), (unsigned int)a.num_batches);
            n_light = min(__ldg(a.hint_in + 4), (unsigned int)a.num_batches);
            const unsigned long long span = __ldg(a.hint_in + 3);
            const unsigned int tiles = __ldg(a.hint_in + 2);
            if (span && tiles) {
                auto pct = [&](int p) { const unsigned long long v = span * (unsigned long long)p / 100ull; return v > 0xfffffffeull ? 0xfffffffeu : (unsigned int)v; };
                thr_s = pct(a.hint_split_pct);
                thr_k = pct(a.hint_keep_pct);
                // heavy = the slowest heavy_pct percent of the timed tiles: walk the histogram down from the top
                const unsigned long long want = (unsigned long long)tiles * (unsigned long long)a.hint_heavy_pct / 100ull;
                unsigned long long seen = 0;
                int bin = kHintBins - 1;
                for (; bin > 0; bin--) {
                    seen += s_hist[bin];
                    if (seen >= want) break;
                }
                thr_h = (unsigned int)bin << kHintBinShift;
                if (thr_h > thr_s) thr_h = thr_s;
                // light = the quickest light_pct percent: walk up from the bottom (bin 0 holds no timed tile)
                const unsigned long long want_l = (unsigned long long)tiles * (unsigned long long)a.hint_light_pct / 100ull;
                unsigned long long seen_l = 0;
                int lb = 1;
                for (; lb < bin; lb++) {
                    seen_l += s_hist[lb];
                    if (seen_l > want_l) break;
                }
                thr_l = a.hint_light_pct > 0 ? (unsigned int)lb << kHintBinShift : 0u;
                if (thr_l >= thr_h) thr_l = 0u;
            }
        }
        s_sched[0] = n_split; s_sched[1] = n_heavy; s_sched[2] = thr_h; s_sched[3] = thr_s; s_sched[4] = thr_k;
        s_sched[5] = (unsigned int)clock();  // the block's warps start together
        s_sched[6] = n_light; s_sched[7] = thr_l;
    }
    __syncthreads();
}
// `ahead` (meaningful in lane 0): the queue index fetched one item ahead, so that the ~1 us round trip of the atomic (one
// address that every warp of the GPU hammers, 14 fetches per warp in a 1080p launch) hides behind the tile. MEASURED SLOWER,
// twice: round 1 without hints (+14 %: a warp holds an item nobody else can take) and round 2 with hints, where the queue
// ends on its quickest tiles (1080p primary 0.217 -> 0.240 ms, 4K 1/8-frame fused pass 0.172 -> 0.198 ms,
// profiles/r2_experiments.md 6). Off unless built with -DRTB_QUEUE_PREFETCH.
__device__ __forceinline__ unsigned int queue_fetch(const TraceArgs& a) { return (unsigned int)atomicAdd(a.work_counter, 1ull); }
#ifdef RTB_QUEUE_PREFETCH
static constexpr bool kQueuePrefetch = true;
#else
static constexpr bool kQueuePrefetch = false;
#endif
// Is tile `batch` inside the rectangle the queue enumerates? A list entry recorded under another rectangle (the camera
// moved, or the frame of the previous launch was assembled in wider groups) may lie outside now: it is dropped -- the fill
// items own everything outside, and every (tile, row) must have exactly ONE owner per launch (a second arrival would
// overshoot the row-assembly counter of its group, and a re-listed tile would be traced twice next time).
__device__ __forceinline__ bool tile_in_rect(const TraceArgs& a, long long batch) {
    const unsigned int b = (unsigned int)batch, kk = b / (unsigned int)a.tiles_x;
    const int tx = (int)(b - kk * (unsigned int)a.tiles_x);
    return (int)kk >= a.in_k0 && (int)kk < a.in_k1 && tx >= a.in_tx0 && tx < a.in_tx1;
}
template <bool PREFETCH>
__device__ __forceinline__ bool next_tile(const TraceArgs& a, const unsigned int* s_sched, int lane, TileWork& tw, unsigned int& ahead) {
    const bool prefetch = PREFETCH && a.hint_out != nullptr;
    for (;;) {
        unsigned long long i = 0;
        if (prefetch) {
            i = __shfl_sync(0xffffffffu, ahead, 0);
            if (lane == 0) ahead = queue_fetch(a);
        } else {
            if (lane == 0) i = queue_fetch(a);
            i = __shfl_sync(0xffffffffu, (unsigned int)i, 0);
        }
        tw.t0 = (unsigned int)clock();
        if (a.fill_first) {
            if (i < a.n_fill) {
                tw.batch = (long long)i;
                tw.rows = kFillItem;
                return true;
            }
            i -= a.n_fill;
        }
        const unsigned long long n_split = s_sched[0], n_heavy = s_sched[1];
        if (i < n_split) {
            // both lists were appended in completion order, i.e. roughly shortest first: walk them backwards (longest first)
            const unsigned int e = __ldg(a.hint_in + kHintHeader + (n_split - 1 - i));
            tw.batch = (long long)(e >> 2);
            tw.rows = 1u << (e & 3u);
            if (tw.batch >= a.num_batches || !tile_in_rect(a, tw.batch)) continue;  // (an index past the end cannot happen; never trust it)
            return true;
        }
        i -= n_split;
        if (i < n_heavy) {
            tw.batch = (long long)__ldg(a.hint_in + kHintHeader + 4 * a.num_batches + (n_heavy - 1 - i));
            tw.rows = 0xFu;
            if (tw.batch >= a.num_batches || !tile_in_rect(a, tw.batch)) continue;
            return true;
        }
        i -= n_heavy;
        const unsigned int wx = (unsigned int)(a.in_tx1 - a.in_tx0);
        const unsigned long long n_main = (unsigned long long)wx * (unsigned int)(a.in_k1 - a.in_k0);
        if (i >= n_main) {  // after the main pass: the light tiles, the slowest of them first
            i -= n_main;
            const unsigned long long n_light = s_sched[6];
            if (i >= n_light) {  // and last the store-only items for what lies outside the scene box's rectangle
                i -= n_light;
                if (a.fill_first || i >= a.n_fill) return false;
                tw.batch = (long long)i;
                tw.rows = kFillItem;
                return true;
            }
            tw.batch = (long long)__ldg(a.hint_in + kHintHeader + 5 * a.num_batches + (n_light - 1 - i));
            tw.rows = 0xFu;
            if (tw.batch >= a.num_batches || !tile_in_rect(a, tw.batch)) continue;
            return true;
        }
        {   // main pass: the tiles inside the rectangle, row-major
            const unsigned int j = (unsigned int)i, jr = j / wx;
            tw.batch = (long long)((unsigned int)(a.in_k0 + jr) * (unsigned int)a.tiles_x + (unsigned int)a.in_tx0 + (j - jr * wx));
        }
        tw.rows = 0xFu;
        if (a.hint_in) {  // rows that a list entry covers were (or will be) traced by that entry's warp
            const unsigned int f = __ldg(a.hint_in + kHintHeader + 6 * a.num_batches + tw.batch);
            const unsigned int covered = ((f & 0xffu) ? 1u : 0u) | ((f & 0xff00u) ? 2u : 0u) | ((f & 0xff0000u) ? 4u : 0u) | ((f & 0xff000000u) ? 8u : 0u);
            tw.rows = 0xFu & ~covered;
            if (!tw.rows) continue;
        }
        return true;
    }
}
// Called by all lanes when the item is done: record the tile's class for the next launch.
__device__ __forceinline__ void record_tile(const TraceArgs& a, const unsigned int* s_sched, int lane, const TileWork& tw) {
    if (!a.hint_out || lane != 0) return;
    const unsigned int dur = (unsigned int)clock() - tw.t0;
    unsigned char* flags = reinterpret_cast<unsigned char*>(a.hint_out + kHintHeader + 6 * a.num_batches) + 4 * tw.batch;
    if (tw.rows == 0xFu) {
        if (dur > 4096u) {  // a tile that traced something (an all-gated tile takes ~2000 cycles): the statistics are over these
            atomicAdd(a.hint_out + 2, 1u);
            atomicAdd(a.hint_out + kHintHist + min(dur >> kHintBinShift, (unsigned int)kHintBins - 1u), 1u);
        }
        unsigned int word = 0u;
        if (dur > s_sched[3]) {
            const unsigned int at = atomicAdd(a.hint_out + 0, 4u);
            for (unsigned int r = 0; r < 4u; r++) a.hint_out[kHintHeader + at + r] = ((unsigned int)tw.batch << 2) | r;
            word = 0x01010101u;
        } else if (dur > s_sched[2]) {
            const unsigned int at = atomicAdd(a.hint_out + 1, 1u);
            a.hint_out[kHintHeader + 4 * a.num_batches + at] = (unsigned int)tw.batch;
            word = 0x01010101u;
        } else if (dur > 4096u && dur <= s_sched[7]) {
            const unsigned int at = atomicAdd(a.hint_out + 4, 1u);
            a.hint_out[kHintHeader + 5 * a.num_batches + at] = (unsigned int)tw.batch;
            word = 0x01010101u;
        }
        *reinterpret_cast<unsigned int*>(flags) = word;
    } else {  // a single row or the uncovered rows of a tile: rows that are still slow stay on the split list
        const bool keep = dur > (__popc(tw.rows) == 1 ? s_sched[4] : s_sched[2]);
        const unsigned int n = keep ? __popc(tw.rows) : 0u;
        unsigned int at = n ? atomicAdd(a.hint_out + 0, n) : 0u;
        for (unsigned int r = 0; r < 4u; r++)
            if (tw.rows & (1u << r)) {
                flags[r] = keep ? 1 : 0;
                if (keep) a.hint_out[kHintHeader + at++] = ((unsigned int)tw.batch << 2) | r;
            }
    }
}
// Called once per warp when it retires: the launch span is the longest life of a persistent warp.
__device__ __forceinline__ void record_span(const TraceArgs& a, const unsigned int* s_sched, int lane) {
    if (a.hint_out && lane == 0) atomicMax(a.hint_out + 3, (unsigned int)clock() - s_sched[5]);
}
// A fill item: 32 tiles (256 x 4 pixels) of one tile row; every pixel of it that lies outside the scene box's rectangle
// gets the values a gated-out pixel gets -- the miss record(s) and `frame_miss` in the 4-byte frame (written to the
// destination itself: groups of tiles outside the rectangle are never row-assembled) -- in 512-byte row segments, and
// the tiles' hint flags are cleared so that a later launch which finds them inside does not read a stale class.
__device__ __forceinline__ void fill_outside(const TraceArgs& a, int lane, unsigned int item, unsigned int frame_miss) {
    const unsigned int segs = ((unsigned int)a.tiles_x + 31u) >> 5;
    const unsigned int kk = item / segs, seg = item - kk * segs;
    const bool row_inside = (int)kk >= a.in_k0 && (int)kk < a.in_k1;
    const int y0 = (int)(global_tile_row(a, kk) * 4u);
    const float4 miss = make_float4(__int_as_float(-1), RTB_T_INIT, 0.0f, 0.0f);
    for (int c = 0; c < 8; c++) {
        const int x = (int)(seg * 256u) + c * 32 + lane;
        const int tx = x >> 3;
        if (x >= a.w || (row_inside && tx >= a.in_tx0 && tx < a.in_tx1)) continue;
        for (int r = 0; r < 4 && y0 + r < a.h; r++) {
            const size_t px = (size_t)(y0 + r) * a.w + x;
            if (a.hits_out) a.hits_out[px] = miss;
            if (a.shadow_hits_out) a.shadow_hits_out[px] = miss;
            if (a.frame_out) a.frame_out[px] = frame_miss;
        }
    }
    if (a.hint_out) {
        const int tx = (int)(seg * 32u) + lane;
        if (tx < a.tiles_x && !(row_inside && tx >= a.in_tx0 && tx < a.in_tx1))
            a.hint_out[kHintHeader + 6 * a.num_batches + (size_t)kk * a.tiles_x + tx] = 0u;
    }
}

__device__ __forceinline__ void store_ray(float4* rays_out, long long i, const Ray& r) {
    rays_out[2 * i] = make_float4(r.ori.x, r.ori.y, r.ori.z, RTB_T_INIT);
    rays_out[2 * i + 1] = make_float4(r.dir.x, r.dir.y, r.dir.z, 0.0f);
}

// ---- persistent-warp trace kernel ------------------------------------------------------------------
// Grid = (resident blocks per SM) x 148 SMs; every warp pulls 32-ray batches from a global queue
// head with one atomicAdd by lane 0 + a shuffle, until the queue is empty.
template <int SRC, bool ANY_HIT, bool SMEM_TOP, bool FAST_BOX = false, bool INNER_EXIT = false>
__global__ void __launch_bounds__(kBlockThreads, (SRC == SRC_PRIMARY && !FAST_BOX) ? RTB_MINB_PRIMARY : RTB_MINB_BATCH) trace_kernel(const TraceArgs a, int smem_count) {
    extern __shared__ float4 smem_pairs[];
    if (SMEM_TOP) {
        for (int i = threadIdx.x; i < smem_count * 4; i += blockDim.x) smem_pairs[i] = a.scene.pairs[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    unsigned int traced = 0;
    __shared__ unsigned int s_sched[8];
    warm_l2(a);
    if (SRC == SRC_PRIMARY) sched_init(a, s_sched);
    unsigned int ahead = 0;
    if (SRC == SRC_PRIMARY && kQueuePrefetch && a.hint_out && lane == 0) ahead = queue_fetch(a);
    for (;;) {
        unsigned long long batch = 0;
        TileWork tw;
        if (SRC == SRC_PRIMARY) {
            if (!next_tile<kQueuePrefetch>(a, s_sched, lane, tw, ahead)) break;
            if (tw.rows == kFillItem) {
                fill_outside(a, lane, (unsigned int)tw.batch, 0xffffffffu);
                continue;
            }
            batch = (unsigned long long)tw.batch;
        } else {
            if (lane == 0) batch = atomicAdd(a.work_counter, 1ull);
            batch = __shfl_sync(0xffffffffu, batch, 0);
            if (batch >= (unsigned long long)a.num_batches) break;
        }
        RTB_TL_BEGIN(a);

        Ray ray;
        float tmax = RTB_T_INIT;
        long long out_index = -1;
        bool active = false;
        int tile_x = 0, tile_y0 = 0;
        long long tile_k = 0;
        if (SRC == SRC_BUFFER) {
            const long long i = (long long)batch * 32 + lane;
            if (i < a.n) {
                const float4 o = __ldg(a.rays_in + 2 * i), d = __ldg(a.rays_in + 2 * i + 1);
                ray.ori = ld3(o);
                ray.dir = ld3(d);
                tmax = o.w;
                out_index = i;
                active = true;
            }
        } else if (SRC == SRC_PRIMARY) {
            int x, y;
            tile_pixel(a, (long long)batch, lane, x, y, tile_x, tile_k);
            tile_y0 = y - (lane >> 3);
            if (x < a.w && y < a.h && ((tw.rows >> (lane >> 3)) & 1u)) {
                out_index = (long long)y * a.w + x;
                if (a.rays_out || !certainly_gated_out(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h))
                    active = primary_ray(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h, ray);
                if (a.rays_out) store_ray(a.rays_out, out_index, ray);
                if (!active) {
                    if (a.hits_out) a.hits_out[out_index] = make_float4(__int_as_float(-1), RTB_T_INIT, 0.0f, 0.0f);
                    if (a.frame_out) pixel_sink(a)[out_index] = 0xffffffffu;
                }
            }
        } else {  // SRC_SHADOW
            const long long i = (long long)batch * 32 + lane;
            if (i < a.n) {
                out_index = i;
                const float4 hin = __ldg(a.hits_in + i);
                if (__float_as_int(hin.x) >= 0) {
                    const float4 o = __ldg(a.rays_in + 2 * i), d = __ldg(a.rays_in + 2 * i + 1);
                    Ray pr;
                    pr.ori = ld3(o);
                    pr.dir = ld3(d);
                    f3 hp;
                    ray = shadow_ray(ld3(a.params.light_pos), pr, hin.y, hp);
                    active = true;
                    if (a.rays_out) store_ray(a.rays_out, i, ray);
                } else {
                    a.hits_out[i] = make_float4(__int_as_float(-1), RTB_T_INIT, 0.0f, 0.0f);
                    if (a.rays_out) {
                        a.rays_out[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        a.rays_out[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
        }
        if (active) {
            traced++;
            const TraceResult r = traverse<ANY_HIT, SMEM_TOP, FAST_BOX, INNER_EXIT>(a.scene, smem_pairs, smem_count, ray, tmax);
            if (SRC != SRC_PRIMARY || a.hits_out) a.hits_out[out_index] = make_float4(__int_as_float(r.idx), r.t, r.u, r.v);
            if (SRC == SRC_PRIMARY && a.frame_out) pixel_sink(a)[out_index] = (unsigned int)r.idx;
        }
        if (SRC == SRC_PRIMARY) {
            if (a.frame_out) finish_tile(a, tile_x, tile_k, tile_y0, lane, __popc(tw.rows));
            __syncwarp();
            record_tile(a, s_sched, lane, tw);
        }
        RTB_TL_END(a, batch, lane);
    }
    if (SRC == SRC_PRIMARY) record_span(a, s_sched, lane);
    retire_ray_count(a, traced, lane);
}

// ---- primary + shadow in ONE launch (BASELINE config 3 / 5) -------------------------------------------
// Per pixel: camera ray + scene gate (vR.cl:1156-1196), closest-hit traversal (the first traverse_bvh call, vR.cl:1238),
// and for every hit the any-hit shadow ray towards the light built exactly as vR.cl:1314,1407-1441. One launch pays one
// ramp-down and one fixed cost where rt_primary_device + rt_shadow_device pay two -- what limits a frame that is split
// over 8 GPUs. Outputs (each optional): the closest-hit records, the shadow any-hit records, and the 4-byte/pixel
// visibility word  vis = -1 (no hit)  |  3*triId + (occluded ? 1 : 0)  with the reference's rule
// occluded = shadow idx >= 0 && shadow t > 0.025 (vR.cl:1444-1449); 3*triId is a multiple of 3, so the word is lossless.
template <bool INNER_EXIT>
__global__ void __launch_bounds__(kBlockThreads) primary_shadow_kernel(const TraceArgs a) {
    const int lane = threadIdx.x & 31;
    const f3 light_pos = ld3(a.params.light_pos);
    unsigned int traced = 0;
    __shared__ unsigned int s_sched[8];
    warm_l2(a);
    sched_init(a, s_sched);
    unsigned int ahead = 0;
    if (kQueuePrefetch && a.hint_out && lane == 0) ahead = queue_fetch(a);
    for (;;) {
        TileWork tw;
        if (!next_tile<kQueuePrefetch>(a, s_sched, lane, tw, ahead)) break;
        if (tw.rows == kFillItem) {
            fill_outside(a, lane, (unsigned int)tw.batch, 0xffffffffu);
            continue;
        }
        const long long batch = tw.batch;
        RTB_TL_BEGIN(a);
        int x, y, tx;
        long long k;
        tile_pixel(a, batch, lane, x, y, tx, k);
        if (x < a.w && y < a.h && ((tw.rows >> (lane >> 3)) & 1u)) {
            const long long px = (long long)y * a.w + x;
            Ray ray;
            TraceResult hit;
            hit.idx = -1; hit.t = RTB_T_INIT; hit.u = hit.v = 0.0f;
            TraceResult sh = hit;
            if (!certainly_gated_out(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h) &&
                primary_ray(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h, ray)) {
                traced++;
                hit = traverse<false, false, false, INNER_EXIT>(a.scene, nullptr, 0, ray, RTB_T_INIT);
                if (hit.idx >= 0) {
                    traced++;
                    f3 hp;
                    const Ray sray = shadow_ray(light_pos, ray, hit.t, hp);
                    sh = traverse<true, false, false, INNER_EXIT>(a.scene, nullptr, 0, sray, RTB_T_INIT);
                }
            }
            if (a.hits_out) a.hits_out[px] = make_float4(__int_as_float(hit.idx), hit.t, hit.u, hit.v);
            if (a.shadow_hits_out) a.shadow_hits_out[px] = make_float4(__int_as_float(sh.idx), sh.t, sh.u, sh.v);
            if (a.frame_out) pixel_sink(a)[px] = hit.idx < 0 ? 0xffffffffu : (unsigned int)(hit.idx + ((sh.idx >= 0 && sh.t > 0.025f) ? 1 : 0));
        }
        if (a.frame_out) finish_tile(a, tx, k, y - (lane >> 3), lane, __popc(tw.rows));
        __syncwarp();
        record_tile(a, s_sched, lane, tw);
        RTB_TL_END(a, batch, lane);
    }
    record_span(a, s_sched, lane);
    retire_ray_count(a, traced, lane);
}

// ---- frame megakernel: the whole raytracer_bvh kernel (volumeRender.cl:1043-1547) ------------------

#define RTB_PI_F 3.14159274101257f
#define RTB_RAY_TRACE_DEPTH 3

// volumeRender.cl:27-53
__device__ __forceinline__ f3 normal_at_tri_point(f3 pNew, f3 p0, f3 p1, f3 p2, f3 vn0, f3 vn1, f3 vn2) {
    const float Det = p0.x * (p1.y * p2.z - p2.y * p1.z) - p1.x * (p0.y * p2.z - p2.y * p0.z) +
                      p2.x * (p0.y * p1.z - p1.y * p0.z);
    const float Det_l0 = pNew.x * (p1.y * p2.z - p2.y * p1.z) - p1.x * (pNew.y * p2.z - p2.y * pNew.z) +
                         p2.x * (pNew.y * p1.z - p1.y * pNew.z);
    const float Det_l1 = p0.x * (pNew.y * p2.z - p2.y * pNew.z) - pNew.x * (p0.y * p2.z - p2.y * p0.z) +
                         p2.x * (p0.y * pNew.z - pNew.y * p0.z);
    const float Det_l2 = p0.x * (p1.y * pNew.z - pNew.y * p1.z) - p1.x * (p0.y * pNew.z - pNew.y * p0.z) +
                         pNew.x * (p0.y * p1.z - p1.y * p0.z);
    const f3 l = mk3(Det_l0 / Det, Det_l1 / Det, Det_l2 / Det);
    return add3(add3(scale3(vn0, l.x), scale3(vn1, l.y)), scale3(vn2, l.z));
}

// volumeRender.cl:1732-1738
__device__ __forceinline__ float ggx_partial_geometry(float cosThetaN, float alpha) {
    const float cosTheta_sqr = clampf(cosThetaN * cosThetaN, 0.0f, 1.0f);
    const float tan2 = (1 - cosTheta_sqr) / cosTheta_sqr;
    return 2 / (1 + sqrtf(1 + alpha * alpha * tan2));
}
// volumeRender.cl:1740-1746
__device__ __forceinline__ float ggx_distribution(float cosThetaNH, float alpha) {
    const float alpha2 = alpha * alpha;
    const float NH_sqr = clampf(cosThetaNH * cosThetaNH, 0.0f, 1.0f);
    const float den = NH_sqr * alpha2 + (1.0f - NH_sqr);
    return alpha2 / (RTB_PI_F * den * den);
}
__device__ __forceinline__ float max0(float v) { return (0.0f < v) ? v : 0.0f; }
// volumeRender.cl:1754-1779 (FresnelSchlick :1748-1751 inlined)
__device__ __forceinline__ f3 cook_torrance_ggx(f3 n, f3 l, f3 v, f3 albedo, f3 f0, float roughness) {
    n = normalize3(n);
    v = normalize3(v);
    l = normalize3(l);
    const f3 h = normalize3(add3(v, l));
    const float NL = dot3(n, l);
    if (NL <= 0.0f) return mk3(0.f, 0.f, 0.f);
    const float NV = dot3(n, v);
    if (NV <= 0.0f) return mk3(0.f, 0.f, 0.f);
    const float NH = dot3(n, h);
    const float HV = dot3(h, v);
    const float roug_sqr = roughness * roughness;
    const float G = ggx_partial_geometry(NV, roug_sqr) * ggx_partial_geometry(NL, roug_sqr);
    const float D = ggx_distribution(NH, roug_sqr);
    const float p = powf(1.0f - clampf(HV, 0.0f, 1.0f), 5.0f);
    const f3 F = mk3(f0.x + (1.0f - f0.x) * p, f0.y + (1.0f - f0.y) * p, f0.z + (1.0f - f0.z) * p);
    const float GD = G * D;
    const float den = NV + 0.001f;
    const f3 specK = mk3(GD * F.x * 0.25f / den, GD * F.y * 0.25f / den, GD * F.z * 0.25f / den);
    const f3 diffK = mk3(clampf(1.0f - F.x, 0.f, 1.f), clampf(1.0f - F.y, 0.f, 1.f), clampf(1.0f - F.z, 0.f, 1.f));
    return mk3(max0(albedo.x * diffK.x * NL / RTB_PI_F + specK.x), max0(albedo.y * diffK.y * NL / RTB_PI_F + specK.y),
               max0(albedo.z * diffK.z * NL / RTB_PI_F + specK.z));
}
// volumeRender.cl:186-195
__device__ __forceinline__ unsigned int rgb_to_int(float r, float g, float b) {
    r = clampf(r, 0.0f, 255.0f);
    g = clampf(g, 0.0f, 255.0f);
    b = clampf(b, 0.0f, 255.0f);
    return ((unsigned int)b << 16) | ((unsigned int)g << 8) | ((unsigned int)r);
}

// One path vertex of raytracer_bvh (volumeRender.cl:1306-1403, 1407-1441, 1493-1497): shading of the hit, the shadow
// ray towards the light and the mirror-reflection ray.
struct PathVertex {
    f3 rez_color;
    Ray shadow;
    Ray reflect;
};
__device__ __forceinline__ PathVertex shade_path_vertex(const SceneView& s, f3 light_pos, const Ray& r, int hit_idx, float hit_t) {
    PathVertex pv;
    const f3 v0 = ld3(__ldg(s.verts + __ldg(s.indices + hit_idx + 0)));
    const f3 v1 = ld3(__ldg(s.verts + __ldg(s.indices + hit_idx + 1)));
    const f3 v2 = ld3(__ldg(s.verts + __ldg(s.indices + hit_idx + 2)));
    const f3 vn0 = ld3(__ldg(s.normals + __ldg(s.normal_indices + hit_idx + 0)));
    const f3 vn1 = ld3(__ldg(s.normals + __ldg(s.normal_indices + hit_idx + 1)));
    const f3 vn2 = ld3(__ldg(s.normals + __ldg(s.normal_indices + hit_idx + 2)));
    f3 vNew;
    pv.shadow = shadow_ray(light_pos, r, hit_t, vNew);  // vNew = o + d*(t-0.001)
    const f3 normal = normalize3(normal_at_tri_point(vNew, v0, v1, v2, vn0, vn1, vn2));
    const f3 l1 = normalize3(sub3(light_pos, vNew));
    const f3 v = normalize3(sub3(r.ori, vNew));
    const f3 n = normalize3(normal);
    const f3 diffuse = ld3(__ldg(s.mat_diffuse + __ldg(s.tri_to_material + hit_idx / 3)));
    const f3 f0 = scale3(mk3(40.f, 40.f, 40.f), (1 / 255.0f));
    pv.rez_color = scale3(cook_torrance_ggx(n, l1, v, diffuse, f0, 0.5f), 3.0f);
    pv.rez_color = add3(pv.rez_color, mul3(mk3(0.3f, 0.3f, 0.3f), diffuse));
    {  // reflect(i, n) = i - 2.0f * n * dot(n, i)
        const float d = dot3(normal, r.dir);
        const f3 refl = sub3(r.dir, scale3(scale3(normal, 2.0f), d));
        pv.reflect = ray_init(add3(vNew, scale3(refl, 0.001f)), refl);
    }
    return pv;
}
// volumeRender.cl:1519-1546: average over the path vertices, scale by the mean shadow coefficient, pack
__device__ __forceinline__ unsigned int resolve_pixel(f3 color, float shadow_coef_sum, int ray_depth) {
    if (ray_depth >= 1) {
        color = div3s(color, (float)ray_depth);
        shadow_coef_sum /= (float)ray_depth;
        color = scale3(color, shadow_coef_sum);
    } else {
        color = mk3(0.f, 0.f, 0.f);
    }
    return rgb_to_int(color.x * 255, color.y * 255, color.z * 255);
}

template <bool SMEM_TOP, bool INNER_EXIT>
__device__ __forceinline__ unsigned int render_pixel(const TraceArgs& a, const float4* smem_pairs, int smem_count, unsigned x,
                                                     unsigned y, unsigned int& traced) {
    const SceneView& s = a.scene;
    const f3 light_pos = ld3(a.params.light_pos);
    Ray r;
    bool continue_path = !certainly_gated_out(a.params, x, y, (unsigned)a.w, (unsigned)a.h) &&
                         primary_ray(a.params, x, y, (unsigned)a.w, (unsigned)a.h, r);
    f3 color = mk3(0.f, 0.f, 0.f);
    int ray_depth = 0;
    float shadow_coef_sum = 0.0f;

    while (continue_path && ray_depth < RTB_RAY_TRACE_DEPTH) {
        traced++;
        const TraceResult hit = traverse<false, SMEM_TOP, false, INNER_EXIT>(s, smem_pairs, smem_count, r, RTB_T_INIT);
        float shadow_coef = 1.0f;
        if (hit.idx >= 0) {
            ray_depth++;
            traced++;
            const PathVertex pv = shade_path_vertex(s, light_pos, r, hit.idx, hit.t);
            {
                const TraceResult sh = traverse<true, SMEM_TOP, false, INNER_EXIT>(s, smem_pairs, smem_count, pv.shadow, RTB_T_INIT);
                if (sh.idx >= 0 && sh.t > 0.025f) shadow_coef = 0.25f;
            }
            color = add3(color, pv.rez_color);
            shadow_coef_sum += shadow_coef;
            r = pv.reflect;
        } else {
            continue_path = false;
        }
    }
    return resolve_pixel(color, shadow_coef_sum, ray_depth);
}

template <bool SMEM_TOP, bool INNER_EXIT = false>
__global__ void __launch_bounds__(kBlockThreads, 8) render_kernel(const TraceArgs a, int smem_count) {
    extern __shared__ float4 smem_pairs[];
    if (SMEM_TOP) {
        for (int i = threadIdx.x; i < smem_count * 4; i += blockDim.x) smem_pairs[i] = a.scene.pairs[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    unsigned int traced = 0;
    __shared__ unsigned int s_sched[8];
    warm_l2(a);
    sched_init(a, s_sched);
    unsigned int ahead = 0;  // no fetch-ahead here: the extra live register costs this kernel 120 bytes of spills
    for (;;) {
        TileWork tw;
        if (!next_tile<false>(a, s_sched, lane, tw, ahead)) break;
        if (tw.rows == kFillItem) {
            fill_outside(a, lane, (unsigned int)tw.batch, 0u);  // a pixel that fails the gate is black (vR.cl:1542)
            continue;
        }
        const long long batch = tw.batch;
        RTB_TL_BEGIN(a);
        int x, y, tx;
        long long k;
        tile_pixel(a, batch, lane, x, y, tx, k);
        if (x < a.w && y < a.h && ((tw.rows >> (lane >> 3)) & 1u))
            pixel_sink(a)[(size_t)y * a.w + x] = render_pixel<SMEM_TOP, INNER_EXIT>(a, smem_pairs, smem_count, (unsigned)x, (unsigned)y, traced);
        finish_tile(a, tx, k, y - (lane >> 3), lane, __popc(tw.rows));
        __syncwarp();
        record_tile(a, s_sched, lane, tw);
        RTB_TL_END(a, batch, lane);
    }
    record_span(a, s_sched, lane);
    retire_ray_count(a, traced, lane);
}

// ---- synthetic incoherent workload (BASELINE config 4; not part of the reference) -------------------

__device__ __forceinline__ unsigned int lowbias32(unsigned int x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void diffuse_flags_kernel(long long n, const float4* hits, int* flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = __float_as_int(hits[i].x) >= 0 ? 1 : 0;
}

// `spp` cosine-weighted hemisphere directions about the geometric normal (flipped to face the
// incoming ray), origin = hit point + n * 1e-3; slot = (rank among valid hits) * spp + s.
__global__ void diffuse_rays_kernel(SceneView s, long long n, const float4* rays, const float4* hits, const int* offsets,
                                    int spp, unsigned int seed, float4* out_rays, long long* out_count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 h = hits[i];
    const int idx = __float_as_int(h.x);
    if (i == n - 1 && out_count) *out_count = ((long long)offsets[i] + (idx >= 0 ? 1 : 0)) * spp;
    if (idx < 0) return;
    const f3 o = ld3(rays[2 * i]), d = ld3(rays[2 * i + 1]);
    const f3 p0 = ld3(s.verts[s.indices[idx]]), p1 = ld3(s.verts[s.indices[idx + 1]]), p2 = ld3(s.verts[s.indices[idx + 2]]);
    f3 nrm = normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));
    if (dot3(nrm, d) > 0.0f) nrm = mk3(-nrm.x, -nrm.y, -nrm.z);
    const f3 hp = add3(o, scale3(d, h.y));
    const f3 org = add3(hp, scale3(nrm, 1e-3f));
    // orthonormal basis (Duff et al. 2017)
    const float sg = copysignf(1.0f, nrm.z);
    const float aa = -1.0f / (sg + nrm.z);
    const float bb = nrm.x * nrm.y * aa;
    const f3 t1 = mk3(1.0f + sg * nrm.x * nrm.x * aa, sg * bb, -sg * nrm.x);
    const f3 t2 = mk3(bb, sg + nrm.y * nrm.y * aa, -nrm.y);
    for (int k = 0; k < spp; k++) {
        const unsigned int h1 = lowbias32((unsigned int)(i * spp + k) + seed);
        const unsigned int h2 = lowbias32(h1 ^ 0x9e3779b9u);
        const float u1 = (float)(h1 >> 8) * (1.0f / 16777216.0f);
        const float u2 = (float)(h2 >> 8) * (1.0f / 16777216.0f);
        const float rr = sqrtf(u1), phi = 6.283185307179586f * u2;
        const float lx = rr * cosf(phi), ly = rr * sinf(phi), lz = sqrtf(fmaxf(0.0f, 1.0f - u1));
        const f3 dir = normalize3(add3(add3(scale3(t1, lx), scale3(t2, ly)), scale3(nrm, lz)));
        const long long slot = (long long)offsets[i] * spp + k;
        out_rays[2 * slot] = make_float4(org.x, org.y, org.z, RTB_T_INIT);
        out_rays[2 * slot + 1] = make_float4(dir.x, dir.y, dir.z, 0.0f);
    }
}

// ---- optional coherence pre-pass for ray buffers: sort key = direction octant | Morton code of the quantised origin |
// coarse direction. Tracing the rays in key order changes nothing per ray (results are scattered back to the caller's
// order) but puts rays that walk the same part of the tree into the same warp.
__device__ __forceinline__ unsigned int spread3(unsigned int v) {  // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void ray_sort_keys_kernel(long long n, const float4* rays, float3 lo, float3 inv_extent, unsigned int* keys, unsigned int* index) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 o = rays[2 * i], d = rays[2 * i + 1];
    const unsigned int octant = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
    const float qx = fminf(fmaxf((o.x - lo.x) * inv_extent.x, 0.f), 0.999f), qy = fminf(fmaxf((o.y - lo.y) * inv_extent.y, 0.f), 0.999f),
                qz = fminf(fmaxf((o.z - lo.z) * inv_extent.z, 0.f), 0.999f);
    const unsigned int morton = spread3((unsigned int)(qx * 128.f)) | (spread3((unsigned int)(qy * 128.f)) << 1) |
                                (spread3((unsigned int)(qz * 128.f)) << 2);  // 21 bits
    const unsigned int dx = (unsigned int)(fminf(fabsf(d.x), 0.999f) * 4.f), dy = (unsigned int)(fminf(fabsf(d.y), 0.999f) * 4.f),
                       dz = (unsigned int)(fminf(fabsf(d.z), 0.999f) * 4.f);
    keys[i] = (octant << 27) | (morton << 6) | (dx << 4) | (dy << 2) | dz;
    index[i] = (unsigned int)i;
}
__global__ void ray_gather_kernel(long long n, const float4* rays, const unsigned int* perm, float4* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned int src = perm[i];
    out[2 * i] = rays[2 * (size_t)src];
    out[2 * i + 1] = rays[2 * (size_t)src + 1];
}
__global__ void hit_scatter_kernel(long long n, const float4* hits_sorted, const unsigned int* perm, float4* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[perm[i]] = hits_sorted[i];
}

// ---- self test: div_hoisted(x, d, div_prepare(d)) and its packed form must equal x / d bit for bit over the admitted window ---
// d: any sign, exponent in [-20, 20]; x: 0 or any sign, exponent in [-100, 61]; mantissas random or edge patterns.
// Range probe: same comparison with the numerator's exponent fixed to `ex` (unbiased) and the divisor's exponent
// in [-20, 20]; used to establish how far below 1.0 the numerator may go before the hoisted sequence (which has no
// FCHK guard) stops matching the compiler's guarded division.
__global__ void selftest_division_range_kernel(unsigned int seed, int iters, int ex, unsigned long long* mismatches) {
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int state = lowbias32(tid ^ (seed * 0x9e3779b9u));
    unsigned long long bad = 0;
    for (int i = 0; i < iters; i++) {
        state = lowbias32(state + 0x6a09e667u);
        const unsigned int a = state;
        state = lowbias32(state + 0xbb67ae85u);
        const unsigned int b = state;
        state = lowbias32(state + 0x3c6ef372u);
        const unsigned int c = state;
        const unsigned int ed = 127u - 20u + (c % 40u);
        const float d = __uint_as_float(((a >> 31) << 31) | (ed << 23) | (a & 0x7FFFFFu));
        float x;
        if (ex >= -126) x = __uint_as_float(((b >> 31) << 31) | ((unsigned int)(ex + 127) << 23) | (b & 0x7FFFFFu));
        else x = __uint_as_float(((b >> 31) << 31) | ((b & 0x7FFFFFu) >> (-126 - ex)));  // denormal with top bit at 2^ex
        const float want = x / d;
        const float r1 = div_prepare(d);
        const float got = div_hoisted(x, d, r1);
        if (__float_as_uint(want) != __float_as_uint(got) && !(want == 0.0f && got == 0.0f)) bad++;
        // the packed form the traversal kernels execute (FMUL2 + 2 FFMA2): both halves, numerators x and -x
        float g0, g1;
        unpack2(div_hoisted2_core(pack2(x, -x), pack2(-d, -d), pack2(r1, r1)), g0, g1);
        if (__float_as_uint(want) != __float_as_uint(g0) && !(want == 0.0f && g0 == 0.0f)) bad++;
        if (__float_as_uint(-want) != __float_as_uint(g1) && !(want == 0.0f && g1 == 0.0f)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

__global__ void selftest_division_kernel(unsigned int seed, int iters, unsigned long long* mismatches) {
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int state = lowbias32(tid ^ (seed * 0x9e3779b9u));
    const unsigned int edge[8] = {0u, 0x7FFFFFu, 0x400000u, 1u, 0x7FFFFEu, 0x555555u, 0x2AAAAAu, 0x7FF000u};
    unsigned long long bad = 0;
    for (int i = 0; i < iters; i++) {
        state = lowbias32(state + 0x6a09e667u);
        const unsigned int a = state;
        state = lowbias32(state + 0xbb67ae85u);
        const unsigned int b = state;
        state = lowbias32(state + 0x3c6ef372u);
        const unsigned int c = state;
        unsigned int md = a & 0x7FFFFFu, mx = b & 0x7FFFFFu;
        if ((c & 7u) == 0u) md = edge[(c >> 3) & 7u];
        if ((c & 0x38u) == 0u) mx = edge[(c >> 6) & 7u];
        const unsigned int ed = 127u - 20u + ((c >> 9) % 41u);   // exponent -20..20
        const unsigned int ex = 127u - 100u + ((c >> 16) % 162u);  // exponent -100..61
        const float d = __uint_as_float(((a >> 31) << 31) | (ed << 23) | md);
        float x = __uint_as_float(((b >> 31) << 31) | (ex << 23) | mx);
        if ((c >> 24) == 0u) x = 0.0f;
        if (fabsf(d) > 1048576.0f) continue;  // exponent 20 with a non-zero mantissa is outside the window
        const float want = x / d;
        const float r1 = div_prepare(d);
        const float got = div_hoisted(x, d, r1);
        if (__float_as_uint(want) != __float_as_uint(got) && !(want == 0.0f && got == 0.0f)) bad++;
        // the packed form the traversal kernels execute (FMUL2 + 2 FFMA2): both halves, numerators x and -x
        float g0, g1;
        unpack2(div_hoisted2_core(pack2(x, -x), pack2(-d, -d), pack2(r1, r1)), g0, g1);
        if (__float_as_uint(want) != __float_as_uint(g0) && !(want == 0.0f && g0 == 0.0f)) bad++;
        if (__float_as_uint(-want) != __float_as_uint(g1) && !(want == 0.0f && g1 == 0.0f)) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}


// Exhaustive form of the self test: EVERY mantissa pair (2^23 x 2^23 = 2^46 quotients) at one exponent / sign pair.
// Scaling either operand by a power of two scales the quotient, every intermediate of the hoisted sequence and every
// rounding boundary by the same power of two (no overflow / underflow inside the admitted window, which the range
// probe establishes), so one exponent pair stands for all of them. blockIdx.y selects the divisor mantissa md0 + y,
// the x dimension strides over all 2^23 numerator mantissas.
__global__ void selftest_division_exhaustive_kernel(unsigned int md0, int ex_x, int ex_d, unsigned int signs,
                                                    unsigned long long* mismatches) {
    const unsigned int md = md0 + blockIdx.y;
    const float d = __uint_as_float(((signs & 1u) << 31) | ((unsigned int)(ex_d + 127) << 23) | (md & 0x7FFFFFu));
    const float r1 = div_prepare(d);
    const f32x2 nd2 = pack2(-d, -d), r2 = pack2(r1, r1);
    const unsigned int xs = ((signs >> 1) & 1u) << 31 | ((unsigned int)(ex_x + 127) << 23);
    unsigned int bad = 0;
    for (unsigned int mx = blockIdx.x * blockDim.x + threadIdx.x; mx < (1u << 23); mx += gridDim.x * blockDim.x) {
        const float x = __uint_as_float(xs | mx);
        const float want = x / d;
        const float got = div_hoisted(x, d, r1);
        float g0, g1;
        unpack2(div_hoisted2_core(pack2(x, -x), nd2, r2), g0, g1);
        bad += (__float_as_uint(want) != __float_as_uint(got)) + (__float_as_uint(want) != __float_as_uint(g0)) +
               (__float_as_uint(-want) != __float_as_uint(g1));
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// ---- validation of an adopted blob (rt_adopt_scene_blob) --------------------------------------------------------------
// A blob uploaded by rt_upload_scene was validated on the host before packing; one that arrives in device memory (NCCL
// broadcast, peer copy) is checked here once: every child reference of every node pair points inside the pair / triangle
// sections, every packed triangle's id indexes the shading arrays, and the sentinel behind the last triangle terminates
// any leaf run -- so a stale or corrupt buffer is rejected instead of read out of bounds or looped over forever.
__global__ void validate_blob_kernel(const float4* pairs, int num_pairs, const float4* tris, int num_tris, int T, int root_ref,
                                     unsigned int* bad) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned int wrong = 0;
    auto ref_ok = [&](int ref) {
        if (ref == kRefPoison) return true;
        return ref >= 0 ? ref < num_pairs : ~ref <= num_tris;
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < num_pairs; i += stride) {
        wrong += !ref_ok(__float_as_int(pairs[4 * i + 1].z));
        wrong += !ref_ok(__float_as_int(pairs[4 * i + 3].z));
    }
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j <= num_tris; j += stride) {
        const int id = __float_as_int(tris[3 * j].w), last = __float_as_int(tris[3 * j + 1].w);
        if (j == num_tris) wrong += (last == 0);  // the sentinel ends every leaf run
        else wrong += (id < 0 || id % 3 != 0 || (T > 0 && id / 3 >= T) || (last != 0 && last != 1));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) wrong += !ref_ok(root_ref);
    if (wrong) atomicAdd(bad, wrong);
}

// ---- SASS probes (tests/test_sass.py reads them with cuobjdump; they are never launched) ---------------------------
// probe_ray_triangle_kernel holds exactly one inlined ray_triangle(); probe_reciprocal_kernel exactly one IEEE 1.0f / x.
// If ptxas contracted any product-sum of Moller-Trumbore into an FFMA, the first kernel would show more FFMAs than the
// division's own fixed-up reciprocal sequence, and fewer than 27 FMUL / 18 FADD.
__global__ void probe_ray_triangle_kernel(const float4* __restrict__ in, float4* __restrict__ out) {
    const float4 o = in[0], d = in[1], a = in[2], b = in[3], c = in[4];
    Ray r;
    r.ori = ld3(o);
    r.dir = ld3(d);
    float u = 0.0f, v = 0.0f;
    const float t = ray_triangle(r, ld3(a), ld3(b), ld3(c), u, v);
    out[0] = make_float4(t, u, v, 0.0f);
}
__global__ void probe_reciprocal_kernel(const float* __restrict__ in, float* __restrict__ out) { out[0] = 1.0f / in[0]; }

}  // namespace rtb
