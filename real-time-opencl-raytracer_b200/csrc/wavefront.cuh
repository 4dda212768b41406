// wavefront.cuh -- the whole raytracer_bvh frame (volumeRender.cl:1043-1547) as a wavefront pipeline.
//
// The reference (and render_kernel in kernels.cuh) runs one thread per pixel through up to three path segments of
// closest-hit + shading + any-hit shadow + reflection: lanes whose path ended idle while their neighbours bounce on,
// and coherent primary rays share a warp with incoherent reflections. Here every stage gets full warps of like work:
//
//   wf_trace_shade<PRIMARY>   camera rays (8x4 tiles) -> closest hit -> shade -> pixel colour, queue shadow + reflection rays
//   wf_trace_shade<BOUNCE> x2 queued reflection rays -> closest hit -> shade -> colour += , queue shadow (+ reflection) rays
//   wf_shadow x3              queued shadow rays -> any hit -> coef += (occluded && t > 0.025 ? 0.25 : 1)
//   wf_resolve                colour / depth * (coef / depth) -> RGBA8
//
// Queues are compacted with warp-aggregated atomics (ballot + popc + one atomicAdd per warp); a queued ray carries its
// pixel index, so the order inside a queue is a scheduling detail only. Per pixel the arithmetic is the megakernel's:
// colour accumulates in bounce order, the coefficients are 1 or 0.25 (their sum is exact in any order), every ray is
// built by the same shade_path_vertex(). Frames are bit-identical to render_kernel's (tests/test_gpu_parity.py).
// The shadow stage of bounce k runs on a second stream concurrently with the trace stage of bounce k+1.
#pragma once
#include "kernels.cuh"

namespace rtb {

struct WavefrontArgs {
    SceneView scene;
    ParamsBlock params;
    int w, h, tiles_x, part, n_parts, band_tile_rows;
    long long num_batches;      // PRIMARY: tiles of this rank's bands
    int tile_order;
    unsigned int order_mul;
    // per-pixel path state (index = y*w+x)
    float4* color;              // xyz = sum of rez_color over the path vertices
    float* coef;                // sum of shadow coefficients
    int* depth;                 // ray_depth
    // queues: 2 x float4 per ray {o.xyz, bits(pixel)} {d.xyz, -}
    const float4* rays_in;      // BOUNCE: reflection rays queued by the previous trace stage
    const unsigned long long* n_in;  // BOUNCE / SHADOW: number of queued rays (device memory)
    float4* next_out;           // reflection rays for the next trace stage (NULL: last bounce)
    unsigned long long* n_next;
    float4* shadow_out;         // shadow rays of this stage
    unsigned long long* n_shadow;
    unsigned long long* work_counter;
    unsigned int* frame_out;    // resolve
};

// Append one ray per participating lane to a queue: one atomicAdd per warp.
__device__ __forceinline__ void queue_push(bool pred, float4* queue, unsigned long long* count, const Ray& r, int pixel) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) {
        const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
        queue[2 * slot] = make_float4(r.ori.x, r.ori.y, r.ori.z, __int_as_float(pixel));
        queue[2 * slot + 1] = make_float4(r.dir.x, r.dir.y, r.dir.z, 0.0f);
    }
}

template <bool PRIMARY>
__global__ void __launch_bounds__(kBlockThreads) wf_trace_shade_kernel(const WavefrontArgs a) {
    const int lane = threadIdx.x & 31;
    const f3 light_pos = ld3(a.params.light_pos);
    const unsigned long long batches =
        PRIMARY ? (unsigned long long)a.num_batches : (__ldg(a.n_in) + 31ull) / 32ull;
    for (;;) {
        unsigned long long batch = 0;
        if (lane == 0) batch = atomicAdd(a.work_counter, 1ull);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        if (batch >= batches) break;
        Ray ray;
        int pixel = -1;
        bool active = false;
        if (PRIMARY) {
            TraceArgs t;  // only the tile-mapping fields are read
            t.tiles_x = a.tiles_x; t.part = a.part; t.n_parts = a.n_parts; t.band_tile_rows = a.band_tile_rows;
            t.num_batches = a.num_batches; t.tile_order = a.tile_order; t.order_mul = a.order_mul;
            int x, y;
            tile_pixel(t, (long long)batch, lane, x, y);
            if (x < a.w && y < a.h) {
                pixel = y * a.w + x;
                active = !certainly_gated_out(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h) &&
                         primary_ray(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h, ray);
                a.color[pixel] = make_float4(0.f, 0.f, 0.f, 0.f);
                a.coef[pixel] = 0.0f;
                a.depth[pixel] = 0;
            }
        } else {
            const unsigned long long i = batch * 32ull + lane;
            if (i < __ldg(a.n_in)) {
                const float4 o = __ldg(a.rays_in + 2 * i), d = __ldg(a.rays_in + 2 * i + 1);
                ray.ori = ld3(o);
                ray.dir = ld3(d);
                pixel = __float_as_int(o.w);
                active = true;
            }
        }
        bool hit = false;
        PathVertex pv;
        if (active) {
            const TraceResult r = traverse<false, false>(a.scene, nullptr, 0, ray, RTB_T_INIT);
            if (r.idx >= 0) {
                hit = true;
                pv = shade_path_vertex(a.scene, light_pos, ray, r.idx, r.t);
                if (PRIMARY) {
                    a.color[pixel] = make_float4(pv.rez_color.x, pv.rez_color.y, pv.rez_color.z, 0.f);
                    a.depth[pixel] = 1;
                } else {  // colour accumulates in bounce order: one ray per pixel per stage, stages are stream-ordered
                    const float4 c = a.color[pixel];
                    const f3 sum = add3(mk3(c.x, c.y, c.z), pv.rez_color);
                    a.color[pixel] = make_float4(sum.x, sum.y, sum.z, 0.f);
                    a.depth[pixel] += 1;
                }
            }
        }
        queue_push(hit, a.shadow_out, a.n_shadow, pv.shadow, pixel);
        if (a.next_out) queue_push(hit, a.next_out, a.n_next, pv.reflect, pixel);
    }
}

// Dense shading of the hits a bounce's trace stage queued: every lane has a path vertex to shade.
__global__ void __launch_bounds__(256) wf_shade_kernel(const WavefrontArgs a, const float4* shade_queue, const unsigned long long* n_shade) {
    const unsigned long long n = __ldg(n_shade);
    const f3 light_pos = ld3(a.params.light_pos);
    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < n; base += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = base + threadIdx.x;
        const bool valid = i < n;
        PathVertex pv;
        int pixel = -1;
        if (valid) {
            const float4 o = __ldg(shade_queue + 3 * i), d = __ldg(shade_queue + 3 * i + 1), h = __ldg(shade_queue + 3 * i + 2);
            Ray ray;
            ray.ori = ld3(o);
            ray.dir = ld3(d);
            pixel = __float_as_int(o.w);
            pv = shade_path_vertex(a.scene, light_pos, ray, __float_as_int(h.x), d.w);
            const float4 c = a.color[pixel];
            const f3 sum = add3(mk3(c.x, c.y, c.z), pv.rez_color);
            a.color[pixel] = make_float4(sum.x, sum.y, sum.z, 0.f);
            a.depth[pixel] += 1;
        }
        queue_push(valid, a.shadow_out, a.n_shadow, pv.shadow, pixel);
        if (a.next_out) queue_push(valid, a.next_out, a.n_next, pv.reflect, pixel);
    }
}

__global__ void __launch_bounds__(kBlockThreads) wf_shadow_kernel(const WavefrontArgs a) {
    const int lane = threadIdx.x & 31;
    const unsigned long long n = __ldg(a.n_in);
    const unsigned long long batches = (n + 31ull) / 32ull;
    for (;;) {
        unsigned long long batch = 0;
        if (lane == 0) batch = atomicAdd(a.work_counter, 1ull);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        if (batch >= batches) break;
        const unsigned long long i = batch * 32ull + lane;
        if (i < n) {
            const float4 o = __ldg(a.rays_in + 2 * i), d = __ldg(a.rays_in + 2 * i + 1);
            Ray ray;
            ray.ori = ld3(o);
            ray.dir = ld3(d);
            const int pixel = __float_as_int(o.w);
            const TraceResult r = traverse<true, false>(a.scene, nullptr, 0, ray, RTB_T_INIT);
            const float c = (r.idx >= 0 && r.t > 0.025f) ? 0.25f : 1.0f;  // volumeRender.cl:1444-1449
            a.coef[pixel] += c;  // one shadow ray per pixel per stage; shadow stages are stream-ordered
        }
    }
}

__global__ void wf_resolve_kernel(const WavefrontArgs a) {
    // same band mapping as the trace stage, one thread per pixel of this rank's bands
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long batch = gid >> 5;
    if (batch >= a.num_batches) return;
    TraceArgs t;
    t.tiles_x = a.tiles_x; t.part = a.part; t.n_parts = a.n_parts; t.band_tile_rows = a.band_tile_rows;
    t.num_batches = a.num_batches; t.tile_order = 0; t.order_mul = 1;
    int x, y;
    tile_pixel(t, batch, (int)(gid & 31), x, y);
    if (x >= a.w || y >= a.h) return;
    const int pixel = y * a.w + x;
    const float4 c = a.color[pixel];
    a.frame_out[pixel] = resolve_pixel(mk3(c.x, c.y, c.z), a.coef[pixel], a.depth[pixel]);
}

}  // namespace rtb
