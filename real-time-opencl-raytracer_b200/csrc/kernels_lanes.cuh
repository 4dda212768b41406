// kernels_lanes.cuh -- "persistent lanes" scheduler around the same per-ray traversal semantics.
//
// trace_kernel (kernels.cuh) gives every warp a batch of 32 rays and runs traverse() to completion:
// lanes whose ray ends early idle until the slowest ray of the batch is done, and lanes drift apart
// between the inner-node loop and the leaf loop (ncu: 16.3 of 32 threads active per instruction).
// Here a lane is the unit of scheduling:
//   * REFILL   lanes without a ray take the next work items from a warp-local chunk of the global
//              queue (one atomicAdd by lane 0 per chunk, ballot + popc give each empty lane its item)
//              as soon as `refill_threshold` lanes are empty
//   * INNER    all lanes that stand on an inner node do node-pair visits in lock step (warp-uniform
//              loop on a ballot); a lane drops out when it reaches a leaf or finishes its ray
//   * LEAF     all lanes that stand on a leaf test that leaf's triangles together
// Per ray, the sequence of box tests, pushes, pops and triangle tests is exactly the one of traverse():
// results are bit-identical; only WHEN a lane executes its next step changes.
#pragma once
#include "kernels.cuh"

namespace rtb {

static constexpr int kLaneChunk = 32;   // work items a warp takes from the global queue per atomicAdd

// WIDE_L2: the node loads ask L2 for 256-byte fills (device_math.cuh ldg256): the instance for scenes that do not fit L2.
template <int SRC, bool ANY_HIT, bool WIDE_L2 = false>
__global__ void __launch_bounds__(kBlockThreads, RTB_MINB_LANES) trace_lanes_kernel(const TraceArgs a, int refill_threshold, int inner_exit_threshold) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    warm_l2(a);
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long total = (SRC == SRC_PRIMARY) ? (unsigned long long)a.num_batches * 32ull : (unsigned long long)a.n;

    // warp-local slice of the work queue
    unsigned long long q_next = 0, q_end = 0;
    bool queue_open = true;

    // lane state
    bool has_ray = false;
    unsigned int traced = 0;
    RayX rx;  // the lane's ray lives in here (ray_of(rx)); a separate Ray would cost 6 more registers per lane
    float tHit = RTB_T_INIT;
    TraceResult res;
    long long out_index = 0;
    int cur = 0, sp = 0;
    int stack[RTB_STACK];
    {
        Ray idle;
        idle.ori = idle.dir = mk3(0.f, 0.f, 1.f);
        rx = ray_prepare(idle, false);
    }
    res.idx = -1; res.t = 0.f; res.u = 0.f; res.v = 0.f;

    for (;;) {
        // ---- REFILL ----------------------------------------------------------------------------
        const unsigned empty = __ballot_sync(FULL, !has_ray);
        const int n_empty = __popc(empty);
        if (queue_open && n_empty >= refill_threshold) {
            const unsigned long long avail = q_end - q_next;
            unsigned long long fresh = 0;
            if ((unsigned long long)n_empty > avail) {  // warp-uniform
                if (lane == 0) fresh = atomicAdd(a.work_counter, (unsigned long long)kLaneChunk);
                fresh = __shfl_sync(FULL, fresh, 0);
            }
            const unsigned rank = __popc(empty & lt_mask);
            unsigned long long item = total;  // "nothing"
            if (!has_ray) {
                if (rank < avail) item = q_next + rank;
                else if ((unsigned long long)n_empty > avail) item = fresh + (rank - avail);
            }
            if ((unsigned long long)n_empty > avail) {
                q_next = fresh + ((unsigned long long)n_empty - avail);
                q_end = fresh + kLaneChunk;
                if (fresh >= total) queue_open = false;
            } else {
                q_next += n_empty;
            }
            if (item < total) {
                bool active = false;
                Ray ray;
                tHit = RTB_T_INIT;
                if (SRC == SRC_BUFFER) {
                    const float4 o = __ldg(a.rays_in + 2 * item), d = __ldg(a.rays_in + 2 * item + 1);
                    ray.ori = ld3(o);
                    ray.dir = ld3(d);
                    tHit = o.w;
                    out_index = (long long)item;
                    active = true;
                } else if (SRC == SRC_PRIMARY) {
                    int x, y;
                    tile_pixel(a, (long long)(item >> 5), (int)(item & 31), x, y);
                    if (x < a.w && y < a.h) {
                        out_index = (long long)y * a.w + x;
                        if (a.rays_out || !certainly_gated_out(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h))
                            active = primary_ray(a.params, (unsigned)x, (unsigned)y, (unsigned)a.w, (unsigned)a.h, ray);
                        if (a.rays_out) store_ray(a.rays_out, out_index, ray);
                        if (!active) a.hits_out[out_index] = make_float4(__int_as_float(-1), RTB_T_INIT, 0.0f, 0.0f);
                    }
                } else {  // SRC_SHADOW
                    out_index = (long long)item;
                    const float4 hin = __ldg(a.hits_in + item);
                    if (__float_as_int(hin.x) >= 0) {
                        const float4 o = __ldg(a.rays_in + 2 * item), d = __ldg(a.rays_in + 2 * item + 1);
                        Ray pr;
                        pr.ori = ld3(o);
                        pr.dir = ld3(d);
                        f3 hp;
                        ray = shadow_ray(ld3(a.params.light_pos), pr, hin.y, hp);
                        active = true;
                        if (a.rays_out) store_ray(a.rays_out, out_index, ray);
                    } else {
                        a.hits_out[item] = make_float4(__int_as_float(-1), RTB_T_INIT, 0.0f, 0.0f);
                        if (a.rays_out) {
                            a.rays_out[2 * item] = make_float4(0.f, 0.f, 0.f, 0.f);
                            a.rays_out[2 * item + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                if (active) {
                    traced++;
                    rx = ray_prepare(ray, a.scene.coords_in_window != 0);
                    cur = a.scene.root_ref;
                    sp = 0;
                    res.idx = -1;
                    res.u = res.v = 0.0f;
                    has_ray = true;
                }
            }
        }
        if (!__any_sync(FULL, has_ray)) {
            if (!queue_open) break;
            continue;
        }
        const bool warp_hoisted = __all_sync(FULL, !has_ray || rx.fast);

        // ---- INNER: node-pair visits in lock step -----------------------------------------------
        for (;;) {
            const bool inner = has_ray && cur >= 0;
            const unsigned m = __ballot_sync(FULL, inner);
            if (m == 0) break;
            if (__popc(m) < inner_exit_threshold && __ballot_sync(FULL, has_ray && cur < 0) != 0) break;
            if (inner) {
                const float4* p = a.scene.pairs + 4 * (size_t)cur;
                float4 q0, q1, q2, q3;
                ldg256<WIDE_L2>(p, q0, q1);
                ldg256<WIDE_L2>(p + 2, q2, q3);
                float t0n, t0f, t1n, t1f;
                if (warp_hoisted || rx.fast) {  // warp_hoisted is uniform: no divergent branch in the common case
                    ray_box_hoisted(rx, q0, q1, t0n, t0f);
                    ray_box_hoisted(rx, q2, q3, t1n, t1f);
                } else {
                    const Ray ray = ray_of(rx);
                    ray_box(ray, q0, q1, t0n, t0f);
                    ray_box(ray, q2, q3, t1n, t1f);
                }
                const bool hit0 = (t0n <= t0f) && (t0f >= RTB_TMIN) && (t0n <= tHit);
                const bool hit1 = (t1n <= t1f) && (t1f >= RTB_TMIN) && (t1n <= tHit);
                int c0 = box_ref(q1), c1 = box_ref(q3);
                if (hit0 && hit1) {
                    if (t0n > t1n) { const int t = c0; c0 = c1; c1 = t; }
                    if (sp >= RTB_STACK) {  // reference: stack overflow returns -1 and drops the hit (vR.cl:914)
                        a.hits_out[out_index] = make_float4(__int_as_float(-1), tHit, 0.0f, 0.0f);
                        has_ray = false;
                    } else {
                        stack[sp++] = c1;
                        cur = c0;
                    }
                } else if (hit0) {
                    cur = c0;
                } else if (hit1) {
                    cur = c1;
                } else if (sp == 0) {
                    a.hits_out[out_index] = make_float4(__int_as_float(res.idx), tHit, res.u, res.v);
                    has_ray = false;
                } else {
                    cur = stack[--sp];
                }
            }
        }

        // ---- LEAF: the triangles of the leaf each lane stands on ------------------------------------
        if (has_ray && cur < 0) {
            bool done = false;
            if (cur == kRefPoison) {  // vR.cl:841,855-856
                res.idx = -1; res.u = res.v = 0.0f;
                done = true;
            } else {
                const float4* tp = a.scene.tris + 3 * (size_t)(~cur);
                const Ray ray = ray_of(rx);
                for (;;) {
                    const float4 ta = __ldg(tp), tb = __ldg(tp + 1), tc = __ldg(tp + 2);
                    float u, v;
                    const float t = ray_triangle(ray, ld3(ta), ld3(tb), ld3(tc), u, v);
                    if (t < tHit && t > RTB_TMIN) {
                        tHit = t;
                        res.idx = __float_as_int(ta.w);
                        res.u = u;
                        res.v = v;
                        if (ANY_HIT) { done = true; break; }
                    }
                    if (__float_as_int(tb.w) != 0) break;
                    tp += 3;
                }
                if (!done) {
                    if (sp == 0) done = true;
                    else cur = stack[--sp];
                }
            }
            if (done) {
                a.hits_out[out_index] = make_float4(__int_as_float(res.idx), tHit, res.u, res.v);
                has_ray = false;
            }
        }
        __syncwarp();
    }
    retire_ray_count(a, traced, lane);
}

}  // namespace rtb
