/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference OpenCL kernel).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so. Nothing under real-time-opencl-raytracer_b200/ links, imports or calls it.
 *
 * Parity status: PINNED. The reference ships no tests / golden vectors. (1) On the GPU box the reference's own,
 * unmodified OpenCL kernel runs through NVIDIA's ICD (oracle/_ref/libref_cl.so, oracle/ref_cl_driver.c) and its
 * frames are compared with orc_render_frame (tests/test_reference_opencl.py): identical coverage, colours identical
 * up to the FMA-contraction / built-in-rounding differences of that OpenCL build (a handful of pixels).
 * (2) Pinned bit-exactly against the reference's own compiled host code (oracle/_ref/ref_host, built from
 * /root/reference unmodified):
 *   - orc_ray_triangle       == spec::RayTriangleIntersection (common.h:193-219), bit-exact
 *   - orc_scene_box_gate     == spec::RayBoxIntersection      (common.h:172-190), bit-exact (non-NaN)
 *   - dot / cross / normalize == vectors_math.cpp:73-84, bit-exact
 *   - the flat BVH the oracle walks is produced by the reference's own SplitBVHBuilder/BVH_Cuda
 * The traversal loop itself (volumeRender.cl:658-1010) is restated line by line and cross-checked
 * against the author's brute-force loop (volumeRender.cl:690-712); per-ray hit indices / t / u / v are
 * oracle-defined (the reference kernel exports pixels only), its frames are checked as described in (1).
 *
 * All data are in the REFERENCE layouts (RayTracer.cpp:942-984):
 *   verts   float4[V]  (w = 1)                     mesh1.vertices
 *   indices int[3T]                                mesh1.indices
 *   nodes   48 B[N] = {float4 min, float4 max, int left, right, tri_off, tri_cnt}   BVH_Cuda.h:12-29
 *   tri_indices int[R], pre-multiplied by 3        BVH_Cuda.h:90-93
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_T_INIT 4294967296.0f /* (float)UINT_MAX, volumeRender.cl:224 */

typedef struct {
    const float* verts;       /* float4[V] */
    const int* indices;       /* int[3T] */
    const float* nodes;       /* 12 x 32-bit words per node */
    const int* tri_indices;   /* int[R] */
    int num_verts, num_tris, num_bvh_tris, num_bvh_nodes;
    /* shading inputs (may be NULL when only orc_trace / orc_primary are used) */
    const float* normals;     /* float4[Vn] */
    const int* normal_indices;/* int[3T] */
    const float* materials;   /* 44 x 32-bit words (176 B) per material, Mesh.h:20-33 */
    const int* tri_to_material; /* int[T] */
} orc_scene;

/* rays: 8 floats each {ox,oy,oz,tmax, dx,dy,dz,reserved}; dir is used as given (the caller has
 * already normalised it, as RayInit volumeRender.cl:205-211 does); tmax = initial *tHit.
 * hits: 4 x 32-bit each {int idx (3*triId or -1), float t, float u, float v}.
 * mode: 0 = closest-hit (needClosestHit=true), 1 = any-hit (false).
 * counters (optional, may be NULL): [0] inner-node visits, [1] leaf visits, [2] triangle tests,
 *                                   [3] max stack depth seen. */
void orc_trace(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits, uint64_t* counters);

/* same as orc_trace, additionally stats[3*i+{0,1,2}] = inner visits, leaf visits, triangle tests of ray i */
void orc_trace_stats(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits, uint32_t* stats);

/* brute-force loop of volumeRender.cl:690-712 (the author's "works !!!" body), same outputs */
void orc_trace_bruteforce(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits);

/* primary rays of volumeRender.cl:1156-1196 for a w x h frame from the 128-byte Params block
 * (8 x float4: a,b,c,campos,light_pos,light_color,aabb_min,aabb_max). Writes rays in rt_trace
 * format (tmax = ORC_T_INIT) and gate[i] = 1 iff the scene-AABB gate passes. */
void orc_primary_rays(const float* params, int w, int h, float* rays, uint8_t* gate);

/* shadow rays of volumeRender.cl:1314,1407-1441 from primary rays + their closest hits.
 * valid[i]=0 (ray untouched) where hits[i].idx < 0. */
void orc_shadow_rays(const float* params, int64_t n, const float* rays, const void* hits, float* out_rays,
                     uint8_t* valid);

/* the whole raytracer_bvh kernel (volumeRender.cl:1043-1547): pixels as uint32 (b<<16|g<<8|r) */
void orc_render_frame(const orc_scene* s, const float* params, int w, int h, uint32_t* out, uint64_t* counters);

/* single-function probes used to pin the oracle against oracle/_ref */
float orc_ray_triangle(const float* o, const float* d, const float* v0, const float* e1, const float* e2, float* u,
                       float* v);
int orc_scene_box_gate(const float* bmin, const float* bmax, const float* org, const float* invdir, float* tmin,
                       float* tmax);
void orc_vec_probe(const float* a, const float* b, float* out7);
void orc_ray_box(const float* o, const float* d, const float* bmin, const float* bmax, float* tspan2);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
