/* oracle/oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement, in plain C, of the reference's only live GPU code:
 *   /root/reference/x64/Release/volumeRender.cl  (abbreviated vR.cl below)
 * Each function cites the lines it follows. Arithmetic is fp32, uncontracted (build with
 * -ffp-contract=off), evaluated left to right exactly as the source is written. OpenCL built-ins
 * are DEFINED as follows (OpenCL leaves their rounding implementation-defined):
 *   dot/cross/normalize  -> the reference host's own definitions, vectors_math.cpp:73-84
 *   fmin/fmax            -> C fminf/fmaxf (NaN-ignoring, like OpenCL)
 *   clamp(x,a,b)         -> fminf(fmaxf(x,a),b)
 *   sqrt/pow             -> sqrtf/powf
 *   convert_uint         -> truncation
 *
 * See oracle.h for who may use this file and for the parity-pinning status.
 */
#include "oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } f3;

static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul3(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 scale3(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline f3 div3s(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
/* vectors_math.cpp:73-75 */
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* vectors_math.cpp:77-79 */
static inline f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* vectors_math.cpp:18-20,81-84: invLen = 1.0f / sqrtf(dot(v,v)); invLen * v */
static inline f3 normalize3(f3 v) {
    float invLen = 1.0f / sqrtf(dot3(v, v));
    return mk3(invLen * v.x, invLen * v.y, invLen * v.z);
}
static inline float clampf(float x, float a, float b) { return fminf(fmaxf(x, a), b); }

#define STACK_SIZE 65 /* vR.cl:636 */
#define TMIN 0.001f   /* vR.cl:640 */
#define RAY_TRACE_DEPTH 3 /* vR.cl:12 */
#define M_PI_F 3.14159274101257f

typedef struct { f3 ori, dir, inv_dir; } Ray; /* vR.cl:198-203 */

/* vR.cl:205-211 RayInit */
static inline void RayInit(Ray* r, f3 o, f3 d) {
    r->ori = o;
    r->dir = normalize3(d);
    /* 1.0/x is a double expression in the source; for fp32 operands the double quotient rounded
     * to fp32 equals the fp32 quotient (53 >= 2*24+2), so plain fp32 division is identical. */
    r->inv_dir = mk3(1.0f / r->dir.x, 1.0f / r->dir.y, 1.0f / r->dir.z);
}

/* vR.cl:236-254 RayBoxIntersection (scene gate, uses inv_dir multiply) */
static inline int RayBoxIntersection(f3 BBMin, f3 BBMax, f3 RayOrg, f3 RayDirInv, float* tmin, float* tmax) {
    float l1 = (BBMin.x - RayOrg.x) * RayDirInv.x;
    float l2 = (BBMax.x - RayOrg.x) * RayDirInv.x;
    *tmin = fminf(l1, l2);
    *tmax = fmaxf(l1, l2);
    l1 = (BBMin.y - RayOrg.y) * RayDirInv.y;
    l2 = (BBMax.y - RayOrg.y) * RayDirInv.y;
    *tmin = fmaxf(fminf(l1, l2), *tmin);
    *tmax = fminf(fmaxf(l1, l2), *tmax);
    l1 = (BBMin.z - RayOrg.z) * RayDirInv.z;
    l2 = (BBMax.z - RayOrg.z) * RayDirInv.z;
    *tmin = fmaxf(fminf(l1, l2), *tmin);
    *tmax = fminf(fmaxf(l1, l2), *tmax);
    return ((*tmax >= *tmin) && (*tmax >= 0.0f));
}

/* vR.cl:257-282 RayTriangleIntersection (Moller-Trumbore, two-sided, no epsilon). u,v exported
 * for the barycentric parity criterion; they are the values of the LAST evaluation. */
static inline float RayTriangleIntersection(const Ray* r, f3 v0, f3 edge1, f3 edge2, float* uo, float* vo) {
    f3 tvec = sub3(r->ori, v0);
    f3 pvec = cross3(r->dir, edge2);
    float det = dot3(edge1, pvec);
    det = 1.0f / det;
    float u = dot3(tvec, pvec) * det;
    if (u < 0.0f || u > 1.0f) return -1.0f;
    f3 qvec = cross3(tvec, edge1);
    float v = dot3(r->dir, qvec) * det;
    if (v < 0.0f || (u + v) > 1.0f) return -1.0f;
    *uo = u;
    *vo = v;
    return dot3(edge2, qvec) * det;
}

/* vR.cl:612-624 ray_box: TRUE division by dir (not inv_dir), NaN-ignoring fmin/fmax */
static inline void ray_box(const Ray* r, const float* mn, const float* mx, float* tmin1, float* tmax1) {
    f3 t0 = mk3((mn[0] - r->ori.x) / r->dir.x, (mn[1] - r->ori.y) / r->dir.y, (mn[2] - r->ori.z) / r->dir.z);
    f3 t1 = mk3((mx[0] - r->ori.x) / r->dir.x, (mx[1] - r->ori.y) / r->dir.y, (mx[2] - r->ori.z) / r->dir.z);
    f3 tmin = mk3(fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z));
    f3 tmax = mk3(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z));
    *tmin1 = fmaxf(fmaxf(tmin.x, tmin.y), tmin.z);
    *tmax1 = fminf(fminf(tmax.x, tmax.y), tmax.z);
}

typedef struct { uint64_t inner, leaf, tris, maxstack; } Counters;

static inline f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

/* vR.cl:658-1010 traverse_bvh, live lines 776-793, 824-867, 901-931, 965-1006.
 * Returns 3*triId of the accepted triangle (closest, or first accepted when !needClosestHit), or -1. */
static int traverse_bvh(const orc_scene* s, const Ray* ray, float* tHit, int needClosestHit, float* uo, float* vo,
                        Counters* c) {
    const float* bvh_nodes = s->nodes; /* float4 bvh_nodes[3*N]; node i = words [12i .. 12i+11] */
    const int num_bvh_nodes = s->num_bvh_nodes;
    int traversalStack[STACK_SIZE];
    int stack_count = 1;
    traversalStack[0] = 0;
    int tri_index = -1;

    while (stack_count > 0) {
        if ((uint64_t)stack_count > c->maxstack) c->maxstack = (uint64_t)stack_count;
        int nodeIndex = traversalStack[stack_count - 1];
        const int* row2 = (const int*)(bvh_nodes + (size_t)nodeIndex * 12 + 8); /* as_int(bvh_nodes[n*3+2].xyzw) */
        int offset_left = row2[0];
        if (offset_left >= 0) { /* inner node, vR.cl:837 */
            int offset_right = row2[1];
            if (offset_right < 0) return -1;              /* vR.cl:841 */
            if (offset_left >= num_bvh_nodes) return -1;  /* vR.cl:855 */
            if (offset_right >= num_bvh_nodes) return -1; /* vR.cl:856 */
            c->inner++;
            float t0x, t0y, t1x, t1y;
            ray_box(ray, bvh_nodes + (size_t)offset_left * 12, bvh_nodes + (size_t)offset_left * 12 + 4, &t0x, &t0y);
            ray_box(ray, bvh_nodes + (size_t)offset_right * 12, bvh_nodes + (size_t)offset_right * 12 + 4, &t1x, &t1y);
            int intersect0 = (t0x <= t0y) && (t0y >= TMIN) && (t0x <= *tHit); /* vR.cl:866 */
            int intersect1 = (t1x <= t1y) && (t1y >= TMIN) && (t1x <= *tHit); /* vR.cl:867 */
            if (intersect0 && intersect1) {
                if (t0x > t1x) { /* vR.cl:903: ties keep LEFT first */
                    int t = offset_left; offset_left = offset_right; offset_right = t;
                }
                traversalStack[stack_count - 1] = offset_right;
                if (stack_count >= STACK_SIZE) return -1; /* vR.cl:914 */
                traversalStack[stack_count] = offset_left;
                ++stack_count;
            } else if (intersect0) {
                traversalStack[stack_count - 1] = offset_left;
            } else if (intersect1) {
                traversalStack[stack_count - 1] = offset_right;
            } else {
                --stack_count;
            }
        } else { /* leaf, vR.cl:965-1001 */
            int node_offset_tris = row2[2];
            int node_num_tris = row2[3];
            c->leaf++;
            for (int i = 0; i < node_num_tris; ++i) {
                int tri1 = s->tri_indices[node_offset_tris + i];
                const float* p0 = s->verts + (size_t)s->indices[tri1 + 0] * 4;
                const float* p1 = s->verts + (size_t)s->indices[tri1 + 1] * 4;
                const float* p2 = s->verts + (size_t)s->indices[tri1 + 2] * 4;
                f3 v0 = ld3(p0);
                f3 e1 = sub3(ld3(p1), v0);
                f3 e2 = sub3(ld3(p2), v0);
                float u = 0.f, v = 0.f;
                c->tris++;
                float t = RayTriangleIntersection(ray, v0, e1, e2, &u, &v);
                if (t < *tHit && t > TMIN) { /* vR.cl:978, strict on both sides */
                    *tHit = t;
                    *uo = u;
                    *vo = v;
                    if (!needClosestHit) return tri1; /* vR.cl:986 */
                    tri_index = tri1;
                }
            }
            --stack_count;
        }
    }
    return tri_index;
}

/* vR.cl:690-712: the author's brute-force loop over all triangles in index order */
static int brute_force(const orc_scene* s, const Ray* ray, float* tHit, int needClosestHit, float* uo, float* vo) {
    int hit_index = -1;
    for (int i = 0; i < s->num_tris; i++) {
        f3 v0 = ld3(s->verts + (size_t)s->indices[i * 3 + 0] * 4);
        f3 e1 = sub3(ld3(s->verts + (size_t)s->indices[i * 3 + 1] * 4), v0);
        f3 e2 = sub3(ld3(s->verts + (size_t)s->indices[i * 3 + 2] * 4), v0);
        float u = 0.f, v = 0.f;
        float t = RayTriangleIntersection(ray, v0, e1, e2, &u, &v);
        if (t < *tHit && t > TMIN) {
            *tHit = t;
            *uo = u;
            *vo = v;
            hit_index = i * 3;
            if (!needClosestHit) return hit_index;
        }
    }
    return hit_index;
}

typedef struct { int idx; float t, u, v; } Hit;

static void trace_impl(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits_v, uint64_t* counters,
                       int brute) {
    Hit* hits = (Hit*)hits_v;
    uint64_t ci = 0, cl = 0, ct = 0, cm = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ci, cl, ct) reduction(max : cm)
    for (int64_t i = 0; i < n; i++) {
        const float* r8 = rays + i * 8;
        Ray r;
        r.ori = mk3(r8[0], r8[1], r8[2]);
        r.dir = mk3(r8[4], r8[5], r8[6]);
        r.inv_dir = mk3(1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z);
        float tHit = r8[3];
        float u = 0.f, v = 0.f;
        Counters c = {0, 0, 0, 0};
        int idx = brute ? brute_force(s, &r, &tHit, mode == 0, &u, &v)
                        : traverse_bvh(s, &r, &tHit, mode == 0, &u, &v, &c);
        hits[i].idx = idx;
        hits[i].t = tHit;
        /* u,v are only meaningful together with an accepted hit; an early `return -1` after an
         * accepted hit (vR.cl:841,855,856,914) leaves tHit updated but idx = -1: report u=v=0. */
        hits[i].u = idx >= 0 ? u : 0.f;
        hits[i].v = idx >= 0 ? v : 0.f;
        ci += c.inner; cl += c.leaf; ct += c.tris;
        if (c.maxstack > cm) cm = c.maxstack;
    }
    if (counters) { counters[0] = ci; counters[1] = cl; counters[2] = ct; counters[3] = cm; }
}

/* per-ray operation trace for scheduler studies (tools/simsched.c): seg[i*max_seg + 2k] = inner visits before the
 * k-th leaf, seg[i*max_seg + 2k+1] = triangle tests in that leaf; nseg[i] = entries used (a trailing run of inner
 * visits with no leaf after it is stored with tri count 0). closest-hit only. */
void orc_trace_segments(const orc_scene* s, int64_t n, const float* rays, uint16_t* seg, int max_seg, int* nseg) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        const float* r8 = rays + i * 8;
        Ray ray;
        ray.ori = mk3(r8[0], r8[1], r8[2]);
        ray.dir = mk3(r8[4], r8[5], r8[6]);
        float tHit = r8[3];
        uint16_t* out = seg + (size_t)i * max_seg;
        int ns = 0, run = 0;
        int stack[STACK_SIZE], count = 1;
        stack[0] = 0;
        while (count > 0) {
            int node = stack[count - 1];
            const int* row2 = (const int*)(s->nodes + (size_t)node * 12 + 8);
            int l = row2[0];
            if (l >= 0) {
                int r = row2[1];
                run++;
                float a0, a1, b0, b1;
                ray_box(&ray, s->nodes + (size_t)l * 12, s->nodes + (size_t)l * 12 + 4, &a0, &a1);
                ray_box(&ray, s->nodes + (size_t)r * 12, s->nodes + (size_t)r * 12 + 4, &b0, &b1);
                int h0 = (a0 <= a1) && (a1 >= TMIN) && (a0 <= tHit), h1 = (b0 <= b1) && (b1 >= TMIN) && (b0 <= tHit);
                if (h0 && h1) {
                    if (a0 > b0) { int t = l; l = r; r = t; }
                    stack[count - 1] = r;
                    stack[count++] = l;
                } else if (h0) stack[count - 1] = l;
                else if (h1) stack[count - 1] = r;
                else --count;
            } else {
                int off = row2[2], cnt = row2[3];
                for (int k = 0; k < cnt; k++) {
                    int tri1 = s->tri_indices[off + k];
                    f3 v0 = ld3(s->verts + (size_t)s->indices[tri1] * 4);
                    f3 e1 = sub3(ld3(s->verts + (size_t)s->indices[tri1 + 1] * 4), v0);
                    f3 e2 = sub3(ld3(s->verts + (size_t)s->indices[tri1 + 2] * 4), v0);
                    float u, v;
                    float t = RayTriangleIntersection(&ray, v0, e1, e2, &u, &v);
                    if (t < tHit && t > TMIN) tHit = t;
                }
                if (ns + 2 <= max_seg) { out[ns++] = (uint16_t)run; out[ns++] = (uint16_t)cnt; }
                run = 0;
                --count;
            }
        }
        if (run && ns + 2 <= max_seg) { out[ns++] = (uint16_t)run; out[ns++] = 0; }
        nseg[i] = ns;
    }
}

/* per-ray visit counts (inner, leaf, tris) for workload analysis: stats[3*i + {0,1,2}] */
void orc_trace_stats(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits_v, uint32_t* stats) {
    Hit* hits = (Hit*)hits_v;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        const float* r8 = rays + i * 8;
        Ray r;
        r.ori = mk3(r8[0], r8[1], r8[2]);
        r.dir = mk3(r8[4], r8[5], r8[6]);
        r.inv_dir = mk3(0, 0, 0);
        float tHit = r8[3], u = 0.f, v = 0.f;
        Counters c = {0, 0, 0, 0};
        int idx = traverse_bvh(s, &r, &tHit, mode == 0, &u, &v, &c);
        hits[i].idx = idx; hits[i].t = tHit; hits[i].u = idx >= 0 ? u : 0.f; hits[i].v = idx >= 0 ? v : 0.f;
        stats[3 * i] = (uint32_t)c.inner; stats[3 * i + 1] = (uint32_t)c.leaf; stats[3 * i + 2] = (uint32_t)c.tris;
    }
}

void orc_trace(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits, uint64_t* counters) {
    trace_impl(s, mode, n, rays, hits, counters, 0);
}
void orc_trace_bruteforce(const orc_scene* s, int mode, int64_t n, const float* rays, void* hits) {
    trace_impl(s, mode, n, rays, hits, NULL, 1);
}

/* vR.cl:1156-1196: one primary ray. Returns the scene-AABB gate. */
static inline int primary_ray(const float* params, unsigned x, unsigned y, unsigned w, unsigned h, Ray* r) {
    f3 a = ld3(params + 0), b = ld3(params + 4), c = ld3(params + 8), campos = ld3(params + 12);
    f3 aabb_min = ld3(params + 24), aabb_max = ld3(params + 28);
    /* (x-0.5)/((float)w): x-0.5 is exact in fp32 for x < 2^23 and the double quotient rounds to the
     * same fp32 value as an fp32 division (see RayInit) -> evaluate in fp32. */
    float xf = ((float)x - 0.5f) / ((float)w);
    float yf = ((float)y - 0.5f) / ((float)h);
    f3 t1 = add3(c, scale3(a, xf));
    f3 t2 = scale3(b, yf);
    f3 image_pos = add3(t1, t2);
    RayInit(r, image_pos, sub3(image_pos, campos));
    float t_min, t_max;
    return RayBoxIntersection(aabb_min, aabb_max, r->ori, r->inv_dir, &t_min, &t_max);
}

void orc_primary_rays(const float* params, int w, int h, float* rays, uint8_t* gate) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            Ray r;
            int g = primary_ray(params, (unsigned)x, (unsigned)y, (unsigned)w, (unsigned)h, &r);
            float* o = rays + ((size_t)y * w + x) * 8;
            o[0] = r.ori.x; o[1] = r.ori.y; o[2] = r.ori.z; o[3] = ORC_T_INIT;
            o[4] = r.dir.x; o[5] = r.dir.y; o[6] = r.dir.z; o[7] = 0.f;
            if (gate) gate[(size_t)y * w + x] = (uint8_t)g;
        }
}

/* vR.cl:1314 + 1407-1441: hitpoint = o + d*(t-0.001); L = normalize(light-hitpoint);
 * shadow ray = RayInit(hitpoint + L*0.001, L) (RayInit normalises L again). */
static inline void shadow_ray(f3 light_pos, const Ray* r, float t, Ray* out) {
    f3 vNew = add3(r->ori, scale3(r->dir, (t - 0.001f)));
    f3 L = normalize3(sub3(light_pos, vNew));
    RayInit(out, add3(vNew, scale3(L, 0.001f)), L);
}

void orc_shadow_rays(const float* params, int64_t n, const float* rays, const void* hits_v, float* out_rays,
                     uint8_t* valid) {
    const Hit* hits = (const Hit*)hits_v;
    f3 light_pos = ld3(params + 16);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float* o = out_rays + i * 8;
        if (hits[i].idx < 0) {
            if (valid) valid[i] = 0;
            memset(o, 0, 32);
            continue;
        }
        Ray r, sr;
        r.ori = ld3(rays + i * 8);
        r.dir = ld3(rays + i * 8 + 4);
        shadow_ray(light_pos, &r, hits[i].t, &sr);
        o[0] = sr.ori.x; o[1] = sr.ori.y; o[2] = sr.ori.z; o[3] = ORC_T_INIT;
        o[4] = sr.dir.x; o[5] = sr.dir.y; o[6] = sr.dir.z; o[7] = 0.f;
        if (valid) valid[i] = 1;
    }
}

/* vR.cl:27-53 get_normal_at_tri_point: barycentric weights by Cramer's rule on POSITIONS */
static inline f3 get_normal_at_tri_point(f3 pNew, f3 p0, f3 p1, f3 p2, f3 vn0, f3 vn1, f3 vn2) {
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

/* vR.cl:1732-1738 */
static inline float GGX_PartialGeometry(float cosThetaN, float alpha) {
    float cosTheta_sqr = clampf(cosThetaN * cosThetaN, 0.0f, 1.0f);
    float tan2 = (1 - cosTheta_sqr) / cosTheta_sqr;
    float GP = 2 / (1 + sqrtf(1 + alpha * alpha * tan2));
    return GP;
}
/* vR.cl:1740-1746 */
static inline float GGX_Distribution(float cosThetaNH, float alpha) {
    float alpha2 = alpha * alpha;
    float NH_sqr = clampf(cosThetaNH * cosThetaNH, 0.0f, 1.0f);
    float den = NH_sqr * alpha2 + (1.0f - NH_sqr);
    return alpha2 / (M_PI_F * den * den);
}
/* vR.cl:1748-1751 */
static inline f3 FresnelSchlick(f3 F0, float cosTheta) {
    float p = powf(1.0f - clampf(cosTheta, 0.0f, 1.0f), 5.0f);
    return mk3(F0.x + (1.0f - F0.x) * p, F0.y + (1.0f - F0.y) * p, F0.z + (1.0f - F0.z) * p);
}
static inline float max0(float v) { return (0.0f < v) ? v : 0.0f; } /* OpenCL max(0.0, v) */
/* vR.cl:1754-1779 CookTorrance_GGX */
static inline f3 CookTorrance_GGX(f3 n, f3 l, f3 v, f3 albedo, f3 f0, float roughness) {
    n = normalize3(n);
    v = normalize3(v);
    l = normalize3(l);
    f3 h = normalize3(add3(v, l));
    float NL = dot3(n, l);
    if (NL <= 0.0f) return mk3(0.f, 0.f, 0.f);
    float NV = dot3(n, v);
    if (NV <= 0.0f) return mk3(0.f, 0.f, 0.f);
    float NH = dot3(n, h);
    float HV = dot3(h, v);
    float roug_sqr = roughness * roughness;
    float G = GGX_PartialGeometry(NV, roug_sqr) * GGX_PartialGeometry(NL, roug_sqr);
    float D = GGX_Distribution(NH, roug_sqr);
    f3 F = FresnelSchlick(f0, HV);
    float GD = G * D;
    float den = NV + 0.001f;
    f3 specK = mk3(GD * F.x * 0.25f / den, GD * F.y * 0.25f / den, GD * F.z * 0.25f / den);
    f3 diffK = mk3(clampf(1.0f - F.x, 0.f, 1.f), clampf(1.0f - F.y, 0.f, 1.f), clampf(1.0f - F.z, 0.f, 1.f));
    return mk3(max0(albedo.x * diffK.x * NL / M_PI_F + specK.x), max0(albedo.y * diffK.y * NL / M_PI_F + specK.y),
               max0(albedo.z * diffK.z * NL / M_PI_F + specK.z));
}

/* vR.cl:186-195 rgbToInt */
static inline uint32_t rgbToInt(float r, float g, float b) {
    r = clampf(r, 0.0f, 255.0f);
    g = clampf(g, 0.0f, 255.0f);
    b = clampf(b, 0.0f, 255.0f);
    return ((uint32_t)b << 16) | ((uint32_t)g << 8) | ((uint32_t)r);
}

/* vR.cl:1043-1547 raytracer_bvh for one pixel */
static uint32_t render_pixel(const orc_scene* s, const float* params, unsigned x, unsigned y, unsigned w, unsigned h,
                             Counters* c) {
    const f3 light_pos = ld3(params + 16);
    Ray r;
    int continue_path = primary_ray(params, x, y, w, h, &r);
    float hit_t = ORC_T_INIT; /* HitRecordInit vR.cl:223-227 */
    int hit_index = -1;
    f3 color = mk3(0.f, 0.f, 0.f);
    int ray_depth = 0;
    float shadow_coef_2 = 1.0f;
    float shadow_coef_2_resault = 0.0f;
    float uu, vv;

    while (continue_path && ray_depth < RAY_TRACE_DEPTH) { /* vR.cl:1236 */
        hit_index = traverse_bvh(s, &r, &hit_t, 1, &uu, &vv, c);
        shadow_coef_2 = 1.0f;
        if (hit_index >= 0) {
            ray_depth++;
            const f3 v0 = ld3(s->verts + (size_t)s->indices[hit_index + 0] * 4); /* vR.cl:1306-1312 */
            const f3 v1 = ld3(s->verts + (size_t)s->indices[hit_index + 1] * 4);
            const f3 v2 = ld3(s->verts + (size_t)s->indices[hit_index + 2] * 4);
            const f3 vn0 = ld3(s->normals + (size_t)s->normal_indices[hit_index + 0] * 4);
            const f3 vn1 = ld3(s->normals + (size_t)s->normal_indices[hit_index + 1] * 4);
            const f3 vn2 = ld3(s->normals + (size_t)s->normal_indices[hit_index + 2] * 4);
            const f3 vNew = add3(r.ori, scale3(r.dir, (hit_t - 0.001f))); /* vR.cl:1314 */
            f3 normal = get_normal_at_tri_point(vNew, v0, v1, v2, vn0, vn1, vn2);
            normal = normalize3(normal);
            f3 l1 = normalize3(sub3(light_pos, vNew)); /* vR.cl:1357-1361 */
            f3 v = normalize3(sub3(r.ori, vNew));
            f3 n = normalize3(normal);
            const float* mat = s->materials + (size_t)s->tri_to_material[hit_index / 3] * 44; /* vR.cl:1374 */
            f3 diffuse = ld3(mat + 12); /* int4 technique, emission, ambient, DIFFUSE */
            f3 f0 = scale3(mk3(40.f, 40.f, 40.f), (1 / 255.0f));
            f3 rez_color = scale3(CookTorrance_GGX(n, l1, v, diffuse, f0, 0.5f), 3.0f); /* vR.cl:1389 */
            f3 ambient_color = mul3(mk3(0.3f, 0.3f, 0.3f), diffuse);                     /* vR.cl:1402 */
            rez_color = add3(rez_color, ambient_color);
            f3 hitpoint = vNew;
            f3 L = normalize3(sub3(light_pos, hitpoint)); /* vR.cl:1407-1409 */
            {
                Ray ray_shadow;
                RayInit(&ray_shadow, add3(hitpoint, scale3(L, 0.001f)), L); /* vR.cl:1441 */
                hit_t = ORC_T_INIT;
                hit_index = -1;
                hit_index = traverse_bvh(s, &ray_shadow, &hit_t, 0, &uu, &vv, c);
                if (hit_index >= 0 && hit_t > 0.025f) shadow_coef_2 = 0.25f; /* vR.cl:1444 */
            }
            color = add3(color, rez_color);
            shadow_coef_2_resault += shadow_coef_2;
            { /* vR.cl:1493-1497: reflect(i,n) = i - 2.0f * n * dot(n,i) */
                float d = dot3(normal, r.dir);
                f3 refl = sub3(r.dir, scale3(scale3(normal, 2.0f), d));
                RayInit(&r, add3(hitpoint, scale3(refl, 0.001f)), refl);
                hit_t = ORC_T_INIT;
                hit_index = -1;
            }
        } else {
            continue_path = 0;
        }
    }
    if (ray_depth >= 1) { /* vR.cl:1519-1538 */
        color = div3s(color, (float)ray_depth);
        shadow_coef_2_resault /= (float)ray_depth;
        color = scale3(color, shadow_coef_2_resault);
    } else {
        color = mk3(0.f, 0.f, 0.f);
    }
    return rgbToInt(color.x * 255, color.y * 255, color.z * 255);
}

void orc_render_frame(const orc_scene* s, const float* params, int w, int h, uint32_t* out, uint64_t* counters) {
    uint64_t ci = 0, cl = 0, ct = 0, cm = 0;
    const int tx = (w + 7) / 8, ty = (h + 7) / 8; /* 8x8 work-groups, RayTracer.cpp:332-336 */
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : ci, cl, ct) reduction(max : cm)
    for (int tile = 0; tile < tx * ty; tile++) {
        Counters c = {0, 0, 0, 0};
        int bx = (tile % tx) * 8, by = (tile / tx) * 8;
        for (int y = by; y < by + 8 && y < h; y++)
            for (int x = bx; x < bx + 8 && x < w; x++)
                out[(size_t)y * w + x] = render_pixel(s, params, (unsigned)x, (unsigned)y, (unsigned)w, (unsigned)h, &c);
        ci += c.inner; cl += c.leaf; ct += c.tris;
        if (c.maxstack > cm) cm = c.maxstack;
    }
    if (counters) { counters[0] = ci; counters[1] = cl; counters[2] = ct; counters[3] = cm; }
}

float orc_ray_triangle(const float* o, const float* d, const float* v0, const float* e1, const float* e2, float* u,
                       float* v) {
    Ray r;
    r.ori = ld3(o);
    r.dir = ld3(d);
    r.inv_dir = mk3(0, 0, 0);
    float uu = 0.f, vv = 0.f;
    float t = RayTriangleIntersection(&r, ld3(v0), ld3(e1), ld3(e2), &uu, &vv);
    if (u) *u = uu;
    if (v) *v = vv;
    return t;
}

int orc_scene_box_gate(const float* bmin, const float* bmax, const float* org, const float* invdir, float* tmin,
                       float* tmax) {
    return RayBoxIntersection(ld3(bmin), ld3(bmax), ld3(org), ld3(invdir), tmin, tmax);
}

void orc_vec_probe(const float* a, const float* b, float* out7) {
    f3 A = ld3(a), B = ld3(b);
    f3 c = cross3(A, B), n = normalize3(A);
    out7[0] = dot3(A, B);
    out7[1] = c.x; out7[2] = c.y; out7[3] = c.z;
    out7[4] = n.x; out7[5] = n.y; out7[6] = n.z;
}

void orc_ray_box(const float* o, const float* d, const float* bmin, const float* bmax, float* tspan2) {
    Ray r;
    r.ori = ld3(o);
    r.dir = ld3(d);
    r.inv_dir = mk3(0, 0, 0);
    ray_box(&r, bmin, bmax, &tspan2[0], &tspan2[1]);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
