// host_capi.cpp -- flat C entry points over the host scene layer (Mesh, SceneGen, ColladaLoader,
// FW::BVH2 / SplitBVHBuilder, BVH_Cuda, Camera) for the Python test / benchmark harness (ctypes).
// Pointers returned by rth_*_ptrs alias the handle's own std::vectors: valid until the handle is
// modified or freed. No GPU code here; the device boundary is include/rtb200.h.
#include <cstring>
#include <new>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "BVH_Cuda.h"
#include "Camera.h"
#include "ColladaLoader.h"
#include "Mesh.h"
#include "SceneGen.h"
#include "SplitBVHBuilder.h"

extern "C" {

// Threads used by the task-parallel SBVH builder (torchrun exports OMP_NUM_THREADS=1 to its workers).
void rth_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int rth_get_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void* rth_mesh_new() { return new (std::nothrow) Mesh(); }
void rth_mesh_free(void* m) { delete (Mesh*)m; }
void rth_mesh_clear(void* m) { ((Mesh*)m)->clear(); }

int rth_mesh_set(void* mh, const float* verts4, int V, const int* indices, int T) {
    Mesh* m = (Mesh*)mh;
    m->clear();
    m->vertices.resize(V);
    memcpy((void*)m->vertices.data(), verts4, sizeof(float4) * (size_t)V);
    m->indices.assign(indices, indices + (size_t)T * 3);
    return 0;
}
void rth_mesh_icosphere(void* m, int subdiv, float radius, float cx, float cy, float cz) {
    scenegen::add_icosphere(*(Mesh*)m, subdiv, radius, cx, cy, cz);
}
void rth_mesh_terrain(void* m, int nquads, float half) { scenegen::add_terrain(*(Mesh*)m, nquads, half); }
void rth_mesh_sphere_field(void* m, int grid, float pitch, int subdiv, float radius) {
    scenegen::add_sphere_field(*(Mesh*)m, grid, pitch, subdiv, radius);
}
void rth_mesh_sticks(void* m, int count, unsigned seed, float extent) { scenegen::add_sticks(*(Mesh*)m, count, seed, extent); }
void rth_mesh_bvh_test(void* m, int which) {
    if (which == 0) scenegen::add_bvh_test0(*(Mesh*)m);
    else scenegen::add_bvh_test1(*(Mesh*)m);
}
void rth_mesh_finish(void* m) { ((Mesh*)m)->finish_synthetic(); }
void rth_mesh_set_diffuse(void* mh, float r, float g, float b) {
    Mesh* m = (Mesh*)mh;
    if (m->materials.empty()) m->materials.push_back(Material());
    m->materials[0].diffuse = float4(r, g, b, 1.0f);
}
int rth_mesh_write_dae(void* m, const char* path) { return scenegen::write_dae(*(Mesh*)m, path) ? 0 : 1; }
int rth_mesh_load_dae(void* mh, const char* path) {
    Mesh* m = (Mesh*)mh;
    ColladaLoader loader;
    if (!loader.load(path)) return 1;
    m->clear();
    m->init(loader);
    return 0;
}

// counts: V, T, Vn, M ; ptrs: vertices, indices, normals, normals_indices, materials, tri->material
void rth_mesh_ptrs(void* mh, int counts[4], const void* ptrs[6], float aabb[6]) {
    Mesh* m = (Mesh*)mh;
    counts[0] = (int)m->vertices.size();
    counts[1] = m->getNumTriangles();
    counts[2] = (int)m->normals.size();
    counts[3] = (int)m->materials.size();
    ptrs[0] = m->vertices.data();
    ptrs[1] = m->indices.data();
    ptrs[2] = m->normals.data();
    ptrs[3] = m->normals_indices.data();
    ptrs[4] = m->materials.data();
    ptrs[5] = m->triangle_index_to_material_index.data();
    aabb[0] = m->scene_aabbox_min.x; aabb[1] = m->scene_aabbox_min.y; aabb[2] = m->scene_aabbox_min.z;
    aabb[3] = m->scene_aabbox_max.x; aabb[4] = m->scene_aabbox_max.y; aabb[5] = m->scene_aabbox_max.z;
}

struct BvhHandle {
    BVH_Cuda flat;
    int duplicates = 0;
    double seconds = 0.0;
};

// Mesh -> FW::BVH2 (SplitBVHBuilder) -> BVH_Cuda. parallel_threshold <= 0 builds on one thread.
void* rth_bvh_build(void* mh, int parallel_threshold) {
    Mesh* m = (Mesh*)mh;
    BvhHandle* h = new (std::nothrow) BvhHandle();
    if (!h) return nullptr;
    FW::BVH2 bvh2;  // built through the same two calls as the reference's initRayTrace
    if (parallel_threshold != 16384) {
        // custom threshold: drive the builder directly
        bvh2.m_scene = m;
        FW::SplitBVHBuilder b(bvh2);
        b.setParallelThreshold(parallel_threshold);
        bvh2.m_root = b.run();
        h->duplicates = b.numDuplicates();
    } else {
        bvh2.setMesh(m);
        h->duplicates = bvh2.numDuplicates();
        h->seconds = bvh2.buildSeconds();
    }
    h->flat.build_from_bvh2(bvh2);
    return h;
}
void* rth_bvh_load(const char* path) {
    BvhHandle* h = new (std::nothrow) BvhHandle();
    if (h && !h->flat.load(path)) {
        delete h;
        h = nullptr;
    }
    return h;
}
int rth_bvh_save(void* h, const char* path) { return ((BvhHandle*)h)->flat.save(path) ? 0 : 1; }
void rth_bvh_free(void* h) { delete (BvhHandle*)h; }
// counts: N, R, duplicates ; ptrs: nodes (48 B each), tri_indices
void rth_bvh_ptrs(void* hh, int counts[3], const void* ptrs[2], double* seconds) {
    BvhHandle* h = (BvhHandle*)hh;
    counts[0] = (int)h->flat.bvh_nodes.size();
    counts[1] = (int)h->flat.tri_indices.size();
    counts[2] = h->duplicates;
    ptrs[0] = h->flat.bvh_nodes.data();
    ptrs[1] = h->flat.tri_indices.data();
    if (seconds) *seconds = h->seconds;
}

// Camera(): default pose, then add_radius(d_radius) and add_rotate(d_alpha, d_beta); fills Params.
void rth_camera_params(float d_radius, float d_alpha, float d_beta, int w, int h, const float light_pos[3],
                       const float light_color[3], const float aabb_min[3], const float aabb_max[3], float out_params[32],
                       float out_eye[3]) {
    Camera cam;
    if (d_radius != 0.0f) cam.add_radius(d_radius);
    if (d_alpha != 0.0f || d_beta != 0.0f) cam.add_rotate(d_alpha, d_beta);
    cam.make_params(w, h, float3(light_pos[0], light_pos[1], light_pos[2]),
                    float3(light_color[0], light_color[1], light_color[2]), float3(aabb_min[0], aabb_min[1], aabb_min[2]),
                    float3(aabb_max[0], aabb_max[1], aabb_max[2]), out_params);
    if (out_eye) {
        out_eye[0] = cam.eye.x; out_eye[1] = cam.eye.y; out_eye[2] = cam.eye.z;
    }
}

}  // extern "C"
