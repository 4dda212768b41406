// Mesh.h -- indexed triangle soup + materials: the scene container the BVH builder and the
// device upload read from.
//
// Public surface and memory layout follow the reference (reference Mesh.h:20-99, Mesh.cpp:10-141):
//   vertices / normals are float4 with w = 1, `indices` and `normals_indices` hold 3 ints per
//   triangle (separate streams), `Material` is the 176-byte block the kernel indexes as
//   {int4 technique, 10 x float4}, `triangle_index_to_material_index[triId]` picks the material.
// These vectors are exactly the arrays handed to rt_upload_scene (include/rtb200.h).
//
// Additions over the reference (needed because synthetic scenes are filled with add_tri, which in
// the reference leaves normals / materials / scene AABB empty -- SURVEY.md B.3):
//   finish_synthetic(), reserve_tris(), add_indexed().
#pragma once
#include <vector>

#include "vecmath.h"

struct Effect;        // ColladaLoader.h
class ColladaLoader;  // ColladaLoader.h

struct Material {
    int4 technique;  // .x = 1 PHONG / 2 COOK_TORRANCE
    float4 emission, ambient, diffuse, specular;
    float4 shininess;  // .x
    float4 reflective;
    float4 reflectivity;  // .x
    float4 transparent;
    float4 transparency;  // .x
    float4 glossiness;    // .x

    static const int PHONG = 1, COOK_TORRANCE = 2;

    Material();  // reference default material (reference Mesh.h:35-39)
    explicit Material(const Effect& fx);
    void set(int tech, float4 emi, float4 amb, float4 diff, float4 spec, float shini, float4 refl, float refl_ty,
             float4 transp, float transp_cy, float gloss);
    void set(const Effect& fx);
};
static_assert(sizeof(Material) == 176, "kernel-side Material is 176 bytes (reference volumeRender.cl:340-357)");

class Mesh {
public:
    std::vector<int> indices;
    std::vector<float4> vertices;
    std::vector<int> normals_indices;
    std::vector<float4> normals;
    std::vector<Material> materials;
    std::vector<int> triangle_index_to_material_index;
    float3 scene_aabbox_min, scene_aabbox_max;

    Mesh() {}

    void init(ColladaLoader& loader);  // bake node transforms, offset indices per geometry
    void add_tri(const float4& v0, const float4& v1, const float4& v2);  // 3 fresh vertices per call

    int* getTrisPtr() { return indices.data(); }
    float4* getVertsPtr() { return vertices.data(); }
    int getNumTriangles() const { return (int)(indices.size() / 3); }

    // --- additions -------------------------------------------------------------------------
    void clear();
    void reserve_tris(size_t ntris, size_t nverts);
    int add_vertex(float x, float y, float z) { vertices.push_back(float4(x, y, z, 1.0f)); return (int)vertices.size() - 1; }
    void add_indexed(int i0, int i1, int i2) { indices.push_back(i0); indices.push_back(i1); indices.push_back(i2); }
    // Fill what the kernel dereferences unconditionally (reference volumeRender.cl:1310-1312,1374):
    // per-vertex area-weighted normals (normals_indices = indices), one default material unless
    // materials exist, tri->material = 0, and the scene AABB.
    void finish_synthetic();
    void compute_scene_aabb();
};
