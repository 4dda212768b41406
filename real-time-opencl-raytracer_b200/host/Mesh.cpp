#include "Mesh.h"

#include "ColladaLoader.h"

Material::Material() {
    set(PHONG, float4(0, 0, 0, 1), float4(0, 0, 0, 1), float4(1, 1, 1, 1), float4(1, 1, 1, 1), 2.0f, float4(1, 1, 1, 1),
        1.0f, float4(1, 1, 1, 1), 0.0f, 1.0f);
}

Material::Material(const Effect& fx) { set(fx); }

void Material::set(int tech, float4 emi, float4 amb, float4 diff, float4 spec, float shini, float4 refl, float refl_ty,
                   float4 transp, float transp_cy, float gloss) {
    technique.x = tech;
    emission = emi;
    ambient = amb;
    diffuse = diff;
    specular = spec;
    shininess.x = shini;
    reflective = refl;
    reflectivity.x = refl_ty;
    transparent = transp;
    transparency.x = transp_cy;
    glossiness.x = gloss;
}

void Material::set(const Effect& fx) {
    set(fx.technique, fx.emission, fx.ambient, fx.diffuse, fx.specular, fx.shininess, fx.reflective, fx.reflectivity,
        fx.transparent, fx.transparency, fx.glossiness);
}

// One Material per <effect>; every geometry's vertex / normal indices are rebased by the number of
// vertices / normals emitted so far; positions and normals go through the geometry's scene-node
// matrix; the scene AABB is accumulated over transformed positions (reference Mesh.cpp:10-78).
void Mesh::init(ColladaLoader& loader) {
    for (size_t e = 0; e < loader.library_effects.size(); ++e) materials.push_back(Material(loader.library_effects[e]));

    int vbase = 0, nbase = 0;
    for (size_t g = 0; g < loader.library_geometries.size(); ++g) {
        Geometry& geo = loader.library_geometries[g];
        for (const PolygonTriangle& p : geo.polygons) {
            for (int k = 0; k < 3; ++k) indices.push_back(vbase + p.vertex_indices.m[k]);
            for (int k = 0; k < 3; ++k) normals_indices.push_back(nbase + p.normal_indices.m[k]);
            triangle_index_to_material_index.push_back(p.effect_index);
        }
        for (float3& pos : geo.float_array_positions) {
            float3 v = loader.get_vertex(pos, (int)g);
            if (vbase < 1) {
                scene_aabbox_min = v;
                scene_aabbox_max = v;
            } else {
                scene_aabbox_min = fminf1(scene_aabbox_min, v);
                scene_aabbox_max = fmaxf1(scene_aabbox_max, v);
            }
            vertices.push_back(float4(v.x, v.y, v.z, 1.0f));
            ++vbase;
        }
        for (float3& nrm : geo.float_array_normals) {
            float3 n = loader.get_normal(nrm, (int)g);
            normals.push_back(float4(n.x, n.y, n.z, 1.0f));
            ++nbase;
        }
    }
}

void Mesh::add_tri(const float4& v0, const float4& v1, const float4& v2) {
    int base = (int)vertices.size();
    add_indexed(base, base + 1, base + 2);
    vertices.push_back(v0);
    vertices.push_back(v1);
    vertices.push_back(v2);
}

void Mesh::clear() {
    indices.clear();
    vertices.clear();
    normals_indices.clear();
    normals.clear();
    materials.clear();
    triangle_index_to_material_index.clear();
    scene_aabbox_min = scene_aabbox_max = float3();
}

void Mesh::reserve_tris(size_t ntris, size_t nverts) {
    indices.reserve(ntris * 3);
    vertices.reserve(nverts);
}

void Mesh::compute_scene_aabb() {
    for (size_t i = 0; i < vertices.size(); ++i) {
        float3 v = make_float3(vertices[i]);
        if (i == 0) {
            scene_aabbox_min = scene_aabbox_max = v;
        } else {
            scene_aabbox_min = fminf1(scene_aabbox_min, v);
            scene_aabbox_max = fmaxf1(scene_aabbox_max, v);
        }
    }
}

void Mesh::finish_synthetic() {
    const size_t T = indices.size() / 3;
    std::vector<float3> acc(vertices.size());
    for (size_t t = 0; t < T; ++t) {
        const int i0 = indices[3 * t], i1 = indices[3 * t + 1], i2 = indices[3 * t + 2];
        float3 p0 = make_float3(vertices[i0]);
        float3 fn = cross(make_float3(vertices[i1]) - p0, make_float3(vertices[i2]) - p0);  // area-weighted
        acc[i0] += fn;
        acc[i1] += fn;
        acc[i2] += fn;
    }
    normals.resize(vertices.size());
    for (size_t i = 0; i < vertices.size(); ++i) {
        float3 n = acc[i];
        float l2 = dot(n, n);
        n = l2 > 0.0f ? n * (1.0f / sqrtf(l2)) : float3(0.0f, 1.0f, 0.0f);
        normals[i] = float4(n.x, n.y, n.z, 1.0f);
    }
    normals_indices = indices;
    if (materials.empty()) materials.push_back(Material());
    triangle_index_to_material_index.assign(T, 0);
    compute_scene_aabb();
}
