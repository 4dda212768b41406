// ColladaLoader.h -- COLLADA 1.4 <polygons> scene loader (the demo's only scene input path).
//
// Public surface follows the reference (reference ColladaLoader.h:59-211): Effect, PolygonTriangle,
// Geometry, SceneNode, and ColladaLoader::{load, get_vertex, get_normal, library_effects,
// library_geometries, library_visual_scenes}; Mesh::init(ColladaLoader&) consumes it.
// Accepted dialect and quirks are the reference's (reference
// ColladaLoader.cpp:13-593), including the rotate-sid order jointOrientX/Y/Z, rotateX, rotateZ,
// rotateY applied with axis = position % 3.
//
// Differences (inputs the reference handles give identical Mesh output):
//   * XML is read by the built-in xml_lite reader instead of the un-vendored pugixml
//   * float arrays are parsed in one linear pass (the reference re-copies the remaining string for
//     every float: O(n^2), 1.6 s for the 123 K floats of an 80 K-triangle sphere)
//   * malformed input (missing <translate>, missing instance_geometry, unknown technique) yields
//     defined results / `false` instead of reading uninitialised memory or throwing
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "Matrix4x4.h"
#include "vecmath.h"

namespace xml_lite { struct Node; }

struct Effect {
    static const int PHONG = 0x001;
    static const int COOK_TORRANCE = 0x002;

    int technique = PHONG;
    float4 emission, ambient, diffuse, specular;
    float shininess = 0.0f;
    float4 reflective;
    float reflectivity = 0.0f;
    float4 transparent;
    float transparency = 0.0f;
    float glossiness = 0.0f;

    void set(int tech, float4 emi, float4 amb, float4 diff, float4 spec, float shini, float4 refl, float refl_ty,
             float4 transp, float transp_cy, float gloss);
};

struct PolygonTriangle {
    int effect_index = 0;
    int3 vertex_indices, normal_indices, uv0_indices;
};

struct Geometry {
    std::vector<float3> float_array_positions;
    std::vector<float3> float_array_normals;
    std::vector<float2> float_array_uv0;
    std::vector<PolygonTriangle> polygons;
};

struct SceneNode {
    int geometry_index = 0;
    Matrix4x4 matrix;
};

// Parse up to `num_floats` whitespace-separated floats; returns how many were read.
int stof_array(const std::string& s, int num_floats, float* pf);

class ColladaLoader {
public:
    std::vector<Effect> library_effects;
    std::vector<Geometry> library_geometries;
    std::vector<SceneNode> library_visual_scenes;

    ColladaLoader() {}
    bool load(const char* filename);
    const std::string& error() const { return error_; }

    // position / direction of geometry `geometry_index` through its scene node's matrix (v * M)
    float3 get_vertex(const float3& v, int geometry_index);
    float3 get_normal(const float3& n, int geometry_index);

private:
    bool load_effects(const xml_lite::Node* lib);
    bool load_effect(const xml_lite::Node* fx, int count);
    bool load_geometries(const xml_lite::Node* collada);
    bool load_geometry(const xml_lite::Node* geo, int count);
    bool load_polygons(const xml_lite::Node* polys, Geometry& g);
    bool load_visual_scenes(const xml_lite::Node* collada);
    bool load_scene_node(const xml_lite::Node* node);
    Matrix4x4 load_scene_node_matrix(const xml_lite::Node* node);
    void compute_geometry_to_scene_index();
    const Matrix4x4& matrix_of(int geometry_index);

    std::unordered_map<std::string, int> effect_name_to_index_;
    std::unordered_map<std::string, int> geometry_id_to_index_;
    std::unordered_map<int, int> geometry_to_scene_index_;
    std::string error_;
};
