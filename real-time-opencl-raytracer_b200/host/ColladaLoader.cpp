#include "ColladaLoader.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "xml_lite.h"

using xml_lite::Node;

namespace {
const double kPi = 3.14159265358979323846;

const Node* child(const Node* n, const char* name) { return n ? n->child(name) : nullptr; }
const char* attr(const Node* n, const char* name) { return n ? n->attribute(name) : ""; }
const std::string kEmpty;
const std::string& text(const Node* n) { return n ? n->text : kEmpty; }
std::string strip_hash(const char* url) { return (url && url[0]) ? std::string(url + 1) : std::string(); }

const char* const kAttrNames[10] = {"emission", "ambient", "diffuse", "specular", "shininess",
                                    "reflective", "reflectivity", "transparent", "transparency", "glossiness"};
const int kAttrFloats[10] = {4, 4, 4, 4, 1, 4, 1, 4, 1, 1};
}  // namespace

int stof_array(const std::string& s, int num_floats, float* pf) {
    if (s.empty() || num_floats < 1 || !pf) return 0;
    const char* p = s.c_str();
    int n = 0;
    while (n < num_floats) {
        char* end = nullptr;
        float v = strtof(p, &end);
        if (end == p) break;
        pf[n++] = v;
        p = end;
    }
    return n;
}

void Effect::set(int tech, float4 emi, float4 amb, float4 diff, float4 spec, float shini, float4 refl, float refl_ty,
                 float4 transp, float transp_cy, float gloss) {
    technique = tech;
    emission = emi;
    ambient = amb;
    diffuse = diff;
    specular = spec;
    shininess = shini;
    reflective = refl;
    reflectivity = refl_ty;
    transparent = transp;
    transparency = transp_cy;
    glossiness = gloss;
}

bool ColladaLoader::load(const char* filename) {
    xml_lite::Document doc;
    if (!doc.load_file(filename)) {
        error_ = doc.error();
        return false;
    }
    const Node* collada = doc.child("COLLADA");
    if (!collada) {
        error_ = "no <COLLADA> root element";
        return false;
    }
    load_effects(child(collada, "library_effects"));
    if (!load_geometries(collada) && !error_.empty()) return false;
    load_visual_scenes(collada);
    compute_geometry_to_scene_index();
    for (const Geometry& g : library_geometries)
        for (const PolygonTriangle& t : g.polygons)
            if (t.effect_index < 0 || (size_t)t.effect_index >= library_effects.size()) {
                error_ = "a <polygons> element names a material that no <effect> defines";
                return false;
            }
    return true;
}

bool ColladaLoader::load_effects(const Node* lib) {
    if (!lib) return false;
    int count = 0;
    for (const Node* fx = lib->child("effect"); fx; fx = fx->next_sibling("effect"), ++count) load_effect(fx, count);
    return true;
}

// <effect name> / profile_COMMON / technique / {cook-torrance | phong} / <attr> / {color | float}
bool ColladaLoader::load_effect(const Node* fx, int count) {
    effect_name_to_index_[attr(fx, "name")] = count;
    const Node* technique = child(child(fx, "profile_COMMON"), "technique");
    Effect e;
    const Node* shader = child(technique, "cook-torrance");
    if (shader) {
        e.technique = Effect::COOK_TORRANCE;
    } else if ((shader = child(technique, "phong"))) {
        e.technique = Effect::PHONG;
    } else {
        fprintf(stderr, "ColladaLoader: effect '%s' has neither <cook-torrance> nor <phong>\n", attr(fx, "name"));
        return false;
    }
    float4 c[10];
    for (int i = 0; i < 10; ++i)
        stof_array(text(child(child(shader, kAttrNames[i]), kAttrFloats[i] == 4 ? "color" : "float")), kAttrFloats[i], c[i].m);
    e.set(e.technique, c[0], c[1], c[2], c[3], c[4].x, c[5], c[6].x, c[7], c[8].x, c[9].x);
    library_effects.push_back(e);
    return true;
}

bool ColladaLoader::load_geometries(const Node* collada) {
    const Node* lib = child(collada, "library_geometries");
    if (!lib) return false;
    int count = 0;
    for (const Node* g = lib->child("geometry"); g; g = g->next_sibling("geometry"), ++count)
        if (!load_geometry(g, count)) return false;
    return true;
}

namespace {
// source of the <input semantic=...> among the first three inputs of <polygons>
std::string input_source(const Node* polys, const char* semantic) {
    const Node* in = child(polys, "input");
    for (int i = 0; i < 3 && in; ++i, in = in->next_sibling("input"))
        if (!strcmp(attr(in, "semantic"), semantic)) return strip_hash(attr(in, "source"));
    return std::string();
}
const Node* float_array_of(const Node* mesh, const std::string& source_id) {
    for (const Node* s = child(mesh, "source"); s; s = s->next_sibling("source"))
        if (source_id == attr(s, "id")) return s->child("float_array");
    return nullptr;
}
// `count` comes from the file: it is trusted only as far as the text can back it (a float needs at least two
// characters, digit + separator) and only whole elements are parsed, so a count that is negative, absurd or not a
// multiple of `comps` can neither throw out of resize() nor write past the vector.
template <class V> bool read_array(const Node* float_array, int comps, std::vector<V>& out) {
    out.clear();
    if (!float_array) return true;  // a missing optional source (e.g. no TEXCOORD) is not an error
    const long long num_floats = atoll(attr(float_array, "count"));
    const std::string& txt = text(float_array);
    if (num_floats < 0 || num_floats > (long long)(txt.size() + 1) / 2 + 1) return false;
    const size_t elems = (size_t)(num_floats / comps);
    out.resize(elems);
    if (elems == 0) return true;
    static_assert(sizeof(V) % sizeof(float) == 0, "V is a plain float vector");
    if (sizeof(V) != sizeof(float) * (size_t)comps) return false;
    stof_array(txt, (int)(elems * comps), out[0].m);  // a short text leaves zeros, as resize() made them
    return true;
}
}  // namespace

bool ColladaLoader::load_geometry(const Node* geo, int count) {
    const Node* mesh = child(geo, "mesh");
    const Node* polys = child(mesh, "polygons");
    geometry_id_to_index_[attr(geo, "id")] = count;

    Geometry g;
    load_polygons(polys, g);

    // VERTEX goes through <vertices id> -> its <input source>; NORMAL / TEXCOORD name a <source> directly
    const std::string vtx_id = input_source(polys, "VERTEX");
    std::string pos_id;
    for (const Node* v = child(mesh, "vertices"); v; v = v->next_sibling("vertices"))
        if (vtx_id == attr(v, "id")) {
            pos_id = strip_hash(attr(v->child("input"), "source"));
            break;
        }
    const bool arrays_ok = read_array(float_array_of(mesh, pos_id), 3, g.float_array_positions) &&
                           read_array(float_array_of(mesh, input_source(polys, "NORMAL")), 3, g.float_array_normals) &&
                           read_array(float_array_of(mesh, input_source(polys, "TEXCOORD")), 2, g.float_array_uv0);
    if (!arrays_ok) {
        error_ = std::string("geometry '") + attr(geo, "id") + "': float_array count does not match its text";
        return false;
    }
    // Index ranges are checked here, before anything (Mesh::init, the SBVH builder) dereferences them.
    for (const PolygonTriangle& t : g.polygons)
        for (int k = 0; k < 3; ++k) {
            const int vi = t.vertex_indices.m[k], ni = t.normal_indices.m[k];
            if (vi < 0 || (size_t)vi >= g.float_array_positions.size() || ni < 0 ||
                (!g.float_array_normals.empty() && (size_t)ni >= g.float_array_normals.size()) ||
                (g.float_array_normals.empty() && ni != 0)) {
                error_ = std::string("geometry '") + attr(geo, "id") + "': <p> index out of range";
                return false;
            }
        }
    library_geometries.push_back(std::move(g));
    return true;
}

// every <p> holds one triangle as "v n t v n t v n t"
bool ColladaLoader::load_polygons(const Node* polys, Geometry& g) {
    long long num = atoll(attr(polys, "count"));
    long long have = 0;  // a count larger than the <p> elements present is clamped: nothing to read there
    for (const Node* q = child(polys, "p"); q; q = q->next_sibling("p")) ++have;
    if (num < 0) num = 0;
    if (num > have) num = have;
    const int effect_index = effect_name_to_index_[attr(polys, "material")];
    g.polygons.resize((size_t)num);
    const Node* p = child(polys, "p");
    for (long long i = 0; i < num; ++i) {
        PolygonTriangle& t = g.polygons[i];
        t.effect_index = effect_index;
        if (p) {
            sscanf(p->text.c_str(), "%d %d %d %d %d %d %d %d %d", &t.vertex_indices.x, &t.normal_indices.x, &t.uv0_indices.x,
                   &t.vertex_indices.y, &t.normal_indices.y, &t.uv0_indices.y, &t.vertex_indices.z, &t.normal_indices.z,
                   &t.uv0_indices.z);
            p = p->next_sibling("p");
        }
    }
    return true;
}

bool ColladaLoader::load_visual_scenes(const Node* collada) {
    const Node* scene = child(child(collada, "library_visual_scenes"), "visual_scene");
    if (!scene) return false;
    for (const Node* n = scene->child("node"); n; n = n->next_sibling("node")) load_scene_node(n);
    return true;
}

bool ColladaLoader::load_scene_node(const Node* node) {
    SceneNode sn;
    sn.geometry_index = geometry_id_to_index_[strip_hash(attr(child(node, "instance_geometry"), "url"))];
    sn.matrix = load_scene_node_matrix(node);
    library_visual_scenes.push_back(sn);
    return true;
}

Matrix4x4 ColladaLoader::load_scene_node_matrix(const Node* node) {
    Matrix4x4 matrix;
    if (const Node* mnode = child(node, "matrix")) {  // column-vector matrix in the file -> transpose
        float v[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        stof_array(mnode->text, 16, v);
        matrix.set(v);
        matrix.transponse();
        return matrix;
    }
    std::unordered_map<std::string, const Node*> by_sid;
    for (const Node* r = child(node, "rotate"); r; r = r->next_sibling("rotate")) by_sid[attr(r, "sid")] = r;

    // The axis actually rotated about is position % 3 (X, Y, Z), so the sid "rotateZ" drives a
    // Y rotation and "rotateY" a Z rotation -- the reference's behaviour, kept.
    static const char* const kOrder[6] = {"jointOrientX", "jointOrientY", "jointOrientZ", "rotateX", "rotateZ", "rotateY"};
    for (int i = 0; i < 6; i++) {
        auto it = by_sid.find(kOrder[i]);
        if (it == by_sid.end()) continue;
        const std::string& txt = it->second->text;  // "ax ay az angle": the angle follows the first 6 characters
        float angle = 0.0f;
        if (txt.size() > 6) stof_array(txt.substr(6), 1, &angle);
        const float rad = (float)(angle * kPi / 180.0f);
        if (i % 3 == 0) matrix.rotateX(rad);
        else if (i % 3 == 1) matrix.rotateY(rad);
        else matrix.rotateZ(rad);
    }
    float tr[3] = {0.0f, 0.0f, 0.0f};
    stof_array(text(child(node, "translate")), 3, tr);
    matrix.translate(tr[0], tr[1], tr[2]);
    return matrix;
}

void ColladaLoader::compute_geometry_to_scene_index() {
    for (size_t i = 0; i < library_geometries.size() && i < library_visual_scenes.size(); ++i)
        geometry_to_scene_index_[library_visual_scenes[i].geometry_index] = (int)i;
}

const Matrix4x4& ColladaLoader::matrix_of(int geometry_index) {
    static const Matrix4x4 identity;
    auto it = geometry_to_scene_index_.find(geometry_index);
    if (it == geometry_to_scene_index_.end())  // the reference's map lookup default-inserts 0
        return library_visual_scenes.empty() ? identity : library_visual_scenes[0].matrix;
    return library_visual_scenes[it->second].matrix;
}

float3 ColladaLoader::get_vertex(const float3& v, int geometry_index) {
    const float4 r = multiply(float4(v.x, v.y, v.z, 1.0f), matrix_of(geometry_index));
    return float3(r.x, r.y, r.z);
}

float3 ColladaLoader::get_normal(const float3& n, int geometry_index) {
    const float4 r = multiply(float4(n.x, n.y, n.z, 0.0f), matrix_of(geometry_index));
    return float3(r.x, r.y, r.z);
}
