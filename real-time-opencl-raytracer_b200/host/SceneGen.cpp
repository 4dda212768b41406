#include "SceneGen.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <unordered_map>

namespace scenegen {

void add_icosphere(Mesh& m, int subdiv, float radius, float cx, float cy, float cz) {
    // unit icosahedron in double, subdivided by edge midpoints re-projected to the sphere
    struct V { double x, y, z; };
    std::vector<V> v;
    std::vector<int> f;
    const double t = (1.0 + std::sqrt(5.0)) / 2.0;
    const double raw[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                               {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    for (auto& r : raw) {
        double l = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        v.push_back({r[0] / l, r[1] / l, r[2] / l});
    }
    const int faces[20][3] = {{0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                              {11, 10, 2}, {10, 7, 6}, {7, 1, 8}, {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8},
                              {3, 8, 9}, {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};
    for (auto& fc : faces) { f.push_back(fc[0]); f.push_back(fc[1]); f.push_back(fc[2]); }

    for (int s = 0; s < subdiv; s++) {
        std::unordered_map<uint64_t, int> mid;
        std::vector<int> nf;
        nf.reserve(f.size() * 4);
        auto midpoint = [&](int a, int b) {
            uint64_t key = a < b ? ((uint64_t)a << 32) | (uint32_t)b : ((uint64_t)b << 32) | (uint32_t)a;
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            V p = {(v[a].x + v[b].x) * 0.5, (v[a].y + v[b].y) * 0.5, (v[a].z + v[b].z) * 0.5};
            double l = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
            v.push_back({p.x / l, p.y / l, p.z / l});
            int idx = (int)v.size() - 1;
            mid.emplace(key, idx);
            return idx;
        };
        for (size_t i = 0; i < f.size(); i += 3) {
            int a = f[i], b = f[i + 1], c = f[i + 2];
            int ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            const int quad[12] = {a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca};
            nf.insert(nf.end(), quad, quad + 12);
        }
        f.swap(nf);
    }
    const int base = (int)m.vertices.size();
    for (const V& p : v)
        m.add_vertex((float)(cx + radius * p.x), (float)(cy + radius * p.y), (float)(cz + radius * p.z));
    for (size_t i = 0; i < f.size(); i += 3) m.add_indexed(base + f[i], base + f[i + 1], base + f[i + 2]);
}

void add_terrain(Mesh& m, int nquads, float half) {
    const int n = nquads + 1;
    const int base = (int)m.vertices.size();
    m.reserve_tris(m.indices.size() / 3 + (size_t)nquads * nquads * 2, m.vertices.size() + (size_t)n * n);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
            const double x = -half + 2.0 * half * i / nquads;
            const double z = -half + 2.0 * half * j / nquads;
            const double y = 6.0 * std::sin(0.15 * x) * std::cos(0.11 * z) + 1.5 * std::sin(0.9 * x + 0.4 * z);
            m.add_vertex((float)x, (float)y, (float)z);
        }
    for (int j = 0; j < nquads; j++)
        for (int i = 0; i < nquads; i++) {
            const int a = base + j * n + i, b = a + 1, c = a + n, d = c + 1;
            m.add_indexed(a, c, b);
            m.add_indexed(b, c, d);
        }
}

void add_sphere_field(Mesh& m, int grid, float pitch, int subdiv, float radius) {
    const float off = 0.5f * pitch * (float)(grid - 1);
    for (int j = 0; j < grid; j++)
        for (int i = 0; i < grid; i++) add_icosphere(m, subdiv, radius, i * pitch - off, 0.0f, j * pitch - off);
    add_icosphere(m, subdiv, radius, 0.0f, 150.0f, 0.0f);
}

static inline uint32_t lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

void add_sticks(Mesh& m, int count, unsigned seed, float extent) {
    uint32_t state = seed * 2654435761u + 12345u;
    auto rnd = [&]() {  // uniform in [0,1)
        state = lowbias32(state + 0x9e3779b9u);
        return (float)(state >> 8) * (1.0f / 16777216.0f);
    };
    for (int i = 0; i < count; i++) {
        float3 p((rnd() - 0.5f) * extent, (rnd() - 0.5f) * extent, (rnd() - 0.5f) * extent);
        float3 d(rnd() - 0.5f, rnd() - 0.5f, rnd() - 0.5f);
        d = normalize(d) * (extent * (0.3f + 0.7f * rnd()));
        float3 w(rnd() - 0.5f, rnd() - 0.5f, rnd() - 0.5f);
        w = normalize(w) * (extent * 0.004f);
        float3 a = p - d * 0.5f, b = p + d * 0.5f, c = p + w;
        m.add_tri(float4(a.x, a.y, a.z, 1.0f), float4(b.x, b.y, b.z, 1.0f), float4(c.x, c.y, c.z, 1.0f));
    }
}

static void tri(Mesh& m, float ax, float ay, float az, float bx, float by, float bz, float cx, float cy, float cz) {
    m.add_tri(float4(ax, ay, az, 1.0f), float4(bx, by, bz, 1.0f), float4(cx, cy, cz, 1.0f));
}

void add_bvh_test0(Mesh& m) {  // four fins + a floor quad at y = -5
    tri(m, -10, 0, 0, -10, 10, 0, -1, 0, 0);
    tri(m, 10, 0, 0, 10, 10, 0, 1, 0, 0);
    tri(m, 0, 0, -10, 0, 10, -10, 0, 0, -1);
    tri(m, 0, 0, 10, 0, 10, 10, 0, 0, 1);
    tri(m, -10, -5, -10, -10, -5, 10, 10, -5, -10);
    tri(m, -10, -5, 10, 10, -5, 10, 10, -5, -10);
}

void add_bvh_test1(Mesh& m) {  // three fins meeting at the origin + a floor quad at y = 0
    tri(m, -10, 0, 0, -10, 10, 0, 0, 0, 0);
    tri(m, 10, 0, 0, 10, 10, 0, 0, 0, 0);
    tri(m, 0, 0, -10, 0, 10, -10, 0, 0, 0);
    tri(m, -10, 0, -10, -10, 0, 10, 10, 0, -10);
    tri(m, -10, 0, 10, 10, 0, 10, 10, 0, -10);
}

bool write_dae(const Mesh& m, const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) return false;
    const size_t T = m.indices.size() / 3;
    const Material mat = m.materials.empty() ? Material() : m.materials[0];
    fprintf(f, "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n");
    fprintf(f, "<COLLADA xmlns=\"http://www.collada.org/2005/11/COLLADASchema\" version=\"1.4.1\">\n");
    fprintf(f, " <library_effects>\n  <effect id=\"fx0\" name=\"mat0\">\n   <profile_COMMON>\n    <technique sid=\"common\">\n     <phong>\n");
    auto color = [&](const char* name, const float4& c) {
        fprintf(f, "      <%s><color>%.9g %.9g %.9g %.9g</color></%s>\n", name, c.x, c.y, c.z, c.w, name);
    };
    auto scalar = [&](const char* name, float v) { fprintf(f, "      <%s><float>%.9g</float></%s>\n", name, v, name); };
    color("emission", mat.emission);
    color("ambient", mat.ambient);
    color("diffuse", mat.diffuse);
    color("specular", mat.specular);
    scalar("shininess", mat.shininess.x);
    color("reflective", mat.reflective);
    scalar("reflectivity", mat.reflectivity.x);
    color("transparent", mat.transparent);
    scalar("transparency", mat.transparency.x);
    scalar("glossiness", mat.glossiness.x);
    fprintf(f, "     </phong>\n    </technique>\n   </profile_COMMON>\n  </effect>\n </library_effects>\n");

    fprintf(f, " <library_geometries>\n  <geometry id=\"geo0\" name=\"geo0\">\n   <mesh>\n");
    fprintf(f, "    <source id=\"geo0-pos\"><float_array id=\"geo0-pos-array\" count=\"%zu\">", m.vertices.size() * 3);
    for (const float4& v : m.vertices) fprintf(f, "%.9g %.9g %.9g ", v.x, v.y, v.z);
    fprintf(f, "</float_array></source>\n");
    fprintf(f, "    <source id=\"geo0-nrm\"><float_array id=\"geo0-nrm-array\" count=\"%zu\">", m.normals.size() * 3);
    for (const float4& v : m.normals) fprintf(f, "%.9g %.9g %.9g ", v.x, v.y, v.z);
    fprintf(f, "</float_array></source>\n");
    fprintf(f, "    <source id=\"geo0-uv\"><float_array id=\"geo0-uv-array\" count=\"2\">0 0</float_array></source>\n");
    fprintf(f, "    <vertices id=\"geo0-vtx\"><input semantic=\"POSITION\" source=\"#geo0-pos\"/></vertices>\n");
    fprintf(f, "    <polygons material=\"mat0\" count=\"%zu\">\n", T);
    fprintf(f, "     <input semantic=\"VERTEX\" source=\"#geo0-vtx\" offset=\"0\"/>\n");
    fprintf(f, "     <input semantic=\"NORMAL\" source=\"#geo0-nrm\" offset=\"1\"/>\n");
    fprintf(f, "     <input semantic=\"TEXCOORD\" source=\"#geo0-uv\" offset=\"2\" set=\"0\"/>\n");
    for (size_t t = 0; t < T; t++)
        fprintf(f, "     <p>%d %d 0 %d %d 0 %d %d 0</p>\n", m.indices[3 * t], m.normals_indices[3 * t], m.indices[3 * t + 1],
                m.normals_indices[3 * t + 1], m.indices[3 * t + 2], m.normals_indices[3 * t + 2]);
    fprintf(f, "    </polygons>\n   </mesh>\n  </geometry>\n </library_geometries>\n");
    fprintf(f, " <library_visual_scenes>\n  <visual_scene id=\"scene0\">\n   <node id=\"node0\" name=\"node0\">\n");
    fprintf(f, "    <matrix>1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1</matrix>\n");
    fprintf(f, "    <instance_geometry url=\"#geo0\"/>\n   </node>\n  </visual_scene>\n </library_visual_scenes>\n");
    fprintf(f, "</COLLADA>\n");
    fclose(f);
    return true;
}

}  // namespace scenegen
