// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A small command-line driver that is linked against the UNMODIFIED reference host objects
// (compiled where they lie under /root/reference by oracle/Makefile, outputs only in oracle/_ref/).
// It lets the tests and the golden-vector generator ask the reference's own code:
//
//   ref_host build  in.bin out.bin   Mesh -> FW::BVH2::setMesh (SplitBVHBuilder::run, reference
//                                    SplitBVHBuilder.cpp:41) -> BVH_Cuda::build_from_bvh2
//                                    (reference BVH_Cuda.h:87) ; dumps bvh_nodes + tri_indices
//   ref_host mt     in.bin out.bin   spec::RayTriangleIntersection (reference common.h:193-219)
//   ref_host box    in.bin out.bin   spec::RayBoxIntersection      (reference common.h:172-190)
//   ref_host vec    in.bin out.bin   dot / cross / normalize       (reference vectors_math.cpp:73-84)
//
// One process handles ONE `build`: the reference flattener keeps function-static counters
// (BVH_Cuda.h:100-101), so it is only correct once per process.
//
// Binary formats (little endian):
//   build in : int32 V, int32 T, float32 verts[V][4], int32 indices[T][3]
//   build out: int32 N, int32 R, 48-byte nodes[N], int32 tri_indices[R]
//   mt in    : int32 n, n x {o[3], d[3], v0[3], e1[3], e2[3]} float32 ; out: n x float32 t
//   box in   : int32 n, n x {bmin[3], bmax[3], org[3], invdir[3]}     ; out: n x {int32 hit, f32 tmin, f32 tmax}
//   vec in   : int32 n, n x {a[3], b[3]} ; out: n x {dot, cross[3], normalize(a)[3]} float32
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "Mesh.h"
#include "BVH2.h"
#include "BVH_Cuda.h"
#include "common.h"

// Mesh.cpp references these two ColladaLoader members (Mesh.cpp:50,71); ColladaLoader.cpp itself
// cannot be linked without pugixml, and `build` never goes through Mesh::init(ColladaLoader&).
float3 ColladaLoader::get_vertex(float3& v, int) { return v; }
float3 ColladaLoader::get_normal(float3& v, int) { return v; }

static std::vector<char> slurp(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "ref_host: cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf(n);
    if (n && fread(buf.data(), 1, n, f) != (size_t)n) { fprintf(stderr, "ref_host: short read\n"); exit(2); }
    fclose(f);
    return buf;
}

static int do_build(const char* in, const char* out) {
    std::vector<char> buf = slurp(in);
    const int* hdr = (const int*)buf.data();
    const int V = hdr[0], T = hdr[1];
    const float* vp = (const float*)(hdr + 2);
    const int* ip = (const int*)(vp + (size_t)V * 4);

    static Mesh mesh;  // static storage, like the reference's global `mesh1` (RayTracer.cpp:41)
    mesh.vertices.resize(V);
    for (int i = 0; i < V; i++) mesh.vertices[i] = float4(vp[4 * i], vp[4 * i + 1], vp[4 * i + 2], vp[4 * i + 3]);
    mesh.indices.assign(ip, ip + (size_t)T * 3);

    // FW::BVH2's constructor reads its own uninitialised m_scene (BVH2.cpp:11-21); the reference
    // only works because `bvh2` is a zero-initialised global (RayTracer.cpp:45). Same here.
    static FW::BVH2 bvh2;
    static BVH_Cuda bvh_cuda;
    bvh2.setMesh(&mesh);
    bvh_cuda.build_from_bvh2(bvh2);

    FILE* f = fopen(out, "wb");
    if (!f) { fprintf(stderr, "ref_host: cannot write %s\n", out); return 2; }
    int N = (int)bvh_cuda.bvh_nodes.size(), R = (int)bvh_cuda.tri_indices.size();
    fwrite(&N, 4, 1, f);
    fwrite(&R, 4, 1, f);
    static_assert(sizeof(BVH_Node_) == 48, "reference node must be 48 bytes");
    fwrite(bvh_cuda.bvh_nodes.data(), sizeof(BVH_Node_), N, f);
    fwrite(bvh_cuda.tri_indices.data(), 4, R, f);
    fclose(f);
    printf("\nref_host build: V=%d T=%d -> N=%d R=%d\n", V, T, N, R);
    return 0;
}

static int do_mt(const char* in, const char* out) {
    std::vector<char> buf = slurp(in);
    const int n = *(const int*)buf.data();
    const float* p = (const float*)(buf.data() + 4);
    std::vector<float> res(n);
    for (int i = 0; i < n; i++, p += 15) {
        spec::Ray r;
        r.o = float3(p[0], p[1], p[2]);
        r.dir = float3(p[3], p[4], p[5]);
        res[i] = spec::RayTriangleIntersection(r, float3(p[6], p[7], p[8]), float3(p[9], p[10], p[11]),
                                               float3(p[12], p[13], p[14]));
    }
    FILE* f = fopen(out, "wb");
    fwrite(res.data(), 4, n, f);
    fclose(f);
    return 0;
}

static int do_box(const char* in, const char* out) {
    std::vector<char> buf = slurp(in);
    const int n = *(const int*)buf.data();
    const float* p = (const float*)(buf.data() + 4);
    FILE* f = fopen(out, "wb");
    for (int i = 0; i < n; i++, p += 12) {
        float tmin, tmax;
        int hit = spec::RayBoxIntersection(float3(p[0], p[1], p[2]), float3(p[3], p[4], p[5]), float3(p[6], p[7], p[8]),
                                           float3(p[9], p[10], p[11]), tmin, tmax);
        fwrite(&hit, 4, 1, f);
        fwrite(&tmin, 4, 1, f);
        fwrite(&tmax, 4, 1, f);
    }
    fclose(f);
    return 0;
}

static int do_vec(const char* in, const char* out) {
    std::vector<char> buf = slurp(in);
    const int n = *(const int*)buf.data();
    const float* p = (const float*)(buf.data() + 4);
    FILE* f = fopen(out, "wb");
    for (int i = 0; i < n; i++, p += 6) {
        float3 a(p[0], p[1], p[2]), b(p[3], p[4], p[5]);
        float d = dot(a, b);
        float3 c = cross(a, b);
        float3 nn = normalize(a);
        float o[7] = {d, c.x, c.y, c.z, nn.x, nn.y, nn.z};
        fwrite(o, 4, 7, f);
    }
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: ref_host {build|mt|box|vec} in.bin out.bin\n");
        return 2;
    }
    if (!strcmp(argv[1], "build")) return do_build(argv[2], argv[3]);
    if (!strcmp(argv[1], "mt")) return do_mt(argv[2], argv[3]);
    if (!strcmp(argv[1], "box")) return do_box(argv[2], argv[3]);
    if (!strcmp(argv[1], "vec")) return do_vec(argv[2], argv[3]);
    fprintf(stderr, "ref_host: unknown mode %s\n", argv[1]);
    return 2;
}
