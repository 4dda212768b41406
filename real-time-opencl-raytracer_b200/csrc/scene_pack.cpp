// scene_pack.cpp -- host-side validation + re-packing of the reference's flat scene arrays into the
// device blob described in scene_blob.h. Runs once per rt_upload_scene; O(N + R + V).
#include "scene_blob.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rtb {

namespace {

struct RefNode {  // reference BVH_Node_ (BVH_Cuda.h:12-29), 48 bytes
    float mn[4], mx[4];
    int32_t left, right, tri_off, tri_cnt;
};
static_assert(sizeof(RefNode) == 48, "reference node is 48 bytes");

enum Kind : uint8_t { kUnseen = 0, kInner = 1, kPoison = 2, kLeaf = 3 };

uint64_t align256(uint64_t x) { return (x + 255u) & ~(uint64_t)255u; }

int fail(char* err, int err_len, const char* fmt, long a = 0, long b = 0) {
    if (err && err_len > 0) snprintf(err, err_len, fmt, a, b);
    return 1;
}

inline float as_float(int32_t v) {
    float f;
    memcpy(&f, &v, 4);
    return f;
}

}  // namespace

int pack_scene(const SceneInputs& in, int top_pairs, uint8_t** out_blob, uint64_t* out_bytes, char* err, int err_len) {
    *out_blob = nullptr;
    *out_bytes = 0;
    if (!in.verts || !in.indices || !in.nodes || in.V <= 0 || in.T <= 0 || in.N <= 0 || in.R < 0 || (in.R > 0 && !in.tri_indices))
        return fail(err, err_len, "rt_upload_scene: verts/indices/nodes must be non-empty");
    const bool shading = in.normals && in.normal_indices && in.materials && in.tri_to_material && in.Vn > 0 && in.M > 0;
    if (!shading && (in.Vn > 0 || in.M > 0))
        return fail(err, err_len, "rt_upload_scene: normals/normal_indices/materials/tri_to_material must be given together");
    const RefNode* nodes = (const RefNode*)in.nodes;
    const int N = in.N, R = in.R;

    for (int64_t i = 0; i < (int64_t)in.T * 3; i++)
        if (in.indices[i] < 0 || in.indices[i] >= in.V) return fail(err, err_len, "indices[%ld] = %ld is outside [0, V)", i, in.indices[i]);
    if (shading) {
        for (int64_t i = 0; i < (int64_t)in.T * 3; i++)
            if (in.normal_indices[i] < 0 || in.normal_indices[i] >= in.Vn)
                return fail(err, err_len, "normal_indices[%ld] = %ld is outside [0, Vn)", i, in.normal_indices[i]);
        for (int i = 0; i < in.T; i++)
            if (in.tri_to_material[i] < 0 || in.tri_to_material[i] >= in.M)
                return fail(err, err_len, "tri_to_material[%ld] = %ld is outside [0, M)", i, in.tri_to_material[i]);
    }

    // ---- classify every reachable node, breadth-first (this is also the pair order of the tree top)
    std::vector<uint8_t> kind(N, kUnseen);
    std::vector<int32_t> depth(N, 0);
    std::vector<int32_t> bfs;  // reachable valid inner nodes in BFS order
    std::vector<int32_t> queue;
    queue.push_back(0);
    depth[0] = 1;
    int max_depth = 1;
    for (size_t head = 0; head < queue.size(); head++) {
        const int i = queue[head];
        if (kind[i] != kUnseen) return fail(err, err_len, "BVH node %ld is reachable twice (cycle or shared subtree)", i);
        const RefNode& n = nodes[i];
        if (depth[i] > max_depth) max_depth = depth[i];
        if (n.left >= 0) {
            if (n.right < 0 || n.left >= N || n.right >= N) {
                kind[i] = kPoison;  // the reference kernel returns -1 when it reaches this node
                continue;
            }
            kind[i] = kInner;
            bfs.push_back(i);
            depth[n.left] = depth[n.right] = depth[i] + 1;
            queue.push_back(n.left);
            queue.push_back(n.right);
        } else {
            kind[i] = kLeaf;
            if (n.tri_cnt > 0 && (n.tri_off < 0 || (int64_t)n.tri_off + n.tri_cnt > R))
                return fail(err, err_len, "leaf node %ld: triangle range exceeds tri_indices (R = %ld)", i, R);
        }
    }

    // ---- pair ids: BFS for the top, DFS pre-order (left first) for the rest
    const int P = (int)bfs.size();
    if (top_pairs < 0) top_pairs = 0;
    if (top_pairs > P) top_pairs = P;
    std::vector<int32_t> pair_of(N, -1);
    std::vector<int32_t> order;  // pair id -> reference node
    order.reserve(P);
    for (int k = 0; k < top_pairs; k++) {
        pair_of[bfs[k]] = k;
        order.push_back(bfs[k]);
    }
    if (top_pairs < P) {
        std::vector<int32_t> stack(1, 0);
        while (!stack.empty()) {
            const int i = stack.back();
            stack.pop_back();
            if (kind[i] != kInner) continue;
            if (pair_of[i] < 0) {
                pair_of[i] = (int)order.size();
                order.push_back(i);
            }
            stack.push_back(nodes[i].right);
            stack.push_back(nodes[i].left);
        }
    }

    // ---- blob sections
    BlobHeader h;
    memset(&h, 0, sizeof h);
    h.magic = kBlobMagic;
    h.version = kBlobVersion;
    h.num_pairs = P;
    h.num_tris = R;
    h.top_pairs = top_pairs;
    h.max_depth = max_depth;
    h.V = in.V;
    h.T = in.T;
    h.Vn = shading ? in.Vn : 0;
    h.M = shading ? in.M : 0;
    h.num_ref_nodes = N;
    uint64_t off = sizeof(BlobHeader);
    h.off_pairs = off;            off = align256(off + (uint64_t)P * 64);
    h.off_tris = off;             off = align256(off + ((uint64_t)R + 1) * 48);
    h.off_verts = off;            off = align256(off + (uint64_t)in.V * 16);
    h.off_indices = off;          off = align256(off + (uint64_t)in.T * 12);
    h.off_normals = off;          off = align256(off + (uint64_t)h.Vn * 16);
    h.off_normal_indices = off;   off = align256(off + (shading ? (uint64_t)in.T * 12 : 0));
    h.off_mat_diffuse = off;      off = align256(off + (uint64_t)h.M * 16);
    h.off_tri_to_material = off;  off = align256(off + (shading ? (uint64_t)in.T * 4 : 0));
    h.total_bytes = off;

    uint8_t* blob = (uint8_t*)calloc(1, off);
    if (!blob) return fail(err, err_len, "out of host memory packing the scene (%ld bytes)", (long)off);

    auto child_ref = [&](int c) -> int32_t {
        if (kind[c] == kInner) return pair_of[c];
        if (kind[c] == kPoison) return kRefPoison;
        return nodes[c].tri_cnt > 0 ? ~nodes[c].tri_off : ~R;  // empty leaf -> sentinel triangle
    };
    h.root_ref = child_ref(0);

    float* pairs = (float*)(blob + h.off_pairs);
    bool in_window = true;
    auto check = [&in_window](const float* c) {
        for (int k = 0; k < 3; k++) {
            const float a = c[k] < 0 ? -c[k] : c[k];
            if (!(c[k] == 0.0f || (a >= 6.617444900424222e-24f && a <= 1.152921504606847e18f))) in_window = false;  // device_math.cuh coord_in_window
        }
    };
    for (int p = 0; p < P; p++) {
        const RefNode& n = nodes[order[p]];
        const RefNode &c0 = nodes[n.left], &c1 = nodes[n.right];
        float* q = pairs + (size_t)p * 16;
        // {min.x, max.x, min.y, max.y} {min.z, max.z, ref, 0} per child: the planes of one axis are adjacent (FFMA2 operands)
        q[0] = c0.mn[0]; q[1] = c0.mx[0]; q[2] = c0.mn[1]; q[3] = c0.mx[1];
        q[4] = c0.mn[2]; q[5] = c0.mx[2]; q[6] = as_float(child_ref(n.left)); q[7] = 0.0f;
        q[8] = c1.mn[0]; q[9] = c1.mx[0]; q[10] = c1.mn[1]; q[11] = c1.mx[1];
        q[12] = c1.mn[2]; q[13] = c1.mx[2]; q[14] = as_float(child_ref(n.right)); q[15] = 0.0f;
        check(c0.mn); check(c0.mx); check(c1.mn); check(c1.mx);
    }
    h.coords_in_window = in_window ? 1 : 0;
    for (int k = 0; k < 3; k++) {
        h.root_min[k] = nodes[0].mn[k];
        h.root_max[k] = nodes[0].mx[k];
    }

    // ---- triangles in tri_indices order; `last` set from the leaves
    float* tris = (float*)(blob + h.off_tris);
    std::vector<uint8_t> owned(R + 1, 0);
    int rc = 0;
    for (int i = 0; i < N && !rc; i++) {
        if (kind[i] != kLeaf) continue;
        const RefNode& n = nodes[i];
        for (int k = 0; k < n.tri_cnt; k++) {
            const int slot = n.tri_off + k;
            if (owned[slot]) { rc = fail(err, err_len, "leaf node %ld overlaps another leaf at tri_indices[%ld]", i, slot); break; }
            owned[slot] = 1;
            const int32_t tri = in.tri_indices[slot];
            if (tri < 0 || (int64_t)tri + 2 >= (int64_t)in.T * 3) {
                rc = fail(err, err_len, "tri_indices[%ld] = %ld is outside the index buffer", slot, tri);
                break;
            }
            const float* p0 = in.verts + (size_t)in.indices[tri + 0] * 4;
            const float* p1 = in.verts + (size_t)in.indices[tri + 1] * 4;
            const float* p2 = in.verts + (size_t)in.indices[tri + 2] * 4;
            float* t = tris + (size_t)slot * 12;
            t[0] = p0[0]; t[1] = p0[1]; t[2] = p0[2]; t[3] = as_float(tri);
            // float4 - float4 as in the kernel (volumeRender.cl:973-974): one fp32 subtraction per lane
            t[4] = p1[0] - p0[0]; t[5] = p1[1] - p0[1]; t[6] = p1[2] - p0[2]; t[7] = as_float(k == n.tri_cnt - 1 ? 1 : 0);
            t[8] = p2[0] - p0[0]; t[9] = p2[1] - p0[1]; t[10] = p2[2] - p0[2]; t[11] = 0.0f;
        }
    }
    if (rc) {
        free(blob);
        return rc;
    }
    for (int k = 0; k <= R; k++)  // sentinel + slots no leaf references: zero geometry, terminates
        if (!owned[k]) {
            float* t = tris + (size_t)k * 12;
            t[3] = as_float(-1);
            t[7] = as_float(1);
        }

    memcpy(blob + h.off_verts, in.verts, (size_t)in.V * 16);
    memcpy(blob + h.off_indices, in.indices, (size_t)in.T * 12);
    if (shading) {
        memcpy(blob + h.off_normals, in.normals, (size_t)in.Vn * 16);
        memcpy(blob + h.off_normal_indices, in.normal_indices, (size_t)in.T * 12);
        float* md = (float*)(blob + h.off_mat_diffuse);
        for (int m = 0; m < in.M; m++)  // Material.diffuse = words 12..15 of the 176-byte block
            memcpy(md + (size_t)m * 4, (const uint8_t*)in.materials + (size_t)m * 176 + 48, 16);
        memcpy(blob + h.off_tri_to_material, in.tri_to_material, (size_t)in.T * 4);
    }
    memcpy(blob, &h, sizeof h);
    *out_blob = blob;
    *out_bytes = off;
    return 0;
}

}  // namespace rtb
