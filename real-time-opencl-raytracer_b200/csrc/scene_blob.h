// scene_blob.h -- layout of the packed device scene ("blob").
//
// rt_upload_scene takes the reference's arrays (48-byte pre-order nodes, tri_indices = 3*triId,
// indexed float4 vertices; reference BVH_Cuda.h:12-29, RayTracer.cpp:942-984) and re-packs them
// into ONE contiguous, position-independent allocation that the traversal kernels read. One
// allocation = one NCCL broadcast to the other GPUs of the box (rt_scene_blob / rt_adopt_scene_blob).
//
//   [BlobHeader 256 B][node pairs 64 B x P][triangles 48 B x (R+1)][verts][indices][normals]
//   [normal_indices][material diffuse float4 x M][tri_to_material]      every section 256-B aligned
//
// Node pair (64 B = 4 x float4, one per INNER node of the reference tree, same topology):
//   q0 = {c0.min.x, c0.max.x, c0.min.y, c0.max.y}   q1 = {c0.min.z, c0.max.z, bits(ref0), 0}
//   q2 = {c1.min.x, c1.max.x, c1.min.y, c1.max.y}   q3 = {c1.min.z, c1.max.z, bits(ref1), 0}
//   (the two planes of an axis are adjacent: one 64-bit operand of the packed fp32 instructions, device_math.cuh)
//   c0 = the reference node's offset_left child, c1 = offset_right child (order preserved: the
//   traversal's "left first on ties" rule depends on it).
//   ref >= 0          : index of the child's own pair (child is an inner node)
//   ref <  0          : child is a leaf; ~ref = index of its first packed triangle
//   ref == kRefPoison : child is an inner node the reference would bail out on
//                       (offset_right < 0 or a child index >= N, volumeRender.cl:841,855-856)
// Pair order: breadth-first for the first `top_pairs` pairs (the tree top, staged in shared memory
// by the kernels that ask for it), depth-first pre-order for the rest.
//
// Packed triangle (48 B = 3 x float4), stored in tri_indices order so a leaf is contiguous:
//   t0 = {v0.xyz, bits(3*triId)}  t1 = {(v1-v0).xyz, bits(last ? 1 : 0)}  t2 = {(v2-v0).xyz, 0}
//   The edges are the same single fp32 subtraction the reference kernel performs per test
//   (volumeRender.cl:973-974), so pre-subtracting is bit-neutral. `last` terminates a leaf.
//   Slot R is a sentinel "null" triangle (all zero, last = 1) that empty leaves point at.
#pragma once
#include <cstdint>

namespace rtb {

static const uint32_t kBlobMagic = 0x42325452u;  // "RT2B"
static const uint32_t kBlobVersion = 5;
static const int32_t kRefPoison = (int32_t)0x80000000;

struct BlobHeader {
    uint32_t magic, version;
    uint64_t total_bytes;
    int32_t root_ref;      // pair index, leaf ref, or kRefPoison
    int32_t num_pairs;     // P
    int32_t num_tris;      // R (packed triangles, excluding the sentinel)
    int32_t top_pairs;     // pairs [0, top_pairs) are in breadth-first order
    int32_t max_depth;     // depth of the reference tree (root = 1)
    int32_t V, T, Vn, M;   // shading-array sizes (Vn = M = 0: no shading data)
    int32_t num_ref_nodes; // N of the reference array
    int32_t coords_in_window;  // all pair-box coordinates are 0 or have magnitude in [2^-77, 2^60]
    int32_t reserved0[1];
    uint64_t off_pairs, off_tris, off_verts, off_indices, off_normals, off_normal_indices, off_mat_diffuse,
        off_tri_to_material;
    float root_min[3], root_max[3];  // bounds of reference node 0 (used to quantise ray origins when sorting rays)
    uint8_t pad[256 - 8 - 8 - 12 * 4 - 8 * 8 - 6 * 4];
};
static_assert(sizeof(BlobHeader) == 256, "header is one 256-byte block");

struct SceneInputs {  // the reference-layout arrays, host pointers
    const float* verts; int V;
    const int32_t* indices; int T;
    const void* nodes; int N;
    const int32_t* tri_indices; int R;
    const float* normals; int Vn;
    const int32_t* normal_indices;
    const void* materials; int M;
    const int32_t* tri_to_material;
};

// Validate + pack into a host buffer (malloc'ed, caller frees with free()). Returns 0 or fills err.
int pack_scene(const SceneInputs& in, int top_pairs, uint8_t** out_blob, uint64_t* out_bytes, char* err, int err_len);

}  // namespace rtb
