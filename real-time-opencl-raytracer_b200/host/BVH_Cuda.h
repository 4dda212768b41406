// BVH_Cuda.h -- the flattened BVH: the exact arrays the device boundary consumes.
//
// Layout contract (reference BVH_Cuda.h:12-29,87-137):
//   BVH_Node_ = 48 bytes {float4 min (w=1), float4 max (w=1), int offset_left, offset_right,
//               offset_tris, num_tris}; nodes in depth-first pre-order, root at 0, so an inner
//               node's left child is always at index+1; leaves have offset_left=offset_right=-1
//   tri_indices[k] = 3 * triangle id (pre-multiplied), leaf ranges are contiguous
// rt_upload_scene (include/rtb200.h) takes bvh_nodes.data() / tri_indices.data() as they are.
//
// Unlike the reference flattener (function-static counters: correct only once per process,
// SURVEY.md Appendix C) this one is re-entrant and iterative.
#pragma once
#include <vector>

#include "BVH2.h"

struct AABB {  // the reference's global (non-FW) box: two float4 corners with w = 1
    float4 min, max;
    AABB() { clear(); }
    void clear();
    void set(const float3& mn, const float3& mx) { min = make_float4(mn); max = make_float4(mx); }
};

struct BVH_Node_ {
    AABB aabb;
    int offset_left, offset_right;
    int offset_tris, num_tris;
    BVH_Node_() : offset_left(-1), offset_right(-1), offset_tris(-1), num_tris(0) {}
};
static_assert(sizeof(BVH_Node_) == 48, "device node is 3 x float4");

class BVH_Cuda {
public:
    std::vector<BVH_Node_> bvh_nodes;
    std::vector<int> tri_indices;

    void clear() { bvh_nodes.clear(); tri_indices.clear(); }
    void build_from_bvh2(FW::BVH2& bvh2);   // tri_indices = 3 * bvh2 triangle ids; nodes = pre-order walk
    int build2(FW::BVHNode* root);          // appends the subtree, returns the index of its root

    // flat-BVH disk cache (addition; the host SBVH build is the slow step): raw little-endian dump
    bool save(const char* path) const;
    bool load(const char* path);
};
