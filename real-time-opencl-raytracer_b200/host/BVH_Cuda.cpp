#include "BVH_Cuda.h"

#include <cmath>
#include <cstdio>
#include <utility>

void AABB::clear() {
    // the reference's "empty" corners are +-e^80 evaluated in float (reference AABB.h:18-21)
    const float big = expf(80.0f);
    min = make_float4(big, big, big, 1.0f);
    max = make_float4(-big, -big, -big, 1.0f);
}

void BVH_Cuda::build_from_bvh2(FW::BVH2& bvh2) {
    clear();
    tri_indices = bvh2.getTriIndices();
    for (int& t : tri_indices) t *= 3;
    if (bvh2.getRoot()) build2(bvh2.getRoot());
}

// Pre-order numbering without recursion: a node gets the next free index when it is popped; its
// right child is pushed first so the left child is popped (and numbered) immediately after it.
int BVH_Cuda::build2(FW::BVHNode* root) {
    const int first = (int)bvh_nodes.size();
    struct Pending { FW::BVHNode* node; int parent; };  // parent < 0: none; else slot whose offset_right to patch
    std::vector<Pending> stack;
    stack.push_back({root, -1});
    while (!stack.empty()) {
        Pending p = stack.back();
        stack.pop_back();
        const int idx = (int)bvh_nodes.size();
        if (p.parent >= 0) bvh_nodes[p.parent].offset_right = idx;
        BVH_Node_ out;
        out.aabb.set(p.node->m_bounds.minf(), p.node->m_bounds.maxf());
        if (!p.node->isLeaf()) {
            out.offset_left = idx + 1;
            stack.push_back({p.node->getChildNode(1), idx});  // right: numbered after the whole left subtree
            stack.push_back({p.node->getChildNode(0), -1});   // left: numbered next
        } else {
            const FW::LeafNode* leaf = static_cast<const FW::LeafNode*>(p.node);
            out.offset_tris = leaf->m_lo;
            out.num_tris = leaf->getNumTriangles();
        }
        bvh_nodes.push_back(out);
    }
    return first;
}

bool BVH_Cuda::save(const char* path) const {
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    const unsigned magic = 0x48564246u;  // "FBVH"
    const int n = (int)bvh_nodes.size(), r = (int)tri_indices.size();
    bool ok = fwrite(&magic, 4, 1, f) == 1 && fwrite(&n, 4, 1, f) == 1 && fwrite(&r, 4, 1, f) == 1;
    ok = ok && fwrite(bvh_nodes.data(), sizeof(BVH_Node_), n, f) == (size_t)n;
    ok = ok && fwrite(tri_indices.data(), 4, r, f) == (size_t)r;
    fclose(f);
    return ok;
}

bool BVH_Cuda::load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    unsigned magic = 0;
    int n = 0, r = 0;
    bool ok = fread(&magic, 4, 1, f) == 1 && magic == 0x48564246u && fread(&n, 4, 1, f) == 1 && fread(&r, 4, 1, f) == 1 &&
              n >= 0 && r >= 0;
    if (ok) {
        bvh_nodes.resize(n);
        tri_indices.resize(r);
        ok = fread(bvh_nodes.data(), sizeof(BVH_Node_), n, f) == (size_t)n && fread(tri_indices.data(), 4, r, f) == (size_t)r;
    }
    fclose(f);
    if (!ok) clear();
    return ok;
}
