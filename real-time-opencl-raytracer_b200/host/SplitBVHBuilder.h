// SplitBVHBuilder.h -- spatial-split BVH (SBVH) builder.
//
// Produces, for the same Mesh, a tree and a triangle-index list BYTE-IDENTICAL (after
// BVH_Cuda::build_from_bvh2) to what the reference's FW::SplitBVHBuilder produces
// (reference SplitBVHBuilder.cpp:41-476): full-sweep SAH object splits on x, y, z with the
// (centroid, triIdx) total order, 128-bin spatial splits with reference un-splitting, SAH leaf
// termination with leaf sizes 1..8, depth limits 64 / 48, right subtree emitted before the left.
// tests/test_host_bvh.py checks that claim against the reference's own compiled objects
// (oracle/_ref/ref_host) and against golden fixtures made from them.
//
// It is NOT the reference's code: the reference recurses on one shared reference stack with a
// function-pointer quicksort (57 s for 1 M triangles, single thread). This builder
//   * sorts (orderable-float key | triIdx) 64-bit integers with std::sort and gathers once,
//   * skips the re-sort when the winning axis is the one the range is already sorted on,
//   * builds large subtrees as OpenMP tasks, each on its own reference stack and scratch, and
//     splices the per-task leaf lists back in the reference's emission order (right, then left).
// Every fp32 expression that feeds a decision keeps the reference's operand order; the translation
// unit is compiled with -ffp-contract=off.
#pragma once
#include <vector>

#include "BVHNode.h"
#include "Platform.h"

namespace FW {

class BVH2;

class SplitBVHBuilder {
public:
    enum { MaxDepth = 64, MaxSpatialDepth = 48, NumSpatialBins = 128 };

    explicit SplitBVHBuilder(BVH2& bvh);
    ~SplitBVHBuilder();
    SplitBVHBuilder(const SplitBVHBuilder&) = delete;
    SplitBVHBuilder& operator=(const SplitBVHBuilder&) = delete;

    BVHNode* run();  // fills bvh.getTriIndices(), returns the root (caller owns the tree)

    int numDuplicates() const { return m_numDuplicates; }
    void setParallelThreshold(int refs) { m_parallelThreshold = refs; }  // <=0: single task

    struct Reference {
        S32 triIdx;
        AABB bounds;
        Reference() : triIdx(-1) {}
    };
    struct NodeSpec {
        S32 numRef;
        AABB bounds;
        NodeSpec() : numRef(0) {}
    };

private:
    struct ObjectSplit {
        F32 sah;
        S32 sortDim, numLeft;
        AABB leftBounds, rightBounds;
        ObjectSplit() : sah(kF32Max), sortDim(0), numLeft(0) {}
    };
    struct SpatialSplit {
        F32 sah;
        S32 dim;
        F32 pos;
        SpatialSplit() : sah(kF32Max), dim(0), pos(0.0f) {}
    };
    struct SpatialBin {
        AABB bounds;
        S32 enter, exit;
    };
    struct SortItem;
    struct Task;  // one reference stack + scratch + local leaf list

    BVHNode* buildNode(Task& t, NodeSpec spec, int level);
    BVHNode* createLeaf(Task& t, const NodeSpec& spec);
    void sortTop(Task& t, int numRef, int dim);
    ObjectSplit findObjectSplit(Task& t, const NodeSpec& spec, F32 nodeSAH);
    void performObjectSplit(Task& t, NodeSpec& left, NodeSpec& right, const NodeSpec& spec, const ObjectSplit& split);
    SpatialSplit findSpatialSplit(Task& t, const NodeSpec& spec, F32 nodeSAH);
    void performSpatialSplit(Task& t, NodeSpec& left, NodeSpec& right, const NodeSpec& spec, const SpatialSplit& split);
    void splitReference(Reference& left, Reference& right, const Reference& ref, int dim, F32 pos) const;
    void spliceChild(Task& parent, Task& child);

    BVH2& m_bvh;
    const Platform& m_platform;
    const int3* m_tris;
    const float4* m_verts;
    F32 m_minOverlap;
    F32 m_splitAlpha;
    int m_numDuplicates;
    int m_parallelThreshold;
};

}  // namespace FW
