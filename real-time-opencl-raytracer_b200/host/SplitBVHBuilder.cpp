// SplitBVHBuilder.cpp -- see SplitBVHBuilder.h for the contract (identical output to the
// reference builder, reference SplitBVHBuilder.cpp:41-476) and for how this implementation differs.
#include "SplitBVHBuilder.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>

#include "BVH2.h"

namespace FW {

// ---------------------------------------------------------------------------------------------
// Per-task state. The reference keeps ONE stack of references for the whole build and always works
// on its top `numRef` entries; a Task is the same thing for one independently built subtree.
// ---------------------------------------------------------------------------------------------
struct SplitBVHBuilder::SortItem {
    uint64_t key;  // (orderable centroid bits << 32) | triIdx
    uint32_t src;  // position inside the node's range before sorting
};

struct SplitBVHBuilder::Task {
    std::vector<Reference> refs;       // reference stack; the node being built owns the top numRef entries
    std::vector<AABB> rightBounds;     // sweep scratch
    std::vector<SortItem> sortItems;   // sort scratch
    std::vector<Reference> gather;     // sort scratch
    std::vector<S32> tris;             // triangle ids in emission order, local to this task
    std::vector<LeafNode*> leaves;     // leaves whose [m_lo,m_hi) is relative to `tris`
    int duplicates = 0;
    int sortedDim = -1;                // axis the top range is currently sorted on (-1: unknown)
    std::unique_ptr<SpatialBin[]> bins;  // [3][NumSpatialBins]

    void prepare(int maxRefs) {
        rightBounds.resize((size_t)std::max(maxRefs, (int)NumSpatialBins) - 1);
        bins.reset(new SpatialBin[3 * NumSpatialBins]);
    }
};

SplitBVHBuilder::SplitBVHBuilder(BVH2& bvh)
    : m_bvh(bvh), m_platform(bvh.getPlatform()), m_tris(nullptr), m_verts(nullptr), m_minOverlap(0.0f),
      m_splitAlpha(1.0e-5f), m_numDuplicates(0), m_parallelThreshold(16384) {}

SplitBVHBuilder::~SplitBVHBuilder() {}

// Move a finished child task's leaves into the parent, rebasing their triangle ranges.
void SplitBVHBuilder::spliceChild(Task& parent, Task& child) {
    const S32 base = (S32)parent.tris.size();
    parent.tris.insert(parent.tris.end(), child.tris.begin(), child.tris.end());
    for (LeafNode* leaf : child.leaves) {
        leaf->m_lo += base;
        leaf->m_hi += base;
        parent.leaves.push_back(leaf);
    }
    parent.duplicates += child.duplicates;
}

BVHNode* SplitBVHBuilder::run() {
    Mesh* scene = m_bvh.getScene();
    m_tris = (const int3*)scene->getTrisPtr();
    m_verts = scene->getVertsPtr();

    // One reference per triangle, bounds grown vertex by vertex; root bounds grown box by box.
    NodeSpec rootSpec;
    rootSpec.numRef = scene->getNumTriangles();
    Task root;
    root.refs.resize(rootSpec.numRef);
    for (int i = 0; i < rootSpec.numRef; i++) {
        Reference& r = root.refs[i];
        r.triIdx = i;
        for (int j = 0; j < 3; j++) r.bounds.grow(make_float3(m_verts[m_tris[i].m[j]]));
        rootSpec.bounds.grow(r.bounds);
    }
    m_minOverlap = rootSpec.bounds.area() * m_splitAlpha;
    root.prepare(rootSpec.numRef);

    BVHNode* rootNode = nullptr;
#pragma omp parallel
#pragma omp single nowait
    rootNode = buildNode(root, rootSpec, 0);

    std::vector<S32>& out = m_bvh.getTriIndices();
    const S32 base = (S32)out.size();
    out.insert(out.end(), root.tris.begin(), root.tris.end());
    if (base)
        for (LeafNode* leaf : root.leaves) {
            leaf->m_lo += base;
            leaf->m_hi += base;
        }
    m_numDuplicates = root.duplicates;
    return rootNode;
}

// Pops the node's references off the stack tail, so a leaf lists them in reverse stack order.
BVHNode* SplitBVHBuilder::createLeaf(Task& t, const NodeSpec& spec) {
    for (int i = 0; i < spec.numRef; i++) {
        t.tris.push_back(t.refs.back().triIdx);
        t.refs.pop_back();
    }
    LeafNode* leaf = new LeafNode(spec.bounds, (int)t.tris.size() - spec.numRef, (int)t.tris.size());
    t.leaves.push_back(leaf);
    return leaf;
}

BVHNode* SplitBVHBuilder::buildNode(Task& t, NodeSpec spec, int level) {
    // Drop degenerate references (negative extent, or at most one non-zero extent): swap-remove
    // against the stack tail, scanning from the tail down.
    {
        const int first = (int)t.refs.size() - spec.numRef;
        for (int i = (int)t.refs.size() - 1; i >= first; i--) {
            const float3 size = t.refs[i].bounds.maxf() - t.refs[i].bounds.minf();
            if (fminf1(size) < 0.0f || fsumf(size) == fmaxf1(size)) {
                t.refs[i] = t.refs.back();
                t.refs.pop_back();
            }
        }
        spec.numRef = (int)t.refs.size() - first;
    }
    t.sortedDim = -1;

    if (spec.numRef <= m_platform.getMinLeafSize() || level >= MaxDepth) return createLeaf(t, spec);

    const F32 area = spec.bounds.area();
    const F32 leafSAH = area * m_platform.getTriangleCost(spec.numRef);
    const F32 nodeSAH = area * m_platform.getNodeCost(2);
    const ObjectSplit object = findObjectSplit(t, spec, nodeSAH);

    SpatialSplit spatial;
    if (level < MaxSpatialDepth) {
        AABB overlap = object.leftBounds;
        overlap.intersect(object.rightBounds);
        if (overlap.area() >= m_minOverlap) spatial = findSpatialSplit(t, spec, nodeSAH);
    }

    const F32 minSAH = fminf1(leafSAH, object.sah, spatial.sah);
    if (minSAH == leafSAH && spec.numRef <= m_platform.getMaxLeafSize()) return createLeaf(t, spec);

    NodeSpec left, right;
    if (minSAH == spatial.sah) performSpatialSplit(t, left, right, spec, spatial);
    if (!left.numRef || !right.numRef) performObjectSplit(t, left, right, spec, object);
    t.duplicates += left.numRef + right.numRef - spec.numRef;

    // Children: RIGHT subtree first (it owns the stack top), then LEFT.
    BVHNode *rightNode = nullptr, *leftNode = nullptr;
    if (m_parallelThreshold > 0 && left.numRef + right.numRef >= m_parallelThreshold) {
        // Build both subtrees as independent tasks on private stacks, then splice their leaf
        // lists back in emission order: everything the right subtree emitted, then the left's.
        Task rt, lt;
        const size_t n = t.refs.size();
        rt.refs.assign(t.refs.begin() + (n - right.numRef), t.refs.end());
        lt.refs.assign(t.refs.begin() + (n - right.numRef - left.numRef), t.refs.begin() + (n - right.numRef));
        t.refs.resize(n - right.numRef - left.numRef);
        rt.prepare(right.numRef);
        lt.prepare(left.numRef);
#pragma omp task shared(rt, rightNode) firstprivate(right, level)
        rightNode = buildNode(rt, right, level + 1);
#pragma omp task shared(lt, leftNode) firstprivate(left, level)
        leftNode = buildNode(lt, left, level + 1);
#pragma omp taskwait
        spliceChild(t, rt);
        spliceChild(t, lt);
    } else {
        rightNode = buildNode(t, right, level + 1);
        leftNode = buildNode(t, left, level + 1);
    }
    return new InnerNode(spec.bounds, leftNode, rightNode);
}

// Sort the top numRef references by (min[dim] + max[dim], triIdx). The comparator is a strict
// total order inside a node (a triangle id appears at most once per node), so any correct sort
// yields the one order the reference's quicksort yields (reference SplitBVHBuilder.cpp:85-94).
void SplitBVHBuilder::sortTop(Task& t, int numRef, int dim) {
    if (numRef < 2) return;
    Reference* r = t.refs.data() + (t.refs.size() - numRef);
    t.sortItems.resize(numRef);
    for (int i = 0; i < numRef; i++) {
        float c = r[i].bounds.minf().m[dim] + r[i].bounds.maxf().m[dim];
        if (c == 0.0f) c = 0.0f;  // -0 and +0 compare equal in the reference: give them one key
        uint32_t u;
        memcpy(&u, &c, 4);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        t.sortItems[i].key = ((uint64_t)u << 32) | (uint32_t)r[i].triIdx;
        t.sortItems[i].src = (uint32_t)i;
    }
    std::sort(t.sortItems.begin(), t.sortItems.end(), [](const SortItem& a, const SortItem& b) { return a.key < b.key; });
    t.gather.resize(numRef);
    for (int i = 0; i < numRef; i++) t.gather[i] = r[t.sortItems[i].src];
    std::copy(t.gather.begin(), t.gather.end(), r);
    t.sortedDim = dim;
}

SplitBVHBuilder::ObjectSplit SplitBVHBuilder::findObjectSplit(Task& t, const NodeSpec& spec, F32 nodeSAH) {
    ObjectSplit split;
    F32 bestTieBreak = kF32Max;
    const int n = spec.numRef;

    for (int dim = 0; dim < 3; dim++) {
        sortTop(t, n, dim);
        const Reference* r = t.refs.data() + (t.refs.size() - n);

        AABB rightBox;  // suffix bounds: rightBounds[i-1] covers r[i..n)
        for (int i = n - 1; i > 0; i--) {
            rightBox.grow(r[i].bounds);
            t.rightBounds[i - 1] = rightBox;
        }
        AABB leftBox;  // prefix bounds, candidate i puts r[0..i) on the left
        for (int i = 1; i < n; i++) {
            leftBox.grow(r[i - 1].bounds);
            const F32 sah = nodeSAH + leftBox.area() * m_platform.getTriangleCost(i) +
                            t.rightBounds[i - 1].area() * m_platform.getTriangleCost(n - i);
            const F32 tieBreak = sqr((F32)i) + sqr((F32)(n - i));
            if (sah < split.sah || (sah == split.sah && tieBreak < bestTieBreak)) {
                split.sah = sah;
                split.sortDim = dim;
                split.numLeft = i;
                split.leftBounds = leftBox;
                split.rightBounds = t.rightBounds[i - 1];
                bestTieBreak = tieBreak;
            }
        }
    }
    return split;
}

void SplitBVHBuilder::performObjectSplit(Task& t, NodeSpec& left, NodeSpec& right, const NodeSpec& spec,
                                         const ObjectSplit& split) {
    if (t.sortedDim != split.sortDim) sortTop(t, spec.numRef, split.sortDim);
    left.numRef = split.numLeft;
    left.bounds = split.leftBounds;
    right.numRef = spec.numRef - split.numLeft;
    right.bounds = split.rightBounds;
}

SplitBVHBuilder::SpatialSplit SplitBVHBuilder::findSpatialSplit(Task& t, const NodeSpec& spec, F32 nodeSAH) {
    const float3 origin = spec.bounds.minf();
    const float3 binSize = (spec.bounds.maxf() - origin) * (1.0f / (F32)NumSpatialBins);
    const float3 invBinSize = 1.0f / binSize;
    SpatialBin* bins = t.bins.get();
    auto bin = [bins](int dim, int i) -> SpatialBin& { return bins[dim * NumSpatialBins + i]; };

    for (int i = 0; i < 3 * NumSpatialBins; i++) {
        bins[i].bounds = AABB();
        bins[i].enter = 0;
        bins[i].exit = 0;
    }

    // Chop every reference into the bins it overlaps, per axis.
    const size_t end = t.refs.size();
    for (size_t refIdx = end - spec.numRef; refIdx < end; refIdx++) {
        const Reference& ref = t.refs[refIdx];
        const int3 firstBin = clamp(make_int3((ref.bounds.minf() - origin) * invBinSize), 0, NumSpatialBins - 1);
        const int3 lastBin = clamp(make_int3((ref.bounds.maxf() - origin) * invBinSize), firstBin, NumSpatialBins - 1);
        for (int dim = 0; dim < 3; dim++) {
            Reference curr = ref;
            for (int i = firstBin.m[dim]; i < lastBin.m[dim]; i++) {
                Reference l, r;
                splitReference(l, r, curr, dim, origin.m[dim] + binSize.m[dim] * (F32)(i + 1));
                bin(dim, i).bounds.grow(l.bounds);
                curr = r;
            }
            bin(dim, lastBin.m[dim]).bounds.grow(curr.bounds);
            bin(dim, firstBin.m[dim]).enter++;
            bin(dim, lastBin.m[dim]).exit++;
        }
    }

    // Pick the cheapest of the 127 planes per axis.
    SpatialSplit split;
    for (int dim = 0; dim < 3; dim++) {
        AABB rightBox;
        for (int i = NumSpatialBins - 1; i > 0; i--) {
            rightBox.grow(bin(dim, i).bounds);
            t.rightBounds[i - 1] = rightBox;
        }
        AABB leftBox;
        int leftNum = 0, rightNum = spec.numRef;
        for (int i = 1; i < NumSpatialBins; i++) {
            leftBox.grow(bin(dim, i - 1).bounds);
            leftNum += bin(dim, i - 1).enter;
            rightNum -= bin(dim, i - 1).exit;
            const F32 sah = nodeSAH + leftBox.area() * m_platform.getTriangleCost(leftNum) +
                            t.rightBounds[i - 1].area() * m_platform.getTriangleCost(rightNum);
            if (sah < split.sah) {
                split.sah = sah;
                split.dim = dim;
                split.pos = origin.m[dim] + binSize.m[dim] * (F32)i;
            }
        }
    }
    return split;
}

// In-place three-way partition of the node's range into [left | straddling | right], then each
// straddler is unsplit to one side or duplicated, whichever gives the lowest SAH.
void SplitBVHBuilder::performSpatialSplit(Task& t, NodeSpec& left, NodeSpec& right, const NodeSpec& spec,
                                          const SpatialSplit& split) {
    std::vector<Reference>& refs = t.refs;
    const int leftStart = (int)refs.size() - spec.numRef;
    int leftEnd = leftStart;
    int rightStart = (int)refs.size();
    left.bounds = right.bounds = AABB();
    t.sortedDim = -1;

    for (int i = leftEnd; i < rightStart; i++) {
        if (refs[i].bounds.maxf().m[split.dim] <= split.pos) {  // entirely left
            left.bounds.grow(refs[i].bounds);
            swap1(refs[i], refs[leftEnd++]);
        } else if (refs[i].bounds.minf().m[split.dim] >= split.pos) {  // entirely right
            right.bounds.grow(refs[i].bounds);
            swap1(refs[i], refs[--rightStart]);
            i--;  // re-examine what was swapped in
        }
    }

    while (leftEnd < rightStart) {
        Reference lref, rref;
        splitReference(lref, rref, refs[leftEnd], split.dim, split.pos);

        AABB lub = left.bounds;   // unsplit to the left
        AABB rub = right.bounds;  // unsplit to the right
        AABB ldb = left.bounds;   // duplicate, left half
        AABB rdb = right.bounds;  // duplicate, right half
        lub.grow(refs[leftEnd].bounds);
        rub.grow(refs[leftEnd].bounds);
        ldb.grow(lref.bounds);
        rdb.grow(rref.bounds);

        const F32 lac = m_platform.getTriangleCost(leftEnd - leftStart);
        const F32 rac = m_platform.getTriangleCost((int)refs.size() - rightStart);
        const F32 lbc = m_platform.getTriangleCost(leftEnd - leftStart + 1);
        const F32 rbc = m_platform.getTriangleCost((int)refs.size() - rightStart + 1);

        const F32 unsplitLeftSAH = lub.area() * lbc + right.bounds.area() * rac;
        const F32 unsplitRightSAH = left.bounds.area() * lac + rub.area() * rbc;
        const F32 duplicateSAH = ldb.area() * lbc + rdb.area() * rbc;
        const F32 minSAH = fminf1(unsplitLeftSAH, unsplitRightSAH, duplicateSAH);

        if (minSAH == unsplitLeftSAH) {
            left.bounds = lub;
            leftEnd++;
        } else if (minSAH == unsplitRightSAH) {
            right.bounds = rub;
            std::swap(refs[leftEnd], refs[--rightStart]);
        } else {
            left.bounds = ldb;
            right.bounds = rdb;
            refs[leftEnd++] = lref;
            refs.push_back(rref);
        }
    }
    left.numRef = leftEnd - leftStart;
    right.numRef = (int)refs.size() - rightStart;
}

// Clip a reference's triangle against the plane x[dim] = pos: vertices go to the side(s) they lie
// on, edge/plane intersections to both; each half is then clamped to the plane and intersected
// with the original reference bounds.
void SplitBVHBuilder::splitReference(Reference& left, Reference& right, const Reference& ref, int dim, F32 pos) const {
    left.triIdx = right.triIdx = ref.triIdx;
    left.bounds = right.bounds = AABB();

    const int3& inds = m_tris[ref.triIdx];
    const float4* v1 = &m_verts[inds.z];
    for (int i = 0; i < 3; i++) {
        const float4* v0 = v1;
        v1 = &m_verts[inds.m[i]];
        const F32 v0p = v0->get(dim);
        const F32 v1p = v1->get(dim);

        if (v0p <= pos) left.bounds.grow(make_float3(v0->x, v0->y, v0->z));
        if (v0p >= pos) right.bounds.grow(make_float3(v0->x, v0->y, v0->z));

        if ((v0p < pos && v1p > pos) || (v0p > pos && v1p < pos)) {
            const float4 p = lerp(*v0, *v1, clamp((pos - v0p) / (v1p - v0p), 0.0f, 1.0f));
            left.bounds.grow(make_float3(p.x, p.y, p.z));
            right.bounds.grow(make_float3(p.x, p.y, p.z));
        }
    }
    left.bounds.maxf().m[dim] = pos;
    right.bounds.minf().m[dim] = pos;
    left.bounds.intersect(ref.bounds);
    right.bounds.intersect(ref.bounds);
}

}  // namespace FW
