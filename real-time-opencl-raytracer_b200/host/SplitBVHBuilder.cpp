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
    SpatialSplit best;
    for (int dim = 0; dim < 3; dim++) {
        AABB suffix;  // bounds of bins [plane, NumSpatialBins), built back to front
        for (int plane = NumSpatialBins - 1; plane >= 1; plane--) {
            suffix.grow(bin(dim, plane).bounds);
            t.rightBounds[plane - 1] = suffix;
        }
        AABB prefix;  // bounds of bins [0, plane)
        int startedLeft = 0, notYetEnded = spec.numRef;
        for (int plane = 1; plane < NumSpatialBins; plane++) {
            const SpatialBin& passed = bin(dim, plane - 1);
            prefix.grow(passed.bounds);
            startedLeft += passed.enter;
            notYetEnded -= passed.exit;
            const F32 sah = nodeSAH + prefix.area() * m_platform.getTriangleCost(startedLeft) +
                            t.rightBounds[plane - 1].area() * m_platform.getTriangleCost(notYetEnded);
            if (sah < best.sah) {
                best.sah = sah;
                best.dim = dim;
                best.pos = origin.m[dim] + binSize.m[dim] * (F32)plane;
            }
        }
    }
    return best;
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

    const int axis = split.dim;
    for (int scan = leftEnd; scan < rightStart;) {
        const AABB& b = refs[scan].bounds;
        if (b.maxf().m[axis] <= split.pos) {  // wholly left of the plane: move to the growing left block
            left.bounds.grow(b);
            swap1(refs[scan], refs[leftEnd]);
            ++leftEnd;
            ++scan;
        } else if (b.minf().m[axis] >= split.pos) {  // wholly right: swap with the last unclassified one, look at it next
            right.bounds.grow(b);
            --rightStart;
            swap1(refs[scan], refs[rightStart]);
        } else {
            ++scan;  // straddles the plane: decided below
        }
    }

    // Straddlers sit in [leftEnd, rightStart). For each, three candidates are costed with the surface-area heuristic:
    // keep it whole on the left, keep it whole on the right, or clip it and reference it from both sides.
    while (leftEnd < rightStart) {
        const Reference& straddler = refs[leftEnd];
        Reference clippedL, clippedR;
        splitReference(clippedL, clippedR, straddler, split.dim, split.pos);

        AABB leftIfWhole = left.bounds, rightIfWhole = right.bounds;  // bounds if the whole reference joins that side
        AABB leftIfClipped = left.bounds, rightIfClipped = right.bounds;
        leftIfWhole.grow(straddler.bounds);
        rightIfWhole.grow(straddler.bounds);
        leftIfClipped.grow(clippedL.bounds);
        rightIfClipped.grow(clippedR.bounds);

        const int nLeft = leftEnd - leftStart, nRight = (int)refs.size() - rightStart;
        const F32 costLeftNow = m_platform.getTriangleCost(nLeft), costRightNow = m_platform.getTriangleCost(nRight);
        const F32 costLeftPlus = m_platform.getTriangleCost(nLeft + 1), costRightPlus = m_platform.getTriangleCost(nRight + 1);

        const F32 sahWholeLeft = leftIfWhole.area() * costLeftPlus + right.bounds.area() * costRightNow;
        const F32 sahWholeRight = left.bounds.area() * costLeftNow + rightIfWhole.area() * costRightPlus;
        const F32 sahBoth = leftIfClipped.area() * costLeftPlus + rightIfClipped.area() * costRightPlus;
        const F32 best = fminf1(sahWholeLeft, sahWholeRight, sahBoth);

        if (best == sahWholeLeft) {  // checked in this order: whole-left wins ties, then whole-right
            left.bounds = leftIfWhole;
            ++leftEnd;
        } else if (best == sahWholeRight) {
            right.bounds = rightIfWhole;
            --rightStart;
            std::swap(refs[leftEnd], refs[rightStart]);
        } else {
            left.bounds = leftIfClipped;
            right.bounds = rightIfClipped;
            refs[leftEnd] = clippedL;
            ++leftEnd;
            refs.push_back(clippedR);  // appended references belong to the right side
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

    const int3& corner = m_tris[ref.triIdx];
    // edges (c -> a), (a -> b), (b -> c): each start vertex is sorted into the side(s) it lies on, each edge that
    // crosses the plane contributes its intersection point to both sides
    const int order[4] = {corner.z, corner.x, corner.y, corner.z};
    for (int e = 0; e < 3; e++) {
        const float4& from = m_verts[order[e]];
        const float4& to = m_verts[order[e + 1]];
        const F32 fromP = from.get(dim), toP = to.get(dim);
        const float3 fromPoint = make_float3(from.x, from.y, from.z);
        if (fromP <= pos) left.bounds.grow(fromPoint);
        if (fromP >= pos) right.bounds.grow(fromPoint);
        const bool crosses = (fromP < pos && toP > pos) || (fromP > pos && toP < pos);
        if (crosses) {
            const float4 hit = lerp(from, to, clamp((pos - fromP) / (toP - fromP), 0.0f, 1.0f));
            const float3 hitPoint = make_float3(hit.x, hit.y, hit.z);
            left.bounds.grow(hitPoint);
            right.bounds.grow(hitPoint);
        }
    }
    left.bounds.maxf().m[dim] = pos;
    right.bounds.minf().m[dim] = pos;
    left.bounds.intersect(ref.bounds);
    right.bounds.intersect(ref.bounds);
}

}  // namespace FW
