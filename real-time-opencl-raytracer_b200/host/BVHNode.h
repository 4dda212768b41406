// BVHNode.h -- pointer-tree node types produced by the SBVH builder.
//
// Public surface follows the reference (reference BVHNode.h:9-54, Util.h:11-33): FW::AABB with
// minf()/maxf()/grow/intersect/area, FW::BVHNode with isLeaf/getChildNode/getNumTriangles/m_bounds,
// InnerNode::m_children and LeafNode::m_lo/m_hi (a [lo,hi) range into BVH2::m_triIndices).
//
// FW::AABB numeric behaviour that the builder's output depends on (kept on purpose):
//   * grow(point) updates BOTH corners with the plain a<b?a:b selects
//   * grow(box) = grow(box.min) then grow(box.max) -- growing by an EMPTY box therefore blows the
//     box up to [-FLT_MAX, +FLT_MAX] (reference Util.h:18-19); empty spatial bins rely on that
//   * area() of an invalid box is 0, otherwise (dx*dy + dy*dz + dz*dx) * 2
#pragma once
#include "vecmath.h"

namespace FW {

typedef int S32;
typedef float F32;
static const float kF32Max = 3.402823466e+38f;

class AABB {
public:
    AABB() : mn_(kF32Max, kF32Max, kF32Max), mx_(-kF32Max, -kF32Max, -kF32Max) {}
    AABB(const float3& mn, const float3& mx) : mn_(mn), mx_(mx) {}

    void grow(const float3& p) { mn_ = fminf1(mn_, p); mx_ = fmaxf1(mx_, p); }
    void grow(const AABB& b) { grow(b.mn_); grow(b.mx_); }
    void intersect(const AABB& b) { mn_ = fmaxf1(mn_, b.mn_); mx_ = fminf1(mx_, b.mx_); }
    bool valid() const { return mn_.x <= mx_.x && mn_.y <= mx_.y && mn_.z <= mx_.z; }
    float volume() const {
        if (!valid()) return 0.0f;
        return (mx_.x - mn_.x) * (mx_.y - mn_.y) * (mx_.z - mn_.z);
    }
    float area() const {
        if (!valid()) return 0.0f;
        float3 d = mx_ - mn_;
        return (d.x * d.y + d.y * d.z + d.z * d.x) * 2.0f;
    }
    float3 midPoint() const { return (mn_ + mx_) * 0.5f; }
    const float3& minf() const { return mn_; }
    const float3& maxf() const { return mx_; }
    float3& minf() { return mn_; }
    float3& maxf() { return mx_; }

private:
    float3 mn_, mx_;
};

class BVHNode {
public:
    virtual ~BVHNode() {}
    virtual bool isLeaf() const = 0;
    virtual S32 getNumChildNodes() const = 0;
    virtual BVHNode* getChildNode(S32 i) const = 0;
    virtual S32 getNumTriangles() const { return 0; }
    float getArea() const { return m_bounds.area(); }
    void deleteSubtree();  // iterative: safe for the 64-deep trees the builder can emit

    AABB m_bounds;
};

class InnerNode : public BVHNode {
public:
    InnerNode(const AABB& bounds, BVHNode* child0, BVHNode* child1) {
        m_bounds = bounds;
        m_children[0] = child0;
        m_children[1] = child1;
    }
    bool isLeaf() const override { return false; }
    S32 getNumChildNodes() const override { return 2; }
    BVHNode* getChildNode(S32 i) const override { return (i == 0 || i == 1) ? m_children[i] : nullptr; }

    BVHNode* m_children[2];
};

class LeafNode : public BVHNode {
public:
    LeafNode(const AABB& bounds, int lo, int hi) : m_lo(lo), m_hi(hi) { m_bounds = bounds; }
    bool isLeaf() const override { return true; }
    S32 getNumChildNodes() const override { return 0; }
    BVHNode* getChildNode(S32) const override { return nullptr; }
    S32 getNumTriangles() const override { return m_hi - m_lo; }

    S32 m_lo, m_hi;
};

}  // namespace FW
