// Platform.h -- SAH cost model handed to the SBVH builder.
//
// Same public surface as the reference's FW::Platform (reference Platform.h:17-47): node and
// triangle costs (1.0 / 1.0 by default), batch rounding, and the min/max leaf-size preference
// that FW::BVH2 sets to 1..8 (reference BVH2.cpp:13).
#pragma once

namespace FW {

typedef int S32;
typedef unsigned int U32;
typedef float F32;

class Platform {
public:
    explicit Platform(float nodeCost = 1.0f, float triCost = 1.0f, S32 nodeBatch = 1, S32 triBatch = 1)
        : node_cost_(nodeCost), tri_cost_(triCost), node_batch_(nodeBatch), tri_batch_(triBatch),
          min_leaf_(1), max_leaf_(0x7FFFFFF) {}

    float getSAHTriangleCost() const { return tri_cost_; }
    float getSAHNodeCost() const { return node_cost_; }

    S32 getTriangleBatchSize() const { return tri_batch_; }
    S32 getNodeBatchSize() const { return node_batch_; }
    void setTriangleBatchSize(S32 n) { tri_batch_ = n; }
    void setNodeBatchSize(S32 n) { node_batch_ = n; }
    S32 roundToTriangleBatchSize(S32 n) const { return round_up(n, tri_batch_); }
    S32 roundToNodeBatchSize(S32 n) const { return round_up(n, node_batch_); }

    // cost(n) = (n rounded up to the batch size, as int) * weight  -- an int*float product
    float getTriangleCost(S32 n) const { return roundToTriangleBatchSize(n) * tri_cost_; }
    float getNodeCost(S32 n) const { return roundToNodeBatchSize(n) * node_cost_; }
    float getCost(int numChildNodes, int numTris) const { return getNodeCost(numChildNodes) + getTriangleCost(numTris); }

    void setLeafPreferences(S32 minSize, S32 maxSize) { min_leaf_ = minSize; max_leaf_ = maxSize; }
    S32 getMinLeafSize() const { return min_leaf_; }
    S32 getMaxLeafSize() const { return max_leaf_; }

private:
    static S32 round_up(S32 n, S32 batch) { return ((n + batch - 1) / batch) * batch; }
    float node_cost_, tri_cost_;
    S32 node_batch_, tri_batch_;
    S32 min_leaf_, max_leaf_;
};

}  // namespace FW
