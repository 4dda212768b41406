#include "BVH2.h"

#include <chrono>

#include "SplitBVHBuilder.h"

namespace FW {

BVH2::BVH2(Mesh* mesh) : m_scene(nullptr), m_root(nullptr), m_numDuplicates(0), m_buildSeconds(0.0) {
    m_platform.setLeafPreferences(1, 8);  // leaf sizes 1..8 (reference BVH2.cpp:13)
    if (mesh) setMesh(mesh);
}

BVH2::~BVH2() { clear(); }

void BVH2::clear() {
    if (m_root) m_root->deleteSubtree();
    m_root = nullptr;
    m_triIndices.clear();
}

void BVH2::setMesh(Mesh* mesh) {
    clear();
    m_scene = mesh;
    if (!mesh) return;
    const auto t0 = std::chrono::steady_clock::now();
    SplitBVHBuilder builder(*this);
    m_root = builder.run();
    m_numDuplicates = builder.numDuplicates();
    m_buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void BVHNode::deleteSubtree() {
    std::vector<BVHNode*> pending(1, this);
    while (!pending.empty()) {
        BVHNode* n = pending.back();
        pending.pop_back();
        for (int i = 0; i < n->getNumChildNodes(); i++) pending.push_back(n->getChildNode(i));
        delete n;
    }
}

}  // namespace FW
