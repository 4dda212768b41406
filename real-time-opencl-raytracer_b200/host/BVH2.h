// BVH2.h -- owner of the SBVH pointer tree + the leaf triangle-index list.
//
// Public surface follows the reference's FW::BVH2 (reference BVH2.h:11-37, BVH2.cpp:11-45):
// setMesh() runs SplitBVHBuilder with leaf preferences 1..8, getRoot() / getTriIndices() expose
// the result that BVH_Cuda flattens. Fixed relative to the reference (outputs unchanged,
// SURVEY.md Appendix C): the constructor honours its `mesh` argument instead of reading an
// uninitialised member, and m_triIndices is cleared on every setMesh().
#pragma once
#include <vector>

#include "BVHNode.h"
#include "Mesh.h"
#include "Platform.h"

namespace FW {

class BVH2 {
public:
    explicit BVH2(Mesh* mesh = nullptr);
    ~BVH2();
    BVH2(const BVH2&) = delete;
    BVH2& operator=(const BVH2&) = delete;

    void setMesh(Mesh* mesh);
    void clear();

    Mesh* getScene() const { return m_scene; }
    const Platform& getPlatform() const { return m_platform; }
    const Platform& getPlatform1() const { return m_platform; }  // the reference's spelling
    BVHNode* getRoot() const { return m_root; }
    std::vector<S32>& getTriIndices() { return m_triIndices; }
    const std::vector<S32>& getTriIndices() const { return m_triIndices; }

    // build statistics of the last setMesh() (additions)
    int numDuplicates() const { return m_numDuplicates; }
    double buildSeconds() const { return m_buildSeconds; }

    Mesh* m_scene;
    Platform m_platform;
    BVHNode* m_root;
    std::vector<S32> m_triIndices;

private:
    int m_numDuplicates;
    double m_buildSeconds;
};

}  // namespace FW
