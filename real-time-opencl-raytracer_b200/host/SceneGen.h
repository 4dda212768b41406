// SceneGen.h -- deterministic synthetic scenes for the benchmark / parity configurations
// (BASELINE.json configs). Not part of the reference: the reference only
// loads one COLLADA file (reference RayTracer.cpp:862); its two hand-built micro-scenes
// (reference RayTracer.cpp:825-856) are reproduced here as fixtures.
#pragma once
#include "Mesh.h"

namespace scenegen {

// Icosphere: 20 * 4^subdiv triangles with shared vertices (subdiv 6 -> 81 920 tris / 40 962 verts).
void add_icosphere(Mesh& m, int subdiv, float radius, float cx, float cy, float cz);

// Displaced grid over x,z in [-half, half]: nquads^2 quads, 2 triangles each, winding (a,c,b),(b,c,d);
// y = 6 sin(0.15x) cos(0.11z) + 1.5 sin(0.9x + 0.4z). nquads = 707 -> 999 698 tris / 501 264 verts.
void add_terrain(Mesh& m, int nquads, float half);

// C4: `grid` x `grid` icospheres at pitch `pitch` on y = 0 plus one at (0, 150, 0), all baked
// into one mesh (the reference has no instancing). grid=11, subdiv=6 -> 122 * 81 920 = 9 994 240 tris.
void add_sphere_field(Mesh& m, int grid, float pitch, int subdiv, float radius);

// Long thin crossing triangles: forces spatial splits / duplicated references in the SBVH.
void add_sticks(Mesh& m, int count, unsigned seed, float extent);

// The reference's hand-built builder smoke scenes (reference RayTracer.cpp:825-856).
void add_bvh_test0(Mesh& m);
void add_bvh_test1(Mesh& m);

// Write the mesh as a COLLADA 1.4 <polygons> document in the dialect ColladaLoader accepts
// (SURVEY.md Appendix B.4): one effect, one geometry, one node with an identity <matrix>.
bool write_dae(const Mesh& m, const char* path);

}  // namespace scenegen
