// MSVC-only header named by the reference (Camera.cpp:3, ColladaLoader.cpp:7).
#include <cmath>
