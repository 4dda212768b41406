// rt_demo.cpp -- headless version of the reference's main() (RayTracer.cpp:574-606): set up the device,
// load / generate a scene, build the SBVH, render `frames` frames of the orbit animation, write a PPM.
//   rt_demo [scene.dae | terrain:<quads> | spheres:<subdiv>] [w h] [frames] [out.ppm]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "RayTracer.h"
#include "SceneGen.h"

int main(int argc, char** argv) {
    const char* scene = argc > 1 ? argv[1] : "terrain:200";
    RayTracer rt;
    if (argc > 3) {
        rt.image_width = atoi(argv[2]);
        rt.image_height = atoi(argv[3]);
    }
    const int frames = argc > 4 ? atoi(argv[4]) : 10;
    const char* out = argc > 5 ? argv[5] : "frame.ppm";

    if (rt.setupCL(0) != SDK_SUCCESS) {
        fprintf(stderr, "setupCL failed: %s\n", rt.last_error().c_str());
        return 1;
    }
    int rc;
    if (!strncmp(scene, "terrain:", 8)) {
        scenegen::add_terrain(rt.mesh1, atoi(scene + 8), 100.0f);
        rc = rt.initRayTraceFromMesh();
    } else if (!strncmp(scene, "spheres:", 8)) {
        scenegen::add_icosphere(rt.mesh1, atoi(scene + 8), 50.0f, 0.0f, 0.0f, 0.0f);
        rc = rt.initRayTraceFromMesh();
    } else {
        rc = rt.initRayTrace(scene);
    }
    if (rc != SDK_SUCCESS) {
        fprintf(stderr, "initRayTrace failed: %s\n", rt.last_error().c_str());
        return 1;
    }
    printf("scene: %d triangles, %zu BVH nodes, SBVH build %.2f s\n", rt.mesh1.getNumTriangles(), rt.bvh_cuda.bvh_nodes.size(),
           rt.last_build_seconds());
    rt.animate = true;
    rt.delta_t = 0.05f;
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; f++) {
        if (rt.updateCamera() != SDK_SUCCESS || rt.raytrace_gpgpu() != SDK_SUCCESS) {
            fprintf(stderr, "frame %d failed: %s\n", f, rt.last_error().c_str());
            return 1;
        }
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%d frames %dx%d in %.3f s -> %.1f fps\n", frames, rt.image_width, rt.image_height, s, frames / s);
    if (!rt.write_ppm(out)) fprintf(stderr, "cannot write %s\n", out);
    rt.cleanup();
    return 0;
}
