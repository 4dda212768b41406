// rt_demo.cpp -- headless version of the reference's main() (RayTracer.cpp:574-606): set up the device,
// load / generate a scene, build the SBVH, render `frames` frames of the orbit animation, write a PPM.
//   rt_demo [scene.dae | terrain:<quads> | spheres:<subdiv>] [w h] [frames] [out.ppm] [--gpus N]
// The loop runs twice: frame by frame as the reference does (update camera, trace, wait, read back), then pipelined
// with two frames in flight (RayTracer::raytrace_gpgpu_begin/_end).
// --gpus N (N > 1): the first loop splits every frame over N GPUs (RayTracer::setupCL(0, N): NCCL scene broadcast, row
// bands, every GPU storing into the same page-locked frame); the second loop renders the same orbit on GPU 0 alone
// through the plain single-context call -- the two checksums must be identical.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/rtb200.h"
#include "RayTracer.h"
#include "SceneGen.h"

// the host-side consumer of a finished frame (stands in for the reference's texture upload + blit, RayTracer.cpp:436-470)
static unsigned long long consume(const std::vector<unsigned int>& frame, unsigned long long acc) {
    unsigned long long s = 0;
    for (unsigned int p : frame) s += p;
    return acc * 1099511628211ull + s;
}

int main(int argc, char** argv) {
    int n_gpus = 1;
    for (int i = 1; i + 1 < argc; i++)
        if (!strcmp(argv[i], "--gpus")) {  // strip the flag: the positional arguments keep their places
            n_gpus = atoi(argv[i + 1]);
            for (int j = i; j + 2 < argc; j++) argv[j] = argv[j + 2];
            argc -= 2;
            break;
        }
    const char* scene = argc > 1 ? argv[1] : "terrain:200";
    RayTracer rt;
    if (argc > 3) {
        rt.image_width = atoi(argv[2]);
        rt.image_height = atoi(argv[3]);
    }
    const int frames = argc > 4 ? atoi(argv[4]) : 10;
    const char* out = argc > 5 ? argv[5] : "frame.ppm";

    if (rt.setupCL(0, n_gpus) != SDK_SUCCESS) {
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
    unsigned long long sum_a = 0, sum_b = 0;
    const Camera start = rt.cam;  // both loops fly the same orbit
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; f++) {
        if (rt.updateCamera() != SDK_SUCCESS || rt.raytrace_gpgpu() != SDK_SUCCESS) {
            fprintf(stderr, "frame %d failed: %s\n", f, rt.last_error().c_str());
            return 1;
        }
        sum_a = consume(rt.out_data, sum_a);
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%d frames %dx%d in %.3f s -> %.1f fps (frame by frame, %d GPU%s)\n", frames, rt.image_width, rt.image_height, s, frames / s,
           rt.gpus(), rt.gpus() > 1 ? "s" : "");
    rt.cam = start;
    if (rt.gpus() > 1) {
        printf("scene broadcast (ncclBroadcast, device time): %.3f ms; last frame, device ms per GPU:", rt.broadcast_ms());
        for (double ms : rt.last_rank_ms()) printf(" %.3f", ms);
        printf("\n");
        // the same orbit on GPU 0 alone, through the single-context entry point
        std::vector<unsigned int> one((size_t)rt.image_width * rt.image_height);
        const auto t2 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; f++) {
            if (rt.updateCamera() != SDK_SUCCESS || rt_render_frame(rt.context(), rt.image_width, rt.image_height, one.data()) != RT_OK) {
                fprintf(stderr, "single-GPU frame %d failed: %s\n", f, rt_last_error(rt.context()));
                return 1;
            }
            sum_b = consume(one, sum_b);
        }
        const double s2 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t2).count();
        printf("%d frames %dx%d in %.3f s -> %.1f fps (frame by frame, GPU 0 alone, pageable frame)\n", frames, rt.image_width, rt.image_height, s2, frames / s2);
        printf("frame checksums: %016llx (%d GPUs) / %016llx (1 GPU) %s\n", sum_a, rt.gpus(), sum_b, sum_a == sum_b ? "(identical)" : "(DIFFERENT)");
        if (sum_a != sum_b) return 2;
        if (!rt.write_ppm(out)) fprintf(stderr, "cannot write %s\n", out);
        rt.cleanup();
        return 0;
    }
    const auto t1 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; f++) {
        if (rt.updateCamera() != SDK_SUCCESS || rt.raytrace_gpgpu_begin() != SDK_SUCCESS) {
            fprintf(stderr, "pipelined frame %d failed: %s\n", f, rt.last_error().c_str());
            return 1;
        }
        if (rt.frames_in_flight() == 2) {  // frame f-1 is consumed while frame f traces
            if (rt.raytrace_gpgpu_end() != SDK_SUCCESS) {
                fprintf(stderr, "pipelined frame %d failed: %s\n", f - 1, rt.last_error().c_str());
                return 1;
            }
            sum_b = consume(rt.out_data, sum_b);
        }
    }
    while (rt.frames_in_flight() > 0) {
        if (rt.raytrace_gpgpu_end() != SDK_SUCCESS) {
            fprintf(stderr, "pipelined drain failed: %s\n", rt.last_error().c_str());
            return 1;
        }
        sum_b = consume(rt.out_data, sum_b);
    }
    const double s1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    printf("%d frames %dx%d in %.3f s -> %.1f fps (two frames in flight)\n", frames, rt.image_width, rt.image_height, s1, frames / s1);
    printf("frame checksums: %016llx / %016llx %s\n", sum_a, sum_b, sum_a == sum_b ? "(identical)" : "(DIFFERENT)");
    if (sum_a != sum_b) return 2;
    if (!rt.write_ppm(out)) fprintf(stderr, "cannot write %s\n", out);
    rt.cleanup();
    return 0;
}
