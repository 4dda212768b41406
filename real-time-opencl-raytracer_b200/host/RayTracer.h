// RayTracer.h -- the frame driver: what the reference keeps in global functions of RayTracer.cpp,
// gathered into the (there empty) `class RayTracer` (reference RayTracer.h:19-24).
//
//   reference (RayTracer.cpp)                              here
//   setupCL()            :2370  context/queue/kernel        RayTracer::setupCL(device, n)  -> rt_create / rt_create_group
//   initRayTrace()       :858   load scene, build SBVH,     RayTracer::initRayTrace(...)   -> SplitBVHBuilder,
//                                flatten, 11 clCreateBuffer                                    BVH_Cuda, rt_upload_scene
//   updateCamera()       :609   Camera -> Params, write     RayTracer::updateCamera()      -> rt_set_params
//   raytrace_gpgpu()     :330   NDRange + finish + readback RayTracer::raytrace_gpgpu()    -> rt_render_frame(_tiled)
//   cleanup()            :1263                              RayTracer::cleanup()           -> rt_destroy
//
// Return convention as in the reference: SDK_SUCCESS (0) or SDK_FAILURE (1); the message is in last_error().
// All device work goes through the C ABI of include/rtb200.h (librtb200.so); no OpenCL, no CPU fallback.
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "BVH_Cuda.h"
#include "Camera.h"
#include "Mesh.h"

struct rt_context;
struct rt_group;

#ifndef SDK_SUCCESS
#define SDK_SUCCESS 0
#define SDK_FAILURE 1
#endif

class RayTracer {
public:
    // the reference's globals (RayTracer.cpp:41-90)
    Mesh mesh1;
    FW::BVH2 bvh2;
    BVH_Cuda bvh_cuda;
    Camera cam;
    float3 light_pos = float3(-23.0f, 200.0f, 3.0f);
    float3 light_color1 = float3(1.0f, 1.0f, 1.0f);
    int image_width = 1024, image_height = 768;  // WIDTH / HEIGHT (RayTracer.cpp:39-40)
    bool animate = false;
    float delta_t = 0.0f;
    std::vector<unsigned int> out_data;  // w*h pixels, (b<<16)|(g<<8)|r, row 0 = bottom row as displayed
    float params[32];                    // the 128-byte Params block of the last updateCamera()

    RayTracer();
    ~RayTracer();
    RayTracer(const RayTracer&) = delete;
    RayTracer& operator=(const RayTracer&) = delete;

    // n_gpus > 1: the frame is split in row bands over GPUs 0..n_gpus-1 of this box (rt_create_group: the scene is
    // broadcast with NCCL over NVLink, every GPU stores its bands straight into out_data). The reference takes
    // devices[0] only (RayTracer.cpp:2128-2131).
    int setupCL(int device_ordinal = 0, int n_gpus = 1);
    int gpus() const { return n_gpus_; }
    // per-GPU device time of the last frame (ms) and the scene broadcast time, n_gpus > 1 only
    std::vector<double> last_rank_ms() const;
    double broadcast_ms() const;
    // scene from a COLLADA file (the reference loads data/collada/cubes2.DAE) ...
    int initRayTrace(const char* collada_path);
    // ... or from a Mesh the caller filled (synthetic scenes); `flat_bvh_cache` (optional) is read if present,
    // written otherwise (the SBVH build is the slow step)
    int initRayTraceFromMesh(const char* flat_bvh_cache = nullptr);
    int updateCamera();
    int raytrace_gpgpu();
    // The same frame without the reference's per-frame stall (clFinish + blocking read, RayTracer.cpp:341-343):
    // _begin enqueues the frame of the last updateCamera() and returns, _end waits for the oldest frame in flight and
    // leaves it in out_data. Two frames may be in flight, so the host updates the camera / consumes frame k while the
    // device traces frame k+1.
    int raytrace_gpgpu_begin();
    int raytrace_gpgpu_end();
    int frames_in_flight() const { return inflight_; }
    int cleanup();

    bool write_ppm(const char* path) const;  // headless replacement of the GLUT texture blit
    const std::string& last_error() const { return err_; }
    rt_context* context() const { return ctx_; }
    double last_build_seconds() const { return build_seconds_; }

private:
    int upload();
    bool size_and_pin(std::vector<unsigned int>& v);  // w*h pixels, page-locked so the kernel stores into it directly
    void unpin(std::vector<unsigned int>& v);
    std::vector<std::pair<void*, size_t>> pinned_;
    std::vector<unsigned int> ring_[2];
    int head_ = 0, inflight_ = 0;
    rt_context* ctx_ = nullptr;
    rt_group* group_ = nullptr;   // n_gpus > 1; ctx_ is then GPU 0's context (borrowed from the group)
    int n_gpus_ = 1;
    std::string err_;
    double build_seconds_ = 0.0;
};
