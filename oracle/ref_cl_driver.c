/* oracle/ref_cl_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Runs the reference's OWN, UNMODIFIED OpenCL program (x64/Release/volumeRender.cl, kernel "raytracer_bvh") through
 * the reference's own launch sequence, so that the CPU oracle and the CUDA path can be compared with what the
 * reference itself computes -- on whatever OpenCL device the box offers. The GPU boxes of this pool ship NVIDIA's
 * OpenCL ICD (libnvidia-opencl.so.1) without an /etc/OpenCL/vendors entry; the ICD loader (libOpenCL.so.1, part of
 * the CUDA toolkit) picks it up through OCL_ICD_FILENAMES, which refcl_init sets if nothing else is configured.
 *
 * What this file restates from the reference host (RayTracer.cpp): context/queue/program/kernel creation
 * (:2050-2433, first platform, first GPU device, clBuildProgram without options), the buffer set of initRayTrace
 * (:942-984), the 17 clSetKernelArg slots (:1228-1260), the 2-D NDRange with 8x8 work-groups rounded up to the
 * frame (:332-336), clFinish and the blocking read of w*h uint32 pixels (:340-343).
 *
 * The kernel source is NOT copied into the repository: it is pulled into this object at build time with .incbin
 * from REF_CL_PATH (= /root/reference/x64/Release/volumeRender.cl, see oracle/Makefile) and handed to
 * clCreateProgramWithSource verbatim; the only file produced is oracle/_ref/libref_cl.so (git-ignored).
 * No OpenCL headers exist in this image: the handful of API entry points and constants used are declared below
 * (values from the public OpenCL 1.2 specification) and resolved with dlsym.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef REF_CL_PATH
#error "build with -DREF_CL_PATH=\"/root/reference/x64/Release/volumeRender.cl\""
#endif
__asm__(".section .rodata\n"
        ".global refcl_source\nrefcl_source:\n"
        ".incbin \"" REF_CL_PATH "\"\n"
        ".global refcl_source_end\nrefcl_source_end:\n"
        ".byte 0\n"
        ".previous\n");
extern const char refcl_source[], refcl_source_end[];

/* ---- minimal OpenCL 1.2 API surface -------------------------------------------------------------- */
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_ulong cl_bitfield;
typedef struct _cl_platform_id* cl_platform_id;
typedef struct _cl_device_id* cl_device_id;
typedef struct _cl_context* cl_context;
typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_mem* cl_mem;
typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel;
typedef struct _cl_event* cl_event;
#define CL_SUCCESS 0
#define CL_TRUE 1
#define CL_DEVICE_TYPE_GPU (1 << 2)
#define CL_DEVICE_TYPE_CPU (1 << 1)
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROFILING_COMMAND_START 0x1282
#define CL_PROFILING_COMMAND_END 0x1283
#define CL_DEVICE_NAME 0x102B
#define CL_PLATFORM_NAME 0x0902

static struct {
    void* lib;
    cl_int (*GetPlatformIDs)(cl_uint, cl_platform_id*, cl_uint*);
    cl_int (*GetPlatformInfo)(cl_platform_id, cl_uint, size_t, void*, size_t*);
    cl_int (*GetDeviceIDs)(cl_platform_id, cl_bitfield, cl_uint, cl_device_id*, cl_uint*);
    cl_int (*GetDeviceInfo)(cl_device_id, cl_uint, size_t, void*, size_t*);
    cl_context (*CreateContext)(const intptr_t*, cl_uint, const cl_device_id*, void*, void*, cl_int*);
    cl_command_queue (*CreateCommandQueue)(cl_context, cl_device_id, cl_bitfield, cl_int*);
    cl_program (*CreateProgramWithSource)(cl_context, cl_uint, const char**, const size_t*, cl_int*);
    cl_int (*BuildProgram)(cl_program, cl_uint, const cl_device_id*, const char*, void*, void*);
    cl_int (*GetProgramBuildInfo)(cl_program, cl_device_id, cl_uint, size_t, void*, size_t*);
    cl_kernel (*CreateKernel)(cl_program, const char*, cl_int*);
    cl_mem (*CreateBuffer)(cl_context, cl_bitfield, size_t, void*, cl_int*);
    cl_int (*SetKernelArg)(cl_kernel, cl_uint, size_t, const void*);
    cl_int (*EnqueueNDRangeKernel)(cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint,
                                   const cl_event*, cl_event*);
    cl_int (*EnqueueReadBuffer)(cl_command_queue, cl_mem, cl_uint, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*);
    cl_int (*EnqueueWriteBuffer)(cl_command_queue, cl_mem, cl_uint, size_t, size_t, const void*, cl_uint, const cl_event*, cl_event*);
    cl_int (*Finish)(cl_command_queue);
    cl_int (*GetEventProfilingInfo)(cl_event, cl_uint, size_t, void*, size_t*);
    cl_int (*ReleaseEvent)(cl_event);
    cl_int (*ReleaseMemObject)(cl_mem);
    cl_int (*ReleaseKernel)(cl_kernel);
    cl_int (*ReleaseProgram)(cl_program);
    cl_int (*ReleaseCommandQueue)(cl_command_queue);
    cl_int (*ReleaseContext)(cl_context);
} cl;

static struct {
    cl_platform_id platform;
    cl_device_id device;
    cl_context context;
    cl_command_queue queue;
    cl_program program;
    cl_kernel kernel;
    /* RayTraceData (RayTracer.cpp:165-231) */
    cl_mem mesh_vertices, mesh_indices, bvh_nodes, bvh_tris_indices, mesh_normals, mesh_normals_indices, mesh_materials,
        mesh_triangle_index_to_material_index, temp, params, out;
    int num_bvh_tris, num_bvh_nodes, out_w, out_h;
    char device_name[256];
    char build_note[256];
    char error[4096];
} S;

static int fail(const char* what, cl_int code) {
    snprintf(S.error, sizeof S.error, "%s failed (%d)", what, (int)code);
    return 1;
}

#define LOAD(name)                                               \
    do {                                                         \
        *(void**)(&cl.name) = dlsym(cl.lib, "cl" #name);         \
        if (!cl.name) { snprintf(S.error, sizeof S.error, "libOpenCL has no cl" #name); return 1; } \
    } while (0)

const char* refcl_last_error(void) { return S.error; }
const char* refcl_device_name(void) { return S.device_name; }
const char* refcl_build_note(void) { return S.build_note; }
long refcl_source_bytes(void) { return (long)(refcl_source_end - refcl_source); }

/* build_options: NULL/"" = exactly the reference (clBuildProgram without options, RayTracer.cpp:2173) */
int refcl_init(const char* build_options) {
    memset(&S, 0, sizeof S);
    if (!getenv("OCL_ICD_FILENAMES") && !getenv("OCL_ICD_VENDORS")) {
        const char* cands[] = {"/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1",
                               "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1", NULL};
        for (int i = 0; cands[i]; i++) {
            FILE* f = fopen(cands[i], "rb");
            if (f) {
                fclose(f);
                setenv("OCL_ICD_FILENAMES", cands[i], 0);
                break;
            }
        }
    }
    cl.lib = dlopen("libOpenCL.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!cl.lib) cl.lib = dlopen("/usr/local/cuda/lib64/libOpenCL.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!cl.lib) { snprintf(S.error, sizeof S.error, "cannot load libOpenCL.so.1: %s", dlerror()); return 1; }
    LOAD(GetPlatformIDs); LOAD(GetPlatformInfo); LOAD(GetDeviceIDs); LOAD(GetDeviceInfo); LOAD(CreateContext);
    LOAD(CreateCommandQueue); LOAD(CreateProgramWithSource); LOAD(BuildProgram); LOAD(GetProgramBuildInfo); LOAD(CreateKernel);
    LOAD(CreateBuffer); LOAD(SetKernelArg); LOAD(EnqueueNDRangeKernel); LOAD(EnqueueReadBuffer); LOAD(EnqueueWriteBuffer);
    LOAD(Finish); LOAD(GetEventProfilingInfo); LOAD(ReleaseEvent); LOAD(ReleaseMemObject); LOAD(ReleaseKernel);
    LOAD(ReleaseProgram); LOAD(ReleaseCommandQueue); LOAD(ReleaseContext);

    cl_int err;
    cl_uint n = 0;
    err = cl.GetPlatformIDs(1, &S.platform, &n); /* first platform, RayTracer.cpp:2060 */
    if (err != CL_SUCCESS || n == 0) return fail("clGetPlatformIDs (no OpenCL platform)", err);
    err = cl.GetDeviceIDs(S.platform, CL_DEVICE_TYPE_GPU, 1, &S.device, &n); /* GPU, else CPU, RayTracer.cpp:2076-2088 */
    if (err != CL_SUCCESS || n == 0) err = cl.GetDeviceIDs(S.platform, CL_DEVICE_TYPE_CPU, 1, &S.device, &n);
    if (err != CL_SUCCESS || n == 0) return fail("clGetDeviceIDs", err);
    cl.GetDeviceInfo(S.device, CL_DEVICE_NAME, sizeof S.device_name, S.device_name, NULL);
    S.context = cl.CreateContext(NULL, 1, &S.device, NULL, NULL, &err);
    if (err != CL_SUCCESS) return fail("clCreateContext", err);
    S.queue = cl.CreateCommandQueue(S.context, S.device, CL_QUEUE_PROFILING_ENABLE, &err); /* profiling added for timing only */
    if (err != CL_SUCCESS) return fail("clCreateCommandQueue", err);
    const char* src = refcl_source;
    const size_t len = (size_t)(refcl_source_end - refcl_source);
    S.program = cl.CreateProgramWithSource(S.context, 1, &src, &len, &err);
    if (err != CL_SUCCESS) return fail("clCreateProgramWithSource", err);
    err = cl.BuildProgram(S.program, 1, &S.device, (build_options && build_options[0]) ? build_options : NULL, NULL, NULL);
    if (err != CL_SUCCESS) {
        /* NVIDIA's compiler rejects `__global __write_only uint *out_data` (volumeRender.cl:1043: "access qualifier can only
         * be used for pipe and image type"; the author's AMD compiler accepted it). The source stays untouched: the retry only
         * adds -D__write_only= on the command line (the qualifier occurs nowhere else in the file). */
        char opts[512];
        snprintf(opts, sizeof opts, "%s -D__write_only=", (build_options && build_options[0]) ? build_options : "");
        err = cl.BuildProgram(S.program, 1, &S.device, opts, NULL, NULL);
        if (err == CL_SUCCESS) snprintf(S.build_note, sizeof S.build_note, "built with the extra option -D__write_only=");
    }
    if (err != CL_SUCCESS) {
        int k = snprintf(S.error, sizeof S.error, "clBuildProgram failed (%d): ", (int)err);
        cl.GetProgramBuildInfo(S.program, S.device, CL_PROGRAM_BUILD_LOG, sizeof S.error - k - 1, S.error + k, NULL);
        return 1;
    }
    S.kernel = cl.CreateKernel(S.program, "raytracer_bvh", &err); /* RayTracer.cpp:2422 */
    if (err != CL_SUCCESS) return fail("clCreateKernel(raytracer_bvh)", err);
    return 0;
}

static void release(cl_mem* m) {
    if (*m) cl.ReleaseMemObject(*m);
    *m = NULL;
}

/* the clCreateBuffer calls of initRayTrace (RayTracer.cpp:942-984) */
int refcl_upload_scene(const float* verts, int V, const int* indices, int T, const void* nodes, int N, const int* tri_indices, int R,
                       const float* normals, int Vn, const int* normal_indices, const void* materials, int M,
                       const int* tri_to_material) {
    cl_int err;
    cl_mem* all[] = {&S.mesh_vertices, &S.mesh_indices, &S.bvh_nodes, &S.bvh_tris_indices, &S.mesh_normals, &S.mesh_normals_indices,
                     &S.mesh_materials, &S.mesh_triangle_index_to_material_index, &S.temp, &S.params};
    for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) release(all[i]);
    const cl_bitfield ro = CL_MEM_READ_ONLY | CL_MEM_COPY_HOST_PTR;
#define MK(dst, bytes, ptr)                                                   \
    do {                                                                      \
        dst = cl.CreateBuffer(S.context, ro, (size_t)(bytes), (void*)(ptr), &err); \
        if (err != CL_SUCCESS) return fail("clCreateBuffer(" #dst ")", err);  \
    } while (0)
    MK(S.mesh_vertices, (size_t)V * 16, verts);
    MK(S.mesh_indices, (size_t)T * 12, indices);
    MK(S.bvh_nodes, (size_t)N * 48, nodes);
    MK(S.bvh_tris_indices, (size_t)R * 4, tri_indices);
    MK(S.mesh_normals, (size_t)Vn * 16, normals);
    MK(S.mesh_normals_indices, (size_t)T * 12, normal_indices);
    MK(S.mesh_materials, (size_t)M * 176, materials);
    MK(S.mesh_triangle_index_to_material_index, (size_t)T * 4, tri_to_material);
    S.temp = cl.CreateBuffer(S.context, CL_MEM_WRITE_ONLY, (64 + 1) * 16, NULL, &err);
    if (err != CL_SUCCESS) return fail("clCreateBuffer(temp)", err);
    float zero[32] = {0};
    MK(S.params, 128, zero);
    S.num_bvh_tris = R;
    S.num_bvh_nodes = N;
    return 0;
}

/* updateCamera's params write (RayTracer.cpp:671) + init_arg1/initCLVolume2 (:1228-1260) + raytrace_gpgpu (:330-344).
 * kernel_ms (optional): device time of the NDRange from OpenCL profiling events. */
int refcl_render(const float* params32, int w, int h, uint32_t* out_pixels, double* kernel_ms) {
    cl_int err;
    if (!S.out || S.out_w != w || S.out_h != h) {
        release(&S.out);
        S.out = cl.CreateBuffer(S.context, CL_MEM_WRITE_ONLY, (size_t)w * h * 4, NULL, &err);
        if (err != CL_SUCCESS) return fail("clCreateBuffer(out)", err);
        S.out_w = w;
        S.out_h = h;
    }
    err = cl.EnqueueWriteBuffer(S.queue, S.params, CL_TRUE, 0, 128, params32, 0, NULL, NULL);
    if (err != CL_SUCCESS) return fail("clEnqueueWriteBuffer(params)", err);
    const cl_uint uw = (cl_uint)w, uh = (cl_uint)h;
    const cl_mem null_mem = NULL; /* the legacy `triangles` buffer is never created in the reference (empty array) */
    const cl_int zero = 0;
    int a = 0;
#define ARG(size, ptr)                                              \
    do {                                                            \
        err = cl.SetKernelArg(S.kernel, a++, (size), (ptr));        \
        if (err != CL_SUCCESS) return fail("clSetKernelArg", err);  \
    } while (0)
    ARG(sizeof(cl_mem), &S.out);
    ARG(sizeof(cl_uint), &uw);
    ARG(sizeof(cl_uint), &uh);
    ARG(sizeof(cl_mem), &null_mem);
    ARG(sizeof(cl_int), &zero); /* number_of_triangles */
    ARG(sizeof(cl_mem), &S.params);
    ARG(sizeof(cl_mem), &S.mesh_vertices);
    ARG(sizeof(cl_mem), &S.mesh_indices);
    ARG(sizeof(cl_mem), &S.bvh_nodes);
    ARG(sizeof(cl_mem), &S.bvh_tris_indices);
    ARG(sizeof(cl_int), &S.num_bvh_tris);
    ARG(sizeof(cl_int), &S.num_bvh_nodes);
    ARG(sizeof(cl_mem), &S.temp);
    ARG(sizeof(cl_mem), &S.mesh_normals);
    ARG(sizeof(cl_mem), &S.mesh_normals_indices);
    ARG(sizeof(cl_mem), &S.mesh_materials);
    ARG(sizeof(cl_mem), &S.mesh_triangle_index_to_material_index);
    const size_t local[2] = {8, 8};
    const size_t global[2] = {((size_t)w + 7) / 8 * 8, ((size_t)h + 7) / 8 * 8};
    cl_event ev = NULL;
    err = cl.EnqueueNDRangeKernel(S.queue, S.kernel, 2, NULL, global, local, 0, NULL, &ev);
    if (err != CL_SUCCESS) return fail("clEnqueueNDRangeKernel", err);
    err = cl.Finish(S.queue);
    if (err != CL_SUCCESS) return fail("clFinish", err);
    if (kernel_ms) {
        cl_ulong t0 = 0, t1 = 0;
        cl.GetEventProfilingInfo(ev, CL_PROFILING_COMMAND_START, sizeof t0, &t0, NULL);
        cl.GetEventProfilingInfo(ev, CL_PROFILING_COMMAND_END, sizeof t1, &t1, NULL);
        *kernel_ms = (double)(t1 - t0) * 1e-6;
    }
    if (ev) cl.ReleaseEvent(ev);
    err = cl.EnqueueReadBuffer(S.queue, S.out, CL_TRUE, 0, (size_t)w * h * 4, out_pixels, 0, NULL, NULL);
    if (err != CL_SUCCESS) return fail("clEnqueueReadBuffer", err);
    return 0;
}

void refcl_shutdown(void) {
    if (!cl.lib) return;
    cl_mem* all[] = {&S.mesh_vertices, &S.mesh_indices, &S.bvh_nodes, &S.bvh_tris_indices, &S.mesh_normals, &S.mesh_normals_indices,
                     &S.mesh_materials, &S.mesh_triangle_index_to_material_index, &S.temp, &S.params, &S.out};
    for (size_t i = 0; i < sizeof all / sizeof all[0]; i++) release(all[i]);
    if (S.kernel) cl.ReleaseKernel(S.kernel);
    if (S.program) cl.ReleaseProgram(S.program);
    if (S.queue) cl.ReleaseCommandQueue(S.queue);
    if (S.context) cl.ReleaseContext(S.context);
    memset(&S, 0, sizeof S);
}
