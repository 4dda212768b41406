#include "RayTracer.h"

#include <cstdio>

#include "../../include/rtb200.h"
#include "ColladaLoader.h"

RayTracer::RayTracer() {
    for (float& p : params) p = 0.0f;
}

RayTracer::~RayTracer() { cleanup(); }

int RayTracer::setupCL(int device_ordinal) {
    if (ctx_) return SDK_SUCCESS;
    if (rt_create(device_ordinal, &ctx_) != RT_OK) {
        err_ = rt_last_error(nullptr);
        ctx_ = nullptr;
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

int RayTracer::initRayTrace(const char* collada_path) {
    ColladaLoader loader;
    if (!loader.load(collada_path)) {
        err_ = std::string("cannot load ") + collada_path + ": " + loader.error();
        return SDK_FAILURE;
    }
    mesh1.clear();
    mesh1.init(loader);
    printf("\n mesh1.faces.size() = %d \n", mesh1.getNumTriangles());
    bvh2.setMesh(&mesh1);
    build_seconds_ = bvh2.buildSeconds();
    bvh_cuda.build_from_bvh2(bvh2);
    return upload();
}

int RayTracer::initRayTraceFromMesh(const char* flat_bvh_cache) {
    if (mesh1.getNumTriangles() < 1) {
        err_ = "initRayTraceFromMesh: mesh1 is empty";
        return SDK_FAILURE;
    }
    if (mesh1.normals.empty() || mesh1.materials.empty()) mesh1.finish_synthetic();
    if (!(flat_bvh_cache && bvh_cuda.load(flat_bvh_cache))) {
        bvh2.setMesh(&mesh1);
        build_seconds_ = bvh2.buildSeconds();
        bvh_cuda.build_from_bvh2(bvh2);
        if (flat_bvh_cache) bvh_cuda.save(flat_bvh_cache);
    }
    return upload();
}

int RayTracer::upload() {
    if (!ctx_) {
        err_ = "initRayTrace before setupCL";
        return SDK_FAILURE;
    }
    const int rc = rt_upload_scene(ctx_, &mesh1.vertices[0].x, (int)mesh1.vertices.size(), mesh1.indices.data(), mesh1.getNumTriangles(),
                                   bvh_cuda.bvh_nodes.data(), (int)bvh_cuda.bvh_nodes.size(), bvh_cuda.tri_indices.data(),
                                   (int)bvh_cuda.tri_indices.size(), &mesh1.normals[0].x, (int)mesh1.normals.size(),
                                   mesh1.normals_indices.data(), mesh1.materials.data(), (int)mesh1.materials.size(),
                                   mesh1.triangle_index_to_material_index.data());
    if (rc != RT_OK) {
        err_ = rt_last_error(ctx_);
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

int RayTracer::updateCamera() {
    cam.make_params(image_width, image_height, light_pos, light_color1, mesh1.scene_aabbox_min, mesh1.scene_aabbox_max, params);
    if (animate) cam.add_rotate(-0.25f * 1.5f * delta_t, 0.0f);  // the reference's orbit (RayTracer.cpp:657)
    if (!ctx_ || rt_set_params(ctx_, params) != RT_OK) {
        err_ = ctx_ ? rt_last_error(ctx_) : "updateCamera before setupCL";
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

int RayTracer::raytrace_gpgpu() {
    out_data.resize((size_t)image_width * image_height);
    if (!ctx_ || rt_render_frame(ctx_, image_width, image_height, out_data.data()) != RT_OK) {
        err_ = ctx_ ? rt_last_error(ctx_) : "raytrace_gpgpu before setupCL";
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

int RayTracer::cleanup() {
    if (ctx_) rt_destroy(ctx_);
    ctx_ = nullptr;
    return SDK_SUCCESS;
}

bool RayTracer::write_ppm(const char* path) const {
    if (out_data.size() != (size_t)image_width * image_height) return false;
    FILE* f = fopen(path, "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", image_width, image_height);
    for (int y = image_height - 1; y >= 0; y--)  // the reference displays row 0 at the bottom (GL texture)
        for (int x = 0; x < image_width; x++) {
            const unsigned int p = out_data[(size_t)y * image_width + x];
            const unsigned char rgb[3] = {(unsigned char)(p & 255), (unsigned char)((p >> 8) & 255), (unsigned char)((p >> 16) & 255)};
            fwrite(rgb, 1, 3, f);
        }
    fclose(f);
    return true;
}
