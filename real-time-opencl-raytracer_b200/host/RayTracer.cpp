#include "RayTracer.h"

#include <cstdio>

#include "../../include/rtb200.h"
#include "ColladaLoader.h"

RayTracer::RayTracer() {
    for (float& p : params) p = 0.0f;
}

RayTracer::~RayTracer() { cleanup(); }

int RayTracer::setupCL(int device_ordinal, int n_gpus) {
    if (ctx_) return SDK_SUCCESS;
    if (n_gpus > 1) {
        if (rt_create_group(n_gpus, nullptr, &group_) != RT_OK) {
            err_ = rt_group_last_error(nullptr);
            group_ = nullptr;
            return SDK_FAILURE;
        }
        ctx_ = rt_group_context(group_, 0);
        n_gpus_ = n_gpus;
        return SDK_SUCCESS;
    }
    if (rt_create(device_ordinal, &ctx_) != RT_OK) {
        err_ = rt_last_error(nullptr);
        ctx_ = nullptr;
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

std::vector<double> RayTracer::last_rank_ms() const {
    std::vector<double> ms;
    double st[RT_GROUP_STATS];
    if (group_ && rt_group_stats(group_, st) == RT_OK)
        for (int r = 0; r < n_gpus_; r++) ms.push_back(st[RT_GROUP_STAT_RANK_KERNEL_MS + r]);
    return ms;
}

double RayTracer::broadcast_ms() const {
    double st[RT_GROUP_STATS];
    return (group_ && rt_group_stats(group_, st) == RT_OK) ? st[RT_GROUP_STAT_BROADCAST_MS] : 0.0;
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
    const float* normals = mesh1.normals.empty() ? nullptr : &mesh1.normals[0].x;
    int rc;
    if (group_)
        rc = rt_group_upload_scene(group_, &mesh1.vertices[0].x, (int)mesh1.vertices.size(), mesh1.indices.data(), mesh1.getNumTriangles(),
                                   bvh_cuda.bvh_nodes.data(), (int)bvh_cuda.bvh_nodes.size(), bvh_cuda.tri_indices.data(),
                                   (int)bvh_cuda.tri_indices.size(), normals, (int)mesh1.normals.size(), mesh1.normals_indices.data(),
                                   mesh1.materials.data(), (int)mesh1.materials.size(), mesh1.triangle_index_to_material_index.data());
    else
        rc = rt_upload_scene(ctx_, &mesh1.vertices[0].x, (int)mesh1.vertices.size(), mesh1.indices.data(), mesh1.getNumTriangles(),
                             bvh_cuda.bvh_nodes.data(), (int)bvh_cuda.bvh_nodes.size(), bvh_cuda.tri_indices.data(),
                             (int)bvh_cuda.tri_indices.size(), normals, (int)mesh1.normals.size(), mesh1.normals_indices.data(),
                             mesh1.materials.data(), (int)mesh1.materials.size(), mesh1.triangle_index_to_material_index.data());
    if (rc != RT_OK) {
        err_ = group_ ? rt_group_last_error(group_) : rt_last_error(ctx_);
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

int RayTracer::updateCamera() {
    cam.make_params(image_width, image_height, light_pos, light_color1, mesh1.scene_aabbox_min, mesh1.scene_aabbox_max, params);
    if (animate) cam.add_rotate(-0.25f * 1.5f * delta_t, 0.0f);  // the reference's orbit (RayTracer.cpp:657)
    if (!ctx_ || (group_ ? rt_group_set_params(group_, params) : rt_set_params(ctx_, params)) != RT_OK) {
        err_ = !ctx_ ? "updateCamera before setupCL" : (group_ ? rt_group_last_error(group_) : rt_last_error(ctx_));
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

void RayTracer::unpin(std::vector<unsigned int>& v) {
    for (size_t i = 0; i < pinned_.size(); i++)
        if (pinned_[i].first == (void*)v.data()) {
            if (ctx_) rt_host_unregister(ctx_, pinned_[i].first);
            pinned_.erase(pinned_.begin() + i);
            return;
        }
}

bool RayTracer::size_and_pin(std::vector<unsigned int>& v) {
    const size_t n = (size_t)image_width * image_height;
    if (v.size() != n) {
        unpin(v);  // before the vector frees the pages
        v.assign(n, 0u);
    }
    for (const auto& p : pinned_)
        if (p.first == (void*)v.data()) return true;
    void* alias = nullptr;
    if (n == 0 || rt_host_register(ctx_, v.data(), n * sizeof(unsigned int), &alias) != RT_OK) return false;  // pageable still works
    pinned_.push_back({(void*)v.data(), n * sizeof(unsigned int)});
    return true;
}

int RayTracer::raytrace_gpgpu() {
    if (!ctx_) {
        err_ = "raytrace_gpgpu before setupCL";
        return SDK_FAILURE;
    }
    while (inflight_ > 0)
        if (raytrace_gpgpu_end() != SDK_SUCCESS) return SDK_FAILURE;
    size_and_pin(out_data);
    if (group_) {  // every GPU stores its row bands straight into out_data (page-locked, portable)
        if (rt_render_frame_tiled(group_, image_width, image_height, out_data.data()) != RT_OK) {
            err_ = rt_group_last_error(group_);
            return SDK_FAILURE;
        }
        return SDK_SUCCESS;
    }
    if (rt_render_frame(ctx_, image_width, image_height, out_data.data()) != RT_OK) {
        err_ = rt_last_error(ctx_);
        return SDK_FAILURE;
    }
    return SDK_SUCCESS;
}

int RayTracer::raytrace_gpgpu_begin() {
    if (!ctx_) {
        err_ = "raytrace_gpgpu_begin before setupCL";
        return SDK_FAILURE;
    }
    if (group_) {
        err_ = "raytrace_gpgpu_begin: frames in flight are a single-GPU feature; use raytrace_gpgpu() with a GPU group";
        return SDK_FAILURE;
    }
    if (inflight_ >= 2) {
        err_ = "raytrace_gpgpu_begin: two frames already in flight";
        return SDK_FAILURE;
    }
    const int slot = (head_ + inflight_) & 1;
    size_and_pin(ring_[slot]);
    if (rt_render_frame_begin(ctx_, image_width, image_height, ring_[slot].data(), slot) != RT_OK) {
        err_ = rt_last_error(ctx_);
        return SDK_FAILURE;
    }
    inflight_++;
    return SDK_SUCCESS;
}

int RayTracer::raytrace_gpgpu_end() {
    if (!ctx_ || inflight_ < 1) {
        err_ = "raytrace_gpgpu_end: no frame in flight";
        return SDK_FAILURE;
    }
    const int slot = head_;
    head_ ^= 1;
    inflight_--;
    if (rt_render_frame_end(ctx_, slot) != RT_OK) {
        err_ = rt_last_error(ctx_);
        return SDK_FAILURE;
    }
    out_data.swap(ring_[slot]);  // both stay page-locked: registration follows the pages, not the vector
    return SDK_SUCCESS;
}

int RayTracer::cleanup() {
    if (ctx_) {
        while (inflight_ > 0) raytrace_gpgpu_end();
        for (const auto& p : pinned_) rt_host_unregister(ctx_, p.first);
        pinned_.clear();
        if (group_) rt_destroy_group(group_);  // owns ctx_
        else rt_destroy(ctx_);
    }
    ctx_ = nullptr;
    group_ = nullptr;
    n_gpus_ = 1;
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
