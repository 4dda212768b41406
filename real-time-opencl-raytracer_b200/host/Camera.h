// Camera.h -- orbit camera (radius / azimuth alpha / elevation beta around `center`).
//
// Same public fields and methods as the reference (reference Camera.h:5-29): the default
// constructor reproduces the reference's start pose (radius 200, alpha 225 deg, beta 45 deg,
// i.e. eye ~ (-100, 141.42, -100) looking at the origin; reference Camera.cpp:6-19).
// make_params() is the reference's updateCamera() (reference RayTracer.cpp:634-671) factored
// into the class: it fills the 128-byte Params block the kernel reads.
#pragma once
#include "vecmath.h"

class Camera {
public:
    float3 eye, center, up;
    float cam_radius, cam_alpha, cam_beta;
    float3 camera_right, camera_up, camera_direction;

    Camera();
    void add_radius(float dr);
    void add_rotate(float da, float db);

    // 8 x float4 (w = 1): a, b, c, campos, light_pos, light_color, scene_aabb_min, scene_aabb_max
    void make_params(int image_width, int image_height, const float3& light_pos, const float3& light_color,
                     const float3& aabb_min, const float3& aabb_max, float out_params[32]) const;

private:
    void place_eye();      // eye from (radius, alpha, beta) around center
    void rebuild_basis();  // right / up / direction from eye, center, up
};
