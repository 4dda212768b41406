#include "Camera.h"

#include <cmath>

namespace {
const double kPi = 3.14159265358979323846;  // M_PI
}

Camera::Camera() : center(0.0f, 0.0f, 0.0f), cam_radius(200.0f), cam_alpha(0.0f), cam_beta(0.0f) {
    // 45*(3+2) degrees of azimuth, 45 degrees of elevation, evaluated in double like the reference
    add_rotate((float)(45 * (3 + 2) * kPi / 180.0f), (float)(45 * kPi / 180.0f));
}

void Camera::add_radius(float dr) {
    cam_radius += dr;
    place_eye();
}

void Camera::add_rotate(float da, float db) {
    cam_alpha += da;
    cam_beta += db;
    // keep beta in [0, 2pi]; flip `up` while the camera is upside down
    if (cam_beta < 0.0f)
        cam_beta += 2.0f * kPi;
    else if (cam_beta > 2.0f * kPi)
        cam_beta -= 2.0f * kPi;
    const bool flipped = (cam_beta > kPi / 2.0f) && (cam_beta < 3.0f * kPi / 2.0f);
    up = float3(0.0f, flipped ? -1.0f : 1.0f, 0.0f);
    place_eye();
}

void Camera::place_eye() {
    const float cb = cosf(cam_beta), sb = sinf(cam_beta);
    eye.x = center.x + cam_radius * cb * cosf(cam_alpha);
    eye.y = center.y + cam_radius * sb;
    eye.z = center.z + cam_radius * cb * sinf(cam_alpha);
    rebuild_basis();
}

void Camera::rebuild_basis() {
    camera_direction = normalize(center - eye);
    camera_right = normalize(cross(camera_direction, up));
    camera_up = normalize(-cross(camera_direction, camera_right));
}

void Camera::make_params(int image_width, int image_height, const float3& light_pos, const float3& light_color,
                         const float3& aabb_min, const float3& aabb_max, float out[32]) const {
    // FOV 60 degrees with the reference's pi ~ 3.1415 (reference RayTracer.cpp:639-640)
    const float FOV = 60.0f;
    const float theta = (float)((FOV * 3.1415 * 0.5) / 180.0f);
    const float half_width = tanf(theta);
    const float aspect = (float)image_width / (float)image_height;
    const float u0 = -half_width * aspect, v0 = -half_width;
    const float u1 = half_width * aspect, v1 = half_width;
    const float dist_to_image = 1.0f;

    const float3 a = (u1 - u0) * camera_right;
    const float3 b = (v1 - v0) * camera_up;
    const float3 c = eye + u0 * camera_right + v0 * camera_up + dist_to_image * camera_direction;

    const float3 rows[8] = {a, b, c, eye, light_pos, light_color, aabb_min, aabb_max};
    for (int i = 0; i < 8; ++i) {
        out[4 * i + 0] = rows[i].x;
        out[4 * i + 1] = rows[i].y;
        out[4 * i + 2] = rows[i].z;
        out[4 * i + 3] = 1.0f;
    }
}
