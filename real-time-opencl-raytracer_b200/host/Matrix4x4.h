// Matrix4x4.h -- 4x4 float matrix, row-major storage, ROW-VECTOR convention (v' = v * M), as used
// by the COLLADA node-transform bake (reference Matrix4x4.h:5-60, Matrix4x4.cpp:22-115).
// rotateX/Y/Z and translate post-multiply (*this = *this * R). `transponse` keeps the
// reference's spelling.
#pragma once
#include "vecmath.h"

class Matrix4x4 {
public:
    float m[16];

    Matrix4x4() { identity(); }
    void identity() {
        for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    }
    void set(const float v[16]) {
        for (int i = 0; i < 16; i++) m[i] = v[i];
    }
    void multiply(const Matrix4x4& r);  // *this = *this * r
    void rotateX(float angle);
    void rotateY(float angle);
    void rotateZ(float angle);
    void rotateXYZ(float rx, float ry, float rz) { rotateX(rx); rotateY(ry); rotateZ(rz); }
    void translate(float tx, float ty, float tz);
    void transponse();
};

// v * M, accumulating in the order j = 0..3 starting from 0.0f
inline float4 multiply(const float4& v, const Matrix4x4& mat) {
    float4 out;
    for (int col = 0; col < 4; col++) {
        float s = 0.0f;
        for (int j = 0; j < 4; j++) s += v.m[j] * mat.m[j * 4 + col];
        out.m[col] = s;
    }
    return out;
}
