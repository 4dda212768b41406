#include "Matrix4x4.h"

#include <cmath>

void Matrix4x4::multiply(const Matrix4x4& r) {
    float out[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += m[i * 4 + k] * r.m[k * 4 + j];
            out[i * 4 + j] = s;
        }
    set(out);
}

void Matrix4x4::rotateX(float angle) {
    const float s = sinf(angle), c = cosf(angle);
    Matrix4x4 r;
    const float v[16] = {1, 0, 0, 0, 0, c, s, 0, 0, -s, c, 0, 0, 0, 0, 1};
    r.set(v);
    multiply(r);
}

void Matrix4x4::rotateY(float angle) {
    const float s = sinf(angle), c = cosf(angle);
    Matrix4x4 r;
    const float v[16] = {c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0, 0, 0, 0, 1};
    r.set(v);
    multiply(r);
}

void Matrix4x4::rotateZ(float angle) {
    const float s = sinf(angle), c = cosf(angle);
    Matrix4x4 r;
    const float v[16] = {c, s, 0, 0, -s, c, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    r.set(v);
    multiply(r);
}

void Matrix4x4::translate(float tx, float ty, float tz) {
    Matrix4x4 r;
    r.m[12] = tx;
    r.m[13] = ty;
    r.m[14] = tz;
    multiply(r);
}

void Matrix4x4::transponse() {
    float out[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[j * 4 + i] = m[i * 4 + j];
    set(out);
}
