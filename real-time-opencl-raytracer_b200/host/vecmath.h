// vecmath.h -- small fp32 vector types for the host scene layer.
//
// Mirrors the *names and numeric behaviour* of the reference's math layer (reference
// vectors_math.h / vectors_math.cpp) so that host code written against the reference
// (float3/float4/int3 with .m[] access, make_float3, dot, cross, normalize, fminf1/fmaxf1 ...)
// keeps compiling, but is written from scratch as header-only inline code.
//
// Behaviour that downstream results depend on (checked bit-exactly against the reference's own
// compiled objects in tests/test_oracle_pinning.py and tests/test_host_bvh.py):
//   * fminf1/fmaxf1 are the plain `a<b?a:b` / `a>b?a:b` selects (reference vectors_math.cpp:126-134),
//     NOT the NaN-ignoring libm fminf/fmaxf
//   * dot is evaluated left to right, normalize multiplies by 1.0f/sqrtf(dot(v,v))
//     (reference vectors_math.cpp:18-20,73-84)
//   * every translation unit using this header is built with -ffp-contract=off
#pragma once
#include <cmath>

struct int2 {
    union { struct { int x, y; }; int m[2]; };
    int2() : x(0), y(0) {}
    int2(int x_, int y_) : x(x_), y(y_) {}
};
struct int3 {
    union { struct { int x, y, z; }; int m[3]; };
    int3() : x(0), y(0), z(0) {}
    int3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
    int& get(int i) { return m[i]; }
    const int& get(int i) const { return m[i]; }
};
struct int4 {
    union { struct { int x, y, z, w; }; int m[4]; };
    int4() : x(0), y(0), z(0), w(0) {}
    int4(int x_, int y_, int z_, int w_) : x(x_), y(y_), z(z_), w(w_) {}
};
struct float2 {
    union { struct { float x, y; }; float m[2]; };
    float2() : x(0), y(0) {}
    float2(float x_, float y_) : x(x_), y(y_) {}
};
struct float3 {
    union { struct { float x, y, z; }; float m[3]; };
    float3() : x(0), y(0), z(0) {}
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float& get(int i) { return m[i]; }
    const float& get(int i) const { return m[i]; }
};
struct float4 {
    union { struct { float x, y, z, w; }; float m[4]; };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float& get(int i) { return m[i]; }
    const float& get(int i) const { return m[i]; }
};
static_assert(sizeof(float4) == 16 && sizeof(float3) == 12 && sizeof(int4) == 16, "POD layout");

// ---- scalars -------------------------------------------------------------------------------
inline float fminf1(float a, float b) { return a < b ? a : b; }
inline float fmaxf1(float a, float b) { return a > b ? a : b; }
inline float fminf1(float a, float b, float c) { return fminf1(fminf1(a, b), c); }
inline float fmaxf1(float a, float b, float c) { return fmaxf1(fmaxf1(a, b), c); }
inline float clamp(float f, float lo, float hi) { return fmaxf1(lo, fminf1(f, hi)); }
inline float lerp(float a, float b, float t) { return a * (1.0f - t) + b * t; }
template <class T> inline T sqr(const T& a) { return a * a; }
template <class T> inline void swap1(T& a, T& b) { T t = a; a = b; b = t; }

// ---- float3 --------------------------------------------------------------------------------
inline float3 make_float3(float x, float y, float z) { return float3(x, y, z); }
inline float3 make_float3(const float4& v) { return float3(v.x, v.y, v.z); }
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator/(float s, float3 a) { return float3(s / a.x, s / a.y, s / a.z); }
inline void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
    return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(float3 v) { return sqrtf(dot(v, v)); }
inline float3 normalize(float3 v) {
    float inv_len = 1.0f / sqrtf(dot(v, v));
    return inv_len * v;
}
inline float3 fminf1(const float3& a, const float3& b) {
    return float3(fminf1(a.x, b.x), fminf1(a.y, b.y), fminf1(a.z, b.z));
}
inline float3 fmaxf1(const float3& a, const float3& b) {
    return float3(fmaxf1(a.x, b.x), fmaxf1(a.y, b.y), fmaxf1(a.z, b.z));
}
inline float fminf1(const float3& a) { return fminf1(a.x, a.y, a.z); }
inline float fmaxf1(const float3& a) { return fmaxf1(a.x, a.y, a.z); }
inline float fsumf(const float3& a) { return a.x + a.y + a.z; }

// ---- float4 --------------------------------------------------------------------------------
inline float4 make_float4(float x, float y, float z, float w) { return float4(x, y, z, w); }
inline float4 make_float4(const float3& v) { return float4(v.x, v.y, v.z, 1.0f); }
inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator*(float4 a, float s) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline float4 lerp(const float4& a, const float4& b, float t) { return a * (1.0f - t) + b * t; }

// ---- int3 ----------------------------------------------------------------------------------
inline int3 make_int3(int x, int y, int z) { return int3(x, y, z); }
// float -> int truncation, as the reference's make_int3(const float3&) (vectors_math.cpp:236-238)
inline int3 make_int3(const float3& a) { return int3((int)a.x, (int)a.y, (int)a.z); }
inline int imin_(int a, int b) { return a < b ? a : b; }
inline int imax_(int a, int b) { return a > b ? a : b; }
inline int3 clamp(const int3& v, int lo, int hi) {
    return int3(imax_(imin_(v.x, hi), lo), imax_(imin_(v.y, hi), lo), imax_(imin_(v.z, hi), lo));
}
inline int3 clamp(const int3& v, const int3& lo, int hi) {
    return int3(imax_(imin_(v.x, hi), lo.x), imax_(imin_(v.y, hi), lo.y), imax_(imin_(v.z, hi), lo.z));
}
