// fp32 vector helpers for the device code.  Operation order follows the reference's
// framework/cuda/vec_math.h where it is observable (v/s = v*(1/s), normalize = v*(1/sqrt(v.v)),
// lerp = a + t*(b-a)); nvcc may still contract a*b+c into FMAs, which the parity tolerances allow for.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb2 {
#define PB2_HD __host__ __device__ __forceinline__
#define PB2_D __device__ __forceinline__

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.318309886183790671538f;

PB2_HD float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
PB2_HD float3 mk3(float s) { return make_float3(s, s, s); }
PB2_HD float3 mk3(float4 v) { return make_float3(v.x, v.y, v.z); }
PB2_HD float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB2_HD float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB2_HD float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
PB2_HD float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PB2_HD float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PB2_HD float3 operator*(float s, float3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
PB2_HD float3 operator/(float3 a, float3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PB2_HD float3 operator/(float3 a, float s) {
    float inv = 1.0f / s;
    return a * inv;
}
PB2_HD float3 operator-(float s, float3 a) { return mk3(s - a.x, s - a.y, s - a.z); }
PB2_HD void operator+=(float3 &a, float3 b) { a.x += b.x, a.y += b.y, a.z += b.z; }
PB2_HD void operator*=(float3 &a, float3 b) { a.x *= b.x, a.y *= b.y, a.z *= b.z; }
PB2_HD void operator*=(float3 &a, float s) { a.x *= s, a.y *= s, a.z *= s; }
PB2_HD void operator/=(float3 &a, float s) {
    float inv = 1.0f / s;
    a *= inv;
}
PB2_HD float2 operator*(float s, float2 a) { return make_float2(a.x * s, a.y * s); }
PB2_HD float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
PB2_HD float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PB2_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PB2_HD float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
PB2_HD float3 cross(float3 a, float3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
PB2_HD float length(float3 v) { return sqrtf(dot(v, v)); }
PB2_HD float3 normalize(float3 v) {
    float inv_len = 1.0f / sqrtf(dot(v, v));
    return v * inv_len;
}
PB2_HD float3 lerp3(float3 a, float3 b, float t) { return a + t * (b - a); }
PB2_HD float lerp1(float a, float b, float t) { return a + t * (b - a); } // optix::Lerp, framework/optix/util.h:181-183
PB2_HD float3 fmin3(float3 a, float3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
PB2_HD float3 fmax3(float3 a, float3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }

// rows of a 3x4 affine matrix applied to a point / vector, and the transposed 3x3 for normals
PB2_HD float3 xf_point(const float4 r0, const float4 r1, const float4 r2, float3 p) {
    return mk3(r0.x * p.x + r0.y * p.y + r0.z * p.z + r0.w, r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w,
               r2.x * p.x + r2.y * p.y + r2.z * p.z + r2.w);
}
PB2_HD float3 xf_vector(const float4 r0, const float4 r1, const float4 r2, float3 v) {
    return mk3(r0.x * v.x + r0.y * v.y + r0.z * v.z, r1.x * v.x + r1.y * v.y + r1.z * v.z, r2.x * v.x + r2.y * v.y + r2.z * v.z);
}
PB2_HD float3 xf_normal_t(const float4 r0, const float4 r1, const float4 r2, float3 n) { // (M^T) n
    return mk3(r0.x * n.x + r1.x * n.y + r2.x * n.z, r0.y * n.x + r1.y * n.y + r2.y * n.z, r0.z * n.x + r1.z * n.y + r2.z * n.z);
}
}// namespace pb2
