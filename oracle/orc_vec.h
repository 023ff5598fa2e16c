/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * Minimal fp32 vector algebra with the SAME operation order as the reference's
 * framework/cuda/vec_math.h, because low-order bits feed discrete decisions (lobe choice,
 * IsZero cut-offs).  The conventions that matter:
 *   v / s        = v * (1.0f / s)            (vec_math.h:425-428, 432-435)
 *   normalize(v) = v * (1.0f / sqrtf(v.v))   (vec_math.h:477-480)
 *   lerp(a,b,t)  = a + t * (b - a)           (vec_math.h:439-441)
 *   dot          = x*x' + y*y' + z*z' left to right
 * Compile with -ffp-contract=off so gcc does not fuse these differently from the reference
 * host build.
 */
#ifndef ORC_VEC_H
#define ORC_VEC_H
#include <cmath>

namespace orc {
struct f2 {
    float x, y;
};
struct f3 {
    float x, y, z;
};
struct f4 {
    float x, y, z, w;
};

constexpr float kPi = 3.14159265358979323846f;     // M_PIf   vec_math.h:42-44
constexpr float kInvPi = 0.318309886183790671538f; // M_1_PIf vec_math.h:48-50

inline f3 mk3(float x, float y, float z) { return f3{ x, y, z }; }
inline f3 mk3(float s) { return f3{ s, s, s }; }
inline f2 mk2(float x, float y) { return f2{ x, y }; }

inline f3 operator+(f3 a, f3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline f3 operator-(f3 a, f3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline f3 operator-(f3 a) { return { -a.x, -a.y, -a.z }; }
inline f3 operator*(f3 a, f3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline f3 operator*(f3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline f3 operator*(float s, f3 a) { return { a.x * s, a.y * s, a.z * s }; }
inline f3 operator/(f3 a, f3 b) { return { a.x / b.x, a.y / b.y, a.z / b.z }; }
inline f3 operator/(f3 a, float s) {
    float inv = 1.0f / s;
    return a * inv;
}
inline f3 operator/(float s, f3 a) { return { s / a.x, s / a.y, s / a.z }; }
inline f3 &operator+=(f3 &a, f3 b) {
    a.x += b.x, a.y += b.y, a.z += b.z;
    return a;
}
inline f3 &operator*=(f3 &a, f3 b) {
    a.x *= b.x, a.y *= b.y, a.z *= b.z;
    return a;
}
inline f3 &operator*=(f3 &a, float s) {
    a.x *= s, a.y *= s, a.z *= s;
    return a;
}
inline f3 &operator/=(f3 &a, float s) {
    float inv = 1.0f / s;
    a *= inv;
    return a;
}
inline f3 operator+(f3 a, float s) { return { a.x + s, a.y + s, a.z + s }; }
inline f3 operator-(float s, f3 a) { return { s - a.x, s - a.y, s - a.z }; }

inline f2 operator*(f2 a, float s) { return { a.x * s, a.y * s }; }
inline f2 operator*(float s, f2 a) { return { a.x * s, a.y * s }; }
inline f2 operator+(f2 a, f2 b) { return { a.x + b.x, a.y + b.y }; }

inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline f3 cross(f3 a, f3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float length(f3 v) { return sqrtf(dot(v, v)); }
inline f3 normalize(f3 v) {
    float inv_len = 1.0f / sqrtf(dot(v, v));
    return v * inv_len;
}
inline f3 lerp(f3 a, f3 b, float t) { return a + t * (b - a); }

/* row-major 4x4, column-vector convention (util::Mat4, framework/util/type.h:73-111) */
struct m44 {
    float e[16];
};
inline m44 identity44() {
    m44 r{};
    r.e[0] = r.e[5] = r.e[10] = r.e[15] = 1.f;
    return r;
}
}// namespace orc
#endif
