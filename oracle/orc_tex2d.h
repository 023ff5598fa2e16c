/* ORACLE — test infrastructure, never linked into or imported by the product.
 *
 * CPU emulation of `tex2D<float4>(tex, x, y)` on a float4 cudaArray with normalised coordinates, element read mode
 * and one address mode for both axes — the texture object CudaTextureManager::GetCudaTextureObject creates
 * (framework/cuda/texture.cpp:60-102) and cuda::Texture::Sample reads (framework/cuda/texture.h:52-54).
 *
 * The texture unit is hardware; the reference tree holds no test for it, so this restates the published rules of the
 * CUDA C++ Programming Guide, appendix "Texture Fetching" (parity unpinned by the reference, pinned against the GPU
 * by tests/test_gpu_kat.py):
 *   addressing (normalised x):  wrap: x - floor(x);  clamp: clamp to [0, 1] (texel indices are clamped to N - 1 afterwards:
 *                               the guide's "[0, 1 - 1/N]" for point sampling; for linear filtering the last half texel then
 *                               blends T[N-1] with itself, which is what the texture unit returns — tests/test_gpu_kat.py);
 *                               mirror: frac(x) when floor(x) is even, 1 - frac(x) when odd;  border: outside -> border colour
 *   nearest:  T[floor(x N)]
 *   linear:   xB = x N - 0.5, i = floor(xB), a = frac(xB) kept in 1.8 fixed point;
 *             (1-a)(1-b) T[i,j] + a(1-b) T[i+1,j] + (1-a) b T[i,j+1] + a b T[i+1,j+1], out-of-range indices wrapped
 *             (wrap) or clamped (clamp, mirror).
 */
#ifndef ORC_TEX2D_H
#define ORC_TEX2D_H
#include <cmath>
#include <cstdint>

namespace orc {
enum { ORC_ADDR_WRAP = 0, ORC_ADDR_CLAMP = 1, ORC_ADDR_MIRROR = 2, ORC_ADDR_BORDER = 3 };

struct Tex2dResult {
    float x, y, z, w;
};
namespace tex2d_detail {
inline float address(float x, int mode, int n) {
    switch (mode) {
        case ORC_ADDR_WRAP: return x - floorf(x);
        case ORC_ADDR_MIRROR: {
            const float fl = floorf(x), fr = x - fl;
            return (static_cast<long long>(fl) & 1) ? 1.f - fr : fr;
        }
        case ORC_ADDR_CLAMP: return fminf(fmaxf(x, 0.f), 1.f);
        default: return x;
    }
}
inline bool index(int &i, int mode, int n) { // false: the border colour applies
    if (i >= 0 && i < n) return true;
    if (mode == ORC_ADDR_WRAP) {
        i %= n;
        if (i < 0) i += n;
        return true;
    }
    if (mode == ORC_ADDR_BORDER) return false;
    i = i < 0 ? 0 : n - 1;
    return true;
}
inline Tex2dResult texel(const float *rgba, int w, int h, int i, int j, int mode) {
    if (!index(i, mode, w) || !index(j, mode, h)) return Tex2dResult{ 1.f, 0.f, 0.f, 0.f }; // borderColor[0] = 1 (texture.cpp:89)
    const float *p = rgba + (static_cast<size_t>(j) * w + i) * 4;
    return Tex2dResult{ p[0], p[1], p[2], p[3] };
}
inline float weight8(float a) { return floorf(a * 256.f + 0.5f) / 256.f; } // 1.8 fixed point, round to nearest
}// namespace tex2d_detail

inline Tex2dResult tex2d(const float *rgba, int w, int h, int address_mode, int linear, float x, float y) {
    using namespace tex2d_detail;
    const float xs = address(x, address_mode, w) * static_cast<float>(w), ys = address(y, address_mode, h) * static_cast<float>(h);
    if (!linear) return texel(rgba, w, h, static_cast<int>(floorf(xs)), static_cast<int>(floorf(ys)), address_mode);
    const float xb = xs - 0.5f, yb = ys - 0.5f;
    const float fi = floorf(xb), fj = floorf(yb);
    const float a = weight8(xb - fi), b = weight8(yb - fj);
    const int i = static_cast<int>(fi), j = static_cast<int>(fj);
    const Tex2dResult t00 = texel(rgba, w, h, i, j, address_mode), t10 = texel(rgba, w, h, i + 1, j, address_mode);
    const Tex2dResult t01 = texel(rgba, w, h, i, j + 1, address_mode), t11 = texel(rgba, w, h, i + 1, j + 1, address_mode);
    auto mix = [&](float c00, float c10, float c01, float c11) {
        return (1.f - a) * (1.f - b) * c00 + a * (1.f - b) * c10 + (1.f - a) * b * c01 + a * b * c11;
    };
    return Tex2dResult{ mix(t00.x, t10.x, t01.x, t11.x), mix(t00.y, t10.y, t01.y, t11.y), mix(t00.z, t10.z, t01.z, t11.z),
                        mix(t00.w, t10.w, t01.w, t11.w) };
}
}// namespace orc
#endif
