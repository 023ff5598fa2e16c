/* ORACLE shim — shadows the reference's framework/cuda/texture.h on the include path.
 *
 * The real header pulls util/texture.h -> util/type.h -> <DirectXMath.h> (Windows SDK, absent
 * here).  This shim keeps the type name and member the BSDF / emitter headers use
 * (Pupil::cuda::Texture::Sample) and restates the body of cuda/texture.h:33-57; the bitmap
 * branch's tex2D<float4> (hardware) goes through the emulation in orc_tex2d.h.
 */
#pragma once
#include "cuda/preprocessor.h"
#include "cuda/vec_math.h"
#include "orc_tex2d.h"

namespace Pupil::util {
enum class ETextureType : unsigned int { RGB = 0, Bitmap, Checkerboard }; // util/texture.h:21-25
}
namespace Pupil::cuda {
struct Texture {
    util::ETextureType type = util::ETextureType::RGB;
    float3 rgb;    // RGB colour, or checkerboard patch1
    float3 patch2; // checkerboard patch2
    struct {
        float4 r0, r1, r2, r3;
    } transform;
    // stands in for cudaTextureObject_t bitmap: the texels and the sampler state of the texture object
    const float *bitmap = nullptr;
    int bitmap_w = 0, bitmap_h = 0, address_mode = 0, filter_mode = 1;

    float3 Sample(float2 texcoord) const noexcept {
        const float4 tex = make_float4(texcoord.x, texcoord.y, 0.f, 1.f);
        float tex_x = dot(transform.r0, tex);
        float tex_y = dot(transform.r1, tex);
        float3 color = rgb;
        if (type == util::ETextureType::Bitmap) {
            const orc::Tex2dResult c = orc::tex2d(bitmap, bitmap_w, bitmap_h, address_mode, filter_mode, tex_x, tex_y);
            color = make_float3(c.x, c.y, c.z);
        } else if (type == util::ETextureType::Checkerboard) {
            tex_x = tex_x - (tex_x > 0.f ? floorf(tex_x) : ceilf(tex_x));
            tex_y = tex_y - (tex_y > 0.f ? floorf(tex_y) : ceilf(tex_y));
            if (tex_x < 0.f) tex_x += 1.f;
            if (tex_y < 0.f) tex_y += 1.f;
            if (tex_x > 0.5f)
                color = tex_y > 0.5f ? rgb : patch2;
            else
                color = tex_y > 0.5f ? patch2 : rgb;
        }
        return color;
    }
};
}// namespace Pupil::cuda
