// Device restatement of the reference's shading arithmetic (paths relative to framework/ in the reference):
//   cuda/random.h                      RNG (TEA init + LCG)            — integer exact
//   optix/util.h:33-183                sampling warps, ONB, reflect/refract, MIS, IsZero
//   render/material/fresnel.h, ggx.h   Fresnel terms, isotropic Smith-GGX with VNDF sampling
//   render/material/bsdf/*.h           the seven BSDFs (Sample / GetBsdf / GetPdf)
//   cuda/texture.h:33-57               RGB + checkerboard textures
//   render/emitter/{area,sphere,env}.h emitters, render/emitter.h:104-137 emitter selection
// The reference's quirks are kept on purpose (they change pixel values): RoughConductor tags its
// lobe DiffuseReflection, Plastic picks its lobe on xi.x and RoughPlastic on xi.y, the constant
// environment samples a hemisphere (pdf 1/2pi) but evaluates with pdf 1/4pi, IsZero uses 1e-6.
// Where the reference reads uninitialised memory the value is defined: BsdfSamplingRecord::wi = 0,
// EmitterSampleRecord::{is_delta = false, distance = 0}, EmitEvalRecord::{pdf = 0, radiance = 0}.
#pragma once
#include "pb2_types.cuh"

namespace pb2 {

constexpr float kEps = 0.000001f;     // optix/util.h:8
constexpr float kMaxDistance = 1e16f; // optix/util.h:9
enum : uint32_t { // render/material/bsdf/bsdf.h:7-24
    kLobeUnknown = 0,
    kLobeDiffuseReflection = 1u << 1,
    kLobeGlossyReflection = 1u << 3,
    kLobeGlossyTransmission = 1u << 4,
    kLobeDeltaReflection = 1u << 5,
    kLobeDeltaTransmission = 1u << 6,
    kLobeDelta = (1u << 5) | (1u << 6)
};

// ---- cuda/random.h:14-40 ----------------------------------------------------------------------------
PB2_HD uint32_t rng_init(uint32_t rounds, uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < rounds; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
PB2_HD float rng_next(uint32_t &s) {
    s = 1664525u * s + 1013904223u;
    return static_cast<float>(s & 0x00FFFFFFu) / 0x01000000;
}

// ---- optix/util.h ----------------------------------------------------------------------------------
PB2_D float3 uniform_sample_triangle(float u1, float u2) { // :33-36
    float su = sqrtf(u1);
    return mk3(1.f - su, su * (1.f - u2), u2 * su);
}
PB2_D float3 uniform_sample_sphere(float u1, float u2) { // :38-43
    float z = 1.f - 2.f * u1;
    float sin_theta = sqrtf(fmaxf(0.f, 1.f - z * z));
    float phi = 2.f * kPi * u2, s, c;
    sincosf(phi, &s, &c);
    return mk3(sin_theta * c, sin_theta * s, z);
}
PB2_D float3 cosine_sample_hemisphere(float u1, float u2) { // :45-54
    float sin_theta = sqrtf(u1);
    float phi = 2.0f * kPi * u2, s, c;
    sincosf(phi, &s, &c);
    return mk3(sin_theta * c, sin_theta * s, sqrtf(fmaxf(0.f, 1.f - sin_theta * sin_theta)));
}
PB2_D float cosine_sample_hemisphere_pdf(float3 v) { return v.z > 0.f ? kInvPi * v.z : 0.f; } // :55-57
PB2_D float3 uniform_sample_hemisphere(float u1, float u2) {                                    // :59-69
    float z = 1.f - 2.f * u1;
    float sin_theta = sqrtf(fmaxf(0.f, 1.f - z * z));
    float phi = 2.0f * kPi * u2, s, c;
    sincosf(phi, &s, &c);
    return mk3(sin_theta * c, sin_theta * s, fabsf(z));
}
PB2_D float uniform_sample_hemisphere_pdf(float3 v) { return v.z > 0.f ? kInvPi * 0.5f : 0.f; } // :70-72
PB2_D float3 reflect_z(float3 v) { return mk3(-v.x, -v.y, v.z); }                                 // :74-78
PB2_D float3 reflect(float3 v, float3 n) { return -v + 2 * dot(v, n) * n; }                       // :80-82
PB2_D float3 refract_z(float3 v, float cos_theta_t, float eta) {                                  // :84-87
    float scale = -(cos_theta_t < 0.f ? 1.f / eta : eta);
    return normalize(mk3(scale * v.x, scale * v.y, cos_theta_t));
}
PB2_D float3 refract(float3 v, float3 n, float cos_theta_t, float eta) { // :89-92
    if (cos_theta_t < 0) eta = 1 / eta;
    return n * (dot(v, n) * eta + cos_theta_t) - v * eta;
}
PB2_D void build_onb(float3 N, float3 &b1, float3 &b2) { // :95-101
    float sign = copysignf(1.f, N.z);
    float a = -1.f / (sign + N.z);
    float b = N.x * N.y * a;
    b1 = mk3(1.f + sign * N.x * N.x * a, sign * b, -sign * N.x);
    b2 = mk3(b, sign + N.y * N.y * a, -N.y);
}
struct Onb { // the frame is built once per shading point and reused by ToLocal / ToWorld (:103-115)
    float3 b1, b2, n;
    PB2_D explicit Onb(float3 N) : n(N) { build_onb(N, b1, b2); }
    PB2_D float3 to_local(float3 v) const { return mk3(dot(v, b1), dot(v, b2), dot(v, n)); }
    PB2_D float3 to_world(float3 v) const { return b1 * v.x + b2 * v.y + n * v.z; }
};
PB2_D float2 sphere_texcoord(float3 p) { // :117-128
    float phi = atan2f(p.y, p.x);
    phi = phi < 0.f ? phi + kPi * 2.f : phi;
    float theta = acosf(p.z);
    return make_float2(phi * kInvPi * 0.5f, theta * kInvPi);
}
PB2_D float luminance(float3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }                  // :161-163
PB2_D float mis_weight(float x, float y) { return x / (x + y); }                                             // :165-167
PB2_D bool is_zero(float v) { return fabsf(v) < kEps; }                                                      // :169-171
PB2_D bool is_zero(float3 v) { return fabsf(v.x) < kEps && fabsf(v.y) < kEps && fabsf(v.z) < kEps; }         // :177-179

// ---- render/material/fresnel.h ---------------------------------------------------------------------
PB2_D float fresnel_dielectric(float eta, float cos_theta_i, float &cos_theta_t) { // :7-25
    float scale = cos_theta_i > 0.f ? 1.f / eta : eta;
    float cos_theta_t2 = 1.f - (1.f - cos_theta_i * cos_theta_i) * (scale * scale);
    if (cos_theta_t2 <= 0.0f) {
        cos_theta_t = 0.0f;
        return 1.0f;
    }
    float o_cos_theta_i = cos_theta_i;
    cos_theta_i = fabsf(cos_theta_i);
    cos_theta_t = sqrtf(fmaxf(0.f, cos_theta_t2));
    float rs = (cos_theta_i - eta * cos_theta_t) / (cos_theta_i + eta * cos_theta_t);
    float rp = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
    cos_theta_t = o_cos_theta_i > 0.f ? -cos_theta_t : cos_theta_t;
    return 0.5f * (rs * rs + rp * rp);
}
PB2_D float fresnel_dielectric(float eta, float cos_theta_i) { // :26-29
    float unused;
    return fresnel_dielectric(eta, cos_theta_i, unused);
}
PB2_D float fresnel_conductor1(float eta, float k, float cos_theta_i) { // :31-49
    float cos_theta_i2 = cos_theta_i * cos_theta_i;
    float sin_theta_i2 = 1.f - cos_theta_i2;
    float sin_theta_i4 = sin_theta_i2 * sin_theta_i2;
    float t1 = eta * eta - k * k - sin_theta_i2;
    float a2pb2 = sqrtf(fmaxf(0.f, t1 * t1 + 4.f * k * k * eta * eta));
    float a = sqrtf(fmaxf(0.f, 0.5f * (a2pb2 + t1)));
    float term1 = a2pb2 + cos_theta_i2;
    float term2 = 2.f * a * cos_theta_i;
    float rs2 = (term1 - term2) / (term1 + term2);
    float term3 = a2pb2 * cos_theta_i2 + sin_theta_i4;
    float term4 = term2 * sin_theta_i2;
    float rp2 = rs2 * (term3 - term4) / (term3 + term4);
    return 0.5f * (rp2 + rs2);
}
PB2_D float3 fresnel_conductor(float3 eta, float3 k, float c) { // :51-56
    return mk3(fresnel_conductor1(eta.x, k.x, c), fresnel_conductor1(eta.y, k.y, c), fresnel_conductor1(eta.z, k.z, c));
}

// ---- render/material/ggx.h (isotropic; GGX_Sample_Visible_Area is defined at :6) -----------------------
PB2_D float ggx_lambda(float3 w, float alpha) { // :10-14
    float a2 = alpha * alpha;
    float3 v2 = w * w;
    return (-1.f + sqrtf(1.f + (v2.x + v2.y) * a2 / v2.z)) / 2.f;
}
PB2_D float ggx_g1(float3 w, float alpha) { return 1.f / (1.f + ggx_lambda(w, alpha)); }                 // :16-18
PB2_D float ggx_g(float3 wi, float3 wo, float alpha) { return ggx_g1(wi, alpha) * ggx_g1(wo, alpha); }   // :20-22
PB2_D float ggx_d(float3 wh, float alpha) {                                                               // :24-29
    float a2 = alpha * alpha;
    float3 v2 = wh * wh;
    float t = (v2.x + v2.y) / a2 + v2.z;
    return 1.f / (kPi * a2 * t * t);
}
PB2_D float ggx_pdf(float3 wo, float3 wh, float alpha) { // :31-37
    return ggx_d(wh, alpha) * ggx_g1(wo, alpha) * dot(wo, wh) / fabsf(wo.z);
}
PB2_D float3 ggx_sample(float3 wo, float alpha, float2 xi) { // :39-57
    float3 vh = normalize(mk3(alpha * wo.x, alpha * wo.y, wo.z));
    float3 T1 = wo.z < 0.9999f ? normalize(cross(mk3(0.f, 0.f, 1.f), vh)) : mk3(1.f, 0.f, 0.f);
    float3 T2 = cross(vh, T1);
    float r = sqrtf(xi.x);
    float phi = 2.f * kPi * xi.y, sp, cp;
    sincosf(phi, &sp, &cp);
    float t1 = r * cp;
    float t2 = r * sp;
    float s = 0.5f * (1.f + vh.z);
    t2 = (1.f - s) * sqrtf(1.f - t1 * t1) + s * t2;
    float3 nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.f, 1.f - t1 * t1 - t2 * t2)) * vh;
    float3 ne = mk3(alpha * nh.x, alpha * nh.y, fmaxf(0.f, nh.z));
    return normalize(ne);
}

// ---- cuda/texture.h:33-57 -----------------------------------------------------------------------------
PB2_D float3 tex_sample(const DevTexture *t, float2 uv) {
    const float4 hdr = __ldg(&t->hdr);
    if (__float_as_int(hdr.x) == PB2_TEX_RGB) return mk3(hdr.y, hdr.z, hdr.w);
    const float4 r0 = __ldg(&t->r0), r1 = __ldg(&t->r1);
    const float4 tex = make_float4(uv.x, uv.y, 0.f, 1.f);
    float tex_x = dot(r0, tex);
    float tex_y = dot(r1, tex);
    if (__float_as_int(hdr.x) == PB2_TEX_BITMAP) { // :52-54, the texture unit filters and wraps
        const cudaTextureObject_t obj = (unsigned long long)__float_as_uint(hdr.y) | ((unsigned long long)__float_as_uint(hdr.z) << 32);
        const float4 c = tex2D<float4>(obj, tex_x, tex_y);
        return mk3(c.x, c.y, c.z);
    }
    const float4 b = __ldg(&t->b);
    tex_x = tex_x - (tex_x > 0.f ? floorf(tex_x) : ceilf(tex_x));
    tex_y = tex_y - (tex_y > 0.f ? floorf(tex_y) : ceilf(tex_y));
    if (tex_x < 0.f) tex_x += 1.f;
    if (tex_y < 0.f) tex_y += 1.f;
    const float3 p1 = mk3(hdr.y, hdr.z, hdr.w), p2 = mk3(b.x, b.y, b.z);
    if (tex_x > 0.5f) return tex_y > 0.5f ? p1 : p2;
    return tex_y > 0.5f ? p2 : p1;
}

// ---- optix::material::Material::LocalBsdf as one flat register record -----------------------------
// c0/c1/c2 are the (up to) three sampled colour textures in slot order (pb2.h, pb2_material).
struct LocalBsdf {
    int type;
    float alpha, eta, int_fdr, specular_sampling_weight;
    bool nonlinear;
    float3 c0, c1, c2;
};
// Material::GetLocalBsdf, render/material/optix_material.h:117-130
// forced_type >= 0: the caller knows every material of the scene has this type (a compile-time constant there), so the switches
// here and in bsdf_sample / bsdf_eval fold to one case
PB2_D LocalBsdf get_local_bsdf(const DevMaterial *m, float2 uv, int forced_type = -1) {
    LocalBsdf b;
    const int4 h0 = __ldg(reinterpret_cast<const int4 *>(m));
    const float2 h1 = __ldg(reinterpret_cast<const float2 *>(m) + 2);
    b.type = forced_type >= 0 ? forced_type : h0.x;
    b.eta = __int_as_float(h0.z);
    b.nonlinear = h0.w != 0;
    b.int_fdr = h1.x, b.specular_sampling_weight = h1.y;
    b.alpha = 0.f;
    b.c0 = b.c1 = b.c2 = mk3(0.f);
    switch (b.type) {
        case PB2_MAT_DIFFUSE: b.c0 = tex_sample(&m->tex[0], uv); break;
        case PB2_MAT_DIELECTRIC:
        case PB2_MAT_PLASTIC:
            b.c0 = tex_sample(&m->tex[0], uv), b.c1 = tex_sample(&m->tex[1], uv);
            break;
        case PB2_MAT_ROUGH_DIELECTRIC:
        case PB2_MAT_ROUGH_PLASTIC:
            b.c0 = tex_sample(&m->tex[0], uv), b.c1 = tex_sample(&m->tex[1], uv), b.alpha = tex_sample(&m->tex[2], uv).x;
            break;
        case PB2_MAT_CONDUCTOR:
            b.c0 = tex_sample(&m->tex[0], uv), b.c1 = tex_sample(&m->tex[1], uv), b.c2 = tex_sample(&m->tex[2], uv);
            break;
        case PB2_MAT_ROUGH_CONDUCTOR:
            b.c0 = tex_sample(&m->tex[0], uv), b.c1 = tex_sample(&m->tex[1], uv), b.c2 = tex_sample(&m->tex[2], uv);
            b.alpha = tex_sample(&m->tex[3], uv).x;
            break;
        default: break;
    }
    return b;
}
// LocalBsdf::GetAlbedo (optix_material.h:93-111): always slot 0 by construction of the slot table
PB2_D float3 local_albedo(const LocalBsdf &b) { return b.type >= PB2_MAT_DIFFUSE && b.type <= PB2_MAT_ROUGH_PLASTIC ? b.c0 : mk3(0.f); }

struct BsdfRec { // BsdfSamplingRecord, render/material/bsdf/bsdf.h:26-36
    float3 wi, wo, f;
    float pdf;
    uint32_t type;
};

// ---- diffuse.h:12-35 : c0 = reflectance ----
PB2_D void diffuse_eval(const LocalBsdf &b, BsdfRec &r) {
    const bool ok = r.wi.z > 0.f && r.wo.z > 0.f;
    r.f = ok ? b.c0 * kInvPi : mk3(0.f);
    r.pdf = ok ? cosine_sample_hemisphere_pdf(r.wi) : 0.f;
}
PB2_D void diffuse_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    float x = rng_next(rng), y = rng_next(rng);
    r.wi = cosine_sample_hemisphere(x, y);
    diffuse_eval(b, r);
    r.type = kLobeDiffuseReflection;
}
// ---- conductor.h:14-35 : c0 = specular_reflectance, c1 = eta, c2 = k ----
PB2_D void conductor_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &) {
    r.wi = reflect_z(r.wo);
    r.pdf = 1.f;
    r.f = b.c0 * fresnel_conductor(b.c1, b.c2, r.wo.z) / fabsf(r.wi.z);
    r.type = kLobeDeltaReflection;
}
// ---- dielectric.h:15-45 : c0 = specular_reflectance, c1 = specular_transmittance ----
PB2_D void dielectric_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    float cos_theta_t;
    float fr = fresnel_dielectric(b.eta, r.wo.z, cos_theta_t);
    if (rng_next(rng) < fr) {
        r.wi = reflect_z(r.wo);
        r.pdf = fr;
        r.f = b.c0 * fr / fabsf(r.wi.z);
        r.type = kLobeDeltaReflection;
    } else {
        r.wi = refract_z(r.wo, cos_theta_t, b.eta);
        r.pdf = 1.f - fr;
        float factor = cos_theta_t < 0.f ? 1.f / b.eta : b.eta;
        r.f = b.c1 * (1.f - fr) * factor * factor / fabsf(r.wi.z);
        r.type = kLobeDeltaTransmission;
    }
}
// ---- rough_conductor.h:15-47 : c0 = specular_reflectance, c1 = eta, c2 = k ----
PB2_D void rough_conductor_eval(const LocalBsdf &b, BsdfRec &r) {
    r.f = mk3(0.f), r.pdf = 0.f;
    if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
    float3 wh = normalize(r.wi + r.wo);
    float3 wh2 = normalize(wh); // GetPdf renormalises (:37-38)
    r.pdf = ggx_pdf(r.wo, wh2, b.alpha) / (4.f * dot(r.wo, wh2));
    float3 fr = fresnel_conductor(b.c1, b.c2, dot(r.wo, wh));
    r.f = b.c0 * ggx_d(wh, b.alpha) * fr * ggx_g(r.wi, r.wo, b.alpha) / (4.f * r.wi.z * r.wo.z);
}
PB2_D void rough_conductor_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    float x = rng_next(rng), y = rng_next(rng);
    r.wi = reflect(r.wo, ggx_sample(r.wo, b.alpha, make_float2(x, y)));
    rough_conductor_eval(b, r);
    r.type = kLobeDiffuseReflection; // sic (:45)
}
// ---- rough_dielectric.h:15-97 : c0 = specular_reflectance, c1 = specular_transmittance ----
PB2_D void rough_dielectric_f(const LocalBsdf &b, BsdfRec &r) {
    r.f = mk3(0.f);
    if (is_zero(r.wo.z)) return;
    float3 wh;
    bool sample_reflect = r.wo.z * r.wi.z > 0.f;
    if (sample_reflect) wh = normalize(r.wo + r.wi);
    else wh = normalize(r.wo + r.wi * (r.wo.z > 0.f ? b.eta : 1.f / b.eta));
    wh = wh * (wh.z > 0.f ? 1.f : -1.f);
    float F = fresnel_dielectric(b.eta, dot(r.wo, wh));
    float G = ggx_g(r.wi, r.wo, b.alpha);
    float D = ggx_d(wh, b.alpha);
    if (sample_reflect) {
        r.f = b.c0 * F * G * D / (4.f * fabsf(r.wi.z) * fabsf(r.wo.z));
    } else {
        float _eta = r.wo.z > 0.f ? b.eta : 1.f / b.eta;
        float sqrt_denom = dot(r.wo, wh) + _eta * dot(r.wi, wh);
        r.f = b.c1 * fabsf((1.f - F) * D * G * dot(r.wi, wh) * dot(r.wo, wh) / (sqrt_denom * sqrt_denom * r.wi.z * r.wo.z));
    }
}
PB2_D void rough_dielectric_pdf(const LocalBsdf &b, BsdfRec &r) {
    r.pdf = 0.f;
    bool sample_reflect = r.wo.z * r.wi.z > 0.f;
    float3 wh;
    float dwh_dwo;
    if (sample_reflect) {
        wh = normalize(r.wo + r.wi);
        dwh_dwo = 1.f / (4.f * dot(r.wi, wh));
    } else {
        float _eta = r.wo.z > 0.f ? b.eta : 1.f / b.eta;
        wh = normalize(r.wo + r.wi * _eta);
        float sqrt_denom = dot(r.wo, wh) + _eta * dot(r.wi, wh);
        dwh_dwo = (_eta * _eta * dot(r.wi, wh)) / (sqrt_denom * sqrt_denom);
    }
    wh = wh * (wh.z > 0.f ? 1.f : -1.f);
    float3 wo = r.wo * (r.wo.z > 0.f ? 1.f : -1.f);
    float F = fresnel_dielectric(b.eta, dot(r.wo, wh));
    r.pdf = fabsf(ggx_pdf(wo, wh, b.alpha) * (sample_reflect ? F : 1.f - F) * dwh_dwo);
}
PB2_D void rough_dielectric_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    float x = rng_next(rng), y = rng_next(rng);
    float3 wo = r.wo * (r.wo.z > 0.f ? 1.f : -1.f);
    float3 wh = ggx_sample(wo, b.alpha, make_float2(x, y));
    float cos_theta_t = 0.f;
    float F = fresnel_dielectric(b.eta, dot(r.wo, wh), cos_theta_t);
    if (rng_next(rng) < F) {
        r.wi = reflect(r.wo, wh);
        r.type = kLobeGlossyReflection;
    } else {
        if (is_zero(cos_theta_t)) return;
        r.wi = refract(r.wo, wh, cos_theta_t, b.eta);
        r.type = kLobeGlossyTransmission;
        if (r.wi.z * r.wo.z >= 0.f) return;
    }
    rough_dielectric_pdf(b, r);
    rough_dielectric_f(b, r);
}
// ---- plastic.h:23-81 and rough_plastic.h:22-86 : c0 = diffuse_reflectance, c1 = specular_reflectance ----
PB2_D float3 plastic_diff(const LocalBsdf &b) { return b.c0 / (1.f - (b.nonlinear ? b.c0 * b.int_fdr : mk3(b.int_fdr))); }
PB2_D float plastic_specular_prob(const LocalBsdf &b, float fresnel_o) {
    return (fresnel_o * b.specular_sampling_weight) /
           (fresnel_o * b.specular_sampling_weight + (1 - fresnel_o) * (1.f - b.specular_sampling_weight));
}
PB2_D void plastic_eval(const LocalBsdf &b, BsdfRec &r) {
    r.f = mk3(0.f), r.pdf = 0.f;
    if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
    float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
    float fresnel_i = fresnel_dielectric(b.eta, r.wi.z);
    r.f = plastic_diff(b) * (1.f - fresnel_i) * (1.f - fresnel_o) * cosine_sample_hemisphere_pdf(r.wi) / (b.eta * b.eta * r.wi.z);
    r.pdf = cosine_sample_hemisphere_pdf(r.wi) * (1.f - plastic_specular_prob(b, fresnel_o));
}
PB2_D void plastic_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    if (r.wo.z <= 0.f) return;
    float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
    float x = rng_next(rng), y = rng_next(rng);
    float specular_prob = plastic_specular_prob(b, fresnel_o);
    if (x < specular_prob) {
        r.type = kLobeDeltaReflection;
        r.wi = reflect_z(r.wo);
        r.f = b.c1 * fresnel_o / r.wi.z;
        r.pdf = specular_prob;
    } else {
        r.type = kLobeDiffuseReflection;
        r.wi = cosine_sample_hemisphere((x - specular_prob) / (1.f - specular_prob), y);
        float fresnel_i = fresnel_dielectric(b.eta, r.wi.z);
        r.f = plastic_diff(b) * (1.f - fresnel_i) * (1.f - fresnel_o) * cosine_sample_hemisphere_pdf(r.wi) / (b.eta * b.eta * r.wi.z);
        r.pdf = cosine_sample_hemisphere_pdf(r.wi) * (1.f - specular_prob);
    }
}
PB2_D void rough_plastic_eval(const LocalBsdf &b, BsdfRec &r) {
    r.f = mk3(0.f), r.pdf = 0.f;
    if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
    float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
    float3 wh = normalize(r.wi + r.wo);
    r.f = b.c1 * fresnel_dielectric(b.eta, dot(wh, r.wo)) * ggx_d(wh, b.alpha) * ggx_g(r.wi, r.wo, b.alpha) / (4.f * r.wo.z * r.wi.z);
    float fresnel_i = fresnel_dielectric(b.eta, r.wi.z);
    r.f += plastic_diff(b) * (1.f - fresnel_i) * (1.f - fresnel_o) * kInvPi / (b.eta * b.eta);
    float specular_prob = plastic_specular_prob(b, fresnel_o);
    float diffuse_prob = 1.f - specular_prob;
    r.pdf = specular_prob * ggx_pdf(r.wo, wh, b.alpha) / (4.f * dot(r.wi, wh));
    r.pdf += diffuse_prob * cosine_sample_hemisphere_pdf(r.wi);
}
PB2_D void rough_plastic_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    r.wi = mk3(0.f);
    if (r.wo.z <= 0.f) return;
    float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
    float specular_prob = plastic_specular_prob(b, fresnel_o);
    float x = rng_next(rng), y = rng_next(rng);
    if (y < specular_prob) {
        y /= specular_prob;
        float3 wh = ggx_sample(r.wo, b.alpha, make_float2(x, y));
        r.wi = reflect(r.wo, wh);
        r.type = kLobeGlossyReflection;
    } else {
        y = (y - specular_prob) / (1.f - specular_prob);
        r.wi = cosine_sample_hemisphere(x, y);
        r.type = kLobeDiffuseReflection;
    }
    rough_plastic_eval(b, r);
}

// LocalBsdf::Sample / Eval, render/material/optix_material.h:70-91
PB2_D void bsdf_sample(const LocalBsdf &b, BsdfRec &r, uint32_t &rng) {
    r.wi = mk3(0.f), r.f = mk3(0.f), r.pdf = 0.f, r.type = kLobeUnknown;
    switch (b.type) {
        case PB2_MAT_DIFFUSE: diffuse_sample(b, r, rng); break;
        case PB2_MAT_DIELECTRIC: dielectric_sample(b, r, rng); break;
        case PB2_MAT_ROUGH_DIELECTRIC: rough_dielectric_sample(b, r, rng); break;
        case PB2_MAT_CONDUCTOR: conductor_sample(b, r, rng); break;
        case PB2_MAT_ROUGH_CONDUCTOR: rough_conductor_sample(b, r, rng); break;
        case PB2_MAT_PLASTIC: plastic_sample(b, r, rng); break;
        case PB2_MAT_ROUGH_PLASTIC: rough_plastic_sample(b, r, rng); break;
        default: break;
    }
}
PB2_D void bsdf_eval(const LocalBsdf &b, BsdfRec &r) {
    r.f = mk3(0.f), r.pdf = 0.f;
    switch (b.type) {
        case PB2_MAT_DIFFUSE: diffuse_eval(b, r); break;
        case PB2_MAT_ROUGH_DIELECTRIC: rough_dielectric_f(b, r), rough_dielectric_pdf(b, r); break;
        case PB2_MAT_ROUGH_CONDUCTOR: rough_conductor_eval(b, r); break;
        case PB2_MAT_PLASTIC: plastic_eval(b, r); break;
        case PB2_MAT_ROUGH_PLASTIC: rough_plastic_eval(b, r); break;
        default: break; // dielectric, conductor: delta lobes evaluate to 0
    }
}

// ---- emitters ------------------------------------------------------------------------------------------
// EnvMapEmitter fields as packed into DevEmitter by pb2_api.cu (to_dev)
struct EnvMapView {
    float3 to_world0, to_world1, to_world2, to_local0, to_local1, to_local2;
    float scale, normalization;
    uint32_t w, h;
    const float *row_cdf, *row_weight, *col_cdf;
    PB2_D explicit EnvMapView(const DevEmitter *e) {
        const float4 p0 = __ldg(&e->p0), p1 = __ldg(&e->p1), p2 = __ldg(&e->p2), n0 = __ldg(&e->n0), n1 = __ldg(&e->n1), n2 = __ldg(&e->n2);
        to_world0 = mk3(p0), to_world1 = mk3(p1), to_world2 = mk3(p2), to_local0 = mk3(n0), to_local1 = mk3(n1), to_local2 = mk3(n2);
        scale = p0.w, w = __float_as_uint(p1.w), h = __float_as_uint(p2.w);
        normalization = __ldg(&e->area);
        const float *tables = reinterpret_cast<const float *>((unsigned long long)__float_as_uint(n0.w) | ((unsigned long long)__float_as_uint(n1.w) << 32));
        row_cdf = tables, row_weight = tables + (h + 1u), col_cdf = tables + (2u * h + 1u);
    }
};
// smallest i in [0, n) with x <= cdf[i]; n when there is none (cdf non-decreasing)
PB2_D uint32_t first_not_less(const float *cdf, uint32_t n, float x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (x <= __ldg(cdf + mid)) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
struct EmitSample { // EmitterSampleRecord, render/emitter/types.h:17-26
    float3 radiance, wi;
    float distance, pdf;
    bool is_delta;
};
// Emitter::SampleDirect: area.h:17-35, sphere.h:14-32, env.h:70-80
PB2_D void emitter_sample_direct(const DevEmitter *e, float3 hit_pos, float3 hit_n, const Onb &frame, float2 xi, EmitSample &out) {
    const int4 h = __ldg(reinterpret_cast<const int4 *>(e));
    const int type = h.x;
    const float area = __int_as_float(h.w);
    out.pdf = 0.f, out.distance = 0.f, out.is_delta = false;
    out.radiance = mk3(0.f), out.wi = mk3(0.f);
    if (type == PB2_EMIT_TRI || type == PB2_EMIT_SPHERE) {
        float3 position, normal;
        float2 tex;
        if (type == PB2_EMIT_TRI) {
            const float4 p0 = __ldg(&e->p0), p1 = __ldg(&e->p1), p2 = __ldg(&e->p2);
            const float4 n0 = __ldg(&e->n0), n1 = __ldg(&e->n1), n2 = __ldg(&e->n2);
            float3 t = uniform_sample_triangle(xi.x, xi.y);
            position = mk3(p0) * t.x + mk3(p1) * t.y + mk3(p2) * t.z;
            normal = normalize(mk3(n0) * t.x + mk3(n1) * t.y + mk3(n2) * t.z);
            tex = make_float2(p0.w, p1.w) * t.x + make_float2(p2.w, n0.w) * t.y + make_float2(n1.w, n2.w) * t.z;
        } else {
            const float4 cr = __ldg(&e->center_r);
            float3 t = uniform_sample_sphere(xi.x, xi.y);
            position = t * cr.w + mk3(cr);
            normal = normalize(t);
            tex = sphere_texcoord(t);
        }
        out.radiance = tex_sample(&e->radiance, tex);
        out.wi = normalize(position - hit_pos);
        float NoL = dot(hit_n, out.wi);
        float LNoL = dot(normal, -out.wi);
        if (NoL > 0.f && LNoL > 0.f) {
            float distance = length(position - hit_pos);
            out.pdf = distance * distance / (LNoL * area);
            out.distance = distance;
        }
    } else if (type == PB2_EMIT_CONST_ENV) {
        float3 local_wi = uniform_sample_hemisphere(xi.x, xi.y);
        out.wi = frame.to_world(local_wi);
        out.pdf = uniform_sample_hemisphere_pdf(local_wi);
        out.distance = kMaxDistance;
        const float4 hdr = __ldg(&e->radiance.hdr);
        out.radiance = mk3(hdr.y, hdr.z, hdr.w);
    } else if (type == PB2_EMIT_ENV_MAP) { // EnvMapEmitter::SampleDirect, env.h:24-48
        const EnvMapView em(e);
        // "first index whose cdf value is >= xi" over the same index ranges as the reference's linear scans (:25-31);
        // the tables are non-decreasing, so a binary search returns the same index
        const uint32_t row_index = first_not_less(em.row_cdf, em.h, xi.x); // candidates 0..h-1, h when none
        const uint32_t row_t = min(row_index, em.h - 1u);                  // the reference reads past its tables for row_index == h
        const uint32_t col_index = first_not_less(em.col_cdf + (size_t)row_t * (em.w + 1u), em.w - 1u, xi.y);
        const float phi = col_index * kPi * 2.f / em.w;
        const float theta = row_index * kPi / em.h;
        const float3 local_wi = mk3(sinf(theta) * sinf(kPi - phi), cosf(theta), sinf(theta) * cosf(kPi - phi));
        out.wi = mk3(dot(em.to_world0, local_wi), dot(em.to_world1, local_wi), dot(em.to_world2, local_wi));
        out.distance = kMaxDistance;
        const float2 tex = make_float2(phi * 0.5f * kInvPi, theta * kInvPi);
        out.radiance = tex_sample(&e->radiance, tex) * em.scale;
        out.pdf = luminance(out.radiance) * __ldg(em.row_weight + row_t) * em.normalization / fmaxf(1e-4f, fabsf(sinf(theta)));
        if (out.pdf < 0.f) out.pdf = 0.f;
    }
}
// Emitter::Eval: area.h:37-45, sphere.h:34-42, env.h:82-85
PB2_D void emitter_eval(const DevEmitter *e, float3 emit_pos, float3 emit_n, float2 emit_uv, float3 scatter_pos, float3 &radiance, float &pdf) {
    const int4 h = __ldg(reinterpret_cast<const int4 *>(e));
    radiance = mk3(0.f), pdf = 0.f;
    if (h.x == PB2_EMIT_TRI || h.x == PB2_EMIT_SPHERE) {
        float3 dir = normalize(scatter_pos - emit_pos);
        float LNoL = dot(emit_n, dir);
        if (LNoL > 0.f) {
            float distance = length(scatter_pos - emit_pos);
            pdf = distance * distance / (LNoL * __int_as_float(h.w));
            radiance = tex_sample(&e->radiance, emit_uv);
        }
    } else if (h.x == PB2_EMIT_CONST_ENV) {
        pdf = 0.25f * kInvPi;
        const float4 hdr = __ldg(&e->radiance.hdr);
        radiance = mk3(hdr.y, hdr.z, hdr.w);
    } else if (h.x == PB2_EMIT_ENV_MAP) { // EnvMapEmitter::Eval, env.h:50-64
        const EnvMapView em(e);
        float3 dir = normalize(emit_pos - scatter_pos);
        dir = mk3(dot(em.to_local0, dir), dot(em.to_local1, dir), dot(em.to_local2, dir));
        const float phi = kPi - atan2f(dir.x, dir.z);
        const float theta = acosf(dir.y);
        const float2 tex = make_float2(phi * 0.5f * kInvPi, theta * kInvPi);
        uint32_t row_index = static_cast<uint32_t>(tex.y * em.h);
        row_index = min(max(row_index, 0u), em.h - 2u);
        radiance = tex_sample(&e->radiance, tex) * em.scale;
        const float w0 = __ldg(em.row_weight + row_index), w1 = __ldg(em.row_weight + row_index + 1);
        pdf = luminance(radiance) * lerp1(w0, w1, tex.y * em.h - 1.f * row_index) * em.normalization / fmaxf(1e-4f, fabsf(sinf(theta)));
    }
}
// Emitter::GetRadiance, render/emitter.h:54-71
PB2_D float3 emitter_radiance(const DevEmitter *e, float2 uv) {
    if (__ldg(&e->type) == PB2_EMIT_CONST_ENV) {
        const float4 hdr = __ldg(&e->radiance.hdr);
        return mk3(hdr.y, hdr.z, hdr.w);
    }
    return tex_sample(&e->radiance, uv); // area, sphere and env map alike (emitter.h:54-71: no `scale` for the env map)
}
// EmitterGroup::SelectOneEmiiter, render/emitter.h:110-136: the first area emitter (in table order) with
// p <= (running fp32 sum of select probabilities up to and including it), else the environment emitter, else the last
// area emitter.  `cdf` holds exactly those running sums (pb2_api.cu, area_select_cdf), so a binary search returns the
// entry the reference's linear scan stops at.  Returns nullptr when the scene has no emitter.
PB2_D const DevEmitter *select_emitter(const DevEmitter *areas, const float *cdf, uint32_t n, const DevEmitter *env, float p) {
    const uint32_t i = first_not_less(cdf, n, p);
    if (i < n) return &areas[i];
    if (env) return env;
    return n ? &areas[n - 1] : nullptr;
}
}// namespace pb2
