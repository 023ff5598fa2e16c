/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * "port" arithmetic backend: the reference's device-side math restated from scratch in plain
 * C++.  Each function cites the reference file:line (relative to /root/reference/framework)
 * it follows.  The twin backend orc_backend_ref.h exposes the same interface but calls the
 * reference's own headers; tests/test_oracle_vs_reference.py holds the two bit-for-bit equal
 * on dense input grids, which is what pins this restatement.
 *
 * Places where the reference reads uninitialised memory are DEFINED here (and identically in
 * the product):
 *   - BsdfSamplingRecord::wi starts as (0,0,0)            (plastic.h:56 returns without writing it)
 *   - EmitterSampleRecord::{is_delta=false, distance=0}   (emitter/types.h:17-26, area.h:17-35)
 *   - EmitEvalRecord::{pdf=0, radiance=0} when the emitter faces away (area.h:37-45)
 */
#ifndef ORC_BACKEND_PORT_H
#define ORC_BACKEND_PORT_H
#include "orc_types.h"
#include "orc_vec.h"
#include "orc_tex2d.h"

namespace orc {

constexpr float kEps = 0.000001f;        // optix/util.h:8
constexpr float kMaxDistance = 1e16f;    // optix/util.h:9

enum : uint32_t { // render/material/bsdf/bsdf.h:7-24
    kLobeUnknown = 0,
    kLobeDiffuseReflection = 1u << 1,
    kLobeGlossyReflection = 1u << 3,
    kLobeGlossyTransmission = 1u << 4,
    kLobeDeltaReflection = 1u << 5,
    kLobeDeltaTransmission = 1u << 6,
    kLobeDelta = (1u << 5) | (1u << 6)
};

struct PortBackend {
    static const char *name() { return "port"; }

    // ---- cuda/random.h:14-40 : TEA init + LCG, integer-exact -------------------------------
    static uint32_t rng_init(uint32_t rounds, uint32_t val0, uint32_t val1) {
        uint32_t v0 = val0, v1 = val1, s0 = 0;
        for (uint32_t n = 0; n < rounds; n++) {
            s0 += 0x9e3779b9u;
            v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
            v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
        }
        return v0;
    }
    static float rng_next(uint32_t &s) {
        s = 1664525u * s + 1013904223u;
        return static_cast<float>(s & 0x00FFFFFFu) / 0x01000000;
    }

    // ---- optix/util.h:33-183 ---------------------------------------------------------------
    static f3 uniform_sample_triangle(float u1, float u2) { // :33-36
        float su = sqrtf(u1);
        return mk3(1.f - su, su * (1.f - u2), u2 * su);
    }
    static f3 uniform_sample_sphere(float u1, float u2) { // :38-43
        float z = 1.f - 2.f * u1;
        float sin_theta = sqrtf(fmaxf(0.f, 1.f - z * z));
        float phi = 2.f * kPi * u2;
        return mk3(sin_theta * cosf(phi), sin_theta * sinf(phi), z);
    }
    static f3 cosine_sample_hemisphere(float u1, float u2) { // :45-54
        float sin_theta = sqrtf(u1);
        float phi = 2.0f * kPi * u2;
        f3 p;
        p.x = sin_theta * cosf(phi);
        p.y = sin_theta * sinf(phi);
        p.z = sqrtf(fmaxf(0.f, 1.f - sin_theta * sin_theta));
        return p;
    }
    static float cosine_sample_hemisphere_pdf(f3 v) { return v.z > 0.f ? kInvPi * v.z : 0.f; } // :55-57
    static f3 uniform_sample_hemisphere(float u1, float u2) { // :59-69
        float z = 1.f - 2.f * u1;
        float sin_theta = sqrtf(fmaxf(0.f, 1.f - z * z));
        float phi = 2.0f * kPi * u2;
        return mk3(sin_theta * cosf(phi), sin_theta * sinf(phi), fabsf(z));
    }
    static float uniform_sample_hemisphere_pdf(f3 v) { return v.z > 0.f ? kInvPi * 0.5f : 0.f; } // :70-72
    static f3 reflect_z(f3 v) { return mk3(-v.x, -v.y, v.z); }                                    // :74-78
    static f3 reflect(f3 v, f3 n) { return -v + 2 * dot(v, n) * n; }                              // :80-82
    static f3 refract_z(f3 v, float cos_theta_t, float eta) {                                     // :84-87
        float scale = -(cos_theta_t < 0.f ? 1.f / eta : eta);
        return normalize(mk3(scale * v.x, scale * v.y, cos_theta_t));
    }
    static f3 refract(f3 v, f3 n, float cos_theta_t, float eta) { // :89-92
        if (cos_theta_t < 0) eta = 1 / eta;
        return n * (dot(v, n) * eta + cos_theta_t) - v * eta;
    }
    static void build_onb(f3 N, f3 &b1, f3 &b2) { // :95-101 (Pixar branchless ONB)
        float sign = copysignf(1.f, N.z);
        float a = -1.f / (sign + N.z);
        float b = N.x * N.y * a;
        b1 = mk3(1.f + sign * N.x * N.x * a, sign * b, -sign * N.x);
        b2 = mk3(b, sign + N.y * N.y * a, -N.y);
    }
    static f3 to_local(f3 v, f3 N) { // :103-108
        f3 b1, b2;
        build_onb(N, b1, b2);
        return mk3(dot(v, b1), dot(v, b2), dot(v, N));
    }
    static f3 to_world(f3 v, f3 N) { // :110-115
        f3 b1, b2;
        build_onb(N, b1, b2);
        return b1 * v.x + b2 * v.y + N * v.z;
    }
    static f2 sphere_texcoord(f3 p) { // :117-128
        float phi = atan2f(p.y, p.x);
        phi = phi < 0.f ? phi + kPi * 2.f : phi;
        float theta = acosf(p.z);
        return mk2(phi * kInvPi * 0.5f, theta * kInvPi);
    }
    static float luminance(f3 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; } // :161-163
    static float mis_weight(float x, float y) { return x / (x + y); }                        // :165-167 balance heuristic
    static bool is_zero(float v) { return fabsf(v) < kEps; }                                 // :169-171
    static bool is_zero(f3 v) { return fabsf(v.x) < kEps && fabsf(v.y) < kEps && fabsf(v.z) < kEps; } // :177-179

    // ---- render/material/fresnel.h ---------------------------------------------------------
    static float fresnel_dielectric(float eta, float cos_theta_i, float &cos_theta_t) { // :7-25
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
    static float fresnel_dielectric(float eta, float cos_theta_i) { // :26-29
        float unused;
        return fresnel_dielectric(eta, cos_theta_i, unused);
    }
    static float fresnel_conductor1(float eta, float k, float cos_theta_i) { // :31-49
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
    static f3 fresnel_conductor(f3 eta, f3 k, float c) { // :51-56
        return mk3(fresnel_conductor1(eta.x, k.x, c), fresnel_conductor1(eta.y, k.y, c), fresnel_conductor1(eta.z, k.z, c));
    }
    static float fresnel_diffuse(float eta) { // :58-85
        if (eta < 1) {
            return -1.4399f * (eta * eta) + 0.7099f * eta + 0.6681f + 0.0636f / eta;
        } else {
            float inv_eta = 1.0f / eta;
            float inv_eta2 = inv_eta * inv_eta;
            float inv_eta3 = inv_eta2 * inv_eta;
            float inv_eta4 = inv_eta3 * inv_eta;
            float inv_eta5 = inv_eta4 * inv_eta;
            return 0.919317f - 3.4793f * inv_eta + 6.75335f * inv_eta2 - 7.80989f * inv_eta3 + 4.98554f * inv_eta4 - 1.36881f * inv_eta5;
        }
    }

    // ---- render/material/ggx.h (isotropic, GGX_Sample_Visible_Area defined at :6) ------------
    static float ggx_lambda(f3 w, float alpha) { // :10-14
        float a2 = alpha * alpha;
        f3 v2 = w * w;
        return (-1.f + sqrtf(1.f + (v2.x + v2.y) * a2 / v2.z)) / 2.f;
    }
    static float ggx_g1(f3 w, float alpha) { return 1.f / (1.f + ggx_lambda(w, alpha)); }        // :16-18
    static float ggx_g(f3 wi, f3 wo, float alpha) { return ggx_g1(wi, alpha) * ggx_g1(wo, alpha); } // :20-22
    static float ggx_d(f3 wh, float alpha) { // :24-29
        float a2 = alpha * alpha;
        f3 v2 = wh * wh;
        float t = (v2.x + v2.y) / a2 + v2.z;
        return 1.f / (kPi * a2 * t * t);
    }
    static float ggx_pdf(f3 wo, f3 wh, float alpha) { // :31-37
        return ggx_d(wh, alpha) * ggx_g1(wo, alpha) * dot(wo, wh) / fabsf(wo.z);
    }
    static f3 ggx_sample(f3 wo, float alpha, f2 xi) { // :39-57 (Heitz 2018 VNDF)
        f3 vh = normalize(mk3(alpha * wo.x, alpha * wo.y, wo.z));
        f3 T1 = wo.z < 0.9999f ? normalize(cross(mk3(0.f, 0.f, 1.f), vh)) : mk3(1.f, 0.f, 0.f);
        f3 T2 = cross(vh, T1);
        float r = sqrtf(xi.x);
        float phi = 2.f * kPi * xi.y;
        float t1 = r * cosf(phi);
        float t2 = r * sinf(phi);
        float s = 0.5f * (1.f + vh.z);
        t2 = (1.f - s) * sqrtf(1.f - t1 * t1) + s * t2;
        f3 nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.f, 1.f - t1 * t1 - t2 * t2)) * vh;
        f3 ne = mk3(alpha * nh.x, alpha * nh.y, fmaxf(0.f, nh.z));
        return normalize(ne);
    }

    // ---- cuda/texture.h:33-57 (bitmap: tex2D emulated by orc_tex2d.h) ---------------------------
    static f3 tex_sample(const orc_texture &t, f2 uv) {
        f4 tex{ uv.x, uv.y, 0.f, 1.f };
        float tex_x = dot(f4{ t.to_uv[0], t.to_uv[1], t.to_uv[2], t.to_uv[3] }, tex);
        float tex_y = dot(f4{ t.to_uv[4], t.to_uv[5], t.to_uv[6], t.to_uv[7] }, tex);
        if (t.type == ORC_TEX_BITMAP) {
            const Tex2dResult c = tex2d(t.bitmap, t.bitmap_w, t.bitmap_h, t.address_mode, t.filter_mode, tex_x, tex_y);
            return mk3(c.x, c.y, c.z);
        }
        if (t.type == ORC_TEX_CHECKERBOARD) {
            tex_x = tex_x - (tex_x > 0.f ? floorf(tex_x) : ceilf(tex_x));
            tex_y = tex_y - (tex_y > 0.f ? floorf(tex_y) : ceilf(tex_y));
            if (tex_x < 0.f) tex_x += 1.f;
            if (tex_y < 0.f) tex_y += 1.f;
            const f3 p1 = mk3(t.a[0], t.a[1], t.a[2]), p2 = mk3(t.b[0], t.b[1], t.b[2]);
            if (tex_x > 0.5f) return tex_y > 0.5f ? p1 : p2;
            return tex_y > 0.5f ? p2 : p1;
        }
        return mk3(t.a[0], t.a[1], t.a[2]);
    }

    // ---- the seven BSDFs: render/material/bsdf/*.h -------------------------------------------
    static f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }
    static void st3(float *p, f3 v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }

    struct Rec {
        f3 wi{ 0.f, 0.f, 0.f }, wo{ 0.f, 0.f, 0.f }, f{ 0.f, 0.f, 0.f };
        float pdf = 0.f;
        uint32_t type = kLobeUnknown;
    };

    // diffuse.h:12-35
    static void diffuse_f(const orc_local_bsdf &b, Rec &r) {
        f3 f = mk3(0.f);
        if (r.wi.z > 0.f && r.wo.z > 0.f) f = ld3(b.reflectance) * kInvPi;
        r.f = f;
    }
    static void diffuse_pdf(const orc_local_bsdf &, Rec &r) {
        float pdf = 0.f;
        if (r.wi.z > 0.f && r.wo.z > 0.f) pdf = cosine_sample_hemisphere_pdf(r.wi);
        r.pdf = pdf;
    }
    static void diffuse_sample(const orc_local_bsdf &b, Rec &r, uint32_t &rng) {
        float x = rng_next(rng), y = rng_next(rng);
        r.wi = cosine_sample_hemisphere(x, y);
        diffuse_pdf(b, r);
        diffuse_f(b, r);
        r.type = kLobeDiffuseReflection;
    }
    // conductor.h:14-35 (delta mirror, Eval == 0)
    static void conductor_sample(const orc_local_bsdf &b, Rec &r, uint32_t &) {
        r.wi = reflect_z(r.wo);
        r.pdf = 1.f;
        f3 fr = fresnel_conductor(ld3(b.eta3), ld3(b.k3), r.wo.z);
        r.f = ld3(b.specular_reflectance) * fr / fabsf(r.wi.z);
        r.type = kLobeDeltaReflection;
    }
    // dielectric.h:15-45 (Eval == 0)
    static void dielectric_sample(const orc_local_bsdf &b, Rec &r, uint32_t &rng) {
        float cos_theta_t;
        float fr = fresnel_dielectric(b.eta, r.wo.z, cos_theta_t);
        if (rng_next(rng) < fr) {
            r.wi = reflect_z(r.wo);
            r.pdf = fr;
            r.f = ld3(b.specular_reflectance) * fr / fabsf(r.wi.z);
            r.type = kLobeDeltaReflection;
        } else {
            r.wi = refract_z(r.wo, cos_theta_t, b.eta);
            r.pdf = 1.f - fr;
            float factor = cos_theta_t < 0.f ? 1.f / b.eta : b.eta;
            r.f = ld3(b.specular_transmittance) * (1.f - fr) * factor * factor / fabsf(r.wi.z);
            r.type = kLobeDeltaTransmission;
        }
    }
    // rough_conductor.h:15-47
    static void rough_conductor_f(const orc_local_bsdf &b, Rec &r) {
        r.f = mk3(0.f);
        if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
        f3 wh = normalize(r.wi + r.wo);
        f3 fr = fresnel_conductor(ld3(b.eta3), ld3(b.k3), dot(r.wo, wh));
        r.f = ld3(b.specular_reflectance) * ggx_d(wh, b.alpha) * fr * ggx_g(r.wi, r.wo, b.alpha) / (4.f * r.wi.z * r.wo.z);
    }
    static void rough_conductor_pdf(const orc_local_bsdf &b, Rec &r) {
        r.pdf = 0.f;
        if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
        f3 wh = normalize(r.wi + r.wo);
        wh = normalize(wh);
        r.pdf = ggx_pdf(r.wo, wh, b.alpha) / (4.f * dot(r.wo, wh));
    }
    static void rough_conductor_sample(const orc_local_bsdf &b, Rec &r, uint32_t &rng) {
        float x = rng_next(rng), y = rng_next(rng);
        r.wi = reflect(r.wo, ggx_sample(r.wo, b.alpha, mk2(x, y)));
        rough_conductor_pdf(b, r);
        rough_conductor_f(b, r);
        r.type = kLobeDiffuseReflection; // sic, rough_conductor.h:45
    }
    // rough_dielectric.h:15-97
    static void rough_dielectric_f(const orc_local_bsdf &b, Rec &r) {
        r.f = mk3(0.f);
        if (is_zero(r.wo.z)) return;
        f3 wh;
        bool sample_reflect = r.wo.z * r.wi.z > 0.f;
        if (sample_reflect) wh = normalize(r.wo + r.wi);
        else wh = normalize(r.wo + r.wi * (r.wo.z > 0.f ? b.eta : 1.f / b.eta));
        wh = wh * (wh.z > 0.f ? 1.f : -1.f);
        float F = fresnel_dielectric(b.eta, dot(r.wo, wh));
        float G = ggx_g(r.wi, r.wo, b.alpha);
        float D = ggx_d(wh, b.alpha);
        if (sample_reflect) {
            r.f = ld3(b.specular_reflectance) * F * G * D / (4.f * fabsf(r.wi.z) * fabsf(r.wo.z));
        } else {
            float _eta = r.wo.z > 0.f ? b.eta : 1.f / b.eta;
            float sqrt_denom = dot(r.wo, wh) + _eta * dot(r.wi, wh);
            r.f = ld3(b.specular_transmittance) *
                  fabsf((1.f - F) * D * G * dot(r.wi, wh) * dot(r.wo, wh) / (sqrt_denom * sqrt_denom * r.wi.z * r.wo.z));
        }
    }
    static void rough_dielectric_pdf(const orc_local_bsdf &b, Rec &r) {
        r.pdf = 0.f;
        bool sample_reflect = r.wo.z * r.wi.z > 0.f;
        f3 wh;
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
        f3 wo = r.wo * (r.wo.z > 0.f ? 1.f : -1.f);
        float F = fresnel_dielectric(b.eta, dot(r.wo, wh));
        r.pdf = fabsf(ggx_pdf(wo, wh, b.alpha) * (sample_reflect ? F : 1.f - F) * dwh_dwo);
    }
    static void rough_dielectric_sample(const orc_local_bsdf &b, Rec &r, uint32_t &rng) {
        float x = rng_next(rng), y = rng_next(rng);
        f3 wo = r.wo * (r.wo.z > 0.f ? 1.f : -1.f);
        f3 wh = ggx_sample(wo, b.alpha, mk2(x, y));
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
    // plastic.h:23-81
    static f3 plastic_diff(const orc_local_bsdf &b) {
        f3 d = ld3(b.reflectance);
        return d / (1.f - (b.nonlinear ? d * b.int_fdr : mk3(b.int_fdr)));
    }
    static float plastic_specular_prob(const orc_local_bsdf &b, float fresnel_o) {
        return (fresnel_o * b.specular_sampling_weight) /
               (fresnel_o * b.specular_sampling_weight + (1 - fresnel_o) * (1.f - b.specular_sampling_weight));
    }
    static void plastic_f(const orc_local_bsdf &b, Rec &r) {
        r.f = mk3(0.f);
        if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
        float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
        float fresnel_i = fresnel_dielectric(b.eta, r.wi.z);
        f3 diff = plastic_diff(b);
        r.f = diff * (1.f - fresnel_i) * (1.f - fresnel_o) * cosine_sample_hemisphere_pdf(r.wi) / (b.eta * b.eta * r.wi.z);
    }
    static void plastic_pdf(const orc_local_bsdf &b, Rec &r) {
        r.pdf = 0.f;
        if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
        float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
        float specular_prob = plastic_specular_prob(b, fresnel_o);
        r.pdf = cosine_sample_hemisphere_pdf(r.wi) * (1.f - specular_prob);
    }
    static void plastic_sample(const orc_local_bsdf &b, Rec &r, uint32_t &rng) {
        if (r.wo.z <= 0.f) return;
        float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
        float x = rng_next(rng), y = rng_next(rng);
        float specular_prob = plastic_specular_prob(b, fresnel_o);
        if (x < specular_prob) {
            r.type = kLobeDeltaReflection;
            r.wi = reflect_z(r.wo);
            r.f = ld3(b.specular_reflectance) * fresnel_o / r.wi.z;
            r.pdf = specular_prob;
        } else {
            r.type = kLobeDiffuseReflection;
            r.wi = cosine_sample_hemisphere((x - specular_prob) / (1.f - specular_prob), y);
            float fresnel_i = fresnel_dielectric(b.eta, r.wi.z);
            f3 diff = plastic_diff(b);
            r.f = diff * (1.f - fresnel_i) * (1.f - fresnel_o) * cosine_sample_hemisphere_pdf(r.wi) / (b.eta * b.eta * r.wi.z);
            r.pdf = cosine_sample_hemisphere_pdf(r.wi) * (1.f - specular_prob);
        }
    }
    // rough_plastic.h:22-86
    static void rough_plastic_f(const orc_local_bsdf &b, Rec &r) {
        r.f = mk3(0.f);
        if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
        float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
        f3 wh = normalize(r.wi + r.wo);
        r.f = ld3(b.specular_reflectance) * fresnel_dielectric(b.eta, dot(wh, r.wo)) * ggx_d(wh, b.alpha) *
              ggx_g(r.wi, r.wo, b.alpha) / (4.f * r.wo.z * r.wi.z);
        float fresnel_i = fresnel_dielectric(b.eta, r.wi.z);
        f3 diff = plastic_diff(b);
        r.f += diff * (1.f - fresnel_i) * (1.f - fresnel_o) * kInvPi / (b.eta * b.eta);
    }
    static void rough_plastic_pdf(const orc_local_bsdf &b, Rec &r) {
        r.pdf = 0.f;
        if (r.wi.z <= 0.f || r.wo.z <= 0.f) return;
        float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
        float specular_prob = plastic_specular_prob(b, fresnel_o);
        float diffuse_prob = 1.f - specular_prob;
        f3 wh = normalize(r.wi + r.wo);
        r.pdf = specular_prob * ggx_pdf(r.wo, wh, b.alpha) / (4.f * dot(r.wi, wh));
        r.pdf += diffuse_prob * cosine_sample_hemisphere_pdf(r.wi);
    }
    static void rough_plastic_sample(const orc_local_bsdf &b, Rec &r, uint32_t &rng) {
        r.wi = mk3(0.f);
        if (r.wo.z <= 0.f) return;
        float fresnel_o = fresnel_dielectric(b.eta, r.wo.z);
        float specular_prob = plastic_specular_prob(b, fresnel_o);
        float x = rng_next(rng), y = rng_next(rng);
        if (y < specular_prob) {
            y /= specular_prob;
            f3 wh = ggx_sample(r.wo, b.alpha, mk2(x, y));
            r.wi = reflect(r.wo, wh);
            r.type = kLobeGlossyReflection;
        } else {
            y = (y - specular_prob) / (1.f - specular_prob);
            r.wi = cosine_sample_hemisphere(x, y);
            r.type = kLobeDiffuseReflection;
        }
        rough_plastic_pdf(b, r);
        rough_plastic_f(b, r);
    }

    // LocalBsdf::Sample / Eval switch, render/material/optix_material.h:70-91
    static void bsdf_sample(const orc_local_bsdf &b, f3 wo, uint32_t &rng, orc_bsdf_result &out) {
        Rec r;
        r.wo = wo;
        switch (b.type) {
            case ORC_MAT_DIFFUSE: diffuse_sample(b, r, rng); break;
            case ORC_MAT_DIELECTRIC: dielectric_sample(b, r, rng); break;
            case ORC_MAT_ROUGH_DIELECTRIC: rough_dielectric_sample(b, r, rng); break;
            case ORC_MAT_CONDUCTOR: conductor_sample(b, r, rng); break;
            case ORC_MAT_ROUGH_CONDUCTOR: rough_conductor_sample(b, r, rng); break;
            case ORC_MAT_PLASTIC: plastic_sample(b, r, rng); break;
            case ORC_MAT_ROUGH_PLASTIC: rough_plastic_sample(b, r, rng); break;
            default: break;
        }
        st3(out.wi, r.wi), st3(out.f, r.f);
        out.pdf = r.pdf, out.sampled_type = r.type, out.rng_after = rng;
    }
    static void bsdf_eval(const orc_local_bsdf &b, f3 wi, f3 wo, f3 &f, float &pdf) {
        Rec r;
        r.wi = wi, r.wo = wo;
        switch (b.type) {
            case ORC_MAT_DIFFUSE: diffuse_f(b, r), diffuse_pdf(b, r); break;
            case ORC_MAT_ROUGH_DIELECTRIC: rough_dielectric_f(b, r), rough_dielectric_pdf(b, r); break;
            case ORC_MAT_ROUGH_CONDUCTOR: rough_conductor_f(b, r), rough_conductor_pdf(b, r); break;
            case ORC_MAT_PLASTIC: plastic_f(b, r), plastic_pdf(b, r); break;
            case ORC_MAT_ROUGH_PLASTIC: rough_plastic_f(b, r), rough_plastic_pdf(b, r); break;
            default: break; // dielectric / conductor: f = 0, pdf = 0 (dielectric.h:21-27, conductor.h:20-26)
        }
        f = r.f, pdf = r.pdf;
    }

    // ---- emitters: render/emitter/{area,sphere,env}.h, render/emitter.h -----------------------
    static void emitter_sample_direct(const orc_emitter &e, f3 hit_pos, f3 hit_n, f2 xi, orc_emit_sample &out) {
        out = orc_emit_sample{};
        f3 position, normal, radiance, wi;
        if (e.type == ORC_EMIT_TRI || e.type == ORC_EMIT_SPHERE) {
            f2 tex;
            if (e.type == ORC_EMIT_TRI) { // area.h:17-35
                f3 t = uniform_sample_triangle(xi.x, xi.y);
                position = ld3(e.pos[0]) * t.x + ld3(e.pos[1]) * t.y + ld3(e.pos[2]) * t.z;
                normal = normalize(ld3(e.nrm[0]) * t.x + ld3(e.nrm[1]) * t.y + ld3(e.nrm[2]) * t.z);
                tex = mk2(e.uv[0][0], e.uv[0][1]) * t.x + mk2(e.uv[1][0], e.uv[1][1]) * t.y + mk2(e.uv[2][0], e.uv[2][1]) * t.z;
            } else { // sphere.h:14-32
                f3 t = uniform_sample_sphere(xi.x, xi.y);
                position = t * e.radius + ld3(e.center);
                normal = normalize(t);
                tex = sphere_texcoord(t);
            }
            radiance = tex_sample(e.radiance, tex);
            wi = normalize(position - hit_pos);
            float NoL = dot(hit_n, wi);
            float LNoL = dot(normal, -wi);
            if (NoL > 0.f && LNoL > 0.f) {
                float distance = length(position - hit_pos);
                out.pdf = distance * distance / (LNoL * e.area);
                out.distance = distance;
            }
        } else if (e.type == ORC_EMIT_CONST_ENV) { // env.h:70-80
            f3 local_wi = uniform_sample_hemisphere(xi.x, xi.y);
            wi = to_world(local_wi, hit_n);
            out.pdf = uniform_sample_hemisphere_pdf(local_wi);
            out.distance = kMaxDistance;
            radiance = ld3(e.radiance.a);
            position = hit_pos + wi * out.distance;
            normal = mk3(0.f); // normalize(center - pos) in the reference; never read by the integrator
        } else if (e.type == ORC_EMIT_ENV_MAP) { // env.h:24-48
            unsigned int row_index = 0;
            for (; row_index < (e.map_h + 1) - 1; ++row_index) {
                if (xi.x <= e.row_cdf[row_index]) break;
            }
            unsigned int col_index = 0;
            for (int i = row_index * (e.map_w + 1); col_index < e.map_w - 1; ++i, ++col_index) {
                if (xi.y <= e.col_cdf[i]) break;
            }
            const float phi = col_index * kPi * 2.f / e.map_w;
            const float theta = row_index * kPi / e.map_h;
            const f3 local_wi = mk3(sinf(theta) * sinf(kPi - phi), cosf(theta), sinf(theta) * cosf(kPi - phi));
            wi = mk3(dot(ld3(e.to_world), local_wi), dot(ld3(e.to_world + 3), local_wi), dot(ld3(e.to_world + 6), local_wi));
            out.distance = kMaxDistance;
            const f2 tex = mk2(phi * 0.5f * kInvPi, theta * kInvPi);
            radiance = tex_sample(e.radiance, tex) * e.scale;
            out.pdf = luminance(radiance) * e.row_weight[row_index] * e.normalization / fmaxf(1e-4f, fabsf(sinf(theta)));
            if (out.pdf < 0.f) out.pdf = 0.f;
            position = hit_pos + wi * out.distance;
            normal = mk3(0.f); // as for the constant environment
        } else {
            return;
        }
        st3(out.radiance, radiance), st3(out.wi, wi), st3(out.pos, position), st3(out.normal, normal);
    }
    static void emitter_eval(const orc_emitter &e, f3 emit_pos, f3 emit_n, f2 emit_uv, f3 scatter_pos, f3 &radiance, float &pdf) {
        radiance = mk3(0.f), pdf = 0.f;
        if (e.type == ORC_EMIT_TRI || e.type == ORC_EMIT_SPHERE) { // area.h:37-45, sphere.h:34-42
            f3 dir = normalize(scatter_pos - emit_pos);
            float LNoL = dot(emit_n, dir);
            if (LNoL > 0.f) {
                float distance = length(scatter_pos - emit_pos);
                pdf = distance * distance / (LNoL * e.area);
                radiance = tex_sample(e.radiance, emit_uv);
            }
        } else if (e.type == ORC_EMIT_CONST_ENV) { // env.h:82-85
            pdf = 0.25f * kInvPi;
            radiance = ld3(e.radiance.a);
        } else if (e.type == ORC_EMIT_ENV_MAP) { // env.h:50-64
            f3 dir = normalize(emit_pos - scatter_pos);
            dir = mk3(dot(ld3(e.to_local), dir), dot(ld3(e.to_local + 3), dir), dot(ld3(e.to_local + 6), dir));
            const float phi = kPi - atan2f(dir.x, dir.z);
            const float theta = acosf(dir.y);
            const f2 tex = mk2(phi * 0.5f * kInvPi, theta * kInvPi);
            unsigned int row_index = static_cast<unsigned int>(tex.y * e.map_h);
            row_index = row_index > e.map_h - 2u ? e.map_h - 2u : row_index; // clamp(row_index, 0u, map_size.y - 2u)
            radiance = tex_sample(e.radiance, tex) * e.scale;
            const float w0 = e.row_weight[row_index], w1 = e.row_weight[row_index + 1], t = tex.y * e.map_h - 1.f * row_index;
            pdf = luminance(radiance) * (w0 + t * (w1 - w0)) * e.normalization / fmaxf(1e-4f, fabsf(sinf(theta)));
        }
    }
    static f3 emitter_radiance(const orc_emitter &e, f2 uv) { // emitter.h:54-71
        if (e.type == ORC_EMIT_CONST_ENV) return ld3(e.radiance.a);
        return tex_sample(e.radiance, uv);
    }
    // EmitterGroup::SelectOneEmiiter, emitter.h:110-136.  Returns an index into `areas`, n for
    // the env emitter, or -1 when the scene has no emitter at all (the reference dereferences
    // a null pointer there).
    static int select_emitter(const orc_emitter *areas, int n, bool has_env, float p) {
        float sum_p = 0.f;
        for (int i = 0; i < n; ++i) {
            if (p <= sum_p + areas[i].select_probability) return i;
            sum_p += areas[i].select_probability;
        }
        return has_env ? n : n - 1;
    }
};
}// namespace orc
#endif
