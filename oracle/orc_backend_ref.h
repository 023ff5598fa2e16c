/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * "reference" arithmetic backend: same interface as orc_backend_port.h, but every function is a
 * thin adaptor onto the reference's OWN headers, included by path from the reference tree
 * (-I$PUPIL_REF/framework) — nothing is copied:
 *     cuda/random.h, cuda/vec_math.h, optix/util.h,
 *     render/material/{fresnel,ggx}.h, render/material/bsdf/bsdf.h (all seven BSDFs),
 *     render/emitter.h (+ emitter/{area,sphere,env}.h)
 * Scaffolding needed to compile them as host C++ with g++:
 *   1. ref_shims/optix.h            — OptiX SDK is not installed
 *   2. ref_shims/cuda/texture.h     — cuts the util/type.h -> DirectXMath chain
 *   3. PUPIL_CPP left undefined     — keeps Emitter::{Eval,SampleDirect} visible (render/emitter.h:27)
 *   4. the `using std::abs ...` prelude below — otherwise unqualified abs(float) binds to
 *      int abs(int) on the host and IsZero()/ggx::Pdf() are silently wrong
 *   5. make_float2 is rewritten to a braced initialiser while cuda/random.h is parsed so that
 *      Random::Next2() draws x first, then y, as nvcc's device code does (checked in the PTX:
 *      x = first LCG output).  g++ evaluates call arguments right-to-left and would swap them.
 * render/material/optix_material.h cannot be compiled (MSVC-only `EMatType::##x` pasting), so
 * the LocalBsdf switch is restated here over the reference's Local structs.
 */
#ifndef ORC_BACKEND_REF_H
#define ORC_BACKEND_REF_H

#include <cmath>
#include <cstdlib>
using std::abs;
using std::acos;
using std::asin;
using std::atan;
using std::atan2;
using std::ceil;
using std::cos;
using std::floor;
using std::pow;
using std::sin;
using std::sqrt;
using std::tan;

#include <vector_types.h>
#include <vector_functions.h>
#include "cuda/vec_math.h"
#define make_float2(...) \
    float2 { __VA_ARGS__ }
#include "cuda/random.h"
#undef make_float2
#include "optix/util.h"
#include "render/material/fresnel.h"
#include "render/material/ggx.h"
#include "render/material/bsdf/bsdf.h"
#include "render/emitter.h"

#include "orc_types.h"
#include "orc_vec.h"
#include <vector>

namespace orc {
namespace P = Pupil::optix;
namespace PM = Pupil::optix::material;

enum : uint32_t { kLobeDelta = (1u << 5) | (1u << 6) };
constexpr float kMaxDistance = Pupil::optix::MAX_DISTANCE;

struct RefBackend {
    static const char *name() { return "reference"; }

    static float3 c(f3 v) { return make_float3(v.x, v.y, v.z); }
    static f3 c(float3 v) { return f3{ v.x, v.y, v.z }; }
    static float3 ld3(const float *p) { return make_float3(p[0], p[1], p[2]); }
    static void st3(float *p, float3 v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }

    static uint32_t rng_init(uint32_t rounds, uint32_t v0, uint32_t v1) {
        Pupil::cuda::Random r;
        r.Init(rounds, v0, v1);
        return r.GetSeed();
    }
    static float rng_next(uint32_t &s) {
        Pupil::cuda::Random r;
        r.SetSeed(s);
        float v = r.Next();
        s = r.GetSeed();
        return v;
    }

    static f3 uniform_sample_triangle(float a, float b) { return c(P::UniformSampleTriangle(a, b)); }
    static f3 uniform_sample_sphere(float a, float b) { return c(P::UniformSampleSphere(a, b)); }
    static f3 cosine_sample_hemisphere(float a, float b) { return c(P::CosineSampleHemisphere(a, b)); }
    static f3 uniform_sample_hemisphere(float a, float b) { return c(P::UniformSampleHemisphere(a, b)); }
    static f3 to_local(f3 v, f3 n) { return c(P::ToLocal(c(v), c(n))); }
    static f3 to_world(f3 v, f3 n) { return c(P::ToWorld(c(v), c(n))); }
    static f2 sphere_texcoord(f3 p) {
        float2 t = P::GetSphereTexcoord(c(p));
        return f2{ t.x, t.y };
    }
    static float luminance(f3 v) { return P::GetLuminance(c(v)); }
    static float mis_weight(float x, float y) { return P::MISWeight(x, y); }
    static bool is_zero(float v) { return P::IsZero(v); }
    static bool is_zero(f3 v) { return P::IsZero(c(v)); }

    static float fresnel_dielectric(float eta, float ci, float &ct) { return PM::fresnel::DielectricReflectance(eta, ci, ct); }
    static f3 fresnel_conductor(f3 eta, f3 k, float ci) { return c(PM::fresnel::ConductorReflectance(c(eta), c(k), ci)); }
    static float fresnel_diffuse(float eta) { return PM::fresnel::DiffuseReflectance(eta); }
    static float ggx_d(f3 wh, float a) { return PM::ggx::D(c(wh), a); }
    static float ggx_g1(f3 w, float a) { return PM::ggx::G1(c(w), a); }
    static float ggx_g(f3 wi, f3 wo, float a) { return PM::ggx::G(c(wi), c(wo), a); }
    static float ggx_pdf(f3 wo, f3 wh, float a) { return PM::ggx::Pdf(c(wo), c(wh), a); }
    static f3 ggx_sample(f3 wo, float a, f2 xi) { return c(PM::ggx::Sample(c(wo), a, float2{ xi.x, xi.y })); }

    static Pupil::cuda::Texture tex(const orc_texture &t) {
        Pupil::cuda::Texture r;
        r.type = static_cast<Pupil::util::ETextureType>(t.type);
        r.rgb = ld3(t.a);
        r.patch2 = ld3(t.b);
        r.transform.r0 = make_float4(t.to_uv[0], t.to_uv[1], t.to_uv[2], t.to_uv[3]);
        r.transform.r1 = make_float4(t.to_uv[4], t.to_uv[5], t.to_uv[6], t.to_uv[7]);
        r.transform.r2 = make_float4(t.to_uv[8], t.to_uv[9], t.to_uv[10], t.to_uv[11]);
        r.transform.r3 = make_float4(t.to_uv[12], t.to_uv[13], t.to_uv[14], t.to_uv[15]);
        r.bitmap = t.bitmap, r.bitmap_w = t.bitmap_w, r.bitmap_h = t.bitmap_h, r.address_mode = t.address_mode, r.filter_mode = t.filter_mode;
        return r;
    }
    static f3 tex_sample(const orc_texture &t, f2 uv) { return c(tex(t).Sample(float2{ uv.x, uv.y })); }

    // LocalBsdf switch restated over the reference's own Local structs
    // (render/material/optix_material.h:70-91 is MSVC-only and cannot be included).
    template<typename Fn>
    static void with_local(const orc_local_bsdf &b, Fn &&fn) {
        switch (b.type) {
            case ORC_MAT_DIFFUSE: {
                PM::Diffuse::Local l;
                l.reflectance = ld3(b.reflectance);
                fn(l);
            } break;
            case ORC_MAT_DIELECTRIC: {
                PM::Dielectric::Local l;
                l.eta = b.eta;
                l.specular_reflectance = ld3(b.specular_reflectance);
                l.specular_transmittance = ld3(b.specular_transmittance);
                fn(l);
            } break;
            case ORC_MAT_ROUGH_DIELECTRIC: {
                PM::RoughDielectric::Local l;
                l.alpha = b.alpha, l.eta = b.eta;
                l.specular_reflectance = ld3(b.specular_reflectance);
                l.specular_transmittance = ld3(b.specular_transmittance);
                fn(l);
            } break;
            case ORC_MAT_CONDUCTOR: {
                PM::Conductor::Local l;
                l.eta = ld3(b.eta3), l.k = ld3(b.k3);
                l.specular_reflectance = ld3(b.specular_reflectance);
                fn(l);
            } break;
            case ORC_MAT_ROUGH_CONDUCTOR: {
                PM::RoughConductor::Local l;
                l.alpha = b.alpha, l.eta = ld3(b.eta3), l.k = ld3(b.k3);
                l.specular_reflectance = ld3(b.specular_reflectance);
                fn(l);
            } break;
            case ORC_MAT_PLASTIC: {
                PM::Plastic::Local l;
                l.eta = b.eta, l.int_fdr = b.int_fdr, l.specular_sampling_weight = b.specular_sampling_weight;
                l.nonlinear = b.nonlinear != 0;
                l.diffuse_reflectance = ld3(b.reflectance);
                l.specular_reflectance = ld3(b.specular_reflectance);
                fn(l);
            } break;
            case ORC_MAT_ROUGH_PLASTIC: {
                PM::RoughPlastic::Local l;
                l.eta = b.eta, l.int_fdr = b.int_fdr, l.specular_sampling_weight = b.specular_sampling_weight;
                l.alpha = b.alpha, l.nonlinear = b.nonlinear != 0;
                l.diffuse_reflectance = ld3(b.reflectance);
                l.specular_reflectance = ld3(b.specular_reflectance);
                fn(l);
            } break;
            default: break;
        }
    }
    static void bsdf_sample(const orc_local_bsdf &b, f3 wo, uint32_t &rng, orc_bsdf_result &out) {
        Pupil::cuda::Random sampler;
        sampler.SetSeed(rng);
        P::BsdfSamplingRecord rec;
        rec.wi = make_float3(0.f);
        rec.wo = c(wo);
        rec.sampler = &sampler;
        with_local(b, [&](auto &l) { l.Sample(rec); });
        rng = sampler.GetSeed();
        st3(out.wi, rec.wi), st3(out.f, rec.f);
        out.pdf = rec.pdf, out.sampled_type = static_cast<uint32_t>(rec.sampled_type), out.rng_after = rng;
    }
    static void bsdf_eval(const orc_local_bsdf &b, f3 wi, f3 wo, f3 &f, float &pdf) {
        P::BsdfSamplingRecord rec;
        rec.wi = c(wi), rec.wo = c(wo);
        with_local(b, [&](auto &l) { l.GetBsdf(rec); l.GetPdf(rec); });
        f = c(rec.f), pdf = rec.pdf;
    }

    static P::Emitter emitter(const orc_emitter &e) {
        P::Emitter r;
        r.type = static_cast<P::EEmitterType>(e.type);
        r.weight = e.weight, r.select_probability = e.select_probability;
        if (e.type == ORC_EMIT_TRI) {
            r.area.radiance = tex(e.radiance);
            r.area.area = e.area;
            r.area.geo.v0.pos = ld3(e.pos[0]), r.area.geo.v1.pos = ld3(e.pos[1]), r.area.geo.v2.pos = ld3(e.pos[2]);
            r.area.geo.v0.normal = ld3(e.nrm[0]), r.area.geo.v1.normal = ld3(e.nrm[1]), r.area.geo.v2.normal = ld3(e.nrm[2]);
            r.area.geo.v0.tex = float2{ e.uv[0][0], e.uv[0][1] }, r.area.geo.v1.tex = float2{ e.uv[1][0], e.uv[1][1] };
            r.area.geo.v2.tex = float2{ e.uv[2][0], e.uv[2][1] };
        } else if (e.type == ORC_EMIT_SPHERE) {
            r.sphere.radiance = tex(e.radiance);
            r.sphere.area = e.area;
            r.sphere.geo.center = ld3(e.center), r.sphere.geo.radius = e.radius;
        } else if (e.type == ORC_EMIT_CONST_ENV) {
            r.const_env.color = ld3(e.radiance.a);
            r.const_env.center = make_float3(0.f);
        } else if (e.type == ORC_EMIT_ENV_MAP) { // the reference's own EnvMapEmitter over host-memory views
            r.env_map.radiance = tex(e.radiance);
            r.env_map.center = make_float3(0.f);
            r.env_map.map_size = make_uint2(e.map_w, e.map_h);
            r.env_map.row_cdf.SetData(reinterpret_cast<CUdeviceptr>(e.row_cdf), static_cast<size_t>(e.map_h) + 1);
            r.env_map.col_cdf.SetData(reinterpret_cast<CUdeviceptr>(e.col_cdf), static_cast<size_t>(e.map_w + 1) * e.map_h);
            r.env_map.row_weight.SetData(reinterpret_cast<CUdeviceptr>(e.row_weight), static_cast<size_t>(e.map_h));
            r.env_map.to_world.r0 = ld3(e.to_world), r.env_map.to_world.r1 = ld3(e.to_world + 3), r.env_map.to_world.r2 = ld3(e.to_world + 6);
            r.env_map.to_local.r0 = ld3(e.to_local), r.env_map.to_local.r1 = ld3(e.to_local + 3), r.env_map.to_local.r2 = ld3(e.to_local + 6);
            r.env_map.normalization = e.normalization;
            r.env_map.scale = e.scale;
        }
        return r;
    }
    static void emitter_sample_direct(const orc_emitter &e, f3 hit_pos, f3 hit_n, f2 xi, orc_emit_sample &out) {
        out = orc_emit_sample{};
        P::Emitter em = emitter(e);
        P::EmitterSampleRecord rec{};
        rec.distance = 0.f, rec.is_delta = false, rec.pdf = 0.f;
        rec.radiance = rec.wi = rec.pos = rec.normal = make_float3(0.f);
        P::LocalGeometry geo;
        geo.position = c(hit_pos), geo.normal = c(hit_n), geo.texcoord = float2{ 0.f, 0.f };
        em.SampleDirect(rec, geo, float2{ xi.x, xi.y });
        st3(out.radiance, rec.radiance), st3(out.wi, rec.wi), st3(out.pos, rec.pos);
        if (e.type != ORC_EMIT_CONST_ENV && e.type != ORC_EMIT_ENV_MAP) st3(out.normal, rec.normal); // env normal depends on the scene centre; unused
        out.distance = rec.distance, out.pdf = rec.pdf, out.is_delta = rec.is_delta;
    }
    static void emitter_eval(const orc_emitter &e, f3 emit_pos, f3 emit_n, f2 emit_uv, f3 scatter_pos, f3 &radiance, float &pdf) {
        P::Emitter em = emitter(e);
        P::EmitEvalRecord rec;
        rec.radiance = make_float3(0.f), rec.pdf = 0.f;
        P::LocalGeometry geo;
        geo.position = c(emit_pos), geo.normal = c(emit_n), geo.texcoord = float2{ emit_uv.x, emit_uv.y };
        em.Eval(rec, geo, c(scatter_pos));
        radiance = c(rec.radiance), pdf = rec.pdf;
    }
    static f3 emitter_radiance(const orc_emitter &e, f2 uv) { return c(emitter(e).GetRadiance(float2{ uv.x, uv.y })); }

    // EmitterGroup::SelectOneEmiiter walks device views (CUdeviceptr + count); on the host the
    // views can wrap ordinary memory, so the reference's own loop runs unmodified.
    static int select_emitter(const orc_emitter *areas, int n, bool has_env, float p) {
        std::vector<P::Emitter> cache; // rebuilt per call: emitter tables in the configs are tiny
        cache.reserve(n);
        for (int i = 0; i < n; ++i) cache.push_back(emitter(areas[i]));
        P::Emitter env_dummy;
        P::EmitterGroup g;
        g.areas.SetData(reinterpret_cast<CUdeviceptr>(cache.data()), static_cast<size_t>(n));
        g.points.SetData(0, 0);
        g.directionals.SetData(0, 0);
        g.env.SetData(has_env ? reinterpret_cast<CUdeviceptr>(&env_dummy) : 0);
        if (n == 0 && !has_env) return -1;
        const P::Emitter &sel = g.SelectOneEmiiter(p);
        if (&sel == &env_dummy) return n;
        return static_cast<int>(&sel - cache.data());
    }
};
}// namespace orc
#endif
