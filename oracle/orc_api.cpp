/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_types.h).
 *
 * C API of the oracle.  This one file is compiled twice (oracle/Makefile):
 *   default            -> oracle/_build/liborc_port.so   arithmetic restated (orc_backend_port.h)
 *   -DORC_BACKEND_REF  -> oracle/_ref/liborc_ref.so      arithmetic from the reference's headers
 * Both export the same symbols, so tests load the two side by side and diff them.
 */
#ifdef ORC_BACKEND_REF
#include "orc_backend_ref.h"
using Backend = orc::RefBackend;
#else
#include "orc_backend_port.h"
using Backend = orc::PortBackend;
#endif
#include "orc_render.h"

using namespace orc;
using SceneT = Scene<Backend>;

static f3 v3(const float *p) { return f3{ p[0], p[1], p[2] }; }
static void put3(float *p, f3 v) { p[0] = v.x, p[1] = v.y, p[2] = v.z; }

extern "C" {
const char *orc_backend_name() { return Backend::name(); }

/* ---------------- per-function known-answer entry points ---------------- */
void orc_rng_stream(uint32_t rounds, uint32_t v0, uint32_t v1, uint32_t n, uint32_t *state0, float *out) {
    uint32_t s = Backend::rng_init(rounds, v0, v1);
    if (state0) *state0 = s;
    for (uint32_t i = 0; i < n; ++i) out[i] = Backend::rng_next(s);
}
void orc_warp(int which, float u1, float u2, float *out3) {
    f3 r{};
    switch (which) {
        case 0: r = Backend::uniform_sample_triangle(u1, u2); break;
        case 1: r = Backend::uniform_sample_sphere(u1, u2); break;
        case 2: r = Backend::cosine_sample_hemisphere(u1, u2); break;
        case 3: r = Backend::uniform_sample_hemisphere(u1, u2); break;
    }
    put3(out3, r);
}
void orc_frame(const float *v, const float *n, float *local3, float *world3) {
    put3(local3, Backend::to_local(v3(v), v3(n)));
    put3(world3, Backend::to_world(v3(v), v3(n)));
}
void orc_sphere_texcoord(const float *p, float *uv2) {
    f2 t = Backend::sphere_texcoord(v3(p));
    uv2[0] = t.x, uv2[1] = t.y;
}
float orc_fresnel_dielectric(float eta, float cos_i, float *cos_t) { return Backend::fresnel_dielectric(eta, cos_i, *cos_t); }
void orc_fresnel_conductor(const float *eta, const float *k, float cos_i, float *out3) { put3(out3, Backend::fresnel_conductor(v3(eta), v3(k), cos_i)); }
float orc_fresnel_diffuse(float eta) { return Backend::fresnel_diffuse(eta); }
void orc_ggx(const float *wi, const float *wo, const float *wh, float alpha, float *out4) {
    out4[0] = Backend::ggx_d(v3(wh), alpha);
    out4[1] = Backend::ggx_g1(v3(wo), alpha);
    out4[2] = Backend::ggx_g(v3(wi), v3(wo), alpha);
    out4[3] = Backend::ggx_pdf(v3(wo), v3(wh), alpha);
}
void orc_ggx_sample(const float *wo, float alpha, float x, float y, float *out3) { put3(out3, Backend::ggx_sample(v3(wo), alpha, f2{ x, y })); }
void orc_tex_sample(const orc_texture *t, float u, float v, float *out3) { put3(out3, Backend::tex_sample(*t, f2{ u, v })); }
void orc_bsdf_sample(const orc_local_bsdf *b, const float *wo, uint32_t rng, orc_bsdf_result *out) {
    *out = orc_bsdf_result{};
    Backend::bsdf_sample(*b, v3(wo), rng, *out);
}
void orc_bsdf_eval(const orc_local_bsdf *b, const float *wi, const float *wo, float *f3out, float *pdf) {
    f3 f;
    Backend::bsdf_eval(*b, v3(wi), v3(wo), f, *pdf);
    put3(f3out, f);
}
void orc_emitter_sample_direct(const orc_emitter *e, const float *hit_pos, const float *hit_n, float x, float y, orc_emit_sample *out) {
    Backend::emitter_sample_direct(*e, v3(hit_pos), v3(hit_n), f2{ x, y }, *out);
}
void orc_emitter_eval(const orc_emitter *e, const float *pos, const float *n, const float *uv, const float *scatter, float *rad3, float *pdf) {
    f3 r;
    Backend::emitter_eval(*e, v3(pos), v3(n), f2{ uv[0], uv[1] }, v3(scatter), r, *pdf);
    put3(rad3, r);
}
int orc_select_emitter(const orc_emitter *areas, int n, int has_env, float p) { return Backend::select_emitter(areas, n, has_env != 0, p); }

/* ---------------- host precompute ---------------- */
void orc_resolve_transform(const orc_transform *t, float *out16) {
    m44 m = resolve_transform(*t);
    std::memcpy(out16, m.e, sizeof(m.e));
}
void orc_load_material(const orc_material *m, float *eta, float *int_fdr, float *spec_weight) {
    DeviceMaterial d = load_material<Backend>(*m);
    *eta = d.eta, *int_fdr = d.int_fdr, *spec_weight = d.specular_sampling_weight;
}

/* ---------------- scene ---------------- */
void *orc_scene_new() { return new SceneT(); }
void orc_scene_free(void *s) { delete static_cast<SceneT *>(s); }
void orc_set_integrator(void *s, int max_depth) { static_cast<SceneT *>(s)->max_depth = max_depth; }
void orc_set_sensor(void *sp, float fov, int fov_axis_x, float near_clip, float far_clip, const orc_transform *to_world, int film_w, int film_h) {
    auto *s = static_cast<SceneT *>(sp);
    s->film_w = film_w, s->film_h = film_h;
    s->cam = make_camera(fov, fov_axis_x != 0, near_clip, far_clip, *to_world, film_w, film_h);
}
int orc_add_shape(void *s, int shape_type, const orc_transform *to_world, const orc_material *mat, int is_emitter, const orc_texture *radiance,
                  int flip_normals, const float *center, float radius, int flip_tex_coords, uint32_t nv, uint32_t nf, const float *pos,
                  const float *nrm, const float *uv, const uint32_t *idx) {
    return static_cast<SceneT *>(s)->add_shape(shape_type, to_world, mat, is_emitter, radiance, flip_normals, center, radius, flip_tex_coords, nv,
                                               nf, pos, nrm, uv, idx);
}
void orc_set_env_const(void *s, float r, float g, float b) { static_cast<SceneT *>(s)->set_env_const(r, g, b); }
void orc_set_env_map(void *s, const float *rgba, uint32_t w, uint32_t h, float scale, const orc_transform *to_world) {
    static_cast<SceneT *>(s)->set_env_map(rgba, w, h, scale, *to_world);
}
/* BuildEnvMapCdfTable (world/emitter.cpp:107-149) on its own: row_cdf[h + 1], row_weight[h], col_cdf[(w + 1) * h] */
float orc_build_env_tables(const float *rgba, uint32_t w, uint32_t h, float *row_cdf, float *row_weight, float *col_cdf) {
    EnvTables t = build_env_tables(rgba, w, h);
    std::memcpy(row_cdf, t.row_cdf.data(), (h + 1) * sizeof(float));
    std::memcpy(row_weight, t.row_weight.data(), h * sizeof(float));
    std::memcpy(col_cdf, t.col_cdf.data(), (size_t)(w + 1) * h * sizeof(float));
    return t.normalization;
}
float orc_texture_weight(const orc_texture *t) { return tex_max_weight(*t); } /* GetWeight, world/emitter.cpp:77-101 */
void orc_finalize(void *s) { static_cast<SceneT *>(s)->finalize(); }

void orc_get_camera(void *sp, float *s2c16, float *c2w16, float *fov_y) {
    auto *s = static_cast<SceneT *>(sp);
    std::memcpy(s2c16, s->cam.sample_to_camera.e, 64), std::memcpy(c2w16, s->cam.camera_to_world.e, 64);
    if (fov_y) *fov_y = s->cam.fov_y;
}
int orc_num_area_emitters(void *s) { return (int)static_cast<SceneT *>(s)->areas.size(); }
void orc_get_area_emitters(void *s, orc_emitter *out) {
    auto &a = static_cast<SceneT *>(s)->areas;
    std::memcpy(out, a.data(), a.size() * sizeof(orc_emitter));
}
int orc_get_env_emitter(void *sp, orc_emitter *out) {
    auto *s = static_cast<SceneT *>(sp);
    if (s->has_env) *out = s->env;
    return s->has_env;
}
void orc_get_instance_xform(void *s, int i, float *out16) { std::memcpy(out16, static_cast<SceneT *>(s)->instances[i].xf.e, 64); }
int orc_num_instances(void *s) { return (int)static_cast<SceneT *>(s)->instances.size(); }
uint64_t orc_num_triangles(void *s) { return static_cast<SceneT *>(s)->tris.size(); }

/* rays: n x 8 floats (ox oy oz tmin dx dy dz tmax).  brute != 0 -> exhaustive loop. */
void orc_trace_closest(void *sp, const float *rays, uint64_t n, orc_hit *hits, int brute, int threads, uint64_t *n_prim_tests) {
    auto *s = static_cast<SceneT *>(sp);
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> next{ 0 }, tests{ 0 };
    auto worker = [&]() {
        uint64_t local = 0;
        for (;;) {
            uint64_t b = next.fetch_add(256);
            if (b >= n) break;
            for (uint64_t i = b; i < std::min(n, b + 256); ++i) {
                const float *r = rays + i * 8;
                hits[i] = brute ? s->trace_brute(v3(r), v3(r + 4), r[3], r[7]) : s->trace_closest(v3(r), v3(r + 4), r[3], r[7], &local);
            }
        }
        tests += local;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    if (n_prim_tests) *n_prim_tests = tests;
}
/* exhaustive loop with the instances flagged in objspace[instance] intersected in OBJECT space (orc_render.h: trace_brute_objspace) */
void orc_trace_closest_objspace(void *sp, const float *rays, uint64_t n, orc_hit *hits, const uint8_t *objspace, int threads) {
    auto *s = static_cast<SceneT *>(sp);
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> next{ 0 };
    auto worker = [&]() {
        for (;;) {
            uint64_t b = next.fetch_add(64);
            if (b >= n) break;
            for (uint64_t i = b; i < std::min(n, b + 64); ++i) {
                const float *r = rays + i * 8;
                hits[i] = s->trace_brute_objspace(v3(r), v3(r + 4), r[3], r[7], objspace);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}
void orc_trace_any(void *sp, const float *rays, uint64_t n, uint8_t *occluded, int brute) {
    auto *s = static_cast<SceneT *>(sp);
    for (uint64_t i = 0; i < n; ++i) {
        const float *r = rays + i * 8;
        occluded[i] = brute ? (s->trace_brute(v3(r), v3(r + 4), r[3], r[7]).inst >= 0) : s->trace_any(v3(r), v3(r + 4), r[3], r[7]);
    }
}
/* LocalGeometry of a given hit (closest-hit program), for the geometry KATs */
void orc_hit_geometry(void *sp, const orc_hit *h, const float *ray8, float *pos3, float *nrm3, float *uv2, int *emitter_index) {
    auto *s = static_cast<SceneT *>(sp);
    SceneT::LocalGeometry g{};
    s->hit_local_geometry(*h, v3(ray8), v3(ray8 + 4), g);
    put3(pos3, g.position), put3(nrm3, g.normal), uv2[0] = g.texcoord.x, uv2[1] = g.texcoord.y;
    const Instance &in = s->instances[h->inst];
    *emitter_index = in.emitter_offset >= 0 ? in.emitter_offset + h->prim : -1;
}
/* primary camera rays for a frame (main.cu:55-78): n = w*h, 8 floats each */
void orc_camera_rays(void *sp, uint32_t random_seed, float *rays) {
    auto *s = static_cast<SceneT *>(sp);
    // re-derive through render_pixel's own code path: trace nothing, just replicate the maths
    const uint32_t w = s->film_w, h = s->film_h;
    const float *s2c = s->cam.sample_to_camera.e, *c2w = s->cam.camera_to_world.e;
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            uint32_t pi = y * w + x;
            uint32_t rng = Backend::rng_init(4, pi, random_seed);
            float jx = Backend::rng_next(rng), jy = Backend::rng_next(rng);
            f4 pf{ (static_cast<float>(x) + jx) / static_cast<float>(w), (static_cast<float>(y) + jy) / static_cast<float>(h), 0.f, 1.f };
            f4 d4{ dot(f4{ s2c[0], s2c[1], s2c[2], s2c[3] }, pf), dot(f4{ s2c[4], s2c[5], s2c[6], s2c[7] }, pf),
                   dot(f4{ s2c[8], s2c[9], s2c[10], s2c[11] }, pf), dot(f4{ s2c[12], s2c[13], s2c[14], s2c[15] }, pf) };
            float inv = 1.0f / d4.w;
            d4 = f4{ d4.x * inv, d4.y * inv, d4.z * inv, 0.f };
            float il = 1.0f / sqrtf(dot(d4, d4));
            d4 = f4{ d4.x * il, d4.y * il, d4.z * il, 0.f };
            f3 d = normalize(f3{ dot(f4{ c2w[0], c2w[1], c2w[2], c2w[3] }, d4), dot(f4{ c2w[4], c2w[5], c2w[6], c2w[7] }, d4),
                                 dot(f4{ c2w[8], c2w[9], c2w[10], c2w[11] }, d4) });
            float *r = rays + (size_t)pi * 8;
            r[0] = c2w[3], r[1] = c2w[7], r[2] = c2w[11], r[3] = 0.001f, r[4] = d.x, r[5] = d.y, r[6] = d.z, r[7] = 1e16f;
        }
}
/* n_frames consecutive PTPass::OnRun calls.  depth_limit <= 0 -> the scene's integrator.max_depth.
 * ray_counts[0] = closest-hit rays, [1] = shadow rays. */
void orc_render(void *sp, uint32_t first_seed, uint32_t n_frames, uint32_t sample_cnt0, int depth_limit, int accumulate, int threads, float *accum4,
                float *frame4, float *albedo3, float *normal3, float *test1, uint64_t *ray_counts) {
    auto *s = static_cast<SceneT *>(sp);
    if (!s->finalized) s->finalize();
    s->render(first_seed, n_frames, sample_cnt0, depth_limit > 0 ? depth_limit : s->max_depth, accumulate, threads, accum4, frame4, albedo3,
              normal3, test1, ray_counts);
}
/* one pixel, one frame — for debugging parity failures */
void orc_render_pixel(void *sp, uint32_t x, uint32_t y, uint32_t seed, int depth_limit, float *radiance3, uint32_t *rays2) {
    auto *s = static_cast<SceneT *>(sp);
    if (!s->finalized) s->finalize();
    auto po = s->render_pixel(x, y, seed, depth_limit > 0 ? depth_limit : s->max_depth);
    put3(radiance3, po.radiance);
    if (rays2) rays2[0] = po.closest_rays, rays2[1] = po.shadow_rays;
}
}
