// Per-function known-answer hooks (pb2_kat): run the DEVICE restatement of one reference function over n
// inputs so tests can compare it with the oracle on the same grids.  Test-only entry point, built into its own
// library (libpb2_kat.so, include/pb2_kat.h): the product library libpb2.so does not contain it.  All arrays are HOST pointers; layouts:
//   "rng"      in0 uint32[n][3] (rounds, v0, v1)                       out uint32/float[n][8]  state, 7 draws
//   "warp"     in0 float[n][2]  (u1,u2)                                out float[n][12] tri, sphere, coshemi, unihemi
//   "frame"    in0 float[n][6]  (v, N)                                 out float[n][8]  to_local, to_world, sphere_uv(N)
//   "fresnel"  in0 float[n][8]  (eta, cos, eta3, k3)                   out float[n][8]  F_diel, cos_t, F_cond.xyz
//   "ggx"      in0 float[n][12] (wi, wo, wh, alpha, xi.x, xi.y)        out float[n][8]  D, G1(wo), G, pdf, sample.xyz
//   "texture"  in0 pb2_texture[n], in1 float[n][2] uv                  out float[n][4]  rgb
//   "bsdf"     in0 pb2_kat_bsdf[n], in1 float[n][8] (wo, wi, rng bits) out float[n][16] sample: wi f pdf type rng | eval: f pdf
//   "emitter"  in0 pb2_emitter[n], in1 float[n][8] (hit_pos, hit_n, xi), in2 float[n][12] (emit_pos, emit_n, uv, scatter)
//                                                                      out float[n][16] sample: radiance wi distance pdf | eval: radiance pdf
//   "sort"     in0 uint64[n] keys, in2 uint32[2] (begin_bit, end_bit)      out uint32[n]  input positions in sorted order (radix_sort.cu)
//   "select"   in0 pb2_emitter[m] (m = in2[0] as uint32, has_env = in2[1]), in1 float[n] p   out int32[n] index (m = env, -1 none)
#include "../scene.cuh"
#include "../pt_math.cuh"
#include "../../../include/pb2_kat.h"
#include <cstring>
#include <string>
#include <vector>

namespace pb2 {
struct KatBsdf { // mirrors LocalBsdf; slots c0..c2 as in pb2_material
    int32_t type;
    float alpha, eta, int_fdr, specular_sampling_weight;
    int32_t nonlinear;
    float c0[3], c1[3], c2[3];
};
DevTexture to_dev(const pb2_texture &t);
DevEmitter to_dev(const pb2_emitter &e);
std::vector<float> area_select_cdf(const DevEmitter *areas, size_t n);
bool radix_sort_pairs(cudaStream_t st, uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt, uint32_t n, int begin_bit, int end_bit);

namespace {
__global__ void k_rng(const uint32_t *in, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = rng_init(in[i * 3], in[i * 3 + 1], in[i * 3 + 2]);
    out[i * 8] = __uint_as_float(s);
    for (int k = 1; k < 8; ++k) out[i * 8 + k] = rng_next(s);
}
__device__ void st3(float *o, float3 v) { o[0] = v.x, o[1] = v.y, o[2] = v.z; }
__global__ void k_warp(const float *in, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = in[i * 2], b = in[i * 2 + 1];
    st3(out + i * 12, uniform_sample_triangle(a, b)), st3(out + i * 12 + 3, uniform_sample_sphere(a, b));
    st3(out + i * 12 + 6, cosine_sample_hemisphere(a, b)), st3(out + i * 12 + 9, uniform_sample_hemisphere(a, b));
}
__global__ void k_frame(const float *in, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 v = mk3(in[i * 6], in[i * 6 + 1], in[i * 6 + 2]), N = mk3(in[i * 6 + 3], in[i * 6 + 4], in[i * 6 + 5]);
    const Onb f(N);
    st3(out + i * 8, f.to_local(v)), st3(out + i * 8 + 3, f.to_world(v));
    const float2 uv = sphere_texcoord(N);
    out[i * 8 + 6] = uv.x, out[i * 8 + 7] = uv.y;
}
__global__ void k_fresnel(const float *in, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + i * 8;
    float ct;
    out[i * 8] = fresnel_dielectric(p[0], p[1], ct);
    out[i * 8 + 1] = ct;
    st3(out + i * 8 + 2, fresnel_conductor(mk3(p[2], p[3], p[4]), mk3(p[5], p[6], p[7]), p[1]));
    out[i * 8 + 5] = out[i * 8 + 6] = out[i * 8 + 7] = 0.f;
}
__global__ void k_ggx(const float *in, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + i * 12;
    const float3 wi = mk3(p[0], p[1], p[2]), wo = mk3(p[3], p[4], p[5]), wh = mk3(p[6], p[7], p[8]);
    const float a = p[9];
    out[i * 8] = ggx_d(wh, a), out[i * 8 + 1] = ggx_g1(wo, a), out[i * 8 + 2] = ggx_g(wi, wo, a), out[i * 8 + 3] = ggx_pdf(wo, wh, a);
    st3(out + i * 8 + 4, ggx_sample(wo, a, make_float2(p[10], p[11])));
    out[i * 8 + 7] = 0.f;
}
__global__ void k_texture(const DevTexture *tex, const float *uv, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    st3(out + i * 4, tex_sample(tex + i, make_float2(uv[i * 2], uv[i * 2 + 1])));
    out[i * 4 + 3] = 0.f;
}
__global__ void k_bsdf(const KatBsdf *kb, const float *in, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KatBsdf &k = kb[i];
    LocalBsdf b;
    b.type = k.type, b.alpha = k.alpha, b.eta = k.eta, b.int_fdr = k.int_fdr, b.specular_sampling_weight = k.specular_sampling_weight;
    b.nonlinear = k.nonlinear != 0;
    b.c0 = mk3(k.c0[0], k.c0[1], k.c0[2]), b.c1 = mk3(k.c1[0], k.c1[1], k.c1[2]), b.c2 = mk3(k.c2[0], k.c2[1], k.c2[2]);
    const float *p = in + i * 8;
    uint32_t rng = __float_as_uint(p[6]);
    BsdfRec r;
    r.wo = mk3(p[0], p[1], p[2]);
    bsdf_sample(b, r, rng);
    float *o = out + i * 16;
    st3(o, r.wi), st3(o + 3, r.f);
    o[6] = r.pdf, o[7] = __uint_as_float(r.type), o[8] = __uint_as_float(rng);
    BsdfRec e;
    e.wo = r.wo, e.wi = mk3(p[3], p[4], p[5]);
    bsdf_eval(b, e);
    st3(o + 9, e.f);
    o[12] = e.pdf, o[13] = o[14] = o[15] = 0.f;
}
__global__ void k_emitter(const DevEmitter *em, const float *in1, const float *in2, uint64_t n, float *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in1 + i * 8, *q = in2 + i * 12;
    const float3 hp = mk3(p[0], p[1], p[2]), hn = mk3(p[3], p[4], p[5]);
    const Onb f(hn);
    EmitSample es;
    emitter_sample_direct(em + i, hp, hn, f, make_float2(p[6], p[7]), es);
    float *o = out + i * 16;
    st3(o, es.radiance), st3(o + 3, es.wi);
    o[6] = es.distance, o[7] = es.pdf;
    float3 rad;
    float pdf;
    emitter_eval(em + i, mk3(q[0], q[1], q[2]), mk3(q[3], q[4], q[5]), make_float2(q[6], q[7]), mk3(q[8], q[9], q[10]), rad, pdf);
    st3(o + 8, rad);
    o[11] = pdf, o[12] = o[13] = o[14] = o[15] = 0.f;
}
__global__ void k_select(const DevEmitter *areas, const float *cdf, uint32_t m, const DevEmitter *env, const float *p, uint64_t n, int32_t *out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevEmitter *e = select_emitter(areas, cdf, m, env, p[i]);
    out[i] = !e ? -1 : (e == env ? (int32_t)m : (int32_t)(e - areas));
}
template<typename T>
DevBuf<T> up(const void *host, size_t count) {
    DevBuf<T> b(count);
    if (count) PB2_CUDA(cudaMemcpy(b.ptr, host, count * sizeof(T), cudaMemcpyHostToDevice));
    return b;
}
}// namespace

int run_kat(const char *what_c, const void *in0, const void *in1, const void *in2, uint64_t n, void *out) {
    const std::string what = what_c;
    const unsigned g = div_up(n, 128);
    auto finish = [&](DevBuf<float> &o) {
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaDeviceSynchronize());
        PB2_CUDA(cudaMemcpy(out, o.ptr, o.bytes(), cudaMemcpyDeviceToHost));
        return PB2_OK;
    };
    if (!n) return PB2_OK;
    if (what == "rng") {
        auto a = up<uint32_t>(in0, n * 3);
        DevBuf<float> o(n * 8);
        k_rng<<<g, 128>>>(a.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "warp") {
        auto a = up<float>(in0, n * 2);
        DevBuf<float> o(n * 12);
        k_warp<<<g, 128>>>(a.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "frame") {
        auto a = up<float>(in0, n * 6);
        DevBuf<float> o(n * 8);
        k_frame<<<g, 128>>>(a.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "fresnel") {
        auto a = up<float>(in0, n * 8);
        DevBuf<float> o(n * 8);
        k_fresnel<<<g, 128>>>(a.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "ggx") {
        auto a = up<float>(in0, n * 12);
        DevBuf<float> o(n * 8);
        k_ggx<<<g, 128>>>(a.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "texture") {
        std::vector<DevTexture> t(n);
        for (uint64_t i = 0; i < n; ++i) t[i] = to_dev(static_cast<const pb2_texture *>(in0)[i]);
        auto a = up<DevTexture>(t.data(), n);
        auto b = up<float>(in1, n * 2);
        DevBuf<float> o(n * 4);
        k_texture<<<g, 128>>>(a.ptr, b.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "bsdf") {
        auto a = up<KatBsdf>(in0, n);
        auto b = up<float>(in1, n * 8);
        DevBuf<float> o(n * 16);
        k_bsdf<<<g, 128>>>(a.ptr, b.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "emitter") {
        std::vector<DevEmitter> e(n);
        for (uint64_t i = 0; i < n; ++i) e[i] = to_dev(static_cast<const pb2_emitter *>(in0)[i]);
        auto a = up<DevEmitter>(e.data(), n);
        auto b = up<float>(in1, n * 8);
        auto c = up<float>(in2, n * 12);
        DevBuf<float> o(n * 16);
        k_emitter<<<g, 128>>>(a.ptr, b.ptr, c.ptr, n, o.ptr);
        return finish(o);
    }
    if (what == "select") {
        const uint32_t m = static_cast<const uint32_t *>(in2)[0], has_env = static_cast<const uint32_t *>(in2)[1];
        std::vector<DevEmitter> e(m + 1);
        for (uint32_t i = 0; i < m; ++i) e[i] = to_dev(static_cast<const pb2_emitter *>(in0)[i]);
        e[m] = DevEmitter{};
        e[m].type = PB2_EMIT_CONST_ENV;
        auto a = up<DevEmitter>(e.data(), m + 1);
        auto b = up<float>(in1, n);
        const std::vector<float> cdf_h = area_select_cdf(e.data(), m);
        auto cdf = up<float>(cdf_h.data(), m);
        DevBuf<int32_t> o(n);
        k_select<<<g, 128>>>(a.ptr, cdf.ptr, m, has_env ? a.ptr + m : nullptr, b.ptr, n, o.ptr);
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaDeviceSynchronize());
        PB2_CUDA(cudaMemcpy(out, o.ptr, o.bytes(), cudaMemcpyDeviceToHost));
        return PB2_OK;
    }
    if (what == "sort") { // radix_sort.cu: in0 = uint64 keys[n], in2 = uint32 {begin_bit, end_bit}; out = uint32[n]: the input positions in sorted order
        const uint32_t begin_bit = static_cast<const uint32_t *>(in2)[0], end_bit = static_cast<const uint32_t *>(in2)[1];
        auto keys = up<uint64_t>(in0, n);
        std::vector<uint32_t> iota(n);
        for (uint64_t i = 0; i < n; ++i) iota[i] = (uint32_t)i;
        auto vals = up<uint32_t>(iota.data(), n);
        DevBuf<uint64_t> keys_alt(n);
        DevBuf<uint32_t> vals_alt(n);
        const bool alt = radix_sort_pairs(nullptr, keys.ptr, keys_alt.ptr, vals.ptr, vals_alt.ptr, (uint32_t)n, (int)begin_bit, (int)end_bit);
        PB2_CUDA(cudaDeviceSynchronize());
        PB2_CUDA(cudaMemcpy(out, alt ? vals_alt.ptr : vals.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        return PB2_OK;
    }
    throw std::runtime_error("pb2_kat: unknown function '" + what + "'");
}
}// namespace pb2

// the C entry point of libpb2_kat.so (include/pb2_kat.h); errors are reported through its own message slot
namespace {
thread_local std::string g_kat_error;
}
extern "C" {
const char *pb2_kat_last_error(void) { return g_kat_error.c_str(); }
int pb2_kat(const char *what, const void *in0, const void *in1, const void *in2, uint64_t n, void *out) {
    if (!what || !out) {
        g_kat_error = "pb2_kat: null";
        return PB2_ERR_ARG;
    }
    try {
        return pb2::run_kat(what, in0, in1, in2, n, out);
    } catch (const pb2::CudaError &e) {
        g_kat_error = e.what();
        return PB2_ERR_CUDA;
    } catch (const std::exception &e) {
        g_kat_error = e.what();
        return PB2_ERR_STATE;
    }
}
}
