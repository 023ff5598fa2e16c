// Wavefront pt-with-MIS integrator: the replacement for the reference's OptiX megakernel
// (__raygen__main / __miss__default / __closesthit__default, example/path_tracer/main.cu:38-234), one
// optixLaunch + cudaStreamSynchronize per sample (example/path_tracer/pt_pass.cpp:51-53).
//
// A render call executes `n_frames` frames (= samples per pixel, one RNG seed each) in batches of S
// frames; a batch is S*W*H paths living in SoA arrays indexed by path slot p = frame*W*H + pixel.
// Per bounce three kernels run over index queues (no host round trip: queue sizes stay on the device,
// grids are sized for the worst case and exit early):
//   extend   closest-hit traversal of the extension rays; appends each path to the queue of the hit
//            material type (0 = miss) with warp-aggregated atomics  -> material-sorted shading
//   shade    hit geometry + BSDF; emission with MIS; Russian roulette; light sampling (emits a shadow ray
//            with its pending contribution); BSDF sampling (emits the next extension ray)
//   shadow   any-hit traversal; unoccluded rays add their contribution to the path radiance
// and one `accumulate` kernel per batch folds the per-path radiance into the accumulation buffer in
// frame order with the reference's running-mean formula (main.cu:190-196).
//
// RNG draw order, depth semantics, asymmetric MIS and every other observable quirk follow main.cu
// line by line; see the comments in k_shade.
#include "scene.cuh"
#include "pt_math.cuh"
#include "traverse.cuh"
#include <algorithm>

namespace pb2 {

#ifndef PB2_SHADE_PREFETCH
#define PB2_SHADE_PREFETCH 0 // 1: prefetch.global.L2, 2: .L1 of the next path's records — measured slower (45.6 vs 43.3 ms per Cornell step), kept for re-measurement
#endif
constexpr int kNumTypes = 8; // queue 0 = miss, 1..7 = EMatType
constexpr int kCtrPerRound = 16;
// per-round counter slots
enum { CTR_EXT = 0, CTR_SHADOW = 1, CTR_MAT0 = 2 /* ..9 */, CTR_WORK_EXT = 10, CTR_WORK_SHADOW = 11 };

// Path state of one batch in flight.  A render call keeps two of them going on two streams ("lanes") so that the
// HBM-bound shade kernel of one batch runs next to the issue-bound trace kernels of the other (profiles/README.md).
struct Lane {
    uint64_t capacity = 0;
    DevBuf<float4> ray, hit, thr, rad, shq; // 32-byte ray record, 16-byte hit, throughput|rng and radiance records per path slot; 48-byte shadow-queue entries
    DevBuf<uint32_t> q_ext[2], q_mat;
    DevBuf<uint32_t> counters; // (max_depth + 2) rounds x kCtrPerRound
    uint32_t rounds_alloc = 0;
    cudaStream_t stream = nullptr;      // lane 0: the scene's stream; lane 1: `own`
    cudaStream_t own = nullptr;
    cudaEvent_t accumulated = nullptr;  // recorded after the lane's k_accumulate: batches fold into the image in frame order
    void ensure(uint64_t paths, uint32_t rounds) {
        if (paths > capacity) {
            capacity = paths;
            ray.alloc(paths * 2), hit.alloc(paths), thr.alloc(paths), rad.alloc(paths), shq.alloc(paths * 3);
            q_ext[0].alloc(paths), q_ext[1].alloc(paths), q_mat.alloc(paths * kNumTypes);
        }
        if (rounds > rounds_alloc) {
            rounds_alloc = rounds;
            counters.alloc((size_t)rounds * kCtrPerRound);
        }
        if (!accumulated) PB2_CUDA(cudaEventCreateWithFlags(&accumulated, cudaEventDisableTiming));
    }
    ~Lane() {
        if (accumulated) cudaEventDestroy(accumulated);
        if (own) cudaStreamDestroy(own);
    }
};
struct Wavefront {
    Lane lane[2];
    DevBuf<unsigned long long> trav_counters, ray_totals;
    struct Ev {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Ev> events;
    cudaEvent_t t0 = nullptr, t1 = nullptr, fork = nullptr, stagger = nullptr, join = nullptr;
    uint64_t launches = 0;
    uint32_t rounds_used = 0, batches = 0, n_extend = 0, n_shade = 0, n_shadow = 0, lanes_used = 1;
    bool stats_pending = false, sorted = false;

    void ensure() {
        if (!trav_counters.ptr) trav_counters.alloc(8), ray_totals.alloc(2);
        if (!t0) {
            PB2_CUDA(cudaEventCreate(&t0));
            PB2_CUDA(cudaEventCreate(&t1));
            PB2_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
            PB2_CUDA(cudaEventCreateWithFlags(&stagger, cudaEventDisableTiming));
            PB2_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
            PB2_CUDA(cudaStreamCreateWithFlags(&lane[1].own, cudaStreamNonBlocking));
        }
    }
    ~Wavefront() {
        for (auto &e : events) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
        if (t0) cudaEventDestroy(t0), cudaEventDestroy(t1), cudaEventDestroy(fork), cudaEventDestroy(stagger), cudaEventDestroy(join);
    }
};
void wavefront_destroy(Wavefront *wf) { delete wf; }

namespace {

// Path state lives in 32-byte records (one DRAM sector each) indexed by path slot p, so the scattered accesses
// of the material-sorted shade kernel and of the dynamically scheduled trace kernels never pay for half-used
// sectors; fields are grouped by WRITER so no kernel writes a partial record it has not read:
//   ray[2p]   = o.xyz | bits(state: depth | lobe << 16)     ray[2p+1] = d.xyz | bsdf pdf      generate / shade
//   hit[p]    = u, v (triangle) or t, - (sphere) | bits(prim) | bits(inst)                   extend
//   thr[p]    = throughput.xyz | bits(rng)                                                    generate / shade
//   rad[p]    = radiance.xyz | -     touched only by the vertices that add to it: emitter hits and misses in shade, unoccluded
//               shadow rays — most vertices add nothing, so the record is its own 16-byte array instead of the second half of a
//               sector that every vertex would read and write (adjacent path slots share the sector, and queues stay in slot order)
// Shadow rays are created and consumed exactly once, so their payload travels with the queue instead:
//   shq[3k]   = o.xyz | tmax      shq[3k+1] = d.xyz | bits(p)      shq[3k+2] = contribution.xyz | -
// (extension rays always use tmin 1e-3 / tmax 1e16 and shadow rays tmin 1e-4: main.cu:82-83, emitter.h:93-96)
// (Packing the three records of a slot into one 128-byte line was measured and is slower: -8 % Msamples/s, the
// streaming kernels then stride over unused sectors; see profiles/README.md.)
struct PathArrays {
    float4 *ray, *hit, *thr, *rad, *shq;
};
constexpr float kExtendTmin = 0.001f, kExtendTmax = 1e16f, kShadowTmin = 0.0001f;
struct FrameParams {
    uint32_t width, height, n_pixels;
    uint32_t max_depth;
    uint32_t first_seed, seed_stride;
    uint32_t frames; // frames in this batch
};

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// warp-aggregated appends, called by all 32 lanes of a converged warp: one atomicAdd per warp (per
// distinct key for the keyed version); lanes with pred == false get no slot.
__device__ __forceinline__ uint32_t warp_append(uint32_t *counter, bool pred) {
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (!mask) return 0;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(mask & lanemask_lt());
}
__device__ __forceinline__ uint32_t warp_append_keyed(uint32_t *counters, uint32_t key, bool pred) {
    const uint32_t peers = __match_any_sync(0xffffffffu, pred ? key : 0xffffffffu);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (pred && (int)(threadIdx.x & 31) == leader) base = atomicAdd(&counters[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & lanemask_lt());
}

// ---- generate: main.cu:50-78 ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_generate(PathArrays pa, FrameParams fp, Camera cam, uint32_t *__restrict__ q_ext, uint32_t n_paths, uint32_t *__restrict__ n_ext) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_ext = n_paths; // size of the first extension queue (was a pageable H2D copy from a stack variable per batch)
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_paths; p += gridDim.x * blockDim.x) {
        const uint32_t frame = p / fp.n_pixels, pixel = p - frame * fp.n_pixels;
        const uint32_t y = pixel / fp.width, x = pixel - y * fp.width;
        uint32_t rng = rng_init(4, pixel, fp.first_seed + frame * fp.seed_stride); // :55
        const float jx = rng_next(rng), jy = rng_next(rng);                         // :58 (x drawn first)
        const float4 pf = make_float4((static_cast<float>(x) + jx) / static_cast<float>(fp.width),
                                      (static_cast<float>(y) + jy) / static_cast<float>(fp.height), 0.f, 1.f);
        float4 d = make_float4(dot(cam.s2c[0], pf), dot(cam.s2c[1], pf), dot(cam.s2c[2], pf), dot(cam.s2c[3], pf)); // :67
        // IEEE division / square root spelled out: the primary ray must not depend on --prec-div / --prec-sqrt (a last-bit
        // change of the direction moves silhouette pixels to another primitive)
        const float inv = __fdiv_rn(1.0f, d.w);                                                                     // :69
        d = make_float4(d.x * inv, d.y * inv, d.z * inv, 0.f);                                                      // :70
        const float inv_len = __fdiv_rn(1.0f, __fsqrt_rn(dot(d, d)));                                               // :71
        d = make_float4(d.x * inv_len, d.y * inv_len, d.z * inv_len, 0.f);
        const float3 dw = mk3(dot(cam.c2w[0], d), dot(cam.c2w[1], d), dot(cam.c2w[2], d));
        const float3 dir = dw * __fdiv_rn(1.0f, __fsqrt_rn(dot(dw, dw))); // :73 normalize
        pa.ray[2 * (size_t)p] = make_float4(cam.c2w[0].w, cam.c2w[1].w, cam.c2w[2].w, __uint_as_float(0u)); // :75-78; state = depth 0
        pa.ray[2 * (size_t)p + 1] = make_float4(dir.x, dir.y, dir.z, 0.f);
        pa.thr[p] = make_float4(1.f, 1.f, 1.f, __uint_as_float(rng));
        pa.rad[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        q_ext[p] = p;
    }
}

// ---- extend / shadow: persistent BVH8 traversal over an index queue (traverse.cuh) ----------------------------
struct ExtendIO {
    SceneView sv;
    PathArrays pa;
    const uint32_t *__restrict__ q_in;
    uint32_t *__restrict__ q_mat, *__restrict__ mat_counts;
    uint32_t n, capacity;
    int sort;
    PB2_D uint32_t size() const { return n; }
    PB2_D uint32_t load(uint32_t i, float3 &o, float3 &d, float &tmin, float &tmax) const {
        const uint32_t p = q_in[i];
        const float4 ro = pa.ray[2 * (size_t)p], rd = pa.ray[2 * (size_t)p + 1];
        o = mk3(ro), d = mk3(rd), tmin = kExtendTmin, tmax = kExtendTmax;
        return p;
    }
    // writes the hit record and appends the path to the queue of the hit material type (0 = miss)
    PB2_D void commit(bool valid, uint32_t p, const RayHit &h, bool hit, uint32_t inst_of_hit) const {
        uint32_t type = 0;
        if (valid) {
            int32_t inst = -1;
            uint32_t prim = 0;
            bool sphere = false;
            if (hit) {
                const float4 *rec = reinterpret_cast<const float4 *>(sv.prims + h.prim_slot);
                prim = __float_as_uint(__ldg(rec).w);
                inst = inst_of_hit != 0xffffffffu ? (int32_t)inst_of_hit : (int32_t)__float_as_uint(__ldg(rec + 1).w);
                sphere = __float_as_uint(__ldg(rec + 2).w) != 0u;
                type = (uint32_t)__ldg(&sv.instances[inst].mat_type) & 7u;
            }
            // 16 bytes are all the shade stage needs: barycentrics for a triangle (its position comes from the vertices), t for a
            // sphere (position = o + t d), and the two ids
            pa.hit[p] = make_float4(sphere ? h.t : h.u, h.v, __uint_as_float(prim), __int_as_float(inst));
        }
        // no queue work here: rays finish in scheduling order, and appending in that order would scatter the next
        // kernel's path-state accesses.  k_bin (sorted mode) or k_shade itself (unsorted) walk the queue in order.
        (void)type;
    }
};
struct ShadowIO {
    PathArrays pa;
    uint32_t n;
    unsigned long long *unoccluded; // counting mode only (roofline bookkeeping), else nullptr
    PB2_D uint32_t size() const { return n; }
    PB2_D uint32_t load(uint32_t i, float3 &o, float3 &d, float &tmin, float &tmax) const {
        const float4 ro = pa.shq[3 * (size_t)i], rd = pa.shq[3 * (size_t)i + 1];
        o = mk3(ro), d = mk3(rd), tmin = kShadowTmin, tmax = ro.w;
        return i;
    }
    PB2_D void commit(bool valid, uint32_t i, const RayHit &, bool hit, uint32_t) const {
        if (valid && !hit) { // main.cu:127 `if (!occluded)`
            const uint32_t p = __float_as_uint(pa.shq[3 * (size_t)i + 1].w);
            const float4 c = pa.shq[3 * (size_t)i + 2];
            float4 r = pa.rad[p];
            r.x += c.x, r.y += c.y, r.z += c.z;
            pa.rad[p] = r;
            if (unoccluded) atomicAdd(unoccluded, 1ull);
        }
    }
};

template<bool COUNT, bool COOP, bool TRIS = false, bool INST = false>
__global__ void __launch_bounds__(128, PB2_TRACE_MINB(COOP)) k_extend(SceneView sv, PathArrays pa, const uint32_t *__restrict__ q_in, const uint32_t *__restrict__ n_in,
                                                uint32_t *__restrict__ q_mat, uint32_t *__restrict__ mat_counts, uint32_t capacity, int sort,
                                                uint32_t *__restrict__ work, unsigned long long *__restrict__ trav, int refill) {
    ExtendIO io{ sv, pa, q_in, q_mat, mat_counts, *n_in, capacity, sort };
    TraceCounters ctr{ 0, 0 };
    trace_persistent<false, COUNT, COOP, TRIS, INST>(sv, io, work, &ctr, refill);
    if (COUNT) {
        atomicAdd(&trav[0], (unsigned long long)ctr.nodes);
        atomicAdd(&trav[1], (unsigned long long)ctr.prims);
    }
}
template<bool COUNT, bool COOP, bool TRIS = false, bool INST = false>
__global__ void __launch_bounds__(128, PB2_TRACE_MINB(COOP)) k_shadow(SceneView sv, PathArrays pa, const uint32_t *__restrict__ n_in, uint32_t *__restrict__ work,
                                                unsigned long long *__restrict__ trav, int refill) {
    ShadowIO io{ pa, *n_in, COUNT ? trav + 4 : nullptr };
    TraceCounters ctr{ 0, 0 };
    trace_persistent<true, COUNT, COOP, TRIS, INST>(sv, io, work, &ctr, refill);
    if (COUNT) {
        atomicAdd(&trav[2], (unsigned long long)ctr.nodes);
        atomicAdd(&trav[3], (unsigned long long)ctr.prims);
    }
}

// ---- closest-hit program: Geometry::GetHitLocalGeometry, framework/render/geometry.h:60-100,176-180 ------------
struct LocalGeometry {
    float3 position, normal;
    float2 texcoord;
};
__device__ __forceinline__ void hit_local_geometry(const DevInstance *in, uint32_t flags, float3 ro, float3 rd, float t, float bu, float bv,
                                                   uint32_t prim, LocalGeometry &g) {
    const float4 i0 = __ldg(&in->inv[0]), i1 = __ldg(&in->inv[1]), i2 = __ldg(&in->inv[2]);
    g.texcoord = make_float2(0.f, 0.f); // defined: the reference leaves it untouched for meshes without uvs
    if (flags & PB2_IF_SPHERE) {
        g.position = ro + t * rd;
        const float3 local_pos = xf_point(i0, i1, i2, g.position);
        g.texcoord = sphere_texcoord(normalize(local_pos));
        g.normal = normalize(xf_normal_t(i0, i1, i2, local_pos));
    } else {
        const float4 x0 = __ldg(&in->xf[0]), x1 = __ldg(&in->xf[1]), x2 = __ldg(&in->xf[2]);
        const float *pos = in->pos, *nrm = in->nrm, *uv = in->uv;
        const uint32_t *idx = in->idx + (size_t)prim * 3;
        const uint32_t v0 = __ldg(idx), v1 = __ldg(idx + 1), v2 = __ldg(idx + 2);
        auto ld3 = [](const float *a, uint32_t v) { return mk3(__ldg(a + (size_t)v * 3), __ldg(a + (size_t)v * 3 + 1), __ldg(a + (size_t)v * 3 + 2)); };
        const float3 p0 = ld3(pos, v0), p1 = ld3(pos, v1), p2 = ld3(pos, v2);
        const float w = 1.f - bu - bv;
        g.position = xf_point(x0, x1, x2, w * p0 + bu * p1 + bv * p2);
        float3 n;
        if (flags & PB2_IF_HAS_NRM) n = w * ld3(nrm, v0) + bu * ld3(nrm, v1) + bv * ld3(nrm, v2);
        else n = cross(p1 - p0, p2 - p0);
        g.normal = normalize(xf_normal_t(i0, i1, i2, n));
        if (flags & PB2_IF_HAS_UV) {
            auto ld2 = [](const float *a, uint32_t v) { return make_float2(__ldg(a + (size_t)v * 2), __ldg(a + (size_t)v * 2 + 1)); };
            g.texcoord = w * ld2(uv, v0) + bu * ld2(uv, v1) + bv * ld2(uv, v2);
            if (flags & PB2_INST_FLIP_TEX) g.texcoord.y = 1.f - g.texcoord.y;
        }
    }
    if (flags & PB2_INST_FLIP_NORMALS) g.normal *= -1.f;
    if (dot(-rd, g.normal) < 0.f && (flags & PB2_IF_TWOSIDED)) g.normal = -g.normal;
}

// ---- shade -------------------------------------------------------------------------------------------------
struct ShadeOut {
    uint32_t *q_ext, *n_ext, *n_shadow;
    float *albedo, *normal, *test; // AOVs (may be null); written by the last frame of the batch only
    uint32_t aov_first_slot;       // first path slot of the frame whose primary hits write the AOVs (its pixels follow); ~0u = none
};

// A shadow ray produced by shade_path; it is appended to the shadow queue by the caller (warp-aggregated).
struct ShadowRay {
    float3 o, d, contrib;
    float tmax;
};
// One path at one hit (or miss).  Returns bit 0: `sh` holds a shadow ray, bit 1: an extension ray was written.
// ONLY >= 0: every instance of the scene has this material type (Scene::only_material_type), the other six BSDFs are compiled out
template<int ONLY>
__device__ __forceinline__ uint32_t shade_path(const SceneView &sv, const PathArrays &pa, const FrameParams &fp, const ShadeOut &out, uint32_t p, ShadowRay &sh) {
    const float4 hit = pa.hit[p]; // u, v (triangle) or t (sphere), prim, inst
    const int32_t inst = __float_as_int(hit.w);
    const float4 ro4 = pa.ray[2 * (size_t)p], rd4 = pa.ray[2 * (size_t)p + 1];
    const float3 ray_o = mk3(ro4), ray_d = mk3(rd4);
    const float4 thr4 = pa.thr[p];
    float3 throughput = mk3(thr4);
    const float bsdf_pdf = rd4.w;
    const uint32_t st = __float_as_uint(ro4.w);
    uint32_t depth = st & 0xffffu;
    const uint32_t sampled_type = st >> 16;
    uint32_t rng = __float_as_uint(thr4.w);
    float3 radiance = mk3(0.f); // what this vertex adds to the path's radiance (at most one term: emission, or the environment on a miss)
    auto add_radiance = [&]() {
        if (radiance.x != 0.f || radiance.y != 0.f || radiance.z != 0.f) {
            float4 r = pa.rad[p];
            r.x += radiance.x, r.y += radiance.y, r.z += radiance.z;
            pa.rad[p] = r;
        }
    };
    // p = frame * n_pixels + pixel; only the AOV frame needs the pixel, so no per-path integer division
    const uint32_t pixel = p - out.aov_first_slot;
    const bool write_aov = depth == 0 && out.aov_first_slot != ~0u && pixel < fp.n_pixels;

    if (inst < 0) { // __miss__default, main.cu:199-215
        float3 env_radiance = mk3(0.f);
        float env_pdf = 0.f;
        if (sv.env) emitter_eval(sv.env, ray_o + normalize(ray_d), mk3(0.f), make_float2(0.f, 0.f), ray_o, env_radiance, env_pdf);
        if (depth == 0) {
            if (write_aov) { // :100-104
                if (out.albedo) out.albedo[pixel * 3] = 0.f, out.albedo[pixel * 3 + 1] = 0.f, out.albedo[pixel * 3 + 2] = 0.f;
                if (out.normal) out.normal[pixel * 3] = 0.f, out.normal[pixel * 3 + 1] = 0.f, out.normal[pixel * 3 + 2] = 0.f;
                if (out.test) out.test[pixel] = rng_next(rng);
            }
        } else if (sv.env) { // :168-172
            const float mis = mis_weight(bsdf_pdf, env_pdf);
            env_radiance *= throughput * mis;
        }
        radiance += env_radiance; // :188
        add_radiance();
        return 0u;
    }
    if (ONLY == 0) return 0u; // the miss queue of the sorted mode holds nothing else

    // __closesthit__default, main.cu:220-234
    const DevInstance *in = sv.instances + inst;
    const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(&in->flags)); // flags, mat_type, emitter_offset, n_tris
    const uint32_t flags = meta.x;
    const uint32_t prim = __float_as_uint(hit.z);
    LocalGeometry geo;
    hit_local_geometry(in, flags, ray_o, ray_d, hit.x, hit.x, hit.y, prim, geo);
    const int emitter_index = (int)meta.z >= 0 ? (int)meta.z + (int)prim : -1;
    const LocalBsdf bsdf = get_local_bsdf(sv.materials + inst, geo.texcoord, ONLY);

    if (depth == 0) { // :90-104
        if (emitter_index >= 0) radiance += emitter_radiance(sv.areas + emitter_index, geo.texcoord);
        if (write_aov) {
            const float3 a = local_albedo(bsdf);
            if (out.albedo) out.albedo[pixel * 3] = a.x, out.albedo[pixel * 3 + 1] = a.y, out.albedo[pixel * 3 + 2] = a.z;
            if (out.normal) out.normal[pixel * 3] = geo.normal.x, out.normal[pixel * 3 + 1] = geo.normal.y, out.normal[pixel * 3 + 2] = geo.normal.z;
        }
        const float test = rng_next(rng); // :104 — one draw per path whether or not the buffer exists
        if (write_aov && out.test) out.test[pixel] = test;
    } else if (emitter_index >= 0) { // :175-185
        const DevEmitter *em = sv.areas + emitter_index;
        float3 le;
        float pdf;
        emitter_eval(em, geo.position, geo.normal, geo.texcoord, ray_o, le, pdf);
        if (!is_zero(pdf)) {
            const float mis = (sampled_type & kLobeDelta) ? 1.f : mis_weight(bsdf_pdf, pdf * __ldg(&em->select_probability));
            radiance += throughput * le * mis;
        }
    }

    // ---- loop head, :106-114 ----
    bool alive = true;
    ++depth;
    if (depth >= fp.max_depth) alive = false;
    if (alive) {
        const float rr = depth > 2 ? 0.95 : 1.0;
        if (rng_next(rng) > rr) alive = false;
        else throughput /= rr;
    }
    add_radiance();
    if (!alive) return 0u;
    const Onb frame_onb(geo.normal);
    const float3 wo = frame_onb.to_local(-ray_d);
    uint32_t emitted = 0;

    // ---- direct light sampling, :117-144 ----
    {
        const float sel = rng_next(rng);
        const float e0 = rng_next(rng), e1 = rng_next(rng);
        const DevEmitter *em = select_emitter(sv.areas, sv.area_cdf, sv.n_areas, sv.env, sel);
        if (em) {
            EmitSample es;
            emitter_sample_direct(em, geo.position, geo.normal, frame_onb, make_float2(e0, e1), es);
            if (es.pdf != 0.f) {
                // the reference traces first and evaluates after; evaluating first lets rays whose
                // contribution is zero anyway be skipped — same image, fewer rays
                BsdfRec rec;
                rec.wi = frame_onb.to_local(es.wi), rec.wo = wo;
                bsdf_eval(bsdf, rec);
                if (!is_zero(rec.f * es.pdf)) {
                    const float NoL = dot(geo.normal, es.wi);
                    if (NoL > 0.f) {
                        const float mis = es.is_delta ? 1.f : mis_weight(es.pdf, rec.pdf);
                        const float pdf_e = es.pdf * __ldg(&em->select_probability);
                        const float3 contrib = throughput * es.radiance * rec.f * NoL * mis / pdf_e;
                        sh.o = geo.position, sh.d = es.wi, sh.tmax = es.distance - 0.0001f, sh.contrib = contrib;
                        emitted |= 1u;
                    }
                }
            }
        }
    }
    // ---- BSDF sampling, :146-166 ----
    {
        BsdfRec rec;
        rec.wo = wo;
        bsdf_sample(bsdf, rec, rng);
        if (!(is_zero(rec.f * fabsf(rec.wi.z)) || is_zero(rec.pdf))) {
            throughput *= rec.f * fabsf(rec.wi.z) / rec.pdf;
            const float3 dir = frame_onb.to_world(rec.wi);
            pa.ray[2 * (size_t)p] = make_float4(geo.position.x, geo.position.y, geo.position.z, __uint_as_float(depth | (rec.type << 16)));
            pa.ray[2 * (size_t)p + 1] = make_float4(dir.x, dir.y, dir.z, rec.pdf);
            pa.thr[p] = make_float4(throughput.x, throughput.y, throughput.z, __uint_as_float(rng));
            emitted |= 2u;
        }
    }
    return emitted;
}

// ---- bin: stable split of the extension queue into the eight per-material queues ------------------------------
// Walks the queue in order and keeps that order inside every material queue (per CTA slice of 256 entries; one
// atomic per present type per CTA), so k_shade's scattered record accesses stay ascending in the path slot.
__global__ void __launch_bounds__(256) k_bin(SceneView sv, PathArrays pa, const uint32_t *__restrict__ q_in, const uint32_t *__restrict__ n_in,
                                             uint32_t *__restrict__ q_mat, uint32_t *__restrict__ mat_counts, uint32_t capacity) {
    __shared__ uint32_t s_cnt[8][kNumTypes], s_base[kNumTypes];
    const uint32_t n = *n_in, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t total = (n + 255u) & ~255u;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        if (threadIdx.x < 8 * kNumTypes) (&s_cnt[0][0])[threadIdx.x] = 0u;
        __syncthreads();
        const bool valid = i < n;
        uint32_t p = 0, type = kNumTypes;
        if (valid) {
            p = q_in[i];
            const int32_t inst = __float_as_int(pa.hit[p].w);
            type = inst < 0 ? 0u : ((uint32_t)__ldg(&sv.instances[inst].mat_type) & 7u);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, type);
        const uint32_t rank = __popc(peers & lanemask_lt());
        if (valid && rank == 0) s_cnt[warp][type] = __popc(peers);
        __syncthreads();
        if (threadIdx.x < kNumTypes) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = s_cnt[w][threadIdx.x];
                s_cnt[w][threadIdx.x] = run;
                run += c;
            }
            s_base[threadIdx.x] = run ? atomicAdd(&mat_counts[threadIdx.x], run) : 0u;
        }
        __syncthreads();
        if (valid) q_mat[(size_t)type * capacity + s_base[type] + s_cnt[warp][type] + rank] = p;
        __syncthreads();
    }
}

// Order-preserving append of a 128-thread CTA to two queues at once: thread order is kept inside the CTA's slice
// (a CTA works on 128 consecutive queue entries, so runs of ascending path slots survive compaction and the next
// kernel's record accesses stay close to sequential), one atomic per queue per CTA.
#ifndef PB2_SHADE_PER_TYPE
#define PB2_SHADE_PER_TYPE 1
#endif
#ifndef PB2_SHADE_WAVES
#define PB2_SHADE_WAVES 2
#endif
#ifndef PB2_SHADE_THREADS
#define PB2_SHADE_THREADS 128
#endif
constexpr uint32_t kShadeThreads = PB2_SHADE_THREADS, kShadeWarps = kShadeThreads / 32, kShadeMask = kShadeThreads - 1;
__device__ __forceinline__ void block_append2(uint32_t *counter_a, uint32_t *counter_b, bool pred_a, bool pred_b, uint32_t &pos_a, uint32_t &pos_b, uint32_t parity) {
    // two copies of the scratch words, used alternately by successive calls (`parity`): the next call writes the other copy
    // and the call after that is two barriers away, so no third barrier is needed to protect the reads below
    __shared__ uint32_t s_cnt[2][2][kShadeWarps], s_base[2][2];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, b = parity & 1u;
    const uint32_t ma = __ballot_sync(0xffffffffu, pred_a), mb = __ballot_sync(0xffffffffu, pred_b);
    if (lane == 0) s_cnt[b][0][warp] = __popc(ma), s_cnt[b][1][warp] = __popc(mb);
    __syncthreads();
    if (threadIdx.x < 2) {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < (int)kShadeWarps; ++w) {
            const uint32_t c = s_cnt[b][threadIdx.x][w];
            s_cnt[b][threadIdx.x][w] = run;
            run += c;
        }
        s_base[b][threadIdx.x] = run ? atomicAdd(threadIdx.x ? counter_b : counter_a, run) : 0u;
    }
    __syncthreads();
    pos_a = s_base[b][0] + s_cnt[b][0][warp] + __popc(ma & lanemask_lt());
    pos_b = s_base[b][1] + s_cnt[b][1][warp] + __popc(mb & lanemask_lt());
}

// MINB = resident CTAs per SM the register allocator must allow: 4 -> 114 registers, no spills; 6 -> 80 registers,
// 44 B of spills; 8 -> 64 registers, 160 B of spills.  The choice is measured, see profiles/.
// SORTED: paths come from the eight per-material queues filled by k_extend (queue t occupies the virtual index
// range [start_t, start_t + round_up(count_t, 128)): a CTA never straddles two material types); otherwise the
// kernel walks the extension queue itself, in order, and branches on the material per path.
template<int MINB, bool SORTED, int ONLY = -1>
__global__ void __launch_bounds__(kShadeThreads, MINB * 128 / kShadeThreads) k_shade(SceneView sv, PathArrays pa, FrameParams fp, const uint32_t *__restrict__ queue,
                                               const uint32_t *__restrict__ counts, uint32_t capacity, ShadeOut out) {
    // MULTI: all eight material queues in one launch (SORTED with ONLY < 0, kept for A/B runs); otherwise one queue — the
    // extension queue itself (unsorted), or material queue ONLY of the sorted mode, one launch per material type of the scene
    constexpr bool MULTI = SORTED && ONLY < 0;
    // first virtual index and length of every material queue, once per CTA in shared memory (as a per-thread array it was spilled:
    // ncu charged the look-up below 18 M local-memory sectors per launch)
    __shared__ uint32_t start[kNumTypes + 1], s_count[kNumTypes];
    if (MULTI) {
        if (threadIdx.x == 0) {
            uint32_t run = 0;
#pragma unroll
            for (int t = 0; t < kNumTypes; ++t) {
                const uint32_t c = counts[t];
                start[t] = run, s_count[t] = c;
                run += (c + kShadeMask) & ~kShadeMask;
            }
            start[kNumTypes] = run;
        }
        __syncthreads();
    }
    if (SORTED && !MULTI) queue += (size_t)ONLY * capacity;
    const uint32_t n_single = MULTI ? 0u : counts[SORTED ? ONLY : 0]; // read once: the loop below would reload it from memory every iteration
    const uint32_t total = MULTI ? start[kNumTypes] : ((n_single + kShadeMask) & ~kShadeMask);

    // queue entry of virtual index vi (MULTI: the eight material queues laid end to end, each padded to 128)
    auto fetch = [&](uint32_t vi, uint32_t &p) -> bool {
        if (MULTI) {
            int t = 0;
#pragma unroll
            for (int k = 1; k < kNumTypes; ++k) t += vi >= start[k] ? 1 : 0;
            const uint32_t local = vi - start[t];
            if (local >= s_count[t]) return false;
            p = queue[(size_t)t * capacity + local];
        } else {
            if (vi >= n_single) return false;
            p = queue[vi];
        }
        return true;
    };
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t p_next = 0, iter = 0;
    bool valid_next = blockIdx.x * blockDim.x + threadIdx.x < total && fetch(blockIdx.x * blockDim.x + threadIdx.x, p_next);
    for (uint32_t vi = blockIdx.x * blockDim.x + threadIdx.x; vi < total; vi += stride) { // total % 128 == 0: CTA-uniform
        uint32_t emitted = 0;
        const uint32_t p = p_next;
        const bool valid = valid_next;
        ShadowRay sh;
#if PB2_SHADE_PREFETCH
        // the next iteration's queue entry is fetched now and its three 32-byte records are requested while this path is
        // shaded: ncu showed a third of k_shade's stall samples waiting on exactly these loads (profiles/r1c_ncu.md)
        valid_next = vi + stride < total && fetch(vi + stride, p_next);
#if PB2_SHADE_PREFETCH < 3
        if (valid_next) {
#if PB2_SHADE_PREFETCH == 1
            prefetch_l2(pa.hit + p_next), prefetch_l2(pa.ray + 2 * (size_t)p_next), prefetch_l2(pa.thr + p_next);
#else
            prefetch_l1(pa.hit + p_next), prefetch_l1(pa.ray + 2 * (size_t)p_next), prefetch_l1(pa.thr + p_next);
#endif
        }
#endif
#endif
        if (valid) emitted = shade_path<ONLY>(sv, pa, fp, out, p, sh);
#if PB2_SHADE_PREFETCH >= 3
        // variant: the queue entry has arrived by now, so the prefetches do not wait for it
        if (valid_next) {
#if PB2_SHADE_PREFETCH == 3
            prefetch_l2(pa.hit + p_next), prefetch_l2(pa.ray + 2 * (size_t)p_next), prefetch_l2(pa.thr + p_next);
#else
            prefetch_l1(pa.hit + p_next), prefetch_l1(pa.ray + 2 * (size_t)p_next), prefetch_l1(pa.thr + p_next);
#endif
        }
#endif
#if !PB2_SHADE_PREFETCH
        valid_next = vi + stride < total && fetch(vi + stride, p_next);
#endif
        uint32_t ps, pe;
        block_append2(out.n_shadow, out.n_ext, emitted & 1u, emitted & 2u, ps, pe, iter++);
        if (emitted & 1u) { // consecutive threads write consecutive 48-byte queue entries
            pa.shq[3 * (size_t)ps] = make_float4(sh.o.x, sh.o.y, sh.o.z, sh.tmax);
            pa.shq[3 * (size_t)ps + 1] = make_float4(sh.d.x, sh.d.y, sh.d.z, __uint_as_float(p));
            pa.shq[3 * (size_t)ps + 2] = make_float4(sh.contrib.x, sh.contrib.y, sh.contrib.z, 0.f);
        }
        if (emitted & 2u) out.q_ext[pe] = p;
    }
}

// adds one batch's per-round queue sizes into the 64-bit ray totals: [0] closest, [1] shadow
__global__ void k_collect_counts(const uint32_t *__restrict__ counters, uint32_t rounds, unsigned long long *__restrict__ totals) {
    unsigned long long c = 0, sh = 0;
    for (uint32_t r = threadIdx.x; r < rounds; r += blockDim.x) {
        c += counters[r * kCtrPerRound + CTR_EXT];
        if (r + 1 < rounds) sh += counters[r * kCtrPerRound + CTR_SHADOW];
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o), sh += __shfl_xor_sync(0xffffffffu, sh, o);
    if (threadIdx.x == 0) atomicAdd(&totals[0], c), atomicAdd(&totals[1], sh); // two lanes may finish a batch at the same time
}

// ---- accumulate: main.cu:190-196 ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_accumulate(const float4 *__restrict__ rad, uint32_t n_pixels, uint32_t frames, uint32_t mode, uint32_t sample_cnt0,
                                                    float4 *__restrict__ accum, float4 *__restrict__ frame_buf) {
    for (uint32_t px = blockIdx.x * blockDim.x + threadIdx.x; px < n_pixels; px += gridDim.x * blockDim.x) {
        // sample_cnt0 == 0 starts over in both accumulating modes: the buffer's previous content is not read
        float3 acc = (mode != 0 && sample_cnt0 > 0) ? mk3(accum[px]) : mk3(0.f);
        float w = mode == 2 ? (sample_cnt0 > 0 ? accum[px].w : 0.f) : 1.f;
        for (uint32_t f = 0; f < frames; ++f) {
            const float3 r = mk3(rad[(size_t)f * n_pixels + px]);
            if (mode == 2) { // plain sum for sample-sharded multi-GPU rendering
                acc += r;
                w += 1.f;
            } else {
                const uint32_t sample_cnt = mode == 1 ? sample_cnt0 + f : 0u;
                if (mode == 1 && sample_cnt > 0) {
                    const float t = 1.f / (sample_cnt + 1.f);
                    acc = lerp3(acc, r, t);
                } else {
                    acc = r;
                }
            }
        }
        const float4 o = make_float4(acc.x, acc.y, acc.z, w);
        accum[px] = o;
        if (frame_buf && mode != 2) frame_buf[px] = o;
    }
}
__global__ void k_finalize_sum(const float4 *__restrict__ sum, float4 *__restrict__ frame, uint64_t n, float inv_spp) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 s = sum[i];
        frame[i] = make_float4(s.x * inv_spp, s.y * inv_spp, s.z * inv_spp, 1.f);
    }
}
}// namespace

void finalize_sum_on(cudaStream_t st, const float4 *sum, float4 *frame, uint64_t n, uint32_t spp) {
    if (!n) return;
    k_finalize_sum<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(sum, frame, n, 1.f / (float)spp);
    PB2_LAUNCH_CHECK();
}
void finalize_sum(Scene &s, const float4 *sum, float4 *frame, uint64_t n, uint32_t spp) { finalize_sum_on(s.stream, sum, frame, n, spp); }

void render(Scene &s, const pb2_launch_params &lp) {
    if (!s.bvh_valid) throw std::runtime_error("pb2_render: call pb2_bvh_build first");
    if (!lp.accum_buffer || !lp.width || !lp.height) throw std::runtime_error("pb2_render: accum_buffer / width / height missing");
    if (lp.max_depth > 0xffffu) throw std::runtime_error("pb2_render: max_depth too large");
    s.check_emitter_ranges();
    s.upload_tables();
    if (!s.wf) s.wf = new Wavefront();
    Wavefront &wf = *s.wf;
    wf.ensure();

    const uint32_t n_pixels = lp.width * lp.height;
    const uint32_t n_frames = std::max(1u, lp.n_frames);
    // paths in flight: measured on the Cornell box at 1080p (profiles/README.md) 2 Mi -> 927, 4 Mi -> 1054, 16 Mi -> 1202,
    // 32 Mi -> 1226, 64 Mi -> 1239 Msamples/s with one batch in flight (later bounces leave short queues; bigger batches
    // amortise their launch gaps and tails).  Round 2, final kernels: 32 Mi -> 1508, 64 Mi -> 1523, 128 Mi -> 1537 Msamples/s.  The default
    // is 128 Mi paths = 18.8 GB of path state out of 180 GB (allocated only as far as a render call needs it), shared by the two lanes.
    const uint64_t target = s.paths_in_flight ? s.paths_in_flight : (128ull << 20);
    const uint32_t n_lanes = (s.two_lanes && n_frames >= 2) ? 2u : 1u;
    const uint32_t S = (uint32_t)std::min<uint64_t>((n_frames + n_lanes - 1) / n_lanes, std::max<uint64_t>(1, target / n_lanes / n_pixels));
    const uint32_t rounds = std::max(1u, lp.max_depth);
    if ((uint64_t)S * n_pixels >= 0xffffffffull) throw std::runtime_error("pb2_render: too many paths in flight");
    if (s.l2_dirty) s.apply_l2_window();
    wf.lane[0].stream = s.stream, wf.lane[1].stream = wf.lane[1].own;
    for (uint32_t l = 0; l < n_lanes; ++l) wf.lane[l].ensure((uint64_t)S * n_pixels, rounds + 1);

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const SceneView sv = s.view();
    const bool coop = s.use_coop_prims();
    const bool tris = s.build_stats.n_spheres == 0; // no analytic spheres: the trace kernels without the sphere branch
    const bool inst = s.n_blas > 0;                  // two-level scene: the trace kernels that follow instance nodes
    // material sorting pays when shading diverges: on by default only for scenes with more than one material type
    const bool sorted = s.sort_by_material == 1 || (s.sort_by_material < 0 && s.n_material_types > 1);
    wf.sorted = sorted, wf.lanes_used = n_lanes;

    for (auto &e : wf.events) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
    wf.events.clear();
    wf.launches = 0, wf.batches = 0, wf.rounds_used = rounds, wf.n_extend = wf.n_shade = wf.n_shadow = 0;
    cudaStream_t st = s.stream; // stream of the batch being issued
    auto stage_begin = [&](int stage) {
        if (!s.profiling) return;
        Wavefront::Ev e{ stage, nullptr, nullptr };
        cudaEventCreate(&e.a), cudaEventCreate(&e.b);
        cudaEventRecord(e.a, st);
        wf.events.push_back(e);
    };
    auto stage_end = [&]() {
        if (s.profiling) cudaEventRecord(wf.events.back().b, st);
    };
    if (s.counting) wf.trav_counters.zero(s.stream);
    wf.ray_totals.zero(s.stream);
    PB2_CUDA(cudaEventRecord(wf.t0, s.stream));
    if (n_lanes == 2) { // lane 1 starts after everything already queued on the scene's stream
        PB2_CUDA(cudaEventRecord(wf.fork, s.stream));
        PB2_CUDA(cudaStreamWaitEvent(wf.lane[1].stream, wf.fork, 0));
    }

    uint32_t sample_cnt = lp.sample_cnt;
    uint32_t batch = 0;
    for (uint32_t f0 = 0; f0 < n_frames; f0 += S, ++batch) {
        Lane &ln = wf.lane[batch % n_lanes];
        st = ln.stream;
        PathArrays pa{ ln.ray.ptr, ln.hit.ptr, ln.thr.ptr, ln.rad.ptr, ln.shq.ptr };
        const uint32_t frames = std::min(S, n_frames - f0);
        const uint32_t n_paths = frames * n_pixels;
        FrameParams fp{ lp.width, lp.height, n_pixels, lp.max_depth, lp.random_seed + f0 * (lp.seed_stride ? lp.seed_stride : 1u),
                        lp.seed_stride ? lp.seed_stride : 1u, frames };
        const unsigned grid_stream = (unsigned)std::min<uint64_t>((n_paths + 255) / 256, (uint64_t)sms * 8);
        const unsigned grid_trace = (unsigned)std::min<uint64_t>((n_paths + 127) / 128, (uint64_t)sms * 16);
        const unsigned grid_shade = (unsigned)std::min<uint64_t>((n_paths + kNumTypes * kShadeThreads + kShadeMask) / kShadeThreads,
                                                                   (uint64_t)sms * PB2_SHADE_WAVES * (s.shade_variant >= 4 && s.shade_variant <= 8 ? s.shade_variant : 6) * 128 / kShadeThreads); // full waves of resident CTAs

        // the second lane starts one kernel late, so that its trace kernels meet the first lane's shade kernels rather
        // than both lanes running the same stage side by side
        if (batch == 1 && n_lanes == 2) PB2_CUDA(cudaStreamWaitEvent(st, wf.stagger, 0));
        PB2_CUDA(cudaMemsetAsync(ln.counters.ptr, 0, ln.counters.bytes(), st));
        stage_begin(0);
        k_generate<<<grid_stream, 256, 0, st>>>(pa, fp, s.cam, ln.q_ext[0].ptr, n_paths, ln.counters.ptr + CTR_EXT);
        PB2_LAUNCH_CHECK();
        stage_end();
        ++wf.launches;

        const bool last_batch = f0 + frames >= n_frames;
        for (uint32_t r = 0; r < rounds; ++r) {
            uint32_t *ctr = ln.counters.ptr + (size_t)r * kCtrPerRound, *ctr_next = ctr + kCtrPerRound;
            uint32_t *q_in = ln.q_ext[r & 1].ptr, *q_out = ln.q_ext[(r + 1) & 1].ptr;
            stage_begin(1);
            {
                auto k = inst ? (s.counting ? (coop ? k_extend<true, true, false, true> : k_extend<true, false, false, true>)
                                            : (coop ? k_extend<false, true, false, true> : k_extend<false, false, false, true>))
                         : s.counting ? (coop ? k_extend<true, true> : k_extend<true, false>)
                         : tris       ? (coop ? k_extend<false, true, true> : k_extend<false, false, true>)
                                      : (coop ? k_extend<false, true> : k_extend<false, false>);
                k<<<grid_trace, 128, 0, st>>>(sv, pa, q_in, ctr + CTR_EXT, ln.q_mat.ptr, ctr + CTR_MAT0, (uint32_t)ln.capacity, sorted ? 1 : 0,
                                              ctr + CTR_WORK_EXT, s.counting ? wf.trav_counters.ptr : nullptr, s.refill_threshold);
            }
            PB2_LAUNCH_CHECK();
            stage_end();
            if (batch == 0 && r == 0 && n_lanes == 2) PB2_CUDA(cudaEventRecord(wf.stagger, st));
            if (sorted) {
                stage_begin(2);
                k_bin<<<grid_stream, 256, 0, st>>>(sv, pa, q_in, ctr + CTR_EXT, ln.q_mat.ptr, ctr + CTR_MAT0, (uint32_t)ln.capacity);
                PB2_LAUNCH_CHECK();
                stage_end();
                ++wf.launches;
            }
            ShadeOut so{ q_out, ctr_next + CTR_EXT, ctr + CTR_SHADOW, (float *)lp.albedo_buffer, (float *)lp.normal_buffer,
                         (float *)lp.test_buffer, last_batch ? (frames - 1) * n_pixels : ~0u };
            stage_begin(2);
            // Kernels specialised by material type (k_shade<.., ONLY>: one BSDF compiled in instead of a seven-way switch).
            // A scene with a single material type (unsorted mode) runs the kernel of that type: k_shade -11 % on the Cornell box.
            // In the sorted mode one launch per material queue (PB2_SHADE_PER_TYPE=2) was measured as well and is 2 % SLOWER on
            // the material grid than the single launch over all queues (eight short launches, each with its own tail), so
            // that mode keeps the single launch (profiles/README.md).  Launches of the per-queue form follow each other on the
            // stream in queue order, so the appended queues keep the order of the single-launch form.
            uint32_t shade_launches = 1;
            auto launch_only = [&](int type, const uint32_t *queue, const uint32_t *counts) {
#define PB2_SHADE_CASE(T) \
    case T: k_shade<6, SORTED_T, T><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, queue, counts, (uint32_t)ln.capacity, so); break;
                if (sorted) {
#define SORTED_T true
                    switch (type) { PB2_SHADE_CASE(0) PB2_SHADE_CASE(1) PB2_SHADE_CASE(2) PB2_SHADE_CASE(3) PB2_SHADE_CASE(4) PB2_SHADE_CASE(5) PB2_SHADE_CASE(6) PB2_SHADE_CASE(7) default: break; }
#undef SORTED_T
                } else {
#define SORTED_T false
                    switch (type) { PB2_SHADE_CASE(1) PB2_SHADE_CASE(2) PB2_SHADE_CASE(3) PB2_SHADE_CASE(4) PB2_SHADE_CASE(5) PB2_SHADE_CASE(6) PB2_SHADE_CASE(7) default: break; }
#undef SORTED_T
                }
#undef PB2_SHADE_CASE
            };
            const bool per_type = PB2_SHADE_PER_TYPE && s.shade_variant == 6 && (sorted ? (PB2_SHADE_PER_TYPE >= 2 && s.material_type_mask > 1u && !(s.material_type_mask & 1u)) : s.only_material_type >= 1);
            if (per_type && sorted) {
                shade_launches = 0;
                for (int t = 0; t < (int)kNumTypes; ++t)
                    if (t == 0 || (s.material_type_mask >> t) & 1u) launch_only(t, ln.q_mat.ptr, ctr + CTR_MAT0), ++shade_launches;
            } else if (per_type) {
                launch_only(s.only_material_type, q_in, ctr + CTR_EXT);
            } else if (sorted) {
                switch (s.shade_variant) {
                    case 4: k_shade<4, true><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, ln.q_mat.ptr, ctr + CTR_MAT0, (uint32_t)ln.capacity, so); break;
                    case 7: k_shade<7, true><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, ln.q_mat.ptr, ctr + CTR_MAT0, (uint32_t)ln.capacity, so); break;
                    case 8: k_shade<8, true><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, ln.q_mat.ptr, ctr + CTR_MAT0, (uint32_t)ln.capacity, so); break;
                    default: k_shade<6, true><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, ln.q_mat.ptr, ctr + CTR_MAT0, (uint32_t)ln.capacity, so); break;
                }
            } else {
                switch (s.shade_variant) {
                    case 4: k_shade<4, false><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, q_in, ctr + CTR_EXT, (uint32_t)ln.capacity, so); break;
                    case 7: k_shade<7, false><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, q_in, ctr + CTR_EXT, (uint32_t)ln.capacity, so); break;
                    case 8: k_shade<8, false><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, q_in, ctr + CTR_EXT, (uint32_t)ln.capacity, so); break;
                    default: k_shade<6, false><<<grid_shade, kShadeThreads, 0, st>>>(sv, pa, fp, q_in, ctr + CTR_EXT, (uint32_t)ln.capacity, so); break;
                }
            }
            PB2_LAUNCH_CHECK();
            stage_end();
            wf.launches += 1 + shade_launches, ++wf.n_extend, ++wf.n_shade;
            if (r + 1 < rounds) { // the last round cannot emit rays (depth >= max_depth)
                stage_begin(3);
                auto k = inst ? (s.counting ? (coop ? k_shadow<true, true, false, true> : k_shadow<true, false, false, true>)
                                            : (coop ? k_shadow<false, true, false, true> : k_shadow<false, false, false, true>))
                         : s.counting ? (coop ? k_shadow<true, true> : k_shadow<true, false>)
                         : tris       ? (coop ? k_shadow<false, true, true> : k_shadow<false, false, true>)
                                      : (coop ? k_shadow<false, true> : k_shadow<false, false>);
                k<<<grid_trace, 128, 0, st>>>(sv, pa, ctr + CTR_SHADOW, ctr + CTR_WORK_SHADOW, s.counting ? wf.trav_counters.ptr : nullptr, s.refill_threshold);
                PB2_LAUNCH_CHECK();
                stage_end();
                ++wf.launches, ++wf.n_shadow;
            }
        }
        // batches fold into the image in frame order (the running mean of main.cu:190-196 depends on it): wait for the
        // previous batch's accumulate, which ran on the other lane
        if (batch > 0 && n_lanes == 2) PB2_CUDA(cudaStreamWaitEvent(st, wf.lane[(batch - 1) % n_lanes].accumulated, 0));
        if (s.gate_pending) { // a multi-GPU reduction is still reading the accumulation buffer (comm.cu): this is its only writer
            PB2_CUDA(cudaStreamWaitEvent(st, s.accumulate_gate, 0));
            s.gate_pending = false;
        }
        stage_begin(4);
        k_accumulate<<<(unsigned)std::min<uint64_t>((n_pixels + 255) / 256, (uint64_t)sms * 8), 256, 0, st>>>(
            ln.rad.ptr, n_pixels, frames, lp.accumulate, sample_cnt, (float4 *)lp.accum_buffer, (float4 *)lp.frame_buffer);
        PB2_LAUNCH_CHECK();
        stage_end();
        if (n_lanes == 2) PB2_CUDA(cudaEventRecord(ln.accumulated, st));
        ++wf.launches;
        if (lp.accumulate != 0) sample_cnt += frames;
        k_collect_counts<<<1, 32, 0, st>>>(ln.counters.ptr, rounds, wf.ray_totals.ptr);
        PB2_LAUNCH_CHECK();
        ++wf.launches;
        ++wf.batches;
    }
    if (n_lanes == 2) { // the scene's stream continues after both lanes
        PB2_CUDA(cudaEventRecord(wf.join, wf.lane[1].stream));
        PB2_CUDA(cudaStreamWaitEvent(s.stream, wf.join, 0));
    }
    PB2_CUDA(cudaEventRecord(wf.t1, s.stream));
    wf.stats_pending = true;
}

void collect_render_stats(Scene &s) {
    if (!s.wf || !s.wf->stats_pending) return;
    Wavefront &wf = *s.wf;
    PB2_CUDA(cudaStreamSynchronize(s.stream));
    pb2_render_stats rs{};
    unsigned long long totals[2];
    PB2_CUDA(cudaMemcpy(totals, wf.ray_totals.ptr, sizeof totals, cudaMemcpyDeviceToHost));
    rs.closest_rays = totals[0], rs.shadow_rays = totals[1];
    rs.kernel_launches = wf.launches;
    PB2_CUDA(cudaEventElapsedTime(&rs.total_ms, wf.t0, wf.t1));
    for (auto &e : wf.events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        (e.stage == 0 ? rs.generate_ms : e.stage == 1 ? rs.extend_ms : e.stage == 2 ? rs.shade_ms : e.stage == 3 ? rs.shadow_ms : rs.accumulate_ms) += ms;
    }
    if (s.counting) {
        unsigned long long h[8];
        PB2_CUDA(cudaMemcpy(h, wf.trav_counters.ptr, sizeof h, cudaMemcpyDeviceToHost));
        rs.shadow_unoccluded = h[4];
        rs.nodes_visited = h[0] + h[2], rs.prims_tested = h[1] + h[3];
        rs.nodes_shadow = h[2], rs.prims_shadow = h[3];
    }
    rs.batches = wf.batches, rs.rounds = wf.rounds_used;
    rs.extend_launches = wf.n_extend, rs.shade_launches = wf.n_shade, rs.shadow_launches = wf.n_shadow;
    rs.other_launches = (uint32_t)(wf.launches - wf.n_extend - wf.n_shade - wf.n_shadow);
    rs.shaded_paths = rs.closest_rays;
    rs.sorted = wf.sorted ? 1u : 0u;
    s.render_stats = rs;
    wf.stats_pending = false;
}
}// namespace pb2
