// Stand-alone trace kernels behind pb2_trace_closest / pb2_trace_any (parity hooks and the traversal
// benchmark of config C4).  The wavefront integrator (wavefront.cu) uses the same trace_persistent<>().
#include "scene.cuh"
#include "traverse.cuh"

namespace pb2 {
namespace {
struct ClosestIO {
    const float4 *__restrict__ rays;
    float4 *__restrict__ hit_tuvp;
    int32_t *__restrict__ hit_inst;
    const PrimRec *prims;
    uint32_t n;
    PB2_D uint32_t size() const { return n; }
    PB2_D uint32_t load(uint32_t i, float3 &o, float3 &d, float &tmin, float &tmax) const {
        const float4 ro = __ldg(rays + 2 * (size_t)i), rd = __ldg(rays + 2 * (size_t)i + 1);
        o = mk3(ro), d = mk3(rd), tmin = ro.w, tmax = rd.w;
        return i;
    }
    PB2_D void commit(bool valid, uint32_t i, const RayHit &h, bool hit, uint32_t inst_of_hit) const {
        if (!valid) return;
        int32_t inst = -1;
        uint32_t prim = 0xffffffffu;
        float t = 0.f;
        if (hit) {
            const float4 *rec = reinterpret_cast<const float4 *>(prims + h.prim_slot);
            prim = __float_as_uint(__ldg(rec).w);
            inst = inst_of_hit != 0xffffffffu ? (int32_t)inst_of_hit : (int32_t)__float_as_uint(__ldg(rec + 1).w);
            t = h.t;
        }
        hit_tuvp[i] = make_float4(t, h.u, h.v, __uint_as_float(prim));
        hit_inst[i] = inst;
    }
};
struct AnyIO {
    const float4 *__restrict__ rays;
    uint32_t *__restrict__ occluded;
    uint32_t n;
    PB2_D uint32_t size() const { return n; }
    PB2_D uint32_t load(uint32_t i, float3 &o, float3 &d, float &tmin, float &tmax) const {
        const float4 ro = __ldg(rays + 2 * (size_t)i), rd = __ldg(rays + 2 * (size_t)i + 1);
        o = mk3(ro), d = mk3(rd), tmin = ro.w, tmax = rd.w;
        return i;
    }
    PB2_D void commit(bool valid, uint32_t i, const RayHit &, bool hit, uint32_t) const {
        if (valid) occluded[i] = hit ? 1u : 0u;
    }
};

template<bool COUNT, bool COOP, bool TRIS = false, bool INST = false>
__global__ void __launch_bounds__(128, PB2_TRACE_MINB(COOP)) k_trace_closest(SceneView sv, ClosestIO io, uint32_t *__restrict__ work, unsigned long long *__restrict__ counters, int thr) {
    TraceCounters ctr{ 0, 0 };
    trace_persistent<false, COUNT, COOP, TRIS, INST>(sv, io, work, &ctr, thr);
    if (COUNT) {
        atomicAdd(&counters[0], (unsigned long long)ctr.nodes);
        atomicAdd(&counters[1], (unsigned long long)ctr.prims);
    }
}
template<bool COUNT, bool COOP, bool TRIS = false, bool INST = false>
__global__ void __launch_bounds__(128, PB2_TRACE_MINB(COOP)) k_trace_any(SceneView sv, AnyIO io, uint32_t *__restrict__ work, unsigned long long *__restrict__ counters, int thr) {
    TraceCounters ctr{ 0, 0 };
    trace_persistent<true, COUNT, COOP, TRIS, INST>(sv, io, work, &ctr, thr);
    if (COUNT) {
        atomicAdd(&counters[0], (unsigned long long)ctr.nodes);
        atomicAdd(&counters[1], (unsigned long long)ctr.prims);
    }
}
unsigned trace_grid(uint64_t n, int block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (n + block - 1) / block;
    const uint64_t cap = (uint64_t)sms * 16; // persistent-ish: 16 CTAs of 128 threads per SM = 2048 threads
    return (unsigned)std::max<uint64_t>(1, std::min(want, cap));
}
}// namespace

template<class K, class IO>
void launch_trace(Scene &s, K kernel_plain, K kernel_count, const IO &io, uint64_t n) {
    if (n >= 0xffffffffull) throw std::runtime_error("pb2_trace: more than 2^32-1 rays in one call");
    const SceneView sv = s.view();
    if (s.l2_dirty) s.apply_l2_window();
    s.trace_work.ensure(1);
    PB2_CUDA(cudaMemsetAsync(s.trace_work.ptr, 0, sizeof(uint32_t), s.stream));
    if (s.counting) {
        DevBuf<unsigned long long> ctr(2);
        ctr.zero(s.stream);
        kernel_count<<<trace_grid(n, 128), 128, 0, s.stream>>>(sv, io, s.trace_work.ptr, ctr.ptr, s.refill_threshold);
        PB2_LAUNCH_CHECK();
        unsigned long long h[2];
        PB2_CUDA(cudaMemcpyAsync(h, ctr.ptr, sizeof h, cudaMemcpyDeviceToHost, s.stream));
        PB2_CUDA(cudaStreamSynchronize(s.stream));
        s.render_stats.nodes_visited = h[0], s.render_stats.prims_tested = h[1];
    } else {
        kernel_plain<<<trace_grid(n, 128), 128, 0, s.stream>>>(sv, io, s.trace_work.ptr, nullptr, s.refill_threshold);
        PB2_LAUNCH_CHECK();
    }
}

void trace_closest_dev(Scene &s, const float4 *rays, uint64_t n, float4 *hit_tuvp, int32_t *hit_inst) {
    if (!s.bvh_valid) throw std::runtime_error("pb2_trace_closest: call pb2_bvh_build first");
    if (!n) return;
    ClosestIO io{ rays, hit_tuvp, hit_inst, s.d_prims.ptr, (uint32_t)n };
    const bool tris = s.build_stats.n_spheres == 0; // the plain kernels drop the sphere branch; the counting ones keep it (same counts)
    if (s.n_blas > 0) { // two-level scene: the kernels that follow instance nodes
        if (s.use_coop_prims()) launch_trace(s, k_trace_closest<false, true, false, true>, k_trace_closest<true, true, false, true>, io, n);
        else launch_trace(s, k_trace_closest<false, false, false, true>, k_trace_closest<true, false, false, true>, io, n);
    } else if (s.use_coop_prims()) launch_trace(s, tris ? k_trace_closest<false, true, true> : k_trace_closest<false, true>, k_trace_closest<true, true>, io, n);
    else launch_trace(s, tris ? k_trace_closest<false, false, true> : k_trace_closest<false, false>, k_trace_closest<true, false>, io, n);
}
void trace_any_dev(Scene &s, const float4 *rays, uint64_t n, uint32_t *occluded) {
    if (!s.bvh_valid) throw std::runtime_error("pb2_trace_any: call pb2_bvh_build first");
    if (!n) return;
    AnyIO io{ rays, occluded, (uint32_t)n };
    const bool tris = s.build_stats.n_spheres == 0;
    if (s.n_blas > 0) {
        if (s.use_coop_prims()) launch_trace(s, k_trace_any<false, true, false, true>, k_trace_any<true, true, false, true>, io, n);
        else launch_trace(s, k_trace_any<false, false, false, true>, k_trace_any<true, false, false, true>, io, n);
    } else if (s.use_coop_prims()) launch_trace(s, tris ? k_trace_any<false, true, true> : k_trace_any<false, true>, k_trace_any<true, true>, io, n);
    else launch_trace(s, tris ? k_trace_any<false, false, true> : k_trace_any<false, false>, k_trace_any<true, false>, io, n);
}
}// namespace pb2
