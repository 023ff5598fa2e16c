// Stand-alone trace kernels behind pb2_trace_closest / pb2_trace_any (parity hooks and the traversal
// benchmark of config C4).  The wavefront integrator (wavefront.cu) uses the same traverse<>().
#include "scene.cuh"
#include "traverse.cuh"

namespace pb2 {
namespace {
template<bool COUNT>
__global__ void __launch_bounds__(128) k_trace_closest(SceneView sv, const float4 *__restrict__ rays, uint64_t n, float4 *__restrict__ hit_tuvp,
                                                       int32_t *__restrict__ hit_inst, unsigned long long *__restrict__ counters) {
    TraceCounters ctr{ 0, 0 };
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
        RayHit h;
        h.t = rd.w, h.u = h.v = 0.f;
        traverse<false, COUNT>(sv, mk3(ro), mk3(rd), ro.w, h, &ctr);
        int32_t inst = -1;
        uint32_t prim = 0xffffffffu;
        if (h.prim_slot != 0xffffffffu) {
            const float4 *rec = reinterpret_cast<const float4 *>(sv.prims + h.prim_slot);
            prim = __float_as_uint(__ldg(rec).w);
            inst = (int32_t)__float_as_uint(__ldg(rec + 1).w);
        } else {
            h.t = 0.f;
        }
        hit_tuvp[i] = make_float4(h.t, h.u, h.v, __uint_as_float(prim));
        hit_inst[i] = inst;
    }
    if (COUNT) {
        atomicAdd(&counters[0], (unsigned long long)ctr.nodes);
        atomicAdd(&counters[1], (unsigned long long)ctr.prims);
    }
}
template<bool COUNT>
__global__ void __launch_bounds__(128) k_trace_any(SceneView sv, const float4 *__restrict__ rays, uint64_t n, uint32_t *__restrict__ occluded,
                                                   unsigned long long *__restrict__ counters) {
    TraceCounters ctr{ 0, 0 };
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 ro = __ldg(rays + 2 * i), rd = __ldg(rays + 2 * i + 1);
        RayHit h;
        h.t = rd.w, h.u = h.v = 0.f;
        occluded[i] = traverse<true, COUNT>(sv, mk3(ro), mk3(rd), ro.w, h, &ctr) ? 1u : 0u;
    }
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

void trace_closest_dev(Scene &s, const float4 *rays, uint64_t n, float4 *hit_tuvp, int32_t *hit_inst) {
    if (!s.bvh_valid) throw std::runtime_error("pb2_trace_closest: call pb2_bvh_build first");
    if (!n) return;
    const SceneView sv = s.view();
    if (s.counting) {
        DevBuf<unsigned long long> ctr(2);
        ctr.zero(s.stream);
        k_trace_closest<true><<<trace_grid(n, 128), 128, 0, s.stream>>>(sv, rays, n, hit_tuvp, hit_inst, ctr.ptr);
        PB2_LAUNCH_CHECK();
        unsigned long long h[2];
        PB2_CUDA(cudaMemcpyAsync(h, ctr.ptr, sizeof h, cudaMemcpyDeviceToHost, s.stream));
        PB2_CUDA(cudaStreamSynchronize(s.stream));
        s.render_stats.nodes_visited = h[0], s.render_stats.prims_tested = h[1];
    } else {
        k_trace_closest<false><<<trace_grid(n, 128), 128, 0, s.stream>>>(sv, rays, n, hit_tuvp, hit_inst, nullptr);
        PB2_LAUNCH_CHECK();
    }
}
void trace_any_dev(Scene &s, const float4 *rays, uint64_t n, uint32_t *occluded) {
    if (!s.bvh_valid) throw std::runtime_error("pb2_trace_any: call pb2_bvh_build first");
    if (!n) return;
    const SceneView sv = s.view();
    if (s.counting) {
        DevBuf<unsigned long long> ctr(2);
        ctr.zero(s.stream);
        k_trace_any<true><<<trace_grid(n, 128), 128, 0, s.stream>>>(sv, rays, n, occluded, ctr.ptr);
        PB2_LAUNCH_CHECK();
        unsigned long long h[2];
        PB2_CUDA(cudaMemcpyAsync(h, ctr.ptr, sizeof h, cudaMemcpyDeviceToHost, s.stream));
        PB2_CUDA(cudaStreamSynchronize(s.stream));
        s.render_stats.nodes_visited = h[0], s.render_stats.prims_tested = h[1];
    } else {
        k_trace_any<false><<<trace_grid(n, 128), 128, 0, s.stream>>>(sv, rays, n, occluded, nullptr);
        PB2_LAUNCH_CHECK();
    }
}
}// namespace pb2
