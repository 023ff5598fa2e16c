// Binned-SAH binary BVH builder over the Morton-ordered primitive sequence (builder 1 of pb2_scene_set_builder).
//
// The primitives arrive sorted along the 63-bit Morton curve (bvh_build.cu, stage 2).  Every node of the binary tree
// owns a contiguous range of that sequence; instead of LBVH's "highest differing Morton bit" the split position is
// the one that minimises the surface-area heuristic  cost = A(L)·|L| + A(R)·|R| :
//   * large nodes (more than kSmall primitives): the range is cut into kBins equal-count bins, bin boxes are
//     accumulated in one pass over the primitives (warp-aggregated atomic min/max), and a prefix/suffix sweep over
//     the kBins - 1 bin boundaries picks the split — one level of all large nodes per pass, breadth first;
//   * small nodes: one thread sweeps every split position exactly and finishes the whole subtree.
// Keeping ranges contiguous means no primitive is ever moved after the radix sort, so a level costs one streaming
// pass; the price is that splits are restricted to Z-order-contiguous sets (the collapse to 8-wide nodes then picks
// which binary nodes survive).  Node boxes come out of the sweep, so no refit pass is needed.
// Produces the same BinTree arrays as the LBVH path: left/right (>= 0 internal, < 0 = ~sorted position), range, lo, hi.
#include "scene.cuh"
#include <cfloat>
#include <vector>

namespace pb2 {
namespace {
constexpr int kBins = 16;
constexpr int kSmall = 16;
constexpr uint32_t kDone = 0xffffffffu;

__device__ __forceinline__ int f2o(float f) { // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ float half_area(float3 lo, float3 hi) {
    const float3 d = hi - lo;
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

struct SahCtx {
    uint32_t n;
    const float4 *slo, *shi; // primitive boxes in sorted order
    int *left, *right;
    int2 *range;
    float4 *nlo, *nhi;
    int *split;          // 0: not split yet; > 0: first index of the right child; -1: subtree finished (small node)
    int *slot_of;        // node -> bin slot of the current level, -1 when the node is not being binned
    uint32_t *node_of;   // primitive -> node it currently belongs to (kDone once it sits in a finished subtree)
    uint32_t *counters;  // [0] next free node id, [1] size of the next level's active list
};

__global__ void k_sah_gather(const uint32_t *__restrict__ sorted, const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi, uint32_t n,
                             float4 *__restrict__ slo, float4 *__restrict__ shi, uint32_t *__restrict__ node_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = sorted[i];
    slo[i] = box_lo[p], shi[i] = box_hi[p];
    node_of[i] = 0u;
}

__global__ void k_sah_fill(int *__restrict__ a, int value, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = value;
}

// One pass over the primitives: descend one level where the node was split in the previous pass, then add the box to the
// bin of the (large, active) node the primitive now belongs to.  bins: [slot][kBins][6] ordered ints (lo.xyz, hi.xyz).
__global__ void __launch_bounds__(256) k_sah_bin(SahCtx c, int *__restrict__ bins) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t key = 0xffffffffu;
    float3 lo = mk3(0.f), hi = mk3(0.f);
    if (i < c.n) {
        uint32_t nd = c.node_of[i];
        if (nd != kDone) {
            const int sp = c.split[nd];
            if (sp != 0) {
                const int child = sp < 0 ? -1 : ((int)i < sp ? c.left[nd] : c.right[nd]);
                nd = child < 0 ? kDone : (uint32_t)child;
                if (nd != kDone && c.split[nd] < 0) nd = kDone; // the child is a finished small subtree
                c.node_of[i] = nd;
            }
            if (nd != kDone) {
                const int slot = c.slot_of[nd];
                if (slot >= 0) {
                    const int2 rg = c.range[nd];
                    const uint32_t bin = (uint32_t)(((uint64_t)(i - (uint32_t)rg.x) * kBins) / (uint32_t)rg.y);
                    key = (uint32_t)slot * kBins + bin;
                    lo = mk3(c.slo[i]), hi = mk3(c.shi[i]);
                }
            }
        }
    }
    // consecutive primitives share (node, bin): reduce inside the warp, one set of atomics per distinct key
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (key == 0xffffffffu) return;
    const int v[6] = { f2o(lo.x), f2o(lo.y), f2o(lo.z), f2o(hi.x), f2o(hi.y), f2o(hi.z) };
    int r[6];
#pragma unroll
    for (int k = 0; k < 3; ++k) r[k] = __reduce_min_sync(peers, v[k]), r[3 + k] = __reduce_max_sync(peers, v[3 + k]);
    if ((threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) {
        int *b = bins + (size_t)key * 6;
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicMin(b + k, r[k]), atomicMax(b + 3 + k, r[3 + k]);
    }
}

// Finishes the subtree of a small node in one thread: exact SAH sweep over every split position, recursively.
// `first_free` is the first of the (cnt - 2) node ids reserved for the subtree below `root`.
__device__ void sah_build_small(const SahCtx &c, int root, int first, int cnt, int first_free) {
    int stack_id[kSmall], stack_first[kSmall], stack_cnt[kSmall];
    int sp = 0;
    stack_id[0] = root, stack_first[0] = first, stack_cnt[0] = cnt, sp = 1;
    while (sp) {
        --sp;
        const int id = stack_id[sp], f = stack_first[sp], m = stack_cnt[sp];
        float3 lo[kSmall], hi[kSmall];
        for (int k = 0; k < m; ++k) lo[k] = mk3(c.slo[f + k]), hi[k] = mk3(c.shi[f + k]);
        float right_area[kSmall];
        float3 rl = lo[m - 1], rh = hi[m - 1];
        for (int k = m - 1; k >= 1; --k) {
            rl = fmin3(rl, lo[k]), rh = fmax3(rh, hi[k]);
            right_area[k] = half_area(rl, rh);
        }
        float3 ll = lo[0], lh = hi[0];
        float best = FLT_MAX;
        int s = 1;
        for (int k = 1; k < m; ++k) { // split before element k: left = [0,k), right = [k,m)
            const float cost = half_area(ll, lh) * k + right_area[k] * (m - k);
            if (cost < best) best = cost, s = k;
            ll = fmin3(ll, lo[k]), lh = fmax3(lh, hi[k]);
        }
        c.nlo[id] = make_float4(ll.x, ll.y, ll.z, 0.f), c.nhi[id] = make_float4(lh.x, lh.y, lh.z, 0.f); // ll/lh now cover [0,m)
        c.range[id] = make_int2(f, m);
        c.split[id] = -1;
        int child[2];
        const int cf[2] = { f, f + s }, cc[2] = { s, m - s };
        for (int k = 0; k < 2; ++k) {
            if (cc[k] == 1) {
                child[k] = ~cf[k];
            } else {
                child[k] = first_free++;
                stack_id[sp] = child[k], stack_first[sp] = cf[k], stack_cnt[sp] = cc[k], ++sp;
            }
        }
        c.left[id] = child[0], c.right[id] = child[1];
    }
}

// One thread per active large node: sweep the bin boundaries, split, create the children.
__global__ void __launch_bounds__(128) k_sah_split(SahCtx c, const int *__restrict__ active, uint32_t n_active, const int *__restrict__ bins,
                                                    int *__restrict__ next_active, int *__restrict__ next_bins) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_active) return;
    const int id = active[a];
    const int2 rg = c.range[id];
    const int first = rg.x, cnt = rg.y;
    const int *b = bins + (size_t)a * kBins * 6;
    float3 lo[kBins], hi[kBins];
#pragma unroll
    for (int k = 0; k < kBins; ++k) lo[k] = mk3(o2f(b[k * 6]), o2f(b[k * 6 + 1]), o2f(b[k * 6 + 2])), hi[k] = mk3(o2f(b[k * 6 + 3]), o2f(b[k * 6 + 4]), o2f(b[k * 6 + 5]));
    auto bound = [&](int k) { return first + (int)(((int64_t)cnt * k) / kBins); }; // first index of bin k
    float right_area[kBins];
    float3 rl = lo[kBins - 1], rh = hi[kBins - 1];
    for (int k = kBins - 1; k >= 1; --k) {
        rl = fmin3(rl, lo[k]), rh = fmax3(rh, hi[k]);
        right_area[k] = half_area(rl, rh);
    }
    float3 ll = lo[0], lh = hi[0];
    float best = FLT_MAX;
    int s = kBins / 2;
    for (int k = 1; k < kBins; ++k) { // split before bin k
        const int nl = bound(k) - first;
        const float cost = half_area(ll, lh) * nl + right_area[k] * (cnt - nl);
        if (cost < best) best = cost, s = k;
        ll = fmin3(ll, lo[k]), lh = fmax3(lh, hi[k]);
    }
    c.nlo[id] = make_float4(ll.x, ll.y, ll.z, 0.f), c.nhi[id] = make_float4(lh.x, lh.y, lh.z, 0.f);
    const int mid = bound(s);
    c.split[id] = mid;
    c.slot_of[id] = -1;
    int child[2];
    const int cf[2] = { first, mid }, cc[2] = { mid - first, first + cnt - mid };
    for (int k = 0; k < 2; ++k) {
        if (cc[k] == 1) {
            child[k] = ~cf[k];
            continue;
        }
        if (cc[k] > kSmall) {
            const int nid = (int)atomicAdd(&c.counters[0], 1u);
            child[k] = nid;
            c.range[nid] = make_int2(cf[k], cc[k]);
            c.split[nid] = 0;
            const uint32_t slot = atomicAdd(&c.counters[1], 1u);
            next_active[slot] = nid;
            c.slot_of[nid] = (int)slot;
            int *nb = next_bins + (size_t)slot * kBins * 6;
            for (int q = 0; q < kBins; ++q) {
                nb[q * 6] = nb[q * 6 + 1] = nb[q * 6 + 2] = 0x7f7fffff;                  // f2o(+FLT_MAX)
                nb[q * 6 + 3] = nb[q * 6 + 4] = nb[q * 6 + 5] = (int)0xff7fffff ^ 0x7fffffff; // f2o(-FLT_MAX)
            }
        } else {
            const int nid = (int)atomicAdd(&c.counters[0], (uint32_t)(cc[k] - 1)); // the node itself + its cc-2 descendants
            child[k] = nid;
            sah_build_small(c, nid, cf[k], cc[k], nid + 1);
        }
    }
    c.left[id] = child[0], c.right[id] = child[1];
}

__global__ void k_sah_root_small(SahCtx c) { sah_build_small(c, 0, 0, (int)c.n, 1); }
}// namespace

bool sah_builder_available() { return true; }

// left/right/range/lo/hi: BinTree arrays of n - 1 internal nodes (root = 0).  `sorted` is the Morton order (input).
void build_binary_sah(cudaStream_t st, uint32_t n, const float4 *box_lo, const float4 *box_hi, const uint32_t *sorted, int *left, int *right,
                      int2 *range, float4 *lo, float4 *hi) {
    if (n < 2) return;
    DevBuf<float4> slo(n), shi(n);
    DevBuf<uint32_t> node_of(n), counters(2);
    DevBuf<int> split(n), slot_of(n);
    const size_t max_large = (size_t)n / kSmall + 2;
    DevBuf<int> active[2] = { DevBuf<int>(max_large), DevBuf<int>(max_large) };
    DevBuf<int> bins[2] = { DevBuf<int>(max_large * kBins * 6), DevBuf<int>(max_large * kBins * 6) };
    k_sah_gather<<<div_up(n, 256), 256, 0, st>>>(sorted, box_lo, box_hi, n, slo.ptr, shi.ptr, node_of.ptr);
    PB2_LAUNCH_CHECK();
    k_sah_fill<<<148 * 8, 256, 0, st>>>(split.ptr, 0, n);
    k_sah_fill<<<148 * 8, 256, 0, st>>>(slot_of.ptr, -1, n);
    PB2_LAUNCH_CHECK();
    SahCtx c{ n, slo.ptr, shi.ptr, left, right, range, lo, hi, split.ptr, slot_of.ptr, node_of.ptr, counters.ptr };
    if (n <= (uint32_t)kSmall) {
        k_sah_root_small<<<1, 1, 0, st>>>(c);
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaStreamSynchronize(st));
        return;
    }
    // level 0: the root is the only active node
    {
        const int root = 0, zero_slot = 0;
        const int2 rg = make_int2(0, (int)n);
        uint32_t init[2] = { 1u, 0u };
        std::vector<int> b0(kBins * 6);
        for (int q = 0; q < kBins; ++q)
            for (int k = 0; k < 6; ++k) b0[q * 6 + k] = k < 3 ? 0x7f7fffff : (int)(0xff7fffffu ^ 0x7fffffffu);
        PB2_CUDA(cudaMemcpyAsync(active[0].ptr, &root, sizeof root, cudaMemcpyHostToDevice, st));
        PB2_CUDA(cudaMemcpyAsync(range, &rg, sizeof rg, cudaMemcpyHostToDevice, st));
        PB2_CUDA(cudaMemcpyAsync(slot_of.ptr, &zero_slot, sizeof zero_slot, cudaMemcpyHostToDevice, st));
        PB2_CUDA(cudaMemcpyAsync(counters.ptr, init, sizeof init, cudaMemcpyHostToDevice, st));
        PB2_CUDA(cudaMemcpyAsync(bins[0].ptr, b0.data(), b0.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        PB2_CUDA(cudaStreamSynchronize(st));
    }
    uint32_t n_active = 1;
    int cur = 0, level = 0;
    while (n_active) {
        if (++level > 4096) throw std::runtime_error("binned SAH builder: runaway depth");
        k_sah_bin<<<div_up(n, 256), 256, 0, st>>>(c, bins[cur].ptr);
        PB2_LAUNCH_CHECK();
        k_sah_split<<<div_up(n_active, 128), 128, 0, st>>>(c, active[cur].ptr, n_active, bins[cur].ptr, active[cur ^ 1].ptr, bins[cur ^ 1].ptr);
        PB2_LAUNCH_CHECK();
        uint32_t h[2];
        PB2_CUDA(cudaMemcpyAsync(h, counters.ptr, sizeof h, cudaMemcpyDeviceToHost, st));
        PB2_CUDA(cudaStreamSynchronize(st));
        n_active = h[1];
        PB2_CUDA(cudaMemsetAsync(counters.ptr + 1, 0, sizeof(uint32_t), st));
        cur ^= 1;
    }
    PB2_CUDA(cudaStreamSynchronize(st));
}
}// namespace pb2
