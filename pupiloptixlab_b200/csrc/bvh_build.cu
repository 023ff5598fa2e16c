// GPU build of the world-space compressed 8-wide BVH.
//
// Replaces the closed-source OptiX acceleration-structure build of the reference
// (GAS::Create framework/world/gas_manager.cpp:69-245, IAS::Create framework/world/ias_manager.cpp:29-114):
// every instance's primitives are transformed to world space, one binary BVH is built over all of
// them (LBVH over 63-bit Morton codes, or binned SAH along the Morton order — bvh_sah.cu) and collapsed top-down into
// 80-byte BVH8 nodes with quantised child boxes; primitive records are rewritten in leaf order.
//
// Stages (all on the scene's stream):
//   1 emit_prims      instance triangles / spheres -> 48 B world-space records + AABBs + scene bounds
//   2 morton + sort   63-bit keys, in-tree onesweep radix sort (radix_sort.cu)
//   3 radix_tree      Karras 2012 binary radix tree (ties broken by index)
//   4 refit           bottom-up AABBs and, for the cost-optimal collapse, the SAH cost tables + decisions of every binary node
//   5 collapse        breadth-first: binary subtree -> up to 8 children along the cheapest cut (or by largest-area expansion),
//                     octant slot assignment, quantisation, leaf primitive copy
#include "scene.cuh"
#include "traverse.cuh"
#include <cfloat>
#include <cuda/atomic>

namespace pb2 {
namespace {

#ifndef PB2_LEAF_MAX
#define PB2_LEAF_MAX 3
#endif
constexpr int kLeafMax = PB2_LEAF_MAX; // primitives per BVH8 leaf slot (unary count in 3 meta bits)

// ---- order-preserving float <-> int for atomicMin / atomicMax ------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ordered_to_float(int i) {
    int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

struct Aabb {
    float3 lo, hi;
};
__device__ __forceinline__ float half_area(float3 lo, float3 hi) {
    float3 d = hi - lo;
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

// ---- stage 1 ---------------------------------------------------------------------------------------------
// One thread per primitive.  inst_first[i] = index of the first primitive of the i-th LISTED instance (n_inst+1 entries);
// inst_ids[i] = which instance of the scene that is (nullptr: the i-th).  A bottom-level build lists one instance with the
// identity transform (object-space records), the top-level build lists the instances that are flattened into it.
__global__ void k_emit_prims(const DevInstance *__restrict__ inst, const uint32_t *__restrict__ inst_first, const uint32_t *__restrict__ inst_ids, uint32_t n_inst,
                             uint32_t n_prims, PrimRec *__restrict__ prims, float4 *__restrict__ box_lo, float4 *__restrict__ box_hi, int *__restrict__ scene_bounds) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = mk3(FLT_MAX), hi = mk3(-FLT_MAX);
    if (g < n_prims) {
        uint32_t a = 0, b = n_inst; // last instance with inst_first <= g
        while (b - a > 1) {
            uint32_t m = (a + b) >> 1;
            if (inst_first[m] <= g) a = m;
            else b = m;
        }
        const uint32_t prim = g - inst_first[a];
        if (inst_ids) a = inst_ids[a];
        const DevInstance &in = inst[a];
        PrimRec r;
        if (in.flags & PB2_IF_SPHERE) {
            // exact bounds of the affinely transformed unit sphere: centre +- row norms
            const float4 r0 = in.xf[0], r1 = in.xf[1], r2 = in.xf[2];
            float3 c = mk3(r0.w, r1.w, r2.w);
            float3 e = mk3(sqrtf(r0.x * r0.x + r0.y * r0.y + r0.z * r0.z), sqrtf(r1.x * r1.x + r1.y * r1.y + r1.z * r1.z),
                           sqrtf(r2.x * r2.x + r2.y * r2.y + r2.z * r2.z));
            lo = c - e, hi = c + e;
            r.v0 = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
            r.e1 = make_float4(0.f, 0.f, 0.f, __uint_as_float(a));
            r.e2 = make_float4(0.f, 0.f, 0.f, __uint_as_float(1u));
        } else {
            const uint32_t i0 = in.idx[prim * 3], i1 = in.idx[prim * 3 + 1], i2 = in.idx[prim * 3 + 2];
            // ix_*: the fixed-rounding sequence of traverse.cuh, so the world-space record is reproducible bit for bit
            float3 p0 = ix_point(in.xf[0], in.xf[1], in.xf[2], mk3(in.pos[i0 * 3], in.pos[i0 * 3 + 1], in.pos[i0 * 3 + 2]));
            float3 p1 = ix_point(in.xf[0], in.xf[1], in.xf[2], mk3(in.pos[i1 * 3], in.pos[i1 * 3 + 1], in.pos[i1 * 3 + 2]));
            float3 p2 = ix_point(in.xf[0], in.xf[1], in.xf[2], mk3(in.pos[i2 * 3], in.pos[i2 * 3 + 1], in.pos[i2 * 3 + 2]));
            lo = fmin3(p0, fmin3(p1, p2)), hi = fmax3(p0, fmax3(p1, p2));
            float3 e1 = ix_sub(p1, p0), e2 = ix_sub(p2, p0);
            // the intersector reconstructs p1 = v0 + e1 in fp32; widen the box by that rounding
            lo = fmin3(lo, fmin3(p0 + e1, p0 + e2)), hi = fmax3(hi, fmax3(p0 + e1, p0 + e2));
            r.v0 = make_float4(p0.x, p0.y, p0.z, __uint_as_float(prim));
            r.e1 = make_float4(e1.x, e1.y, e1.z, __uint_as_float(a));
            r.e2 = make_float4(e2.x, e2.y, e2.z, __uint_as_float(0u));
        }
        prims[g] = r;
        box_lo[g] = make_float4(lo.x, lo.y, lo.z, 0.f);
        box_hi[g] = make_float4(hi.x, hi.y, hi.z, 0.f);
    }
    // block reduction of the scene bounds, then 6 atomics per block
    __shared__ float s_lo[3][32], s_hi[3][32];
    float v[6] = { lo.x, lo.y, lo.z, hi.x, hi.y, hi.z };
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
            v[3 + k] = fmaxf(v[3 + k], __shfl_xor_sync(0xffffffffu, v[3 + k], o));
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 3; ++k) s_lo[k][warp] = v[k], s_hi[k][warp] = v[3 + k];
    __syncthreads();
    if (warp == 0) {
        for (int k = 0; k < 3; ++k) {
            float a = lane < nw ? s_lo[k][lane] : FLT_MAX, b = lane < nw ? s_hi[k][lane] : -FLT_MAX;
            for (int o = 16; o; o >>= 1) a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)), b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
            if (lane == 0) {
                atomicMin(&scene_bounds[k], float_to_ordered(a));
                atomicMax(&scene_bounds[3 + k], float_to_ordered(b));
            }
        }
    }
}

// Instance leaves of a top-level build: one record per placement of a mesh that has its own bottom-level tree.
//   record: v0.w = bits(root node of that tree in the scene's node array), e1.w = bits(instance id), e2.w = bits(2)
// The collapse turns each of them into an instance node (traverse.cuh); the record itself is never intersected.
__global__ void k_emit_instance_leaves(uint32_t n, uint32_t first, const float4 *__restrict__ lo, const float4 *__restrict__ hi, const uint32_t *__restrict__ root_node,
                                       const uint32_t *__restrict__ inst_id, PrimRec *__restrict__ prims, float4 *__restrict__ box_lo, float4 *__restrict__ box_hi,
                                       int *__restrict__ scene_bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PrimRec r;
    r.v0 = make_float4(0.f, 0.f, 0.f, __uint_as_float(root_node[i]));
    r.e1 = make_float4(0.f, 0.f, 0.f, __uint_as_float(inst_id[i]));
    r.e2 = make_float4(0.f, 0.f, 0.f, __uint_as_float(2u));
    prims[first + i] = r;
    const float4 l = lo[i], h = hi[i];
    box_lo[first + i] = l, box_hi[first + i] = h;
    atomicMin(&scene_bounds[0], float_to_ordered(l.x)), atomicMin(&scene_bounds[1], float_to_ordered(l.y)), atomicMin(&scene_bounds[2], float_to_ordered(l.z));
    atomicMax(&scene_bounds[3], float_to_ordered(h.x)), atomicMax(&scene_bounds[4], float_to_ordered(h.y)), atomicMax(&scene_bounds[5], float_to_ordered(h.z));
}

// ---- stage 2 -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread21(uint32_t v) { // 21 bits -> every third bit
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__global__ void k_morton(const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi, const int *__restrict__ scene_bounds, uint32_t n,
                         uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, int low_bit) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const float3 slo = mk3(ordered_to_float(scene_bounds[0]), ordered_to_float(scene_bounds[1]), ordered_to_float(scene_bounds[2]));
    const float3 shi = mk3(ordered_to_float(scene_bounds[3]), ordered_to_float(scene_bounds[4]), ordered_to_float(scene_bounds[5]));
    const float3 ext = fmax3(shi - slo, mk3(1e-30f));
    const float3 c = (mk3(box_lo[g]) + mk3(box_hi[g])) * 0.5f;
    const float3 u = (c - slo) / ext;
    const float s = 2097152.f; // 2^21
    uint32_t x = min(2097151u, (uint32_t)fmaxf(0.f, u.x * s)), y = min(2097151u, (uint32_t)fmaxf(0.f, u.y * s)),
             z = min(2097151u, (uint32_t)fmaxf(0.f, u.z * s));
    // bits below low_bit are not sorted (Scene::morton_bits) and must not take part in the radix tree either: equal keys fall back to the index
    keys[g] = (spread21(x) << 2 | spread21(y) << 1 | spread21(z)) & ~((1ull << low_bit) - 1ull);
    vals[g] = g;
}

// ---- stage 3: Karras 2012 -----------------------------------------------------------------------------
// Binary nodes 0..n-2; a child reference >= 0 is an internal node, < 0 is the leaf holding the sorted
// primitive ~ref.  range = (first, count) over sorted primitive positions.
struct BinTree {
    int *left, *right, *parent; // parent: [0, n-1) internal, [n-1, 2n-1) leaves
    int2 *range;
    float4 *lo, *hi; // internal node boxes
    uint32_t n;      // number of primitives
};
__device__ __forceinline__ int delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll(a ^ b);
}
__global__ void k_radix_tree(const uint64_t *__restrict__ keys, BinTree t) {
    const int n = t.n, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int s = lmax >> 1; s >= 1; s >>= 1)
        if (delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int div = 2, tt = (l + div - 1) / div;; div <<= 1, tt = (l + div - 1) / div) {
        if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
        if (tt <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int lc = (first == gamma) ? ~gamma : gamma;
    const int rc = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    t.left[i] = lc, t.right[i] = rc;
    t.range[i] = make_int2(first, last - first + 1);
    t.parent[lc >= 0 ? lc : (n - 1) + ~lc] = i;
    t.parent[rc >= 0 ? rc : (n - 1) + ~rc] = i;
    if (i == 0) t.parent[0] = -1;
}

// ---- stage 4 ---------------------------------------------------------------------------------------------
// weight_src != nullptr (top-level build with instance leaves): range.y becomes a WEIGHTED primitive count in which an instance
// leaf counts kLeafMax + 1, so no subtree that holds one is ever folded into a leaf slot and the collapse reaches it on its own.
__device__ __forceinline__ int leaf_weight(const PrimRec *prims, uint32_t p) { return __float_as_uint(prims[p].e2.w) == 2u ? kLeafMax + 1 : 1; }
// COST: the same bottom-up sweep also fills the tables of the cost-optimal collapse (Ylitie, Karras, Laine 2017, "Efficient
// incoherent ray traversal on GPUs through compressed wide BVHs", section 4.1).  For a binary node n and a budget of i slots,
//   C(n, 1) = A(n) count(n) c_prim                       when the subtree fits a leaf slot (count <= kLeafMax)
//           = A(n) c_node + min_k C(l, k) + C(r, 8 - k)   otherwise: n becomes a wide node of its own
//   C(n, i) = min(C(n, i - 1), min_k C(l, k) + C(r, i - k)),   i = 2 .. 7
// with A the half surface area and C(leaf, i) = A c_prim.  A node keeps C(n, 1..7) (32 bytes) for its parent and one word of
// decisions for the top-down pass of k_collapse: three bits per budget i = 2 .. 8 holding the k of the best split, 0 = "the
// answer for i - 1 is at least as good".  The greedy largest-area expansion this replaces (collapse = 0) spends a wide node on
// every subtree of 4 .. 8 primitives it meets; on the 30 M-triangle terrain that gave 6.8 primitives per 80-byte node.
struct CostTab {
    float4 *c;      // [2 * node], [2 * node + 1]: C(node, 1..4), C(node, 5..7)
    uint32_t *word; // decisions
    float c_prim;   // cost of one primitive test in units of one wide-node test
};
__device__ __forceinline__ void load_cost(const CostTab &ct, int ref, float3 lo, float3 hi, int weight, float C[7]) {
    if (ref >= 0) {
        const float4 a = __ldcg(&ct.c[2 * (size_t)ref]), b = __ldcg(&ct.c[2 * (size_t)ref + 1]);
        C[0] = a.x, C[1] = a.y, C[2] = a.z, C[3] = a.w, C[4] = b.x, C[5] = b.y, C[6] = b.z;
    } else {
        const float v = half_area(lo, hi) * (weight > kLeafMax ? 1.f : ct.c_prim); // an instance leaf costs a node step whatever the budget
#pragma unroll
        for (int i = 0; i < 7; ++i) C[i] = v;
    }
}
// One thread per INTERNAL node.  The threads of the nodes whose two children are leaves start climbing; a node with one leaf and
// one internal child is finished by whoever finishes that child, without synchronisation; only a node with two internal
// children needs the arrival counter — an acquire-release increment, the second thread to arrive goes on (it sees the first one's
// stores).  Half the counters of a sweep that starts at the leaves never get touched, and a thread re-reads mostly what it wrote
// itself.  (Measured on the 30 M-triangle terrain, boxes + cost tables: 5.1 ms starting at the leaves.)
// Small CTAs: most threads leave at once and a few climb for many levels; a CTA's slot is held until its last thread is done, so
// the fewer threads share a slot with a long climber the better.
#ifndef PB2_REFIT_BLOCK
#define PB2_REFIT_BLOCK 64
#endif
constexpr int kRefitBlock = PB2_REFIT_BLOCK;
// the internal nodes with two leaf children, packed (warp-aggregated append; their order does not matter): k_refit starts one
// thread at each, so its warps begin with 32 working lanes instead of the ~8 a launch over all internal nodes leaves
// (ncu, 30 M triangles: 3.6 of 32 lanes active on average, 2.5 G warp instructions)
__global__ void __launch_bounds__(1024) k_refit_seeds(BinTree t, int *__restrict__ seeds, uint32_t *__restrict__ n_seeds) {
    __shared__ uint32_t s_warp[32], s_base; // one atomic per CTA: a million warps adding to one word serialise (0.65 ms for 30 M nodes)
    const int n = t.n, i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool seed = i < n - 1 && t.left[i] < 0 && t.right[i] < 0;
    const uint32_t m = __ballot_sync(0xffffffffu, seed), lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        const uint32_t c = s_warp[lane];
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += up;
        }
        s_warp[lane] = incl - c;
        if (lane == 31) s_base = incl ? atomicAdd(n_seeds, incl) : 0u;
    }
    __syncthreads();
    if (seed) seeds[s_base + s_warp[warp] + __popc(m & ((1u << lane) - 1u))] = i;
}
template<bool COST>
__global__ void __launch_bounds__(kRefitBlock) k_refit(BinTree t, const uint32_t *__restrict__ sorted, const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi,
                        int *__restrict__ arrive, const PrimRec *__restrict__ weight_src, CostTab ct, const int *__restrict__ seeds, const uint32_t *__restrict__ n_seeds) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= *n_seeds) return;
    int node = seeds[tid]; // both children are leaves; nodes with an internal child are reached by the thread that finishes it (or the later of two)
    for (;;) {
        const int lc = t.left[node], rc = t.right[node];
        float3 llo, lhi, rlo, rhi;
        if (lc >= 0) llo = mk3(__ldcg(&t.lo[lc])), lhi = mk3(__ldcg(&t.hi[lc]));
        else llo = mk3(box_lo[sorted[~lc]]), lhi = mk3(box_hi[sorted[~lc]]);
        if (rc >= 0) rlo = mk3(__ldcg(&t.lo[rc])), rhi = mk3(__ldcg(&t.hi[rc]));
        else rlo = mk3(box_lo[sorted[~rc]]), rhi = mk3(box_hi[sorted[~rc]]);
        const float3 lo = fmin3(llo, rlo), hi = fmax3(lhi, rhi);
        t.lo[node] = make_float4(lo.x, lo.y, lo.z, 0.f);
        t.hi[node] = make_float4(hi.x, hi.y, hi.z, 0.f);
        if (weight_src || COST) {
            const int wl = lc >= 0 ? __ldcg(&t.range[lc]).y : weight_src ? leaf_weight(weight_src, sorted[~lc]) : 1;
            const int wr = rc >= 0 ? __ldcg(&t.range[rc]).y : weight_src ? leaf_weight(weight_src, sorted[~rc]) : 1;
            if (weight_src) t.range[node].y = wl + wr;
            if (COST) {
                float Cl[7], Cr[7], C[7];
                load_cost(ct, lc, llo, lhi, wl, Cl);
                load_cost(ct, rc, rlo, rhi, wr, Cr);
                float dist[9];
                uint32_t kb[9];
#pragma unroll
                for (int i = 2; i <= 8; ++i) {
                    dist[i] = FLT_MAX, kb[i] = 1u;
#pragma unroll
                    for (int k = 1; k <= 7; ++k) {
                        if (k >= i || i - k > 7) continue;
                        const float v = Cl[k - 1] + Cr[i - k - 1];
                        if (v < dist[i]) dist[i] = v, kb[i] = (uint32_t)k;
                    }
                }
                const float A = half_area(lo, hi);
                const int cnt = wl + wr;
                C[0] = cnt <= kLeafMax ? A * (float)cnt * ct.c_prim : A + dist[8];
                uint32_t word = kb[8] << 18;
#pragma unroll
                for (int i = 2; i <= 7; ++i) {
                    if (dist[i] < C[i - 2]) C[i - 1] = dist[i], word |= kb[i] << (3 * (i - 2));
                    else C[i - 1] = C[i - 2];
                }
                ct.c[2 * (size_t)node] = make_float4(C[0], C[1], C[2], C[3]);
                ct.c[2 * (size_t)node + 1] = make_float4(C[4], C[5], C[6], 0.f);
                ct.word[node] = word;
            }
        }
        const int up = t.parent[node];
        if (up < 0) return;
        if (t.left[up] >= 0 && t.right[up] >= 0 &&
            cuda::atomic_ref<int, cuda::thread_scope_device>(arrive[up]).fetch_add(1, cuda::memory_order_acq_rel) == 0)
            return; // two internal children and this is the first of them to finish
        node = up;
    }
}

// parent links of a binary tree that came without them (the clustering and the sweep builders): what k_refit<true> climbs
__global__ void __launch_bounds__(256) k_parents(BinTree t, int root) {
    const int n = t.n, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int lc = t.left[i], rc = t.right[i];
    t.parent[lc >= 0 ? lc : (n - 1) + ~lc] = i;
    t.parent[rc >= 0 ? rc : (n - 1) + ~rc] = i;
    if (i == root) t.parent[i] = -1;
}

// ---- stage 5: collapse -----------------------------------------------------------------------------------
struct CollapseCtx {
    BinTree t;
    const uint32_t *sorted;
    const float4 *box_lo, *box_hi;
    const PrimRec *prims_in;
    uint32_t *dst_of_sorted; // leaf-order slot of the primitive at each sorted position
    Bvh8Node *nodes;
    uint32_t *counters; // [0] nodes allocated, [1] prims allocated, [2] next-level queue size, [3] max depth
    float *sah;         // [0] accumulated SAH cost (un-normalised)
    bool inst_leaves;   // top-level build: primitive records of kind 2 become instance nodes
    const uint32_t *word; // decisions of the cost-optimal collapse (k_refit<true>), nullptr: greedy largest-area expansion
};
struct Ref {
    float3 lo, hi;
    int ref;   // >= 0 internal binary node, < 0 leaf (~sorted position)
    int count; // primitives below (range.y of an internal node: the only field of `range` the collapse reads — the clustering
               // builder's subtrees are not contiguous ranges of the sorted order)
};
__device__ __forceinline__ Ref load_ref(const CollapseCtx &c, int ref) {
    Ref r;
    r.ref = ref;
    if (ref >= 0) {
        r.lo = mk3(c.t.lo[ref]), r.hi = mk3(c.t.hi[ref]);
        r.count = c.t.range[ref].y;
    } else {
        const uint32_t p = c.sorted[~ref];
        r.lo = mk3(c.box_lo[p]), r.hi = mk3(c.box_hi[p]);
        r.count = c.inst_leaves ? leaf_weight(c.prims_in, p) : 1;
    }
    return r;
}
// sorted positions of the (at most kLeafMax) primitives below a small subtree, left to right
__device__ __forceinline__ int collect_leaves(const CollapseCtx &c, int ref, int *out) {
    int stack[kLeafMax + 1], sp = 0, n = 0;
    stack[sp++] = ref;
    while (sp > 0 && n < kLeafMax) {
        const int r = stack[--sp];
        if (r < 0) out[n++] = ~r;
        else stack[sp++] = c.t.right[r], stack[sp++] = c.t.left[r];
    }
    return n;
}
// primitive records into BVH leaf order: prims_out[dst_of_sorted[i]] = prims_in[sorted[i]]
__global__ void __launch_bounds__(256) k_scatter_prims(const PrimRec *__restrict__ prims_in, const uint32_t *__restrict__ sorted,
                                                        const uint32_t *__restrict__ dst_of_sorted, uint32_t n, PrimRec *__restrict__ prims_out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t d = dst_of_sorted[i];
        if (d == 0xffffffffu) continue; // an instance leaf: it became a node, not a primitive
        const float4 *src = reinterpret_cast<const float4 *>(prims_in + sorted[i]);
        float4 *dst = reinterpret_cast<float4 *>(prims_out + d);
        const float4 a = __ldg(src), b = __ldg(src + 1), cc = __ldg(src + 2);
        dst[0] = a, dst[1] = b, dst[2] = cc;
    }
}
// work item: x = wide node index, y = binary ref, z = depth
#ifndef PB2_COLLAPSE_MINB
#define PB2_COLLAPSE_MINB 6
#endif
__global__ void __launch_bounds__(128, PB2_COLLAPSE_MINB) k_collapse(CollapseCtx c, const uint4 *__restrict__ q_in, uint32_t n_in, uint4 *__restrict__ q_out) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned warp_mask = __ballot_sync(0xffffffffu, w < n_in); // a prefix of the warp: work items are dense
    if (w >= n_in) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint4 item = q_in[w];
    const Ref self = load_ref(c, (int)item.y);
    Ref ch[8];
    int n = 0;
    const bool is_instance = c.inst_leaves && self.ref < 0 && self.count > kLeafMax; // writes an instance node, has no children
    if (is_instance) {
    } else if (c.word && self.ref >= 0 && self.count > kLeafMax) {
        // the cut below this node that the tables of k_refit<true> found cheapest: hand the eight slots down the binary tree
        int st_ref[8], st_i[8], sp = 0;
        st_ref[sp] = self.ref, st_i[sp++] = 8;
        while (sp > 0) {
            const int m = st_ref[--sp];
            int i = st_i[sp];
            uint32_t k = 0;
            if (m >= 0 && i > 1) {
                const uint32_t w = c.word[m];
                while (i > 1 && (k = (w >> (3 * (i - 2))) & 7u) == 0u) --i;
            }
            if (m < 0 || i == 1) {
                ch[n++] = load_ref(c, m);
                continue;
            }
            st_ref[sp] = c.t.right[m], st_i[sp++] = i - (int)k;
            st_ref[sp] = c.t.left[m], st_i[sp++] = (int)k;
        }
    } else if (self.ref >= 0 && self.count > 1) {
        ch[n++] = load_ref(c, c.t.left[self.ref]);
        ch[n++] = load_ref(c, c.t.right[self.ref]);
        // expand the child with the largest surface area while slots remain; a child can be expanded
        // when it is an internal binary node
        while (n < 8) {
            int best = -1;
            float best_area = -1.f;
            for (int i = 0; i < n; ++i) {
                if (ch[i].ref < 0) continue;
                float a = half_area(ch[i].lo, ch[i].hi);
                if (a > best_area) best_area = a, best = i;
            }
            if (best < 0) break;
            const int b = ch[best].ref;
            ch[best] = load_ref(c, c.t.left[b]);
            ch[n++] = load_ref(c, c.t.right[b]);
        }
    } else {
        ch[n++] = self; // a single primitive (n_prims == 1) or a pure leaf subtree
    }
    // classify
    int n_inner = 0, n_leaf_prims = 0;
    for (int i = 0; i < n; ++i) {
        if (ch[i].count > kLeafMax) ++n_inner;
        else n_leaf_prims += ch[i].count;
    }
    // ---- octant slot assignment (greedy on dot(centroid offset, octant direction)) ----
    // Same greedy order as a plain triple loop (children ascending, slots ascending, first maximum wins); the centroid
    // offsets are hoisted and the bookkeeping lives in bit masks so the 8 x 8 inner loops unroll into registers
    // (ncu: the loop was 40 % of the kernel's instructions, most of them local-memory traffic).
    const float3 centre = (self.lo + self.hi) * 0.5f;
    float offx[8], offy[8], offz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float3 off = i < n ? (ch[i].lo + ch[i].hi) * 0.5f - centre : mk3(0.f);
        offx[i] = off.x, offy[i] = off.y, offz[i] = off.z;
    }
    uint32_t slot_of_packed = 0, slot_used = 0, done = 0; // 4 bits per child | bit per slot | bit per child
    for (int round = 0; round < n; ++round) {
        float best = -FLT_MAX;
        int bi = 0, bs = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i >= n || (done >> i & 1u)) continue;
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) {
                if (slot_used >> sl & 1u) continue;
                const float cost = ((sl & 4) ? offx[i] : -offx[i]) + ((sl & 2) ? offy[i] : -offy[i]) + ((sl & 1) ? offz[i] : -offz[i]);
                if (cost > best) best = cost, bi = i, bs = sl;
            }
        }
        slot_of_packed |= (uint32_t)bs << (4 * bi), slot_used |= 1u << bs, done |= 1u << bi;
    }
    int child_in_slot[8];
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) child_in_slot[sl] = -1;
    for (int i = 0; i < n; ++i) child_in_slot[(slot_of_packed >> (4 * i)) & 7u] = i;

    // ---- allocate children / primitive range ----
    // One atomic per warp and counter instead of four per thread: millions of threads adding to the same three words
    // serialise in L2 (the per-thread version spent most of the kernel's 10 ms per 30 M triangles there).
    uint32_t child_base = 0, prim_base = 0, q_base = 0;
    {
        __syncwarp(warp_mask);
        uint32_t in_incl = (uint32_t)n_inner, pr_incl = (uint32_t)n_leaf_prims;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(warp_mask, in_incl, d), b = __shfl_up_sync(warp_mask, pr_incl, d);
            if ((int)lane >= d) in_incl += a, pr_incl += b;
        }
        const int last = 31 - __clz(warp_mask);
        const uint32_t total_in = __shfl_sync(warp_mask, in_incl, last), total_pr = __shfl_sync(warp_mask, pr_incl, last);
        uint32_t b0 = 0, b1 = 0, b2 = 0;
        if (lane == 0) {
            if (total_in) b0 = atomicAdd(&c.counters[0], total_in), b2 = atomicAdd(&c.counters[2], total_in);
            if (total_pr) b1 = atomicAdd(&c.counters[1], total_pr);
            atomicMax(&c.counters[3], item.z + 1);
        }
        b0 = __shfl_sync(warp_mask, b0, 0), b1 = __shfl_sync(warp_mask, b1, 0), b2 = __shfl_sync(warp_mask, b2, 0);
        child_base = b0 + in_incl - (uint32_t)n_inner, q_base = b2 + in_incl - (uint32_t)n_inner;
        prim_base = b1 + pr_incl - (uint32_t)n_leaf_prims;
    }

    // ---- quantisation frame ----
    const float3 ext = self.hi - self.lo;
    auto exp_of = [](float e) -> int { // smallest ex with 255 * 2^ex >= e
        if (!(e > 0.f)) return -126;
        int ex;
        (void)frexpf(e * 1.000001f / 255.f, &ex); // = m * 2^ex with m in [0.5,1), hence 255 * 2^ex > e
        return max(-126, min(127, ex));
    };
    const int ex = exp_of(ext.x), ey = exp_of(ext.y), ez = exp_of(ext.z);
    const float sx = __int_as_float((ex + 127) << 23), sy = __int_as_float((ey + 127) << 23), sz = __int_as_float((ez + 127) << 23);
    // 1 / scale is a power of two as well: multiplying by it gives the same value as the division (both exact scalings)
    auto inv_pow2 = [](int e, float scale) { return e <= 126 ? __int_as_float((127 - e) << 23) : 1.f / scale; };
    const float isx = inv_pow2(ex, sx), isy = inv_pow2(ey, sy), isz = inv_pow2(ez, sz);
    const float3 p = self.lo;

    uint32_t imask = 0, meta[8], qlo[3][8], qhi[3][8];
    uint32_t inner_seen = 0, prim_off = 0;
    float sah_local = half_area(self.lo, self.hi); // node visit cost 1
    for (int s = 0; s < 8; ++s) {
        const int i = child_in_slot[s];
        if (i < 0) {
            meta[s] = 0;
            for (int k = 0; k < 3; ++k) qlo[k][s] = 255, qhi[k][s] = 0; // inverted box: can never be hit
            continue;
        }
        const Ref &r = ch[i];
        auto qfloor = [](float v, float org, float scale, float inv_scale) -> uint32_t {
            float q = floorf((v - org) * inv_scale);
            q = fminf(fmaxf(q, 0.f), 255.f);
            if (q > 0.f && org + q * scale > v) q -= 1.f; // rounding of (v - org) must not shrink the box
            return (uint32_t)q;
        };
        auto qceil = [](float v, float org, float scale, float inv_scale) -> uint32_t {
            float q = ceilf((v - org) * inv_scale);
            q = fminf(fmaxf(q, 0.f), 255.f);
            if (q < 255.f && org + q * scale < v) q += 1.f;
            return (uint32_t)q;
        };
        qlo[0][s] = qfloor(r.lo.x, p.x, sx, isx), qlo[1][s] = qfloor(r.lo.y, p.y, sy, isy), qlo[2][s] = qfloor(r.lo.z, p.z, sz, isz);
        qhi[0][s] = qceil(r.hi.x, p.x, sx, isx), qhi[1][s] = qceil(r.hi.y, p.y, sy, isy), qhi[2][s] = qceil(r.hi.z, p.z, sz, isz);
        if (r.count > kLeafMax) {
            imask |= 1u << s;
            meta[s] = (1u << 5) | (24u + s);
            q_out[q_base + inner_seen] = make_uint4(child_base + inner_seen, (uint32_t)r.ref, item.z + 1, 0u);
            ++inner_seen;
        } else {
            const uint32_t unary = r.count == 1 ? 1u : r.count == 2 ? 3u : 7u;
            meta[s] = (unary << 5) | prim_off;
            // the 48-byte records are moved by k_scatter_prims afterwards (one thread per record instead of a dependent
            // gather loop per wide node: the loop was the kernel's largest single stall)
            int pos[kLeafMax];
            const int found = collect_leaves(c, r.ref, pos);
            for (int k = 0; k < found; ++k) c.dst_of_sorted[pos[k]] = prim_base + prim_off + k;
            prim_off += r.count;
            sah_local += half_area(r.lo, r.hi) * r.count;
        }
    }
    // child nodes must sit at child_base + (number of internal slots below): q_out was filled in slot order ✓
    auto pack4 = [](const uint32_t *v) { return v[0] | v[1] << 8 | v[2] << 16 | v[3] << 24; };
    Bvh8Node node;
    node.n0 = make_float4(p.x, p.y, p.z, __uint_as_float((uint32_t)(ex + 127) | (uint32_t)(ey + 127) << 8 | (uint32_t)(ez + 127) << 16 | imask << 24));
    node.n1 = make_uint4(child_base, prim_base, pack4(meta), pack4(meta + 4));
    node.n2 = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    node.n3 = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    node.n4 = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    if (is_instance) { // traverse.cuh: tag in the exponent / mask word, bottom-level root and instance id in n1
        const PrimRec rec = c.prims_in[c.sorted[~self.ref]];
        node.n0 = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xffffffffu));
        node.n1 = make_uint4(__float_as_uint(rec.v0.w), __float_as_uint(rec.e1.w), 0u, 0u);
        node.n2 = node.n3 = node.n4 = make_uint4(0u, 0u, 0u, 0u);
        sah_local = 0.f;
    }
    c.nodes[item.x] = node;
    __syncwarp(warp_mask);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const float other = __shfl_down_sync(warp_mask, sah_local, d);
        if (lane + d < 32u && (warp_mask >> (lane + d) & 1u)) sah_local += other;
    }
    if (lane == 0) atomicAdd(c.sah, sah_local);
}
}// namespace

// bvh_ploc.cu: SAH-driven bottom-up clustering producing the same BinTree arrays; returns the root's node index
int build_binary_ploc(cudaStream_t st, uint32_t n, const float4 *box_lo, const float4 *box_hi, const uint32_t *sorted, int *left, int *right, int2 *range,
                      float4 *lo, float4 *hi, int radius, uint32_t *rounds_out);
// radix_sort.cu: in-tree onesweep sort of (64-bit key, 32-bit value) pairs; true = the result is in the alt buffers
bool radix_sort_pairs(cudaStream_t st, uint64_t *keys, uint64_t *keys_alt, uint32_t *vals, uint32_t *vals_alt, uint32_t n, int begin_bit, int end_bit);
// bvh_sah.cu: binned-SAH binary tree producing the same BinTree arrays + `sorted` permutation
bool sah_builder_available();
void build_binary_sah(cudaStream_t st, uint32_t n, const float4 *box_lo, const float4 *box_hi, const uint32_t *sorted, int *left, int *right, int2 *range,
                      float4 *lo, float4 *hi);

// a level's nodes, built with indices relative to the level, into their place in the scene's node array
__global__ void __launch_bounds__(256) k_relocate_nodes(const Bvh8Node *__restrict__ src, uint32_t n, uint32_t node_offset, uint32_t prim_offset, Bvh8Node *__restrict__ dst) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Bvh8Node nd = src[i];
        if (__float_as_uint(nd.n0.w) != 0xffffffffu) nd.n1.x += node_offset, nd.n1.y += prim_offset; // instance nodes hold scene-wide indices already
        dst[node_offset + i] = nd;
    }
}
// largest vertex index of an index buffer (pb2_scene_add_mesh rejects meshes that point past their vertex arrays)
__global__ void __launch_bounds__(256) k_max_index(const uint32_t *__restrict__ idx, uint64_t n, uint32_t *__restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, __ldg(idx + i));
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31u) == 0u) atomicMax(out, m);
}
uint32_t max_index_dev(const uint32_t *idx, uint64_t n, cudaStream_t st) {
    if (!n) return 0;
    DevBuf<uint32_t> d(1);
    d.zero(st);
    k_max_index<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(idx, n, d.ptr);
    PB2_LAUNCH_CHECK();
    uint32_t h = 0;
    PB2_CUDA(cudaMemcpyAsync(&h, d.ptr, sizeof h, cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaStreamSynchronize(st));
    return h;
}

namespace {
constexpr uint32_t kBlasMinTris = 64; // smaller shared meshes are flattened into the top level: an instance step costs more than their triangles

struct LevelOut {
    DevBuf<Bvh8Node> nodes; // relative indices; relocated into the scene's array by the caller
    uint32_t n_nodes = 0, n_prims_out = 0, depth = 0;
    float sah = 0.f, lo[3] = { 0, 0, 0 }, hi[3] = { 0, 0, 0 };
};

void refit_seeds(cudaStream_t st, const BinTree &t, DevBuf<int> &seeds, DevBuf<uint32_t> &n_seeds) {
    seeds.ensure(t.n / 2 + 1), n_seeds.ensure(1);
    n_seeds.zero(st);
    k_refit_seeds<<<div_up(t.n - 1, 1024), 1024, 0, st>>>(t, seeds.ptr, n_seeds.ptr);
    PB2_LAUNCH_CHECK();
}
// 36 bytes per binary node for the cost-optimal collapse; a scene too large for them falls back to the greedy cut instead of failing
bool alloc_cost_tables(DevBuf<float4> &cost_c, DevBuf<uint32_t> &cost_word, uint32_t n) {
    try {
        cost_c.alloc(2 * (size_t)n), cost_word.alloc(n);
        return true;
    } catch (const CudaError &) {
        cudaGetLastError();
        cost_c.release(), cost_word.release();
        return false;
    }
}

// Stages 2-5 over n emitted records (prims_in, box_lo / box_hi, bounds): Morton sort, binary tree, collapse, scatter of the
// records into prims_out (their final place).  inst_leaves: records of kind 2 become instance nodes (top level of a two-level scene).
void build_level(Scene &s, cudaStream_t st, uint32_t n, uint32_t n_inst_leaves, DevBuf<PrimRec> &prims_in, DevBuf<float4> &box_lo, DevBuf<float4> &box_hi,
                 DevBuf<int> &bounds, PrimRec *prims_out, LevelOut &out) {
    const bool inst_leaves = n_inst_leaves > 0;
    DevBuf<uint32_t> sorted(n);
    DevBuf<int> left(n), right(n), parent(2 * (size_t)n);
    DevBuf<int2> range(n);
    DevBuf<float4> nlo(n), nhi(n);
    BinTree t{ left.ptr, right.ptr, parent.ptr, range.ptr, nlo.ptr, nhi.ptr, n };
    int root_ref = 0; // binary node the collapse starts from (node 0 for the top-down builders)
    DevBuf<int> seeds;          // internal nodes with two leaf children: where k_refit starts
    DevBuf<uint32_t> n_seeds;
    DevBuf<float4> cost_c;      // cost-optimal collapse (Scene::collapse = 1, LBVH): filled by k_refit<true>, read by k_collapse
    DevBuf<uint32_t> cost_word;
    {
        // 63-bit Morton keys, sorted by the in-tree onesweep radix sort (radix_sort.cu: eight 8-bit passes; measured against
        // cub::DeviceRadixSort on the 30 M-triangle terrain: 380 vs 347 us per pass, the whole build 12.9 vs 12.8 ms)
        DevBuf<uint64_t> keys(n), keys_sorted(n);
        DevBuf<uint32_t> vals(n);
        // Key bits that are sorted: all 63, or (morton_bits = 0, the default) log2(n) + 21 of them rounded up to whole 8-bit sort passes —
        // seven bits per axis beyond what n uniformly spread primitives need to fall into cells of their own; primitives that
        // still share a key keep their index order.  30 M triangles: 6 passes instead of 8, the same tree, 0.75 ms less.
        int morton_bits = s.morton_bits;
        if (morton_bits <= 0) {
            int lg = 0;
            while ((1ull << lg) < n) ++lg;
            morton_bits = 8 * std::min(8, std::max(3, (lg + 21 + 7) / 8)) - 1;
        }
        const int low_bit = 63 - std::min(63, morton_bits);
        k_morton<<<div_up(n, 256), 256, 0, st>>>(box_lo.ptr, box_hi.ptr, bounds.ptr, n, keys_sorted.ptr, sorted.ptr, low_bit);
        PB2_LAUNCH_CHECK();
        if (radix_sort_pairs(st, keys_sorted.ptr, keys.ptr, sorted.ptr, vals.ptr, n, low_bit, 63)) { // an odd pass count leaves the result in the scratch buffers
            PB2_CUDA(cudaMemcpyAsync(keys_sorted.ptr, keys.ptr, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
            PB2_CUDA(cudaMemcpyAsync(sorted.ptr, vals.ptr, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        }
        // a top level with instance leaves needs the weighted counts of k_refit: it always takes the LBVH path (it is small)
        if (n > 1 && s.builder == 2 && !inst_leaves) {
            // bottom-up clustering by the surface area of the union (bvh_ploc.cu); node boxes come out of the merges
            root_ref = build_binary_ploc(st, n, box_lo.ptr, box_hi.ptr, sorted.ptr, left.ptr, right.ptr, range.ptr, nlo.ptr, nhi.ptr, s.ploc_radius, nullptr);
        } else if (n > 1 && s.builder == 1 && !inst_leaves && sah_builder_available()) {
            // binned SAH over the Morton-ordered sequence (bvh_sah.cu); node boxes come out of the sweep
            build_binary_sah(st, n, box_lo.ptr, box_hi.ptr, sorted.ptr, left.ptr, right.ptr, range.ptr, nlo.ptr, nhi.ptr);
        } else if (n > 1) {
            k_radix_tree<<<div_up(n - 1, 256), 256, 0, st>>>(keys_sorted.ptr, t);
            PB2_LAUNCH_CHECK();
            DevBuf<int> arrive(n);
            arrive.zero(st);
            refit_seeds(st, t, seeds, n_seeds);
            const unsigned refit_grid = div_up(n / 2 + 1, kRefitBlock); // at most every other internal node has two leaf children
            if (s.collapse == 1 && alloc_cost_tables(cost_c, cost_word, n)) {
                const CostTab ct{ cost_c.ptr, cost_word.ptr, (float)s.collapse_prim_cost_pct * 0.01f };
                k_refit<true><<<refit_grid, kRefitBlock, 0, st>>>(t, sorted.ptr, box_lo.ptr, box_hi.ptr, arrive.ptr, inst_leaves ? prims_in.ptr : nullptr, ct, seeds.ptr, n_seeds.ptr);
            } else {
                k_refit<false><<<refit_grid, kRefitBlock, 0, st>>>(t, sorted.ptr, box_lo.ptr, box_hi.ptr, arrive.ptr, inst_leaves ? prims_in.ptr : nullptr, CostTab{}, seeds.ptr, n_seeds.ptr);
            }
            PB2_LAUNCH_CHECK();
        }
        if (n > 1 && s.collapse == 1 && !cost_word.ptr && s.builder != 0 && !inst_leaves && alloc_cost_tables(cost_c, cost_word, n)) {
            // the builders that bring their own boxes: one more bottom-up sweep for the cost tables (it recomputes the same boxes)
            k_parents<<<div_up(n - 1, 256), 256, 0, st>>>(t, root_ref);
            PB2_LAUNCH_CHECK();
            DevBuf<int> arrive(n);
            arrive.zero(st);
            refit_seeds(st, t, seeds, n_seeds);
            const CostTab ct{ cost_c.ptr, cost_word.ptr, (float)s.collapse_prim_cost_pct * 0.01f };
            k_refit<true><<<div_up(n / 2 + 1, kRefitBlock), kRefitBlock, 0, st>>>(t, sorted.ptr, box_lo.ptr, box_hi.ptr, arrive.ptr, nullptr, ct, seeds.ptr, n_seeds.ptr);
            PB2_LAUNCH_CHECK();
            PB2_CUDA(cudaStreamSynchronize(st));
        }
        PB2_CUDA(cudaStreamSynchronize(st)); // tmp / arrive / keys go out of scope
    }
    // ---- collapse ----
    out.nodes.alloc((size_t)n + n_inst_leaves + 1); // upper bound: one wide node per binary internal node (+ root) + one instance node per instance leaf
    DevBuf<uint32_t> counters(4);
    DevBuf<float> sah(1);
    sah.zero(st);
    uint32_t init_counters[4] = { 1, 0, 0, 0 };
    PB2_CUDA(cudaMemcpyAsync(counters.ptr, init_counters, sizeof init_counters, cudaMemcpyHostToDevice, st));
    DevBuf<uint4> qa(n), qb(n);
    uint4 root = make_uint4(0u, n > 1 ? (uint32_t)root_ref : (uint32_t)~0, 0u, 0u); // n == 1: leaf ref ~0
    PB2_CUDA(cudaMemcpyAsync(qa.ptr, &root, sizeof root, cudaMemcpyHostToDevice, st));
    DevBuf<uint32_t> dst_of_sorted(n);
    PB2_CUDA(cudaMemsetAsync(dst_of_sorted.ptr, 0xff, (size_t)n * sizeof(uint32_t), st)); // records that get no leaf slot (instance leaves) stay ~0
    CollapseCtx cc{ t, sorted.ptr, box_lo.ptr, box_hi.ptr, prims_in.ptr, dst_of_sorted.ptr, out.nodes.ptr, counters.ptr, sah.ptr, inst_leaves, cost_word.ptr };
    uint32_t n_in = 1;
    uint4 *q_in = qa.ptr, *q_out = qb.ptr;
    uint32_t host_counters[4] = { 0, 0, 0, 0 };
    while (n_in) {
        k_collapse<<<div_up(n_in, 128), 128, 0, st>>>(cc, q_in, n_in, q_out);
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaMemcpyAsync(host_counters, counters.ptr, sizeof host_counters, cudaMemcpyDeviceToHost, st));
        PB2_CUDA(cudaStreamSynchronize(st));
        n_in = host_counters[2];
        PB2_CUDA(cudaMemsetAsync(counters.ptr + 2, 0, sizeof(uint32_t), st));
        std::swap(q_in, q_out);
    }
    out.n_nodes = host_counters[0], out.n_prims_out = host_counters[1], out.depth = host_counters[3];
    if (out.n_prims_out != n - n_inst_leaves) throw std::runtime_error("pb2_bvh_build: collapse lost primitives");
    if (out.n_prims_out) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        k_scatter_prims<<<(unsigned)std::min<uint64_t>(div_up(n, 256), (uint64_t)sms * 8), 256, 0, st>>>(prims_in.ptr, sorted.ptr, dst_of_sorted.ptr, n, prims_out);
        PB2_LAUNCH_CHECK();
    }
    float sah_host = 0.f;
    int hb[6];
    PB2_CUDA(cudaMemcpyAsync(&sah_host, sah.ptr, sizeof(float), cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaMemcpyAsync(hb, bounds.ptr, sizeof hb, cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; ++k) out.lo[k] = ordered_to_float(hb[k]), out.hi[k] = ordered_to_float(hb[3 + k]);
    const float d[3] = { out.hi[0] - out.lo[0], out.hi[1] - out.lo[1], out.hi[2] - out.lo[2] };
    const float root_area = d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
    out.sah = root_area > 0.f ? sah_host / root_area : 0.f;
}

void init_bounds(DevBuf<int> &bounds, cudaStream_t st) {
    int init[6];
    float mx = FLT_MAX, mn = -FLT_MAX;
    int imx, imn;
    memcpy(&imx, &mx, 4), memcpy(&imn, &mn, 4);
    imn = imn ^ 0x7fffffff; // ordered encoding of -FLT_MAX
    for (int k = 0; k < 3; ++k) init[k] = imx, init[3 + k] = imn;
    PB2_CUDA(cudaMemcpyAsync(bounds.ptr, init, sizeof init, cudaMemcpyHostToDevice, st));
    PB2_CUDA(cudaStreamSynchronize(st)); // `init` is a stack array
}
}// namespace

// Which meshes get a bottom-level tree of their own (Scene::instancing): 1 (default) = meshes placed more than once, as the
// reference shares one GAS between the instances of a shape (gas_manager.cpp:10 RefGAS); 2 = every mesh, so that a transform
// edit never rebuilds more than the top level (ias_manager.cpp:116-151 IAS::Update); 0 = none, everything flattened to world space.
static bool wants_blas(const Scene &s, const Mesh &m, uint32_t uses) {
    if (s.instancing == 0 || m.n_tris < kBlasMinTris) return false;
    return s.instancing >= 2 ? uses >= 1 : uses >= 2;
}

void build_bvh(Scene &s) {
    cudaStream_t st = s.stream;
    s.upload_tables();
    s.bvh_valid = false;
    const uint32_t n_inst = (uint32_t)s.h_inst.size();
    struct Events { // destroyed on every way out, exceptions included
        cudaEvent_t e0 = nullptr, e1 = nullptr, t0 = nullptr;
        ~Events() {
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            if (t0) cudaEventDestroy(t0);
        }
    } ev;
    PB2_CUDA(cudaEventCreate(&ev.e0));
    PB2_CUDA(cudaEventCreate(&ev.e1));
    PB2_CUDA(cudaEventCreate(&ev.t0));
    cudaEvent_t &e0 = ev.e0, &e1 = ev.e1, &t0 = ev.t0;
    PB2_CUDA(cudaEventRecord(e0, st));

    // ---- which mesh of which instance ----
    const std::vector<int> &mesh_of_inst = s.h_inst_mesh; // -1: analytic sphere
    std::vector<uint32_t> uses(s.meshes.size(), 0);
    for (uint32_t i = 0; i < n_inst; ++i)
        if (mesh_of_inst[i] >= 0) ++uses[mesh_of_inst[i]];
    // a finished top-level-only update keeps the bottom-level trees; anything else starts over
    const bool tlas_only = s.blas_valid && s.n_blas > 0;
    if (!tlas_only) {
        for (auto &m : s.meshes) m->blas = Blas{};
        s.n_blas = 0;
    }
    pb2_build_stats stats{};
    std::vector<std::unique_ptr<LevelOut>> levels; // bottom-level trees built in this call, in mesh order
    std::vector<size_t> level_mesh;
    uint64_t blas_prims = 0, blas_nodes_cap = 0;
    if (!tlas_only) {
        for (size_t m = 0; m < s.meshes.size(); ++m)
            if (wants_blas(s, *s.meshes[m], uses[m])) blas_prims += s.meshes[m]->n_tris;
    } else {
        for (auto &m : s.meshes)
            if (m->blas.valid) blas_prims += m->blas.n_prims;
    }
    // ---- top-level census ----
    std::vector<uint32_t> flat_ids, first;
    std::vector<uint32_t> leaf_inst;
    uint64_t n_flat = 0, n_sph = 0;
    for (uint32_t i = 0; i < n_inst; ++i) {
        const int m = mesh_of_inst[i];
        const bool blas = m >= 0 && (tlas_only ? s.meshes[m]->blas.valid : wants_blas(s, *s.meshes[m], uses[m]));
        if (blas) {
            leaf_inst.push_back(i);
            continue;
        }
        flat_ids.push_back(i), first.push_back((uint32_t)n_flat);
        n_flat += s.h_inst[i].n_tris;
        if (s.h_inst[i].flags & PB2_IF_SPHERE) ++n_sph;
    }
    first.push_back((uint32_t)n_flat);
    const uint64_t n_top = n_flat + leaf_inst.size();
    if (blas_prims + n_top >= 0x7fffffffull) throw std::runtime_error("pb2_bvh_build: more than 2^31-1 primitives");
    stats.n_spheres = n_sph;
    if (n_top == 0) {
        s.n_nodes = s.n_prims = 0, s.root = 0, s.n_blas = 0, s.blas_valid = false;
        s.d_nodes.release(), s.d_prims.release();
        s.build_stats = stats;
        s.bvh_valid = true;
        return;
    }
    // ---- primitive array: [bottom-level records, object space][top-level records, world space] ----
    const uint32_t top_prim_offset = (uint32_t)blas_prims;
    if (!tlas_only || s.d_prims.n < blas_prims + n_flat) {
        if (tlas_only) throw std::runtime_error("pb2_bvh_build: top-level update with a changed primitive count (rebuild the scene)");
        s.d_prims.alloc(blas_prims + n_flat);
    }
    // ---- bottom-level trees ----
    uint32_t max_blas_depth = 0;
    if (!tlas_only) {
        uint32_t prim_cursor = 0;
        for (size_t m = 0; m < s.meshes.size(); ++m) {
            Mesh &mesh = *s.meshes[m];
            if (!wants_blas(s, mesh, uses[m])) continue;
            const uint32_t n = mesh.n_tris;
            DevInstance ident{};
            ident.xf[0] = ident.inv[0] = make_float4(1, 0, 0, 0), ident.xf[1] = ident.inv[1] = make_float4(0, 1, 0, 0), ident.xf[2] = ident.inv[2] = make_float4(0, 0, 1, 0);
            ident.pos = mesh.pos.ptr, ident.nrm = mesh.nrm.ptr, ident.uv = mesh.uv.ptr, ident.idx = mesh.idx.ptr, ident.n_tris = n, ident.emitter_offset = -1;
            DevBuf<DevInstance> d_ident(1);
            d_ident.upload(&ident, 1, st);
            const uint32_t firsts[2] = { 0u, n };
            DevBuf<uint32_t> d_first(2);
            d_first.upload(firsts, 2, st);
            PB2_CUDA(cudaStreamSynchronize(st));
            DevBuf<PrimRec> prims_in(n);
            DevBuf<float4> box_lo(n), box_hi(n);
            DevBuf<int> bounds(6);
            init_bounds(bounds, st);
            k_emit_prims<<<div_up(n, 256), 256, 0, st>>>(d_ident.ptr, d_first.ptr, nullptr, 1u, n, prims_in.ptr, box_lo.ptr, box_hi.ptr, bounds.ptr);
            PB2_LAUNCH_CHECK();
            auto lv = std::make_unique<LevelOut>();
            build_level(s, st, n, 0, prims_in, box_lo, box_hi, bounds, s.d_prims.ptr + prim_cursor, *lv);
            mesh.blas.valid = true, mesh.blas.prim_offset = prim_cursor, mesh.blas.n_prims = n, mesh.blas.n_nodes = lv->n_nodes, mesh.blas.depth = lv->depth;
            for (int k = 0; k < 3; ++k) mesh.blas.lo[k] = lv->lo[k], mesh.blas.hi[k] = lv->hi[k];
            prim_cursor += n;
            blas_nodes_cap += lv->n_nodes;
            levels.push_back(std::move(lv)), level_mesh.push_back(m);
            ++s.n_blas;
        }
        uint32_t node_cursor = 0;
        for (size_t k = 0; k < levels.size(); ++k) {
            s.meshes[level_mesh[k]]->blas.node_offset = node_cursor;
            node_cursor += levels[k]->n_nodes;
        }
        s.top_node_offset = node_cursor;
    }
    for (auto &m : s.meshes)
        if (m->blas.valid) max_blas_depth = std::max(max_blas_depth, m->blas.depth), stats.n_nodes += m->blas.n_nodes, stats.n_triangles += m->blas.n_prims;

    // ---- top level ----
    PB2_CUDA(cudaEventRecord(t0, st));
    const uint32_t n = (uint32_t)n_top, n_leaves = (uint32_t)leaf_inst.size();
    DevBuf<PrimRec> prims_in(n);
    DevBuf<float4> box_lo(n), box_hi(n);
    DevBuf<int> bounds(6);
    init_bounds(bounds, st);
    if (n_flat) {
        DevBuf<uint32_t> d_first(first.size()), d_ids(flat_ids.size());
        d_first.upload(first.data(), first.size(), st);
        d_ids.upload(flat_ids.data(), flat_ids.size(), st);
        k_emit_prims<<<div_up(n_flat, 256), 256, 0, st>>>(s.d_inst.ptr, d_first.ptr, d_ids.ptr, (uint32_t)flat_ids.size(), (uint32_t)n_flat, prims_in.ptr, box_lo.ptr,
                                                          box_hi.ptr, bounds.ptr);
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaStreamSynchronize(st));
    }
    if (n_leaves) { // world-space box of every placement: the eight corners of the bottom-level root box through the instance transform
        std::vector<float4> lo(n_leaves), hi(n_leaves);
        std::vector<uint32_t> roots(n_leaves);
        for (uint32_t k = 0; k < n_leaves; ++k) {
            const DevInstance &in = s.h_inst[leaf_inst[k]];
            const Blas &b = s.meshes[mesh_of_inst[leaf_inst[k]]]->blas;
            float l[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, h[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
            for (int c = 0; c < 8; ++c) {
                const float p[3] = { c & 1 ? b.hi[0] : b.lo[0], c & 2 ? b.hi[1] : b.lo[1], c & 4 ? b.hi[2] : b.lo[2] };
                for (int r = 0; r < 3; ++r) {
                    const float4 row = in.xf[r];
                    const float v = row.x * p[0] + row.y * p[1] + row.z * p[2] + row.w;
                    l[r] = std::min(l[r], v), h[r] = std::max(h[r], v);
                }
            }
            // the traversal transforms the RAY, in fp32: pad the box by a few ulps of its largest coordinate
            float pad = 0.f;
            for (int r = 0; r < 3; ++r) pad = std::max(pad, std::max(std::fabs(l[r]), std::fabs(h[r])));
            pad *= 4e-7f;
            lo[k] = make_float4(l[0] - pad, l[1] - pad, l[2] - pad, 0.f), hi[k] = make_float4(h[0] + pad, h[1] + pad, h[2] + pad, 0.f);
            roots[k] = b.node_offset;
        }
        DevBuf<float4> d_lo(n_leaves), d_hi(n_leaves);
        DevBuf<uint32_t> d_roots(n_leaves), d_ids(n_leaves);
        d_lo.upload(lo.data(), n_leaves, st), d_hi.upload(hi.data(), n_leaves, st);
        d_roots.upload(roots.data(), n_leaves, st), d_ids.upload(leaf_inst.data(), n_leaves, st);
        k_emit_instance_leaves<<<div_up(n_leaves, 256), 256, 0, st>>>(n_leaves, (uint32_t)n_flat, d_lo.ptr, d_hi.ptr, d_roots.ptr, d_ids.ptr, prims_in.ptr, box_lo.ptr,
                                                                      box_hi.ptr, bounds.ptr);
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaStreamSynchronize(st));
    }
    LevelOut top;
    build_level(s, st, n, n_leaves, prims_in, box_lo, box_hi, bounds, s.d_prims.ptr + top_prim_offset, top);
    if (!tlas_only) {
        // node array: [bottom-level trees][top level + room to grow: a top-level update after transform edits may need a few more
        // wide nodes than this build did]
        const uint64_t top_cap = std::min<uint64_t>(n_top + n_leaves + 1, (uint64_t)top.n_nodes + top.n_nodes / 2 + 64);
        s.d_nodes.alloc((uint64_t)s.top_node_offset + top_cap);
        for (size_t k = 0; k < levels.size(); ++k) {
            const Blas &b = s.meshes[level_mesh[k]]->blas;
            k_relocate_nodes<<<div_up(levels[k]->n_nodes, 256), 256, 0, st>>>(levels[k]->nodes.ptr, levels[k]->n_nodes, b.node_offset, b.prim_offset, s.d_nodes.ptr);
            PB2_LAUNCH_CHECK();
        }
        s.blas_valid = s.n_blas > 0;
    } else if ((uint64_t)s.top_node_offset + top.n_nodes > s.d_nodes.n) { // the update outgrew its room: build everything again
        s.blas_valid = false;
        build_bvh(s);
        return;
    }
    k_relocate_nodes<<<div_up(top.n_nodes, 256), 256, 0, st>>>(top.nodes.ptr, top.n_nodes, s.top_node_offset, top_prim_offset, s.d_nodes.ptr);
    PB2_LAUNCH_CHECK();
    PB2_CUDA(cudaEventRecord(e1, st));
    PB2_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f, top_ms = 0.f;
    PB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    PB2_CUDA(cudaEventElapsedTime(&top_ms, t0, e1));

    s.root = s.top_node_offset;
    s.n_nodes = s.top_node_offset + top.n_nodes;
    s.n_prims = top_prim_offset + top.n_prims_out;
    stats.n_triangles += n_flat - n_sph;
    stats.n_prims = stats.n_triangles + n_sph;
    stats.n_nodes += top.n_nodes;
    stats.bvh_bytes = (uint64_t)s.n_nodes * sizeof(Bvh8Node) + (uint64_t)s.n_prims * sizeof(PrimRec);
    stats.build_ms = ms;
    stats.sah_cost = top.sah;
    stats.max_depth = top.depth + (s.n_blas ? 1 + max_blas_depth : 0);
    stats.n_blas = s.n_blas, stats.n_instance_leaves = n_leaves, stats.top_level_ms = top_ms;
    s.build_stats = stats;
    if (stats.max_depth > PB2_STACK_SIZE - 2) throw std::runtime_error("pb2_bvh_build: tree deeper than the traversal stack (" + std::to_string(stats.max_depth) + " levels)");
    s.bvh_valid = true;
    if (s.l2_persist_mb > 0) s.l2_dirty = true; // the node array moved
}
}// namespace pb2
