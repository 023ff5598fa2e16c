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
//   2 morton + sort   63-bit keys, cub::DeviceRadixSort::SortPairs
//   3 radix_tree      Karras 2012 binary radix tree (ties broken by index)
//   4 refit           bottom-up AABBs with per-node arrival counters
//   5 collapse        breadth-first: binary subtree -> up to 8 children by largest-area expansion,
//                     octant slot assignment, quantisation, leaf primitive copy
#include "scene.cuh"
#include "traverse.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>

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
// One thread per primitive.  inst_first[i] = index of the first primitive of instance i (n_inst+1 entries).
__global__ void k_emit_prims(const DevInstance *__restrict__ inst, const uint32_t *__restrict__ inst_first, uint32_t n_inst, uint32_t n_prims,
                             PrimRec *__restrict__ prims, float4 *__restrict__ box_lo, float4 *__restrict__ box_hi, int *__restrict__ scene_bounds) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = mk3(FLT_MAX), hi = mk3(-FLT_MAX);
    if (g < n_prims) {
        uint32_t a = 0, b = n_inst; // last instance with inst_first <= g
        while (b - a > 1) {
            uint32_t m = (a + b) >> 1;
            if (inst_first[m] <= g) a = m;
            else b = m;
        }
        const DevInstance &in = inst[a];
        const uint32_t prim = g - inst_first[a];
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
                         uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
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
    keys[g] = spread21(x) << 2 | spread21(y) << 1 | spread21(z);
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
__global__ void k_refit(BinTree t, const uint32_t *__restrict__ sorted, const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi,
                        int *__restrict__ arrive) {
    const int n = t.n, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int node = t.parent[(n - 1) + j];
    while (node >= 0) {
        if (atomicAdd(&arrive[node], 1) == 0) return; // first child to arrive leaves; the second one continues
        __threadfence();
        const int lc = t.left[node], rc = t.right[node];
        float3 llo, lhi, rlo, rhi;
        if (lc >= 0) llo = mk3(__ldcg(&t.lo[lc])), lhi = mk3(__ldcg(&t.hi[lc]));
        else llo = mk3(box_lo[sorted[~lc]]), lhi = mk3(box_hi[sorted[~lc]]);
        if (rc >= 0) rlo = mk3(__ldcg(&t.lo[rc])), rhi = mk3(__ldcg(&t.hi[rc]));
        else rlo = mk3(box_lo[sorted[~rc]]), rhi = mk3(box_hi[sorted[~rc]]);
        const float3 lo = fmin3(llo, rlo), hi = fmax3(lhi, rhi);
        t.lo[node] = make_float4(lo.x, lo.y, lo.z, 0.f);
        t.hi[node] = make_float4(hi.x, hi.y, hi.z, 0.f);
        __threadfence();
        node = t.parent[node];
    }
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
        r.count = 1;
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
        const float4 *src = reinterpret_cast<const float4 *>(prims_in + sorted[i]);
        float4 *dst = reinterpret_cast<float4 *>(prims_out + dst_of_sorted[i]);
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
    if (self.ref >= 0 && self.count > 1) {
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
// bvh_sah.cu: binned-SAH binary tree producing the same BinTree arrays + `sorted` permutation
bool sah_builder_available();
void build_binary_sah(cudaStream_t st, uint32_t n, const float4 *box_lo, const float4 *box_hi, const uint32_t *sorted, int *left, int *right, int2 *range,
                      float4 *lo, float4 *hi);

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

void build_bvh(Scene &s) {
    cudaStream_t st = s.stream;
    s.upload_tables();
    s.bvh_valid = false;
    s.n_nodes = s.n_prims = 0;
    s.build_stats = pb2_build_stats{};

    const uint32_t n_inst = (uint32_t)s.h_inst.size();
    std::vector<uint32_t> first(n_inst + 1, 0);
    uint64_t total = 0, n_sph = 0;
    for (uint32_t i = 0; i < n_inst; ++i) {
        first[i] = (uint32_t)total;
        total += s.h_inst[i].n_tris;
        if (s.h_inst[i].flags & PB2_IF_SPHERE) ++n_sph;
    }
    first[n_inst] = (uint32_t)total;
    if (total >= 0x7fffffffull) throw std::runtime_error("pb2_bvh_build: more than 2^31-1 primitives");
    const uint32_t n = (uint32_t)total;
    s.build_stats.n_prims = n, s.build_stats.n_spheres = n_sph, s.build_stats.n_triangles = n - n_sph;
    if (n == 0) {
        s.bvh_valid = true;
        return;
    }
    cudaEvent_t e0, e1;
    PB2_CUDA(cudaEventCreate(&e0));
    PB2_CUDA(cudaEventCreate(&e1));
    PB2_CUDA(cudaEventRecord(e0, st));

    DevBuf<uint32_t> d_first(n_inst + 1);
    d_first.upload(first.data(), n_inst + 1, st);
    DevBuf<PrimRec> prims_in(n);
    DevBuf<float4> box_lo(n), box_hi(n);
    DevBuf<int> bounds(6);
    {
        int init[6];
        float mx = FLT_MAX, mn = -FLT_MAX;
        int imx, imn;
        memcpy(&imx, &mx, 4), memcpy(&imn, &mn, 4);
        imn = imn ^ 0x7fffffff; // ordered encoding of -FLT_MAX
        for (int k = 0; k < 3; ++k) init[k] = imx, init[3 + k] = imn;
        PB2_CUDA(cudaMemcpyAsync(bounds.ptr, init, sizeof init, cudaMemcpyHostToDevice, st));
    }
    k_emit_prims<<<div_up(n, 256), 256, 0, st>>>(s.d_inst.ptr, d_first.ptr, n_inst, n, prims_in.ptr, box_lo.ptr, box_hi.ptr, bounds.ptr);
    PB2_LAUNCH_CHECK();

    DevBuf<uint32_t> sorted(n);
    DevBuf<int> left(n), right(n), parent(2 * (size_t)n);
    DevBuf<int2> range(n);
    DevBuf<float4> nlo(n), nhi(n);
    BinTree t{ left.ptr, right.ptr, parent.ptr, range.ptr, nlo.ptr, nhi.ptr, n };
    int root_ref = 0; // binary node the collapse starts from (node 0 for the top-down builders)

    {
        DevBuf<uint64_t> keys(n), keys_sorted(n);
        DevBuf<uint32_t> vals(n);
        k_morton<<<div_up(n, 256), 256, 0, st>>>(box_lo.ptr, box_hi.ptr, bounds.ptr, n, keys.ptr, vals.ptr);
        PB2_LAUNCH_CHECK();
        size_t tmp_bytes = 0;
        PB2_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.ptr, keys_sorted.ptr, vals.ptr, sorted.ptr, (int)n, 0, 63, st));
        DevBuf<uint8_t> tmp(tmp_bytes);
        PB2_CUDA(cub::DeviceRadixSort::SortPairs(tmp.ptr, tmp_bytes, keys.ptr, keys_sorted.ptr, vals.ptr, sorted.ptr, (int)n, 0, 63, st));
        if (n > 1 && s.builder == 2) {
            // bottom-up clustering by the surface area of the union (bvh_ploc.cu); node boxes come out of the merges
            root_ref = build_binary_ploc(st, n, box_lo.ptr, box_hi.ptr, sorted.ptr, left.ptr, right.ptr, range.ptr, nlo.ptr, nhi.ptr, s.ploc_radius, nullptr);
        } else if (n > 1 && s.builder == 1 && sah_builder_available()) {
            // binned SAH over the Morton-ordered sequence (bvh_sah.cu); node boxes come out of the sweep
            build_binary_sah(st, n, box_lo.ptr, box_hi.ptr, sorted.ptr, left.ptr, right.ptr, range.ptr, nlo.ptr, nhi.ptr);
        } else if (n > 1) {
            k_radix_tree<<<div_up(n - 1, 256), 256, 0, st>>>(keys_sorted.ptr, t);
            PB2_LAUNCH_CHECK();
            DevBuf<int> arrive(n);
            arrive.zero(st);
            k_refit<<<div_up(n, 256), 256, 0, st>>>(t, sorted.ptr, box_lo.ptr, box_hi.ptr, arrive.ptr);
            PB2_LAUNCH_CHECK();
        }
        PB2_CUDA(cudaStreamSynchronize(st)); // tmp / arrive / keys go out of scope
    }

    // ---- collapse ----
    DevBuf<Bvh8Node> nodes(n); // upper bound: one wide node per binary internal node (+ root)
    s.d_prims.alloc(n);
    DevBuf<uint32_t> counters(4);
    DevBuf<float> sah(1);
    sah.zero(st);
    uint32_t init_counters[4] = { 1, 0, 0, 0 };
    PB2_CUDA(cudaMemcpyAsync(counters.ptr, init_counters, sizeof init_counters, cudaMemcpyHostToDevice, st));
    DevBuf<uint4> qa(n), qb(n);
    uint4 root = make_uint4(0u, n > 1 ? (uint32_t)root_ref : (uint32_t)~0, 0u, 0u); // n == 1: leaf ref ~0
    PB2_CUDA(cudaMemcpyAsync(qa.ptr, &root, sizeof root, cudaMemcpyHostToDevice, st));
    DevBuf<uint32_t> dst_of_sorted(n);
    CollapseCtx cc{ t, sorted.ptr, box_lo.ptr, box_hi.ptr, prims_in.ptr, dst_of_sorted.ptr, nodes.ptr, counters.ptr, sah.ptr };
    uint32_t n_in = 1;
    uint4 *q_in = qa.ptr, *q_out = qb.ptr;
    uint32_t host_counters[4];
    while (n_in) {
        k_collapse<<<div_up(n_in, 128), 128, 0, st>>>(cc, q_in, n_in, q_out);
        PB2_LAUNCH_CHECK();
        PB2_CUDA(cudaMemcpyAsync(host_counters, counters.ptr, sizeof host_counters, cudaMemcpyDeviceToHost, st));
        PB2_CUDA(cudaStreamSynchronize(st));
        n_in = host_counters[2];
        PB2_CUDA(cudaMemsetAsync(counters.ptr + 2, 0, sizeof(uint32_t), st));
        std::swap(q_in, q_out);
    }
    s.n_nodes = host_counters[0], s.n_prims = host_counters[1];
    if (s.n_prims != n) throw std::runtime_error("pb2_bvh_build: collapse lost primitives");
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        k_scatter_prims<<<(unsigned)std::min<uint64_t>(div_up(n, 256), (uint64_t)sms * 8), 256, 0, st>>>(prims_in.ptr, sorted.ptr, dst_of_sorted.ptr, n, s.d_prims.ptr);
        PB2_LAUNCH_CHECK();
    }
    // shrink the node array to its final size
    s.d_nodes.alloc(s.n_nodes);
    PB2_CUDA(cudaMemcpyAsync(s.d_nodes.ptr, nodes.ptr, s.n_nodes * sizeof(Bvh8Node), cudaMemcpyDeviceToDevice, st));
    float sah_host = 0.f, root_area = 1.f;
    int hb[6];
    PB2_CUDA(cudaMemcpyAsync(&sah_host, sah.ptr, sizeof(float), cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaMemcpyAsync(hb, bounds.ptr, sizeof hb, cudaMemcpyDeviceToHost, st));
    PB2_CUDA(cudaEventRecord(e1, st));
    PB2_CUDA(cudaStreamSynchronize(st));
    {
        float d[3];
        for (int k = 0; k < 3; ++k) d[k] = ordered_to_float(hb[3 + k]) - ordered_to_float(hb[k]);
        root_area = d[0] * d[1] + d[1] * d[2] + d[2] * d[0];
    }
    float ms = 0.f;
    PB2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    s.build_stats.n_nodes = s.n_nodes;
    s.build_stats.bvh_bytes = (uint64_t)s.n_nodes * sizeof(Bvh8Node) + (uint64_t)s.n_prims * sizeof(PrimRec);
    s.build_stats.build_ms = ms;
    s.build_stats.sah_cost = root_area > 0.f ? sah_host / root_area : 0.f;
    s.build_stats.max_depth = host_counters[3];
    if (host_counters[3] > PB2_STACK_SIZE - 2) throw std::runtime_error("pb2_bvh_build: wide tree deeper than the traversal stack (" + std::to_string(host_counters[3]) + " levels)");
    s.bvh_valid = true;
    if (s.l2_persist_mb > 0) s.l2_dirty = true; // the node array moved
}
}// namespace pb2
