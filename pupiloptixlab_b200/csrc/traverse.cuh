// BVH8 traversal: the replacement for optixTrace (closest hit: example/path_tracer/main.cu:80-85,161-166;
// any hit with TERMINATE_ON_FIRST_HIT: framework/render/emitter.h:91-100).  No RT cores on B200, so
// this is a software traversal of the 80-byte compressed wide nodes built by bvh_build.cu:
//   * one 5 x 128-bit fetch per node, one 3 x 128-bit fetch per primitive,
//   * octant-ordered child visit without sorting (slot ^ ray-octant gives the visit priority),
//   * a compressed traversal stack of (node-group base, hit mask) pairs.
// Intersection arithmetic (OptiX-internal in the reference) is fp32 Moeller-Trumbore on world-space
// triangles and an analytic unit sphere in the instance's object space; a hit needs tmin < t < tmax.
// It is spelled out with round-to-nearest intrinsics (ix_* below) so that nvcc cannot re-associate or
// contract it: t, u, v depend only on this sequence, which a CPU checker can reproduce with fmaf().
#pragma once
#include "pb2_types.cuh"

namespace pb2 {

#ifndef PB2_STACK_SIZE
#define PB2_STACK_SIZE 32
#endif

struct RayHit {
    float t, u, v;
    uint32_t prim_slot; // index into SceneView::prims, 0xffffffff = miss
};

struct TraceCounters {
    uint32_t nodes, prims;
};

PB2_D uint32_t byte_of(uint32_t w, int i) { return (w >> (i * 8)) & 0xffu; }

// fixed-rounding building blocks (the CPU checker used by the tests spells out the same sequence with fmaf)
PB2_D float ix_dot(float3 a, float3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x))); }
PB2_D float3 ix_cross(float3 a, float3 b) {
    return mk3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
PB2_D float3 ix_sub(float3 a, float3 b) { return mk3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
PB2_D float ix_row_point(float4 r, float3 p) { return __fmaf_rn(r.z, p.z, __fmaf_rn(r.y, p.y, __fmaf_rn(r.x, p.x, r.w))); }
PB2_D float ix_row_vector(float4 r, float3 v) { return __fmaf_rn(r.z, v.z, __fmaf_rn(r.y, v.y, __fmul_rn(r.x, v.x))); }
PB2_D float3 ix_point(float4 r0, float4 r1, float4 r2, float3 p) { return mk3(ix_row_point(r0, p), ix_row_point(r1, p), ix_row_point(r2, p)); }
PB2_D float3 ix_vector(float4 r0, float4 r1, float4 r2, float3 v) { return mk3(ix_row_vector(r0, v), ix_row_vector(r1, v), ix_row_vector(r2, v)); }

// Tests one primitive record.  Returns true when it is hit closer than `hit.t`.
PB2_D bool intersect_prim(const SceneView &sv, uint32_t slot, float3 o, float3 d, float tmin, RayHit &hit) {
    const float4 *rec = reinterpret_cast<const float4 *>(sv.prims + slot);
    const float4 a = __ldg(rec), b = __ldg(rec + 1), c = __ldg(rec + 2);
    if (__float_as_uint(c.w) == 0u) {
        const float3 v0 = mk3(a), e1 = mk3(b), e2 = mk3(c);
        const float3 pvec = ix_cross(d, e2);
        const float det = ix_dot(e1, pvec);
        if (det == 0.f) return false;
        const float inv = __fdiv_rn(1.f, det);
        const float3 tvec = ix_sub(o, v0);
        const float u = __fmul_rn(ix_dot(tvec, pvec), inv);
        if (u < 0.f || u > 1.f) return false;
        const float3 qvec = ix_cross(tvec, e1);
        const float v = __fmul_rn(ix_dot(d, qvec), inv);
        if (v < 0.f || __fadd_rn(u, v) > 1.f) return false;
        const float t = __fmul_rn(ix_dot(e2, qvec), inv);
        if (!(t > tmin && t < hit.t)) return false;
        hit.t = t, hit.u = u, hit.v = v, hit.prim_slot = slot;
        return true;
    } else {
        const DevInstance *in = sv.instances + __float_as_uint(b.w);
        const float4 r0 = __ldg(&in->inv[0]), r1 = __ldg(&in->inv[1]), r2 = __ldg(&in->inv[2]);
        const float3 oo = ix_point(r0, r1, r2, o), dd = ix_vector(r0, r1, r2, d);
        const float qa = ix_dot(dd, dd), qb = ix_dot(oo, dd), qc = __fadd_rn(ix_dot(oo, oo), -1.f);
        const float disc = __fmaf_rn(qb, qb, -__fmul_rn(qa, qc));
        if (!(disc >= 0.f) || qa == 0.f) return false;
        const float sq = __fsqrt_rn(disc);
        const float t0 = __fdiv_rn(__fsub_rn(-qb, sq), qa), t1 = __fdiv_rn(__fadd_rn(-qb, sq), qa);
        float t;
        if (t0 > tmin && t0 < hit.t) t = t0;
        else if (t1 > tmin && t1 < hit.t) t = t1;
        else return false;
        hit.t = t, hit.u = 0.f, hit.v = 0.f, hit.prim_slot = slot;
        return true;
    }
}

// Closest hit (ANY = false) or first hit (ANY = true).  `hit.t` must come in as tmax.
template<bool ANY, bool COUNT>
PB2_D bool traverse(const SceneView &sv, float3 o, float3 d, float tmin, RayHit &hit, TraceCounters *ctr) {
    hit.prim_slot = 0xffffffffu;
    if (sv.n_nodes == 0) return false;
    auto safe_inv = [](float x) { return 1.f / (fabsf(x) > 1e-30f ? x : copysignf(1e-30f, x)); };
    const float3 idir = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    // octant: bit set <=> direction component >= 0.  Children were placed so that slot ^ oct is
    // larger for nearer children; the highest set bit of the hit mask is visited first.
    const bool px = idir.x >= 0.f, py = idir.y >= 0.f, pz = idir.z >= 0.f; // from idir so that -0.0 stays consistent
    const uint32_t oct = (px ? 4u : 0u) | (py ? 2u : 0u) | (pz ? 1u : 0u);
    const uint32_t oct4 = oct * 0x01010101u;

    uint2 stack[PB2_STACK_SIZE];
    int sp = 0;
    uint2 G = make_uint2(0u, 0x80000000u); // (child base, hit bits 31..24 | imask 7..0): the root as a group of one

    for (;;) {
        // ---- pop the nearest pending internal child of G ----
        const uint32_t bit = 31u - __clz(G.y);
        G.y &= ~(1u << bit);
        const uint32_t slot = (bit - 24u) ^ oct;
        const uint32_t rel = __popc(G.y & 0xffu & ((1u << slot) - 1u));
        const uint32_t node_idx = G.x + rel;
        if (G.y & 0xff000000u) stack[sp++] = G;

        const Bvh8Node *np = sv.nodes + node_idx;
        const float4 n0 = __ldg(&np->n0);
        const uint4 n1 = __ldg(&np->n1), n2 = __ldg(&np->n2), n3 = __ldg(&np->n3), n4 = __ldg(&np->n4);
        if (COUNT) ++ctr->nodes;

        const uint32_t ebits = __float_as_uint(n0.w);
        const float sx = __uint_as_float((ebits & 0xffu) << 23), sy = __uint_as_float(((ebits >> 8) & 0xffu) << 23),
                    sz = __uint_as_float(((ebits >> 16) & 0xffu) << 23);
        const float3 adj = mk3(sx * idir.x, sy * idir.y, sz * idir.z);
        const float3 org = mk3((n0.x - o.x) * idir.x, (n0.y - o.y) * idir.y, (n0.z - o.z) * idir.z);
        // far planes are pushed out by a few ulps so rounding can never cull a box the ray touches
        constexpr float kFar = 1.0000004f;
        const float3 adj_f = adj * kFar, org_f = org * kFar;

        uint32_t hitmask = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t meta4 = half ? n1.w : n1.z;
            const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
            const uint32_t inner_mask4 = (is_inner4 >> 4) * 0xffu;
            const uint32_t bit_index4 = (meta4 ^ (oct4 & inner_mask4)) & 0x1f1f1f1fu;
            const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
            const uint32_t qlx = half ? n2.y : n2.x, qly = half ? n2.w : n2.z, qlz = half ? n3.y : n3.x;
            const uint32_t qhx = half ? n3.w : n3.z, qhy = half ? n4.y : n4.x, qhz = half ? n4.w : n4.z;
            // near / far plane bytes per axis depend only on the ray's sign
            const uint32_t nx = px ? qlx : qhx, fx = px ? qhx : qlx;
            const uint32_t ny = py ? qly : qhy, fy = py ? qhy : qly;
            const uint32_t nz = pz ? qlz : qhz, fz = pz ? qhz : qlz;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float tnx = (float)byte_of(nx, j) * adj.x + org.x, tfx = (float)byte_of(fx, j) * adj_f.x + org_f.x;
                const float tny = (float)byte_of(ny, j) * adj.y + org.y, tfy = (float)byte_of(fy, j) * adj_f.y + org_f.y;
                const float tnz = (float)byte_of(nz, j) * adj.z + org.z, tfz = (float)byte_of(fz, j) * adj_f.z + org_f.z;
                const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
                const float tf = fminf(fminf(tfx, tfy), fminf(tfz, hit.t));
                if (tn <= tf) hitmask |= byte_of(child_bits4, j) << byte_of(bit_index4, j);
            }
        }
        G = make_uint2(n1.x, (hitmask & 0xff000000u) | (ebits >> 24));
        uint32_t T = hitmask & 0x00ffffffu;

        // ---- primitives of the hit leaf slots ----
        while (T) {
            const uint32_t i = __ffs(T) - 1;
            T &= T - 1;
            if (COUNT) ++ctr->prims;
            if (intersect_prim(sv, n1.y + i, o, d, tmin, hit)) {
                if (ANY) return true;
            }
        }
        // ---- next group ----
        if (!(G.y & 0xff000000u)) {
            if (sp == 0) break;
            G = stack[--sp];
        }
    }
    return hit.prim_slot != 0xffffffffu;
}
}// namespace pb2
