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

#ifndef PB2_PRIM_BRANCHLESS
#define PB2_PRIM_BRANCHLESS 1
#endif
#ifndef PB2_APPROX_IDIR
#define PB2_APPROX_IDIR 1
#endif
// quantised plane bytes -> float through PRMT + the FMA that follows instead of I2F.U8 (node_step)
#ifndef PB2_NODE_PRMT
#define PB2_NODE_PRMT 1
#endif
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
PB2_D void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
PB2_D void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
// BRANCHLESS (triangles): the four early-outs folded into one predicate — see below; used where all 32 lanes test primitives.
// TRIS: the scene has no analytic spheres (pb2_build_stats::n_spheres == 0), the record kind is not looked at and the sphere
// branch is compiled out: +1.7 % Msamples/s on the Cornell box, +3.9 % on the 30 M-triangle terrain (profiles/README.md).
template<bool BRANCHLESS = false, bool TRIS = false>
PB2_D bool intersect_prim(const SceneView &sv, uint32_t slot, float3 o, float3 d, float tmin, RayHit &hit) {
    const float4 *rec = reinterpret_cast<const float4 *>(sv.prims + slot);
    const float4 a = __ldg(rec), b = __ldg(rec + 1), c = __ldg(rec + 2);
    if (TRIS || __float_as_uint(c.w) == 0u) {
        const float3 v0 = mk3(a), e1 = mk3(b), e2 = mk3(c);
        const float3 pvec = ix_cross(d, e2);
        const float det = ix_dot(e1, pvec);
        if (BRANCHLESS && PB2_PRIM_BRANCHLESS) {
            // Same arithmetic and the same accept / reject decisions as the early-out form below (each test is its literal
            // negation, so NaNs fall the same way), evaluated without branches: in the warp-cooperative loop all 32 lanes hold a
            // primitive and some lane survives every early-out anyway.  Measured (profiles/README.md): +1 % on the terrain; in
            // the per-lane loop of small scenes (12 of 32 lanes active) it gains nothing, so that loop keeps the early-outs.
            const float inv = __frcp_rn(det);
            const float3 tvec = ix_sub(o, v0);
            const float u = __fmul_rn(ix_dot(tvec, pvec), inv);
            const float3 qvec = ix_cross(tvec, e1);
            const float v = __fmul_rn(ix_dot(d, qvec), inv);
            const float t = __fmul_rn(ix_dot(e2, qvec), inv);
            const bool ok = (det != 0.f) & !(u < 0.f || u > 1.f) & !(v < 0.f || __fadd_rn(u, v) > 1.f) & (t > tmin && t < hit.t);
            if (ok) hit.t = t, hit.u = u, hit.v = v, hit.prim_slot = slot;
            return ok;
        }
        if (det == 0.f) return false;
        const float inv = __frcp_rn(det); // correctly rounded 1 / det, the same value as __fdiv_rn(1.f, det) in fewer instructions
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

// ---- persistent, dynamically refilled traversal ------------------------------------------------------------
// A plain one-ray-per-thread loop leaves a warp waiting for its slowest ray (ncu, 30 M triangles, incoherent rays:
// 5.8 of 32 threads active per instruction).  Here every lane owns a traversal state that survives across
// rays and the warp advances in explicit lock step: per iteration every busy lane pops and tests ONE wide node
// and then the primitives of the leaf slots it hit; finished lanes commit their result, and as soon as fewer
// than `refill_threshold` lanes are still busy the idle lanes take the next rays from a global work counter
// (one atomic per warp) — "persistent threads with replacement" (Aila & Laine 2009), with the difference
// that the warp reconverges after every node step.  Measured on the 30 M-triangle terrain (profiles/):
// letting lanes run ahead for 2 / 4 / 8 / 16 node steps between reconvergence points is 8 % / 22 % / 33 % /
// 34 % slower than reconverging every step.
//
// IO supplies the rays and takes the results:
//   uint32_t size() const;
//   uint32_t load(uint32_t i, float3 &o, float3 &d, float &tmin, float &tmax) const;   returns a token (e.g. the path slot)
//        that is handed back to commit()
//   void commit(bool valid, uint32_t token, const RayHit &h, bool hit, uint32_t inst);
//        called by ALL 32 lanes of the warp, converged, whenever at least one lane finished a ray
//        (valid = this lane did), so implementations may use full-mask warp collectives.  inst: the instance of the hit when it
//        lies in a bottom-level tree (two-level scenes), else ~0u = the instance id stored in the primitive record
#ifndef PB2_REFILL_THRESHOLD
#define PB2_REFILL_THRESHOLD 26
#endif

struct RayState {
    float3 o, d, idir;
    float tmin;
    RayHit hit;
    uint2 G;           // current node group: (child base, hit bits 31..24 | imask 7..0)
    uint32_t T;        // pending primitive bits of the current node
    uint32_t prim_base;
    uint32_t oct;
    int sp;
    // two-level scenes only (INST): the instance whose bottom-level tree the ray is in (~0u at the top level) and the instance
    // of the nearest hit so far (~0u: take it from the primitive record — top-level primitives carry their instance id)
    uint32_t cur_inst, hit_inst;
};

// reciprocal direction and octant of r.d (again after an instance transform)
// direction components below this magnitude are treated as this magnitude by the slab tests: 1 / d stays far enough from the
// top of the fp32 range for the products of node_step (32768 * scale / d) in any scene smaller than 1e15 units
constexpr float kMinDir = 1e-18f;
PB2_D void ray_set_dir(RayState &r, float3 d) {
#if PB2_APPROX_IDIR
    // 1 / d feeds the slab tests only (never t, u, v), which are conservative by construction: the one-instruction hardware
    // reciprocal (<= 1 ulp) is enough, and kFar below carries the extra ulp.  ncu charged the three correctly rounded
    // reciprocals 9 % of k_extend's instructions and 10 % of its stall samples on the Cornell box.
    auto safe_inv = [](float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(x) > kMinDir ? x : copysignf(kMinDir, x)));
        return r;
    };
#else
    auto safe_inv = [](float x) { return __frcp_rn(fabsf(x) > kMinDir ? x : copysignf(kMinDir, x)); }; // IEEE whatever the compile flags say
#endif
    r.d = d;
    r.idir = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    r.oct = (r.idir.x >= 0.f ? 4u : 0u) | (r.idir.y >= 0.f ? 2u : 0u) | (r.idir.z >= 0.f ? 1u : 0u);
}
PB2_D void ray_begin(RayState &r, float3 o, float3 d, float tmin, float tmax, bool empty_scene, uint32_t root = 0u) {
    r.o = o, r.tmin = tmin;
    ray_set_dir(r, d);
    r.hit.t = tmax, r.hit.u = r.hit.v = 0.f, r.hit.prim_slot = 0xffffffffu;
    r.G = make_uint2(root, empty_scene ? 0u : 0x80000000u);
    r.T = 0u, r.prim_base = 0u, r.sp = 0;
    r.cur_inst = r.hit_inst = 0xffffffffu;
}
PB2_D bool ray_has_nodes(const RayState &r) { return (r.G.y & 0xff000000u) != 0u || r.sp > 0; }

// pops the nearest pending internal node, tests its eight children, leaves the hit children in G / T
// Traversal stack of one lane: the first kSmemStack entries live in shared memory (entry-major, so the 32 lanes of a warp hit
// 32 different banks), the rest in local memory.  ncu on the Cornell box counted 1.6 GB of DRAM traffic per k_extend launch
// against 0.9 GB of ray / hit records: the difference was the local-memory stack (one 8-byte push per ray allocates a sector).
#ifndef PB2_SMEM_STACK
#define PB2_SMEM_STACK 8
#endif
constexpr int kSmemStack = PB2_SMEM_STACK;
struct TravStack {
    uint2 *sm; // this thread's column of the shared part (stride: CTA size)
    uint2 *lo; // local-memory part
    PB2_D void push(int &sp, uint2 v) const {
        if (sp < kSmemStack) sm[sp * 128] = v;
        else lo[sp - kSmemStack] = v;
        ++sp;
    }
    PB2_D uint2 pop(int &sp) const {
        --sp;
        return sp < kSmemStack ? sm[sp * 128] : lo[sp - kSmemStack];
    }
};
// Two-level scenes (INST): a top-level child may be an INSTANCE NODE — an 80-byte slot in the node array whose exponent / mask
// word is 0xffffffff (no real node has an exponent byte of 255), holding the root of a bottom-level tree (n1.x) and the
// instance that places it in the world (n1.y).  Popping it takes the ray into that instance's object space: o' = M^-1 o,
// d' = M^-1 d (not renormalised, so t keeps its meaning), a sentinel goes on the stack, and the world-space ray waits in shared
// memory until the sentinel is popped again.  This is the reference's IAS -> GAS step (ias_manager.cpp:29-114): one tree per
// shared mesh, any number of placements.
constexpr uint32_t kInstanceNodeTag = 0xffffffffu, kStackSentinel = 0xffffffffu;
struct WorldRay { // this lane's column of the saved world-space rays (stride: CTA size)
    float *p;
    PB2_D void save(float3 o, float3 d) const { p[0] = o.x, p[128] = o.y, p[256] = o.z, p[384] = d.x, p[512] = d.y, p[640] = d.z; }
    PB2_D void load(float3 &o, float3 &d) const { o = mk3(p[0], p[128], p[256]), d = mk3(p[384], p[512], p[640]); }
};
// pops the nearest pending internal node, tests its eight children, leaves the hit children in G / T; returns false when no
// wide node was tested (INST: the pop ended a bottom-level tree with nothing left above it, or entered one)
template<bool INST = false>
PB2_D bool node_step(const SceneView &sv, RayState &r, const TravStack &stack, const WorldRay &world = WorldRay{ nullptr }) {
    if (!(r.G.y & 0xff000000u)) {
        r.G = stack.pop(r.sp);
        if (INST && r.G.x == kStackSentinel) { // the bottom-level tree is exhausted: back to the world-space ray
            float3 o, d;
            world.load(o, d);
            r.o = o;
            ray_set_dir(r, d);
            r.cur_inst = 0xffffffffu;
            r.G.y = 0u, r.T = 0u;
            if (r.sp <= 0) return false;
            r.G = stack.pop(r.sp);
        }
    }
    const uint32_t bit = 31u - __clz(r.G.y);
    r.G.y &= ~(1u << bit);
    const uint32_t slot = (bit - 24u) ^ r.oct;
    const uint32_t rel = __popc(r.G.y & 0xffu & ((1u << slot) - 1u));
    const Bvh8Node *np = sv.nodes + (r.G.x + rel);
    if (r.G.y & 0xff000000u) stack.push(r.sp, r.G);

    const float4 n0 = __ldg(&np->n0);
    const uint4 n1 = __ldg(&np->n1);
    const uint32_t ebits = __float_as_uint(n0.w);
    if (INST && ebits == kInstanceNodeTag) {
        const DevInstance *in = sv.instances + n1.y;
        const float4 r0 = __ldg(&in->inv[0]), r1 = __ldg(&in->inv[1]), r2 = __ldg(&in->inv[2]);
        world.save(r.o, r.d);
        r.o = ix_point(r0, r1, r2, r.o);
        ray_set_dir(r, ix_vector(r0, r1, r2, r.d));
        r.cur_inst = n1.y;
        stack.push(r.sp, make_uint2(kStackSentinel, 0u));
        r.G = make_uint2(n1.x, 0x80000000u);
        r.T = 0u;
        return false;
    }
    const uint4 n2 = __ldg(&np->n2), n3 = __ldg(&np->n3), n4 = __ldg(&np->n4);
    const bool px = r.idir.x >= 0.f, py = r.idir.y >= 0.f, pz = r.idir.z >= 0.f;
    const uint32_t oct4 = r.oct * 0x01010101u;
    const float sx = __uint_as_float((ebits & 0xffu) << 23), sy = __uint_as_float(((ebits >> 8) & 0xffu) << 23),
                sz = __uint_as_float(((ebits >> 16) & 0xffu) << 23);
    const float3 adj = mk3(sx * r.idir.x, sy * r.idir.y, sz * r.idir.z);
    const float3 org = mk3((n0.x - r.o.x) * r.idir.x, (n0.y - r.o.y) * r.idir.y, (n0.z - r.o.z) * r.idir.z);
    constexpr float kFar = PB2_APPROX_IDIR ? 1.0000007f : 1.0000004f; // far planes pushed out by a few ulps: rounding can never cull a touched box
    const float3 adj_f = adj * kFar, org_f = org * kFar;
#if PB2_NODE_PRMT
    // Plane byte b -> float without the conversion unit: PRMT drops b into mantissa bits 8..15 of 2^15, which reads 32768 + b,
    // and the bias moves into the addend of the multiply-add that follows: (32768 + b) adj + (org - 32768 adj).  The addend is
    // rounded DOWN for near planes and UP for far planes, so the box only ever grows (by at most 2^-8 of a quantisation step).
    // I2F.U8 runs on the quarter-rate XU pipe (ncu: the busiest pipe of the trace kernels, 48 conversions per wide node).
    const float3 org_n = mk3(__fmaf_rd(-32768.f, adj.x, org.x), __fmaf_rd(-32768.f, adj.y, org.y), __fmaf_rd(-32768.f, adj.z, org.z));
    const float3 org_ff = mk3(__fmaf_ru(-32768.f, adj_f.x, org_f.x), __fmaf_ru(-32768.f, adj_f.y, org_f.y), __fmaf_ru(-32768.f, adj_f.z, org_f.z));
    const uint32_t bias = sv.plane_bias; // 2^15, from the constant bank (SceneView)
#define PB2_PLANE(w, j) __uint_as_float(__byte_perm((w), bias, 0x7504u | ((j) << 4)))
#else
    const float3 org_n = org, org_ff = org_f;
#define PB2_PLANE(w, j) ((float)byte_of((w), (j)))
#endif
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
        const uint32_t nx = px ? qlx : qhx, fx = px ? qhx : qlx;
        const uint32_t ny = py ? qly : qhy, fy = py ? qhy : qly;
        const uint32_t nz = pz ? qlz : qhz, fz = pz ? qhz : qlz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float tnx = PB2_PLANE(nx, j) * adj.x + org_n.x, tfx = PB2_PLANE(fx, j) * adj_f.x + org_ff.x;
            const float tny = PB2_PLANE(ny, j) * adj.y + org_n.y, tfy = PB2_PLANE(fy, j) * adj_f.y + org_ff.y;
            const float tnz = PB2_PLANE(nz, j) * adj.z + org_n.z, tfz = PB2_PLANE(fz, j) * adj_f.z + org_ff.z;
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, r.hit.t));
            if (tn <= tf) hitmask |= byte_of(child_bits4, j) << byte_of(bit_index4, j);
        }
    }
#undef PB2_PLANE
    r.G = make_uint2(n1.x, (hitmask & 0xff000000u) | (ebits >> 24));
    r.T = hitmask & 0x00ffffffu;
    r.prim_base = n1.y;
    return true;
}

// resident 128-thread CTAs per SM the trace kernels are compiled for: 9 (56 registers) with the per-lane primitive loop,
// 8 (64 registers) with the warp-cooperative one
#ifndef PB2_TRACE_MINB_PLAIN
#define PB2_TRACE_MINB_PLAIN 9
#endif
#ifndef PB2_TRACE_MINB_COOP
#define PB2_TRACE_MINB_COOP 8
#endif
#define PB2_TRACE_MINB(COOP) ((COOP) ? PB2_TRACE_MINB_COOP : PB2_TRACE_MINB_PLAIN)
constexpr int kPairsPerLane = 8;   // primitives one lane contributes per cooperative round
constexpr int kTraceWarps = 4;     // the trace kernels run 128-thread CTAs
struct CoopShared {                // per warp
    float4 o[32], d[32];           // ray origin | tmin, direction | current hit.t of each lane
    unsigned long long best[32];   // (t bits << 32) | pair index of the nearest hit found this round
    uint32_t base[32];             // first primitive slot of each lane's current node
    uint16_t pairs[32 * kPairsPerLane]; // (owner lane << 5) | leaf bit
};

// COOP = false: every lane tests the primitives of its own leaf slots one after the other.
// COOP = true : the (ray, primitive) pairs of the whole warp are tested 32 at a time (see below).  Measured on B200
// (profiles/README.md): +7 % / +10 % Mrays/s for incoherent closest-hit / any-hit rays on the 30 M-triangle terrain, where
// few lanes reach a leaf per step; -17 % on the 36-triangle Cornell box, where every lane does and the bookkeeping only
// adds instructions.  The host picks per scene (Scene::coop_prims).
// INST: the scene has bottom-level trees behind instance nodes (Scene::n_blas > 0); compiled out otherwise.
template<bool ANY, bool COUNT, bool COOP, bool TRIS, bool INST = false, class IO>
PB2_D void trace_persistent(const SceneView &sv, IO &io, uint32_t *__restrict__ work_counter, TraceCounters *ctr, int refill_threshold = PB2_REFILL_THRESHOLD) {
    constexpr uint32_t kFull = 0xffffffffu;
    __shared__ float s_world[INST ? 6 * 128 : 1];
    const WorldRay world{ s_world + (INST ? threadIdx.x : 0) };
    __shared__ CoopShared s_coop[COOP ? kTraceWarps : 1];
    CoopShared &sm = s_coop[COOP ? threadIdx.x >> 5 : 0];
    const uint32_t n = io.size();
    const uint32_t lane = threadIdx.x & 31u;
    __shared__ uint2 s_stack[(kSmemStack > 0 ? kSmemStack : 1) * 128]; // 128-thread CTAs (launch bounds of the trace kernels)
    uint2 stack_local[PB2_STACK_SIZE - kSmemStack];
    const TravStack stack{ s_stack + threadIdx.x, stack_local };
    RayState r;
    r.T = 0u, r.G = make_uint2(0u, 0u), r.sp = 0;
    uint32_t ray = 0;
    bool busy = false;
    bool exhausted = false; // warp-uniform: the work counter has run past the end
    for (;;) {              // every lane of the warp stays in this loop until the whole warp is done
        // ---- refill: when fewer than `refill_threshold` lanes are busy, idle lanes take the next rays ----
        const uint32_t busy_mask = __ballot_sync(kFull, busy);
        if (!exhausted && __popc(busy_mask) < refill_threshold) {
            const uint32_t idle = ~busy_mask;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(work_counter, (uint32_t)__popc(idle));
            base = __shfl_sync(kFull, base, 0);
            const uint32_t mine = base + __popc(idle & ((1u << lane) - 1u));
            if (!busy && mine < n) {
                float3 o, d;
                float tmin, tmax;
                ray = io.load(mine, o, d, tmin, tmax);
                ray_begin(r, o, d, tmin, tmax, sv.n_nodes == 0, sv.root);
                busy = true;
            }
            exhausted = base + __popc(idle) >= n;
        } else if (busy_mask == 0u) {
            break;
        }
        // ---- one node step, then the primitives it uncovered ----
        if (busy && ray_has_nodes(r)) {
            const bool tested = node_step<INST>(sv, r, stack, world);
            if (COUNT && tested) ++ctr->nodes;
            while (!COOP && r.T) {
                const uint32_t i = __ffs(r.T) - 1;
                r.T &= r.T - 1;
                if (COUNT) ++ctr->prims;
                if (intersect_prim<false, TRIS>(sv, r.prim_base + i, r.o, r.d, r.tmin, r.hit)) {
                    if (INST) r.hit_inst = r.cur_inst;
                    if (ANY) r.T = 0u, r.G.y = 0u, r.sp = 0;
                }
            }
        }
        // ---- primitives, warp-cooperatively ----
        // Leaf slots hit by a node step hold 0..24 primitives and most lanes hold none, so a per-lane loop runs at the
        // length of the longest list (ncu, Cornell box: the primitive tests are half of all instructions at 12.6 of 32
        // lanes).  Here the (ray, primitive) pairs of the whole warp are laid out in shared memory in owner order and
        // tested 32 at a time, one pair per lane whatever ray it belongs to; the nearest hit per ray is agreed on with a
        // 64-bit atomicMin on (t bits, pair index) and handed back to the owning lane by shuffle.
        while (COOP) {
            const uint32_t T = busy ? r.T : 0u;
            if (!__any_sync(kFull, T != 0u)) break;
            const uint32_t cnt = min((uint32_t)__popc(T), (uint32_t)kPairsPerLane);
            uint32_t incl = cnt;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) {
                const uint32_t up = __shfl_up_sync(kFull, incl, dlt);
                if ((int)lane >= dlt) incl += up;
            }
            const uint32_t total = __shfl_sync(kFull, incl, 31), excl = incl - cnt;
            sm.o[lane] = make_float4(r.o.x, r.o.y, r.o.z, r.tmin);
            sm.d[lane] = make_float4(r.d.x, r.d.y, r.d.z, r.hit.t);
            sm.base[lane] = r.prim_base;
            sm.best[lane] = ~0ull;
            {
                uint32_t tt = T;
                for (uint32_t j = 0; j < cnt; ++j) {
                    const uint32_t b = __ffs(tt) - 1;
                    tt &= tt - 1;
                    sm.pairs[excl + j] = (uint16_t)((lane << 5) | b);
                }
                if (busy) r.T = tt; // primitives beyond the per-round cap wait for the next round
            }
            __syncwarp();
            for (uint32_t c = 0; c < total; c += 32u) {
                const uint32_t k = c + lane;
                RayHit h;
                h.t = 0.f, h.u = 0.f, h.v = 0.f, h.prim_slot = 0xffffffffu;
                if (k < total) {
                    const uint32_t e = sm.pairs[k], own = e >> 5;
                    if (!ANY || sm.best[own] == ~0ull) {
                        const float4 ro = sm.o[own], rd = sm.d[own];
                        h.t = rd.w;
                        if (COUNT) ++ctr->prims;
                        if (intersect_prim<true, TRIS>(sv, sm.base[own] + (e & 31u), mk3(ro), mk3(rd), ro.w, h))
                            atomicMin(&sm.best[own], ((unsigned long long)__float_as_uint(h.t) << 32) | k); // t > 0: bit order = value order
                    }
                }
                __syncwarp();
                // the owner of a ray whose best pair sits in this chunk fetches the hit from the lane that tested it
                const unsigned long long b = sm.best[lane];
                const uint32_t bk = (uint32_t)b;
                const bool won = busy && b != ~0ull && bk >= c && bk < c + 32u;
                const uint32_t src = won ? bk - c : lane;
                const float wt = __shfl_sync(kFull, h.t, src), wu = __shfl_sync(kFull, h.u, src), wv = __shfl_sync(kFull, h.v, src);
                const uint32_t ws = __shfl_sync(kFull, h.prim_slot, src);
                if (won) {
                    r.hit.t = wt, r.hit.u = wu, r.hit.v = wv, r.hit.prim_slot = ws;
                    if (INST) r.hit_inst = r.cur_inst; // the pairs of this round belong to the tree the owner is in right now
                    if (ANY) r.T = 0u, r.G.y = 0u, r.sp = 0;
                }
            }
            __syncwarp();
        }
        __syncwarp();
        const bool done = busy && !ray_has_nodes(r);
        if (__any_sync(kFull, done)) {
            io.commit(done, ray, r.hit, r.hit.prim_slot != 0xffffffffu, INST ? r.hit_inst : 0xffffffffu);
            if (done) busy = false;
        }
    }
}
}// namespace pb2
