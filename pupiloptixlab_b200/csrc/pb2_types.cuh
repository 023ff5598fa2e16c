// Device-side data layout of the pb2 back end (see DESIGN.md "Data layout in HBM").
#pragma once
#include "../../include/pb2.h"
#include "vecmath.cuh"

namespace pb2 {

// ---- textures / materials / emitters -------------------------------------------------------------
// 64 B, float4-aligned.  hdr = (type bits, a.x, a.y, a.z): a constant-RGB texture costs one 16 B load.
struct DevTexture {
    float4 hdr; // x = __int_as_float(type), yzw = rgb | patch1
    float4 b;   // xyz = patch2
    float4 r0, r1;
};
// 288 B: one per instance (the SBT hit-record payload of the reference, example/path_tracer/type.h:35-39)
struct DevMaterial {
    int32_t type, twosided;
    float eta;
    int32_t nonlinear;
    float int_fdr, specular_sampling_weight;
    int32_t pad0, pad1;
    DevTexture tex[4];
};
// 192 B
struct DevEmitter {
    int32_t type;
    float weight, select_probability, area;
    DevTexture radiance;
    float4 p0, p1, p2; // TriArea world positions; w = uv0.x, uv0.y, uv1.x
    float4 n0, n1, n2; // TriArea normals;         w = uv1.y, uv2.x, uv2.y
    float4 center_r;   // Sphere centre.xyz, radius
};

// ---- instances --------------------------------------------------------------------------------------
#define PB2_IF_SPHERE 0x100u
#define PB2_IF_HAS_NRM 0x200u
#define PB2_IF_HAS_UV 0x400u
#define PB2_IF_TWOSIDED 0x800u
// 144 B: RenderObject geometry + transform (framework/world/render_object.h:11-38)
struct DevInstance {
    float4 xf[3];  // object -> world rows
    float4 inv[3]; // world -> object rows
    const float *pos, *nrm, *uv; // object-space vertex attributes of the shared mesh (nrm / uv may be null)
    const uint32_t *idx;         // 3 per triangle
    uint32_t flags;              // PB2_INST_* | PB2_IF_*
    int32_t mat_type;            // EMatType (queue sort key)
    int32_t emitter_offset;      // -1: not an emitter
    uint32_t n_tris;             // 1 for a sphere
};
static_assert(sizeof(DevInstance) == 144, "DevInstance layout");

// ---- BVH8 ---------------------------------------------------------------------------------------------
// 80-byte compressed wide node (after Ylitie, Karras, Laine 2017), read as five 16-byte words:
//   n0: origin.xyz (fp32), [ex, ey, ez, imask] bytes
//   n1: child_base_idx, prim_base_idx, meta[0..3], meta[4..7]
//   n2: qlo_x[8] qlo_y[8]   n3: qlo_z[8] qhi_x[8]   n4: qhi_y[8] qhi_z[8]
// meta[i]: 0 empty | internal: 0b001<<5 | (24+i) | leaf: unary(count 1..3)<<5 | first prim offset (0..23)
struct Bvh8Node {
    float4 n0;
    uint4 n1, n2, n3, n4;
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

// 48-byte primitive record, world space, in BVH leaf order.
//   triangle: v0.xyz | bits(prim id),  e1.xyz | bits(instance id),  e2.xyz | 0
//   sphere  : -      | 0            ,  -      | bits(instance id),  -      | 1     (analytic unit sphere of that instance)
struct PrimRec {
    float4 v0, e1, e2;
};
static_assert(sizeof(PrimRec) == 48, "PrimRec must be 48 bytes");

struct Camera {
    float4 s2c[4];
    float4 c2w[4];
};

// everything a kernel needs to know about the scene, passed by value
struct SceneView {
    const Bvh8Node *nodes;
    const PrimRec *prims;
    const DevInstance *instances;
    const DevMaterial *materials;
    const DevEmitter *areas;
    const DevEmitter *env; // nullptr: no environment emitter
    const float *area_cdf; // n_areas running fp32 sums of select_probability, in table order (select_emitter)
    uint32_t n_areas;
    uint32_t n_nodes;
    uint32_t n_prims;
    uint32_t root; // node the traversal starts at: the top-level tree's root (it sits behind the bottom-level trees in `nodes`)
    uint32_t plane_bias; // 0x47000000 (2^15): operand of the PRMT that turns plane bytes into floats (node_step).  A kernel parameter
                         // on purpose: PRMT takes ONE immediate, and a literal here makes the compiler route the selectors through registers
};
}// namespace pb2
