/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of PupilOptixLab's pt-with-MIS hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (pupiloptixlab_b200/) never includes, links or calls it.
 *
 * This header: plain-C POD structs shared by the two oracle libraries
 *   oracle/_build/liborc_port.so  — "port": everything restated in this directory
 *   oracle/_ref/liborc_ref.so     — "reference": the same driver, but RNG / BSDF /
 *                                   Fresnel / GGX / emitter / sampling arithmetic comes from
 *                                   the reference's own headers, compiled from /root/reference
 * and by the ctypes binding in tests/orc.py.
 *
 * Enum values follow the reference:
 *   material type  = Pupil::EMatType        (framework/render/material/predefine.h:15-22,
 *                                            framework/decl/material_decl.inl:3-9)
 *   texture type   = util::ETextureType     (framework/util/texture.h:21-25)
 *   emitter type   = optix::EEmitterType    (framework/render/emitter/types.h:7-15)
 *   lobe type      = optix::EBsdfLobeType   (framework/render/material/bsdf/bsdf.h:7-24)
 */
#ifndef ORC_TYPES_H
#define ORC_TYPES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MAT_UNKNOWN = 0, ORC_MAT_DIFFUSE = 1, ORC_MAT_DIELECTRIC = 2, ORC_MAT_ROUGH_DIELECTRIC = 3,
       ORC_MAT_CONDUCTOR = 4, ORC_MAT_ROUGH_CONDUCTOR = 5, ORC_MAT_PLASTIC = 6, ORC_MAT_ROUGH_PLASTIC = 7,
       ORC_MAT_TWOSIDED = 8 };
enum { ORC_TEX_RGB = 0, ORC_TEX_BITMAP = 1, ORC_TEX_CHECKERBOARD = 2 };
enum { ORC_EMIT_NONE = 0, ORC_EMIT_TRI = 1, ORC_EMIT_SPHERE = 2, ORC_EMIT_CONST_ENV = 3, ORC_EMIT_ENV_MAP = 4 };
enum { ORC_SHAPE_OBJ = 1, ORC_SHAPE_SPHERE = 2, ORC_SHAPE_CUBE = 3, ORC_SHAPE_RECTANGLE = 4 }; /* resource/shape.h:28-31 */
enum { ORC_XF_IDENTITY = 0, ORC_XF_MATRIX16 = 1, ORC_XF_MATRIX9 = 2, ORC_XF_LOOKAT = 3, ORC_XF_SRT = 4 };

/* util::Texture (framework/util/texture.h:27-65); `to_uv` is the full row-major 4x4 of util::Transform (only rows
 * 0,1 are read on the device, cuda/texture.h:34-36).  A bitmap is float4 texels, row 0 first, borrowed from the
 * caller; address_mode / filter_mode = util::ETextureAddressMode / ETextureFilterMode (texture.h:10-20). */
typedef struct orc_texture {
    int32_t type;
    float a[3]; /* rgb  | checkerboard patch1 (= xml color0, resource/scene.cpp:170-172) */
    float b[3]; /*        checkerboard patch2 (= xml color1) */
    float to_uv[16];
    int32_t bitmap_w, bitmap_h, address_mode, filter_mode;
    const float *bitmap;
} orc_texture;

/* resource::Material (framework/resource/material.h:16-83) flattened: one struct, all slots. */
typedef struct orc_material {
    int32_t type;     /* EMatType of the nested bsdf */
    int32_t twosided; /* wrapped in <bsdf type="twosided"> */
    float int_ior, ext_ior;
    int32_t nonlinear;
    orc_texture alpha, eta, k, reflectance /* diffuse: reflectance; plastics: diffuse_reflectance */,
        specular_reflectance, specular_transmittance;
} orc_material;

/* <transform name="to_world"> as written in the XML, resolved by the oracle exactly like
 * resource/xml/util_loader.cpp:128-191 does. */
typedef struct orc_transform {
    int32_t kind;
    float m[16];      /* MATRIX16: 16 values; MATRIX9: first 9 */
    float origin[3], target[3], up[3]; /* LOOKAT */
    int32_t has_scale, has_rotate, has_translate; /* SRT */
    float scale[3], axis[3], angle, translate[3];
} orc_transform;

/* device-side evaluated BSDF (optix::material::*::Local) as one flat record */
typedef struct orc_local_bsdf {
    int32_t type;
    float alpha, eta, int_fdr, specular_sampling_weight;
    int32_t nonlinear;
    float eta3[3], k3[3];
    float reflectance[3]; /* diffuse.reflectance / plastic.diffuse_reflectance */
    float specular_reflectance[3], specular_transmittance[3];
} orc_local_bsdf;

typedef struct orc_bsdf_result {
    float wi[3], f[3], pdf;
    uint32_t sampled_type;
    uint32_t rng_after;
} orc_bsdf_result;

/* optix::Emitter (framework/render/emitter.h:13-23): tri area, sphere, constant env, env map */
typedef struct orc_emitter {
    int32_t type;
    float weight, select_probability;
    orc_texture radiance; /* const env: radiance.a = color */
    float area;
    float pos[3][3], nrm[3][3], uv[3][2]; /* TriArea v0..v2 (world space) */
    float center[3], radius;              /* Sphere */
    /* EnvMapEmitter (framework/render/emitter/env.h:6-22); tables borrowed from the owner (orc scene / caller):
     * row_cdf[map_h + 1 (+1 pad)], row_weight[map_h (+1 pad)], col_cdf[(map_w + 1) * (map_h (+1 pad row))] — the pad
     * entries repeat the last row: the reference reads one row past its tables when xi.x > row_cdf[map_h - 1]. */
    float scale, normalization;
    uint32_t map_w, map_h;
    float to_world[9], to_local[9];
    const float *row_cdf, *row_weight, *col_cdf;
} orc_emitter;

typedef struct orc_emit_sample {
    float radiance[3], wi[3], pos[3], normal[3], distance, pdf;
    int32_t is_delta;
} orc_emit_sample;

typedef struct orc_hit {
    float t, u, v;
    int32_t inst, prim; /* inst = -1: miss */
} orc_hit;

#ifdef __cplusplus
}
#endif
#endif
