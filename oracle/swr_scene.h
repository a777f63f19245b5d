/*
 * swr_scene.h -- TEST INFRASTRUCTURE ONLY (oracle side).
 *
 * One flat C description of "a draw call plus its stock shaders and render
 * targets", shared by the two CPU checkers in this directory:
 *
 *   oracle/swr_oracle.c   -- a from-scratch C restatement of the reference draw path
 *   oracle/ref_driver.cpp -- a thin driver around the UNMODIFIED reference sources
 *                            (compiled in place from /root/reference, output oracle/_ref/)
 *
 * Both export   int <prefix>_draw(swr_scene *s)   with identical semantics, so the
 * tests can diff them against each other and against the CUDA path.  Nothing
 * in the product (softwarerenderer_b200/, include/) includes or links this.
 *
 * Enum values follow the reference: DrawMode {Point, Line, Triangle}
 * (VertexProcessor.h:42-46), CullMode {None, CCW, CW} (VertexProcessor.h:49-53),
 * RasterMode {Span, Block, Adaptive} (Rasterizer.h:45-49).
 */
#ifndef SWR_SCENE_H
#define SWR_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SWR_B200_H   /* (tests/hostcheck includes both headers; the ids are identical) */
/* Stock vertex shaders (mirrored 1:1 by softwarerenderer_b200/csrc/stock_shaders.cuh). */
enum {
    SWR_VS_POS_COLOR = 0,     /* {x,y,z,r,g,b}: pos passthrough, w=1, 3 avars   (Benchmark.cpp:38-48)   */
    SWR_VS_MVP_COLOR = 1,     /* {x,y,z,r,g,b}: clip = mvp*(x,y,z,1), 3 avars                              */
    SWR_VS_MVP_NORMAL_UV = 2  /* {pos3,normal3,uv2}: clip = mvp*(pos,1), 3 avars=normal, 2 pvars=uv (Box.cpp:75-87) */
};

/* Stock pixel shaders. */
enum {
    SWR_PS_FLAT = 0,          /* A=3: color[i] = 1                                  (Benchmark.cpp:14-26)  */
    SWR_PS_COUNT_ID = 1,      /* count[i]++, prim_id[i] = emission ordinal (last writer wins)              */
    SWR_PS_GOURAUD = 2,       /* A=3: color[i] = r<<16|g<<8|b                    (RasterizerTest.cpp:37-47) */
    SWR_PS_GOURAUD_DEPTH = 3, /* Z, A=3: if (z < depth[i]) { depth[i]=z; color[i]=rgb; }                   */
    SWR_PS_VARY_DUMP = 4,     /* Z, W, A=3, P=2: vary[k][i] = z,w,invw,a0,a1,a2,p0,p1; count[i]++           */
    SWR_PS_TEXTURED = 5,      /* W, A=3, P=2: color[i] = texture[nearest(u,v) wrapped]                     */
    SWR_PS_TEXTURED_ANISO = 6 /* W, P=2: Box.cpp:39-62 verbatim: perspective derivatives + Texture::sample  */
};
#endif

/* Emission ordinal of a primitive: batch * SWR_ORDINAL_STRIDE + slot, where batch
 * counts the 1024-primitive flushes of VertexProcessor.cpp:110-116 and slot is the
 * primitive's position in that batch's output index list (original slots first,
 * clipper fan extras appended, VertexProcessor.cpp:252-261).  A clipped triangle is at
 * most a 9-gon in general position, but PolyClipper.cpp:64-74 duplicates vertices that lie
 * exactly on a plane, so both checkers and the CUDA path allow SWR_MAX_POLY = 12 vertices
 * (10 fan triangles); 10240 = 1024 * 10 is then the largest slot count of one batch. */
#ifndef SWR_MAX_POLY
#define SWR_MAX_POLY 12
#define SWR_ORDINAL_STRIDE 10240u
#endif
#define SWR_VARY_PLANES 8

typedef struct swr_scene {
    /* geometry (borrowed) */
    const void *vertices;
    int32_t stride;
    int32_t num_vertices;
    const int32_t *indices;
    int64_t index_count;

    /* fixed-function state */
    int32_t draw_mode, cull_mode, raster_mode;
    int32_t vp_x, vp_y, vp_w, vp_h;
    int32_t sc_x, sc_y, sc_w, sc_h;
    float depth_n, depth_f;

    /* stock shaders + uniforms */
    int32_t vs_kind, ps_kind;
    float mvp[16];               /* row-major: clip.x = m[0]*x + m[1]*y + m[2]*z + m[3]*w */
    const uint32_t *texture;     /* tex_w * tex_h texels, power-of-two sides */
    int32_t tex_w, tex_h;

    /* render targets (borrowed; any may be NULL if the pixel shader does not use it) */
    int32_t width, height;
    uint32_t *color;
    float *depth;
    uint32_t *count;
    uint32_t *prim_id;
    float *vary;                 /* SWR_VARY_PLANES planes of width*height floats */

    /* optional dump of the primitive stream handed to the rasterizer, in emission
     * order: SWR_STREAM_FLOATS floats per primitive = { ordinal (uint32 bits),
     * vertex count (uint32 bits), then x,y,z,w of up to 3 screen-space vertices, pad }.
     * Ignored when stream == NULL; stream_len counts every primitive even past cap. */
    float *stream;
    int64_t stream_cap;
    int64_t stream_len;

    /* results */
    uint64_t fragments;          /* drawPixel invocations */
    uint64_t primitives_out;     /* primitives handed to the rasterizer with index != -1 */
} swr_scene;

#define SWR_STREAM_FLOATS 16

#ifdef __cplusplus
}
#endif
#endif
