/*
 * swr_oracle.c -- TEST INFRASTRUCTURE ONLY.  The parity oracle ("port").
 *
 * A from-scratch, single-threaded C restatement of the reference draw path
 *   VertexProcessor::drawElements -> clip / transform / cull -> TriangleEquations ->
 *   Block / Span / Adaptive triangles, DDA lines, points -> PixelShader::drawPixel
 * written as per-primitive pure functions (the shape a GPU thread evaluates), not as the
 * reference's index-list mutation.  Every function cites the reference file:line it follows.
 *
 * PINNED: tests/test_oracle.py checks this file against (a) the known answers of SURVEY.md
 * section 4 / tests/golden/known_answers.json, which were produced by the unmodified reference
 * build (oracle/ref_driver.cpp + /root/reference sources -> oracle/_ref/libswr_ref.so), and
 * (b) that build itself, buffer for buffer, whenever oracle/_ref/libswr_ref.so is present.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (oracle/Makefile).  fp32 everywhere, IEEE
 * division, no FMA contraction, operations in the reference's order.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "swr_scene.h"

#define MAX_AVARS 16            /* IRasterizer.h:36 */
#define MAX_PVARS 16            /* IRasterizer.h:39 */
#define BLOCK 8                 /* IRasterizer.h:33 */
#define BATCH 1024              /* VertexProcessor.cpp:110 */

typedef struct {                /* IRasterizer.h:42-53 */
    float x, y, z, w;
    float avar[MAX_AVARS];
    float pvar[MAX_PVARS];
} vtx;

typedef struct { float a, b, c; int tie; } edge_eq;      /* EdgeEquation.h:31-35 */
typedef struct { float a, b, c; } plane_eq;              /* ParameterEquation.h:31-34 */

typedef struct {                /* TriangleEquations.h:35-45 */
    float area2;
    edge_eq e[3];
    plane_eq z, invw, avar[MAX_AVARS], pvar[MAX_PVARS];
} tri_eq;

static float plane_eval_fwd(const plane_eq *p, float x, float y) { return p->a * x + p->b * y + p->c; }

typedef struct {                /* PixelData.h:38-58 (what drawPixel may read) */
    int x, y;
    float z, w, invw;
    float avar[MAX_AVARS], pvar[MAX_PVARS], ptmp[MAX_PVARS];
} frag;

typedef struct {                /* compile-time traits of the stock shaders */
    int nA_vs, nP_vs;           /* VertexShader::AVarCount / PVarCount */
    int nA, nP, useZ, useW;     /* PixelShader::AVarCount / PVarCount / InterpolateZ / InterpolateW */
} traits;

typedef struct {
    swr_scene *s;
    traits t;
    float px, py, ox, oy;       /* VertexProcessor.cpp:50-53 */
    int minX, minY, maxX, maxY; /* Rasterizer.h:81-87 (max exclusive) */
    uint32_t ordinal;
    const tri_eq *cur_eq;       /* PixelData::equations of the fragment being shaded */
    /* mip chain of the textured_aniso shader (Texture.h:220-293) */
    uint32_t *mip[16];
    int mip_w[16], mip_h[16], mip_levels, max_aniso;
} ctx;

/* ------------------------------------------------------------------ stock shaders */
static void vs_mvp(const float *m, float x, float y, float z, vtx *o)
{
    const float w = 1.0f;
    o->x = m[0] * x + m[1] * y + m[2] * z + m[3] * w;
    o->y = m[4] * x + m[5] * y + m[6] * z + m[7] * w;
    o->z = m[8] * x + m[9] * y + m[10] * z + m[11] * w;
    o->w = m[12] * x + m[13] * y + m[14] * z + m[15] * w;
}

/* VertexProcessor.cpp:134-150 (attribute fetch) + the stock processVertex bodies. */
static void shade_vertex(const ctx *c, int index, vtx *o)
{
    const swr_scene *s = c->s;
    const float *d = (const float *)((const char *)s->vertices + (size_t)s->stride * (size_t)index);
    switch (s->vs_kind) {
    case SWR_VS_POS_COLOR:
        o->x = d[0]; o->y = d[1]; o->z = d[2]; o->w = 1.0f;
        o->avar[0] = d[3]; o->avar[1] = d[4]; o->avar[2] = d[5];
        break;
    case SWR_VS_MVP_COLOR:
        vs_mvp(s->mvp, d[0], d[1], d[2], o);
        o->avar[0] = d[3]; o->avar[1] = d[4]; o->avar[2] = d[5];
        break;
    default: /* SWR_VS_MVP_NORMAL_UV */
        vs_mvp(s->mvp, d[0], d[1], d[2], o);
        o->avar[0] = d[3]; o->avar[1] = d[4]; o->avar[2] = d[5];
        o->pvar[0] = d[6]; o->pvar[1] = d[7];
        break;
    }
}

static uint32_t pack_rgb(const frag *p)  /* RasterizerTest.cpp:39-45 */
{
    int r = (int)(p->avar[0] * 255);
    int g = (int)(p->avar[1] * 255);
    int b = (int)(p->avar[2] * 255);
    return (uint32_t)(r << 16 | g << 8 | b);
}

/* ------------------------------------------------------------------ Texture.h restated */
static int tex_r(uint32_t c) { return (int)((c >> 16) & 0xFF); }
static int tex_g(uint32_t c) { return (int)((c >> 8) & 0xFF); }
static int tex_b(uint32_t c) { return (int)(c & 0xFF); }
static uint32_t tex_pack(int r, int g, int b) { return ((uint32_t)(r & 0xFF) << 16) | ((uint32_t)(g & 0xFF) << 8) | (uint32_t)(b & 0xFF); }
static int tex_u8(float v) { return (int)v & 0xFF; }      /* (Uint8)float the way x86 code does it */

/* Texture.h:220-293: 2x2 box filter down to 1x1 */
static int tex_build_mips(ctx *c, const uint32_t *base, int w, int h)
{
    c->mip_levels = 0;
    uint32_t *lv = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)w * (size_t)h);
    if (!lv) return -4;
    memcpy(lv, base, sizeof(uint32_t) * (size_t)w * (size_t)h);
    c->mip[0] = lv; c->mip_w[0] = w; c->mip_h[0] = h; c->mip_levels = 1;
    while ((w > 1 || h > 1) && c->mip_levels < 16) {
        int nw = w / 2 > 1 ? w / 2 : 1, nh = h / 2 > 1 ? h / 2 : 1;
        const uint32_t *src = c->mip[c->mip_levels - 1];
        uint32_t *dst = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)nw * (size_t)nh);
        if (!dst) return -4;
        for (int y = 0; y < nh; y++)
            for (int x = 0; x < nw; x++) {
                uint32_t p00 = src[(y * 2) * w + (x * 2)];
                uint32_t p10 = x * 2 + 1 < w ? src[(y * 2) * w + (x * 2 + 1)] : p00;
                uint32_t p01 = y * 2 + 1 < h ? src[(y * 2 + 1) * w + (x * 2)] : p00;
                uint32_t p11 = (x * 2 + 1 < w && y * 2 + 1 < h) ? src[(y * 2 + 1) * w + (x * 2 + 1)] : p00;
                dst[y * nw + x] = tex_pack((tex_r(p00) + tex_r(p10) + tex_r(p01) + tex_r(p11)) >> 2,
                                           (tex_g(p00) + tex_g(p10) + tex_g(p01) + tex_g(p11)) >> 2,
                                           (tex_b(p00) + tex_b(p10) + tex_b(p01) + tex_b(p11)) >> 2);
            }
        c->mip[c->mip_levels] = dst; c->mip_w[c->mip_levels] = nw; c->mip_h[c->mip_levels] = nh;
        c->mip_levels++;
        w = nw; h = nh;
    }
    return 0;
}

static void tex_free_mips(ctx *c)
{
    for (int i = 0; i < c->mip_levels; ++i) free(c->mip[i]);
    c->mip_levels = 0;
}

static uint32_t tex_lerp(uint32_t c1, uint32_t c2, float t)  /* Texture.h:191-196 */
{
    return tex_pack(tex_r(c1) + tex_u8((tex_r(c2) - tex_r(c1)) * t),
                    tex_g(c1) + tex_u8((tex_g(c2) - tex_g(c1)) * t),
                    tex_b(c1) + tex_u8((tex_b(c2) - tex_b(c1)) * t));
}

static int tex_bilerp_channel(int c00, int c10, int c01, int c11, float fx, float fy)  /* Texture.h:199-204 */
{
    return tex_u8(c00 * (1 - fx) * (1 - fy) + c10 * fx * (1 - fy) + c01 * (1 - fx) * fy + c11 * fx * fy);
}

static uint32_t tex_bilinear(const ctx *c, int mip, float u, float v)  /* Texture.h:145-180 */
{
    if (mip < 0 || mip >= c->mip_levels) return 0;
    int w = c->mip_w[mip], h = c->mip_h[mip];
    const uint32_t *px = c->mip[mip];
    float fpx = u * (w - 1), fpy = v * (h - 1);
    int x0 = (int)floorf(fpx); if (x0 < 0) x0 = 0;
    int y0 = (int)floorf(fpy); if (y0 < 0) y0 = 0;
    int x1 = x0 + 1 < w - 1 ? x0 + 1 : w - 1;
    int y1 = y0 + 1 < h - 1 ? y0 + 1 : h - 1;
    float fx = fpx - x0, fy = fpy - y0;
    uint32_t c00 = px[y0 * w + x0], c10 = px[y0 * w + x1], c01 = px[y1 * w + x0], c11 = px[y1 * w + x1];
    return tex_pack(tex_bilerp_channel(tex_r(c00), tex_r(c10), tex_r(c01), tex_r(c11), fx, fy),
                    tex_bilerp_channel(tex_g(c00), tex_g(c10), tex_g(c01), tex_g(c11), fx, fy),
                    tex_bilerp_channel(tex_b(c00), tex_b(c10), tex_b(c01), tex_b(c11), fx, fy));
}

static uint32_t tex_trilinear(const ctx *c, float u, float v, float rho)  /* Texture.h:123-143 */
{
    float lod = log2f(rho > 1e-6f ? rho : 1e-6f);
    float top = (float)(c->mip_levels - 1);
    lod = lod < 0.0f ? 0.0f : (top < lod ? top : lod);
    int lodBase = (int)floorf(lod);
    int lodNext = lodBase + 1 < c->mip_levels - 1 ? lodBase + 1 : c->mip_levels - 1;
    float lodFrac = lod - lodBase;
    lodFrac = lodFrac < 0.0f ? 0.0f : (1.0f < lodFrac ? 1.0f : lodFrac);
    uint32_t cb = tex_bilinear(c, lodBase, u, v);
    if (lodBase != lodNext) return tex_lerp(cb, tex_bilinear(c, lodNext, u, v), lodFrac);
    return cb;
}

static float tex_wrap(float x) { x = fmodf(x, 1.0f); if (x < 0) x += 1.0f; return x; }  /* Texture.h:41-44 */

static uint32_t tex_sample(const ctx *c, float u, float v, float dudx, float dvdx, float dudy, float dvdy)  /* Texture.h:35-118 */
{
    if (c->mip_levels <= 0) return 0;
    u = tex_wrap(u);
    v = tex_wrap(v);
    float dudx_s = dudx * c->mip_w[0], dvdx_s = dvdx * c->mip_h[0];
    float dudy_s = dudy * c->mip_w[0], dvdy_s = dvdy * c->mip_h[0];
    float dx_len = sqrtf(dudx_s * dudx_s + dvdx_s * dvdx_s);
    float dy_len = sqrtf(dudy_s * dudy_s + dvdy_s * dvdy_s);
    dx_len = dx_len < 1e-6f ? 1e-6f : dx_len;
    dy_len = dy_len < 1e-6f ? 1e-6f : dy_len;
    float major_len = dx_len < dy_len ? dy_len : dx_len;
    float minor_len = dy_len < dx_len ? dy_len : dx_len;
    float ratio = major_len / minor_len;
    ratio = (float)c->max_aniso < ratio ? (float)c->max_aniso : ratio;
    int num = (int)ceilf(ratio);
    if (num < 1) num = 1;
    if (num <= 1) return tex_trilinear(c, u, v, major_len);
    float mdu, mdv;
    if (dx_len > dy_len) { mdu = dudx / dx_len; mdv = dvdx / dx_len; }
    else { mdu = dudy / dy_len; mdv = dvdy / dy_len; }
    float r = 0, g = 0, b = 0;
    float step = 1.0f / num;
    for (int i = 0; i < num; ++i) {
        float t = (i + 0.5f) * step - 0.5f;
        float su = tex_wrap(u + mdu * major_len * t);
        float sv = tex_wrap(v + mdv * major_len * t);
        uint32_t sc = tex_trilinear(c, su, sv, minor_len);
        r += tex_r(sc); g += tex_g(sc); b += tex_b(sc);
    }
    r = r / num; r = 255.0f < r ? 255.0f : r;
    g = g / num; g = 255.0f < g ? 255.0f : g;
    b = b / num; b = 255.0f < b ? 255.0f : b;
    return tex_pack(tex_u8(r), tex_u8(g), tex_u8(b));
}

/* PixelData::computePerspectiveDerivatives (PixelData.h:128-145) */
static void persp_derivs(const tri_eq *q, int var, int x, int y, float *ddx, float *ddy)
{
    float val = plane_eval_fwd(&q->pvar[var], x + 0.5f, y + 0.5f);
    float iw = plane_eval_fwd(&q->invw, x + 0.5f, y + 0.5f);
    float dvar_dx = q->pvar[var].a, dvar_dy = q->pvar[var].b;
    float dinvw_dx = q->invw.a, dinvw_dy = q->invw.b;
    *ddx = (iw * dvar_dx - val * dinvw_dx) / (iw * iw);
    *ddy = (iw * dvar_dy - val * dinvw_dy) / (iw * iw);
}

/* The stock drawPixel bodies (same as ref_driver.cpp / stock_shaders.cuh). */
static void draw_pixel(ctx *c, const frag *p)
{
    swr_scene *s = c->s;
    size_t n = (size_t)s->width * (size_t)s->height;
    size_t i = (size_t)(p->x + s->width * p->y);
    s->fragments++;
    switch (s->ps_kind) {
    case SWR_PS_FLAT:
        s->color[i] = 1;
        break;
    case SWR_PS_COUNT_ID:
        s->count[i]++;
        s->prim_id[i] = c->ordinal;
        break;
    case SWR_PS_GOURAUD:
        s->color[i] = pack_rgb(p);
        break;
    case SWR_PS_GOURAUD_DEPTH:
        if (p->z < s->depth[i]) {
            s->depth[i] = p->z;
            s->color[i] = pack_rgb(p);
        }
        break;
    case SWR_PS_VARY_DUMP:
        s->vary[0 * n + i] = p->z;
        s->vary[1 * n + i] = p->w;
        s->vary[2 * n + i] = p->invw;
        s->vary[3 * n + i] = p->avar[0];
        s->vary[4 * n + i] = p->avar[1];
        s->vary[5 * n + i] = p->avar[2];
        s->vary[6 * n + i] = p->pvar[0];
        s->vary[7 * n + i] = p->pvar[1];
        s->count[i]++;
        break;
    case SWR_PS_TEXTURED: {
        int tx = (int)floorf(p->pvar[0] * (float)s->tex_w) & (s->tex_w - 1);
        int ty = (int)floorf(p->pvar[1] * (float)s->tex_h) & (s->tex_h - 1);
        s->color[i] = s->texture[ty * s->tex_w + tx];
        break;
    }
    default: { /* SWR_PS_TEXTURED_ANISO: Box.cpp:49-61 */
        float dudx, dudy, dvdx, dvdy;
        persp_derivs(c->cur_eq, 0, p->x, p->y, &dudx, &dudy);
        persp_derivs(c->cur_eq, 1, p->x, p->y, &dvdx, &dvdy);
        s->color[i] = tex_sample(c, p->pvar[0], p->pvar[1], dudx, dvdx, dudy, dvdy);
        break;
    }
    }
}

static int set_traits(int vs, int ps, traits *t)
{
    memset(t, 0, sizeof(*t));
    switch (vs) {
    case SWR_VS_POS_COLOR: case SWR_VS_MVP_COLOR: t->nA_vs = 3; break;
    case SWR_VS_MVP_NORMAL_UV: t->nA_vs = 3; t->nP_vs = 2; break;
    default: return -1;
    }
    switch (ps) {
    case SWR_PS_FLAT: t->nA = 3; break;
    case SWR_PS_COUNT_ID: break;
    case SWR_PS_GOURAUD: t->nA = 3; break;
    case SWR_PS_GOURAUD_DEPTH: t->nA = 3; t->useZ = 1; break;
    case SWR_PS_VARY_DUMP: t->nA = 3; t->nP = 2; t->useZ = 1; t->useW = 1; break;
    case SWR_PS_TEXTURED: t->nA = 3; t->nP = 2; t->useW = 1; break;
    case SWR_PS_TEXTURED_ANISO: t->nP = 2; t->useW = 1; break;
    default: return -2;
    }
    return 0;
}

/* ------------------------------------------------------------------ clipping */
/* VertexProcessor.cpp:122-132: strict '<' on w-x, x+w, ... */
static int outcode(const vtx *v)
{
    int m = 0;
    if (v->w - v->x < 0) m |= 0x01;
    if (v->x + v->w < 0) m |= 0x02;
    if (v->w - v->y < 0) m |= 0x04;
    if (v->y + v->w < 0) m |= 0x08;
    if (v->w - v->z < 0) m |= 0x10;
    if (v->z + v->w < 0) m |= 0x20;
    return m;
}

/* The six planes in the fixed order +X,-X,+Y,-Y,+Z,-Z (VertexProcessor.cpp:187-192,237-242). */
static const float PLANES[6][4] = {
    { -1, 0, 0, 1 }, { 1, 0, 0, 1 }, { 0, -1, 0, 1 }, { 0, 1, 0, 1 }, { 0, 0, -1, 1 }, { 0, 0, 1, 1 },
};

/* PolyClipper.cpp:56,62 / LineClipper.cpp:35-36: a*x + b*y + c*z + d*w, left to right. */
static float plane_dist(const float *p, const vtx *v)
{
    return p[0] * v->x + p[1] * v->y + p[2] * v->z + p[3] * v->w;
}

/* PolyClipper.h:34-48: v0*(1-t) + v1*t over x,y,z,w and the VERTEX shader's var counts. */
static void lerp_vtx(vtx *o, const vtx *a, const vtx *b, float t, int nA, int nP)
{
    float s = 1.0f - t;
    o->x = a->x * s + b->x * t;
    o->y = a->y * s + b->y * t;
    o->z = a->z * s + b->z * t;
    o->w = a->w * s + b->w * t;
    for (int i = 0; i < nA; ++i) o->avar[i] = a->avar[i] * s + b->avar[i] * t;
    for (int i = 0; i < nP; ++i) o->pvar[i] = a->pvar[i] * s + b->pvar[i] * t;
}

static int sgn(float v) { return (0.0f < v) - (v < 0.0f); }  /* PolyClipper.h:91-94 */

/* One Sutherland-Hodgman pass (PolyClipper.cpp:45-81).  Returns the new vertex count, or -1
 * if the polygon would exceed SWR_MAX_POLY (parity is declared undefined there). */
static int clip_poly_plane(const vtx *in, int n, vtx *out, const float *plane, int nA, int nP)
{
    int m = 0;
    const vtx *prev = &in[0];
    float dprev = plane_dist(plane, prev);
    for (int i = 1; i <= n; ++i) {
        const vtx *cur = &in[i == n ? 0 : i];
        float d = plane_dist(plane, cur);
        if (dprev >= 0) {
            if (m >= SWR_MAX_POLY) return -1;
            out[m++] = *prev;
        }
        if (sgn(d) != sgn(dprev)) {
            float t = d < 0 ? dprev / (dprev - d) : -dprev / (d - dprev);
            if (m >= SWR_MAX_POLY) return -1;
            lerp_vtx(&out[m++], prev, cur, t, nA, nP);
        }
        prev = cur;
        dprev = d;
    }
    return m;
}

/* VertexProcessor.cpp:347-377: perspective divide, viewport (y flipped), depth range. */
static void to_screen(const ctx *c, vtx *v)
{
    float invw = 1.0f / v->w;
    v->x *= invw;
    v->y *= invw;
    v->z *= invw;
    v->x = (c->px * v->x + c->ox);
    v->y = (c->py * -v->y + c->oy);
    v->z = 0.5f * (c->s->depth_f - c->s->depth_n) * v->z + 0.5f * (c->s->depth_n + c->s->depth_f);
}

/* ------------------------------------------------------------------ triangle setup */
/* EdgeEquation.h:37-43 */
static void edge_init(edge_eq *e, const vtx *v0, const vtx *v1)
{
    e->a = v0->y - v1->y;
    e->b = v1->x - v0->x;
    e->c = -(e->a * (v0->x + v1->x) + e->b * (v0->y + v1->y)) / 2;
    e->tie = e->a != 0 ? e->a > 0 : e->b > 0;
}

static int edge_test(const edge_eq *e, float v) { return v > 0 || (v == 0 && e->tie); }  /* EdgeEquation.h:58-61 */

/* ParameterEquation.h:36-48 */
static void plane_init(plane_eq *p, float p0, float p1, float p2, const edge_eq *e, float factor)
{
    p->a = factor * (p0 * e[0].a + p1 * e[1].a + p2 * e[2].a);
    p->b = factor * (p0 * e[0].b + p1 * e[1].b + p2 * e[2].b);
    p->c = factor * (p0 * e[0].c + p1 * e[1].c + p2 * e[2].c);
}

static float plane_eval(const plane_eq *p, float x, float y) { return p->a * x + p->b * y + p->c; }

/* TriangleEquations.h:47-71.  Returns 0 when area2 <= 0 (triangle rejected). */
static int tri_setup(tri_eq *q, const vtx *v0, const vtx *v1, const vtx *v2, int nA, int nP)
{
    edge_init(&q->e[0], v1, v2);
    edge_init(&q->e[1], v2, v0);
    edge_init(&q->e[2], v0, v1);
    q->area2 = q->e[0].c + q->e[1].c + q->e[2].c;
    if (q->area2 <= 0) return 0;
    float factor = 1.0f / q->area2;
    plane_init(&q->z, v0->z, v1->z, v2->z, q->e, factor);
    float i0 = 1.0f / v0->w, i1 = 1.0f / v1->w, i2 = 1.0f / v2->w;
    plane_init(&q->invw, i0, i1, i2, q->e, factor);
    for (int i = 0; i < nA; ++i) plane_init(&q->avar[i], v0->avar[i], v1->avar[i], v2->avar[i], q->e, factor);
    for (int i = 0; i < nP; ++i) plane_init(&q->pvar[i], v0->pvar[i] * i0, v1->pvar[i] * i1, v2->pvar[i] * i2, q->e, factor);
    return 1;
}

/* ------------------------------------------------------------------ fragment state */
/* PixelData.h:61-82 */
static void frag_init(frag *p, const tri_eq *q, float x, float y, const traits *t)
{
    if (t->useZ) p->z = plane_eval(&q->z, x, y);
    if (t->useW || t->nP > 0) {
        p->invw = plane_eval(&q->invw, x, y);
        p->w = 1.0f / p->invw;
    }
    for (int i = 0; i < t->nA; ++i) p->avar[i] = plane_eval(&q->avar[i], x, y);
    for (int i = 0; i < t->nP; ++i) {
        p->ptmp[i] = plane_eval(&q->pvar[i], x, y);
        p->pvar[i] = p->ptmp[i] * p->w;
    }
}

/* PixelData.h:84-125: one step along x (dir 0, adds plane.a) or y (dir 1, adds plane.b). */
static void frag_step(frag *p, const tri_eq *q, int dir, const traits *t)
{
    if (t->useZ) p->z = p->z + (dir ? q->z.b : q->z.a);
    if (t->useW || t->nP > 0) {
        p->invw = p->invw + (dir ? q->invw.b : q->invw.a);
        p->w = 1.0f / p->invw;
    }
    for (int i = 0; i < t->nA; ++i) p->avar[i] = p->avar[i] + (dir ? q->avar[i].b : q->avar[i].a);
    for (int i = 0; i < t->nP; ++i) {
        p->ptmp[i] = p->ptmp[i] + (dir ? q->pvar[i].b : q->pvar[i].a);
        p->pvar[i] = p->ptmp[i] * p->w;
    }
}

static float fmin3(float a, float b, float c) { float m = b < a ? b : a; return c < m ? c : m; }
static float fmax3(float a, float b, float c) { float m = a < b ? b : a; return m < c ? c : m; }

/* ------------------------------------------------------------------ Block mode */
/* PixelShaderBase.h:55-94: 64-pixel walk with incremental edges and varyings. */
static void block_draw(ctx *c, const tri_eq *q, int x, int y, int test_edges)
{
    const traits *t = &c->t;
    float xf = x + 0.5f, yf = y + 0.5f;
    frag row;
    c->cur_eq = q;
    float er[3] = { 0, 0, 0 };
    frag_init(&row, q, xf, yf, t);
    if (test_edges)
        for (int k = 0; k < 3; ++k) er[k] = q->e[k].a * xf + q->e[k].b * yf + q->e[k].c;

    for (int yy = y; yy < y + BLOCK; ++yy) {
        frag p = row;
        float ev[3] = { er[0], er[1], er[2] };
        for (int xx = x; xx < x + BLOCK; ++xx) {
            if (!test_edges || (edge_test(&q->e[0], ev[0]) && edge_test(&q->e[1], ev[1]) && edge_test(&q->e[2], ev[2]))) {
                p.x = xx;
                p.y = yy;
                draw_pixel(c, &p);
            }
            frag_step(&p, q, 0, t);
            if (test_edges)
                for (int k = 0; k < 3; ++k) ev[k] = ev[k] + q->e[k].a;
        }
        frag_step(&row, q, 1, t);
        if (test_edges)
            for (int k = 0; k < 3; ++k) er[k] = er[k] + q->e[k].b;
    }
}

/* Rasterizer.h:224-306 */
static void tri_block(ctx *c, const vtx *v0, const vtx *v1, const vtx *v2)
{
    tri_eq q;
    if (!tri_setup(&q, v0, v1, v2, c->t.nA, c->t.nP)) return;

    int minX = (int)fmin3(v0->x, v1->x, v2->x), maxX = (int)fmax3(v0->x, v1->x, v2->x);
    int minY = (int)fmin3(v0->y, v1->y, v2->y), maxY = (int)fmax3(v0->y, v1->y, v2->y);
    if (minX < c->minX) minX = c->minX;
    if (maxX > c->maxX) maxX = c->maxX;
    if (minY < c->minY) minY = c->minY;
    if (maxY > c->maxY) maxY = c->maxY;
    minX &= ~(BLOCK - 1); maxX &= ~(BLOCK - 1);
    minY &= ~(BLOCK - 1); maxY &= ~(BLOCK - 1);

    const float s = BLOCK - 1;
    int stepsX = (maxX - minX) / BLOCK + 1;
    int stepsY = (maxY - minY) / BLOCK + 1;

    for (int i = 0; i < stepsX * stepsY; ++i) {
        int x = minX + (i % stepsX) * BLOCK;
        int y = minY + (i / stepsX) * BLOCK;
        float xf = x + 0.5f, yf = y + 0.5f;
        int in[4][3], all = 0, odd = 1;
        for (int k = 0; k < 3; ++k) {
            const edge_eq *e = &q.e[k];
            float e00 = e->a * xf + e->b * yf + e->c;   /* EdgeData.h:37-42 */
            float e01 = e00 + e->b * s;                 /* stepY(s) */
            float e10 = e00 + e->a * s;                 /* stepX(s) */
            float e11 = e01 + e->a * s;
            in[0][k] = edge_test(e, e00); in[1][k] = edge_test(e, e01);
            in[2][k] = edge_test(e, e10); in[3][k] = edge_test(e, e11);
        }
        for (int j = 0; j < 4; ++j) {
            all += in[j][0] && in[j][1] && in[j][2];
            /* Rasterizer.h:287-290: C++ chained '==', i.e. (t0 == t1) == t2 */
            odd = odd && ((in[j][0] == in[j][1]) == in[j][2]);
        }
        if (all == 0) {
            if (!odd) block_draw(c, &q, x, y, 1);
        } else if (all == 4) {
            block_draw(c, &q, x, y, 0);
        } else {
            block_draw(c, &q, x, y, 1);
        }
    }
}

/* ------------------------------------------------------------------ Span mode */
/* PixelShaderBase.h:96-112 */
static void span_draw(ctx *c, const tri_eq *q, int x, int y, int x2)
{
    frag p;
    p.y = y;
    c->cur_eq = q;
    frag_init(&p, q, x + 0.5f, y + 0.5f, &c->t);
    while (x < x2) {
        p.x = x;
        draw_pixel(c, &p);
        frag_step(&p, q, 0, &c->t);
        x++;
    }
}

/* Rasterizer.h:360-385 (top vertex v0, flat bottom v1 (left) - v2 (right)) */
static void span_bottom_flat(ctx *c, const tri_eq *q, float v0x, float v0y, float v1x, float v1y, float v2x, float v2y)
{
    float inv1 = (v1x - v0x) / (v1y - v0y);
    float inv2 = (v2x - v0x) / (v2y - v0y);
    for (int sy = (int)(v0y + 0.5f); sy < (int)(v1y + 0.5f); sy++) {
        float dy = (sy - v0y) + 0.5f;
        float cx1 = v0x + inv1 * dy + 0.5f;
        float cx2 = v0x + inv2 * dy + 0.5f;
        int xl = (int)cx1, xr = (int)cx2;
        if (xl < c->minX) xl = c->minX;
        if (xr > c->maxX) xr = c->maxX;
        span_draw(c, q, xl, sy, xr);
    }
}

/* Rasterizer.h:387-411 (flat top v0 (left) - v1 (right), bottom vertex v2) */
static void span_top_flat(ctx *c, const tri_eq *q, float v0x, float v0y, float v1x, float v1y, float v2x, float v2y)
{
    float inv1 = (v2x - v0x) / (v2y - v0y);
    float inv2 = (v2x - v1x) / (v2y - v1y);
    for (int sy = (int)(v2y - 0.5f); sy > (int)(v0y - 0.5f); sy--) {
        float dy = (sy - v2y) + 0.5f;
        float cx1 = v2x + inv1 * dy + 0.5f;
        float cx2 = v2x + inv2 * dy + 0.5f;
        int xl = (int)cx1, xr = (int)cx2;
        if (xl < c->minX) xl = c->minX;
        if (xr > c->maxX) xr = c->maxX;
        span_draw(c, q, xl, sy, xr);
    }
}

/* Rasterizer.h:308-358 */
static void tri_span(ctx *c, const vtx *v0, const vtx *v1, const vtx *v2)
{
    tri_eq q;
    if (!tri_setup(&q, v0, v1, v2, c->t.nA, c->t.nP)) return;

    const vtx *t = v0, *m = v1, *b = v2, *tmp;
    if (t->y > m->y) { tmp = t; t = m; m = tmp; }
    if (m->y > b->y) { tmp = m; m = b; b = tmp; }
    if (t->y > m->y) { tmp = t; t = m; m = tmp; }

    float dy = (b->y - t->y);
    float iy = (m->y - t->y);

    if (m->y == t->y) {
        const vtx *l = m, *r = t;
        if (l->x > r->x) { tmp = l; l = r; r = tmp; }
        span_top_flat(c, &q, l->x, l->y, r->x, r->y, b->x, b->y);
    } else if (m->y == b->y) {
        const vtx *l = m, *r = b;
        if (l->x > r->x) { tmp = l; l = r; r = tmp; }
        span_bottom_flat(c, &q, t->x, t->y, l->x, l->y, r->x, r->y);
    } else {
        float v4y = m->y;
        float v4x = t->x + ((b->x - t->x) / dy) * iy;
        float lx = m->x, ly = m->y, rx = v4x, ry = v4y;
        if (lx > rx) { float f; f = lx; lx = rx; rx = f; f = ly; ly = ry; ry = f; }
        span_bottom_flat(c, &q, t->x, t->y, lx, ly, rx, ry);
        span_top_flat(c, &q, lx, ly, rx, ry, b->x, b->y);
    }
}

/* Rasterizer.h:413-428 (the literals 0.4 / 1.6 are doubles) */
static void tri_adaptive(ctx *c, const vtx *v0, const vtx *v1, const vtx *v2)
{
    float minX = fmin3(v0->x, v1->x, v2->x), maxX = fmax3(v0->x, v1->x, v2->x);
    float minY = fmin3(v0->y, v1->y, v2->y), maxY = fmax3(v0->y, v1->y, v2->y);
    float orient = (maxX - minX) / (maxY - minY);
    if (orient > 0.4 && orient < 1.6) tri_block(c, v0, v1, v2);
    else tri_span(c, v0, v1, v2);
}

static void draw_triangle(ctx *c, const vtx *v0, const vtx *v1, const vtx *v2)
{
    switch (c->s->raster_mode) {   /* Rasterizer.h:430-445 */
    case 0: tri_span(c, v0, v1, v2); break;
    case 1: tri_block(c, v0, v1, v2); break;
    default: tri_adaptive(c, v0, v1, v2); break;
    }
}

/* ------------------------------------------------------------------ lines and points */
static int scissor_test(const ctx *c, float x, float y)  /* Rasterizer.h:144-147 */
{
    return x >= c->minX && x < c->maxX && y >= c->minY && y < c->maxY;
}

/* Rasterizer.h:160-173: pvars are passed un-divided, w only if InterpolateW. */
static void frag_from_vertex(const ctx *c, const vtx *v, frag *p)
{
    p->x = (int)v->x;
    p->y = (int)v->y;
    if (c->t.useZ) p->z = v->z;
    if (c->t.useW) { p->w = v->w; p->invw = 1.0f / v->w; }
    for (int i = 0; i < c->t.nA; ++i) p->avar[i] = v->avar[i];
    for (int i = 0; i < c->t.nP; ++i) p->pvar[i] = v->pvar[i];
}

static void draw_point(ctx *c, const vtx *v)  /* Rasterizer.h:149-158 */
{
    frag p;
    memset(&p, 0, sizeof(p));
    if (!scissor_test(c, v->x, v->y)) return;
    frag_from_vertex(c, v, &p);
    draw_pixel(c, &p);
}

/* Rasterizer.h:175-222: DDA with steps = max(|dxi|, |dyi|), last endpoint not drawn. */
static void draw_line(ctx *c, const vtx *v0, const vtx *v1)
{
    const traits *t = &c->t;
    int adx = abs((int)v1->x - (int)v0->x);
    int ady = abs((int)v1->y - (int)v0->y);
    int steps = adx > ady ? adx : ady;
    vtx step, v = *v0;
    memset(&step, 0, sizeof(step));
    step.x = (v1->x - v0->x) / steps;
    step.y = (v1->y - v0->y) / steps;
    if (t->useZ) step.z = (v1->z - v0->z) / steps;
    if (t->useW) step.w = (v1->w - v0->w) / steps;
    for (int i = 0; i < t->nA; ++i) step.avar[i] = (v1->avar[i] - v0->avar[i]) / steps;
    for (int i = 0; i < t->nP; ++i) step.pvar[i] = (v1->pvar[i] - v0->pvar[i]) / steps;

    while (steps-- > 0) {
        frag p;
        memset(&p, 0, sizeof(p));
        frag_from_vertex(c, &v, &p);
        if (scissor_test(c, v.x, v.y)) draw_pixel(c, &p);
        v.x += step.x;
        v.y += step.y;
        if (t->useZ) v.z += step.z;
        if (t->useW) v.w += step.w;
        for (int i = 0; i < t->nA; ++i) v.avar[i] += step.avar[i];
        for (int i = 0; i < t->nP; ++i) v.pvar[i] += step.pvar[i];
    }
}

/* ------------------------------------------------------------------ emission */
typedef struct { uint32_t ordinal; int n; vtx v[3]; } prim;   /* one primitive handed to the rasterizer */

static void record(ctx *c, const prim *p)
{
    swr_scene *s = c->s;
    s->primitives_out++;
    if (s->stream) {
        if (s->stream_len < s->stream_cap) {
            float *r = s->stream + s->stream_len * SWR_STREAM_FLOATS;
            uint32_t un = (uint32_t)p->n;
            memset(r, 0, sizeof(float) * SWR_STREAM_FLOATS);
            memcpy(r + 0, &p->ordinal, 4);
            memcpy(r + 1, &un, 4);
            for (int k = 0; k < p->n; ++k) {
                r[2 + 4 * k + 0] = p->v[k].x; r[2 + 4 * k + 1] = p->v[k].y;
                r[2 + 4 * k + 2] = p->v[k].z; r[2 + 4 * k + 3] = p->v[k].w;
            }
        }
        s->stream_len++;
    }
}

static void rasterize(ctx *c, const prim *p)
{
    c->ordinal = p->ordinal;
    record(c, p);
    if (p->n == 3) draw_triangle(c, &p->v[0], &p->v[1], &p->v[2]);
    else if (p->n == 2) draw_line(c, &p->v[0], &p->v[1]);
    else draw_point(c, &p->v[0]);
}

/* VertexProcessor.cpp:319-345 on screen-space corners.  Returns 0 if culled; may swap v0/v2. */
static int cull_or_orient(const ctx *c, prim *p)
{
    vtx *v0 = &p->v[0], *v1 = &p->v[1], *v2 = &p->v[2];
    float facing = (v0->x - v1->x) * (v2->y - v1->y) - (v2->x - v1->x) * (v0->y - v1->y);
    if (facing < 0) {
        if (c->s->cull_mode == 2) return 0;           /* CullMode::CW */
    } else {
        if (c->s->cull_mode == 1) return 0;           /* CullMode::CCW */
        vtx tmp = *v0; *v0 = *v2; *v2 = tmp;
    }
    return 1;
}

/* One input triangle -> up to SWR_MAX_POLY-2 screen-space triangles: fan (p0, p[k-1], p[k]).
 * VertexProcessor.cpp:217-263 + :347-377 + :319-345, evaluated per primitive.  out[0] is the
 * triangle that keeps the original slot, out[1..] are the extras appended to the batch tail.
 * alive[k] tells whether fan triangle k survives culling. Returns the fan size (0 if fully clipped). */
static int process_triangle(const ctx *c, const int32_t *idx, prim *out, int *alive)
{
    vtx a[SWR_MAX_POLY], b[SWR_MAX_POLY];
    vtx *in = a, *outp = b;
    int n = 3, mask = 0;
    for (int k = 0; k < 3; ++k) {
        shade_vertex(c, idx[k], &in[k]);
        mask |= outcode(&in[k]);
    }
    for (int pl = 0; pl < 6; ++pl) {
        if (!(mask & (1 << pl))) continue;
        if (n < 3) break;                                  /* PolyClipper.cpp:47-48 */
        n = clip_poly_plane(in, n, outp, PLANES[pl], c->t.nA_vs, c->t.nP_vs);
        if (n < 0) return 0;
        vtx *sw = in; in = outp; outp = sw;
    }
    if (n < 3) return 0;                                   /* VertexProcessor.cpp:244-250 */
    for (int k = 0; k < n; ++k) to_screen(c, &in[k]);
    for (int k = 0; k + 2 < n; ++k) {
        out[k].n = 3;
        out[k].v[0] = in[0];
        out[k].v[1] = in[k + 1];
        out[k].v[2] = in[k + 2];
        alive[k] = cull_or_orient(c, &out[k]);
    }
    return n - 2;
}

/* One input line (VertexProcessor.cpp:167-215 + LineClipper.cpp:30-56). Returns 0 if dropped. */
static int process_line(const ctx *c, const int32_t *idx, prim *out)
{
    vtx v0, v1;
    shade_vertex(c, idx[0], &v0);
    shade_vertex(c, idx[1], &v1);
    int m0 = outcode(&v0), m1 = outcode(&v1), mask = m0 | m1;
    float t0 = 0.0f, t1 = 1.0f;
    for (int pl = 0; pl < 6; ++pl) {
        if (!(mask & (1 << pl))) continue;
        float d0 = plane_dist(PLANES[pl], &v0);
        float d1 = plane_dist(PLANES[pl], &v1);
        int n0 = d0 < 0, n1 = d1 < 0;
        if (n0 && n1) return 0;
        if (n0) {
            float t = -d0 / (d1 - d0);
            t0 = t0 < t ? t : t0;                          /* std::max(t0, t) */
        } else {
            float t = d0 / (d0 - d1);
            t1 = t < t1 ? t : t1;                          /* std::min(t1, t) */
        }
    }
    out->n = 2;
    out->v[0] = v0;
    out->v[1] = v1;
    if (m0) lerp_vtx(&out->v[0], &v0, &v1, t0, c->t.nA_vs, c->t.nP_vs);
    if (m1) lerp_vtx(&out->v[1], &v0, &v1, t1, c->t.nA_vs, c->t.nP_vs);
    to_screen(c, &out->v[0]);
    to_screen(c, &out->v[1]);
    return 1;
}

/* One input point (VertexProcessor.cpp:152-165). */
static int process_point(const ctx *c, const int32_t *idx, prim *out)
{
    shade_vertex(c, idx[0], &out->v[0]);
    if (outcode(&out->v[0])) return 0;
    out->n = 1;
    to_screen(c, &out->v[0]);
    return 1;
}

static int ctx_init(ctx *c, swr_scene *s)
{
    memset(c, 0, sizeof(*c));
    c->s = s;
    int rc = set_traits(s->vs_kind, s->ps_kind, &c->t);
    if (rc) return rc;
    if (c->t.nA > c->t.nA_vs || c->t.nP > c->t.nP_vs) return -3;
    c->px = s->vp_w / 2.0f;                                /* VertexProcessor.cpp:50-53 */
    c->py = s->vp_h / 2.0f;
    c->ox = (s->vp_x + c->px);
    c->oy = (s->vp_y + c->py);
    c->minX = s->sc_x; c->minY = s->sc_y;                  /* Rasterizer.h:81-87 */
    c->maxX = s->sc_x + s->sc_w; c->maxY = s->sc_y + s->sc_h;
    s->fragments = 0;
    s->primitives_out = 0;
    s->stream_len = 0;
    if (s->ps_kind == SWR_PS_TEXTURED_ANISO) {
        if (s->draw_mode != 2 || !s->texture) return -5;    /* Box.cpp's shader dereferences p.equations */
        c->max_aniso = 8;                                   /* Texture.h:14 default */
        return tex_build_mips(c, s->texture, s->tex_w, s->tex_h);
    }
    return 0;
}

/* VertexProcessor::drawElements (VertexProcessor.cpp:78-120) for the stock shaders. */
int oracle_draw(swr_scene *s)
{
    ctx c;
    int rc = ctx_init(&c, s);
    if (rc) return rc;

    const int per = s->draw_mode + 1;                      /* indices per primitive */
    const int64_t nprim = s->index_count / per;
    prim fan[SWR_MAX_POLY];
    int alive[SWR_MAX_POLY];
    prim *extras = NULL;
    if (s->draw_mode == 2) {
        extras = (prim *)malloc(sizeof(prim) * (size_t)BATCH * (SWR_MAX_POLY - 3));
        if (!extras) return -4;
    }

    for (int64_t base = 0, batch = 0; base < nprim; base += BATCH, ++batch) {
        int64_t cnt = nprim - base < BATCH ? nprim - base : BATCH;
        uint32_t ord0 = (uint32_t)batch * SWR_ORDINAL_STRIDE;
        uint32_t slot_extra = (uint32_t)cnt;               /* extras are appended after the cnt original slots */
        int64_t nextra = 0;

        for (int64_t i = 0; i < cnt; ++i) {
            const int32_t *idx = s->indices + (base + i) * per;
            if (s->draw_mode == 2) {
                int n = process_triangle(&c, idx, fan, alive);
                /* original slot: rasterized now, in slot order */
                if (n > 0 && alive[0]) {
                    fan[0].ordinal = ord0 + (uint32_t)i;
                    rasterize(&c, &fan[0]);
                }
                /* extras: deferred to the batch tail, numbered in append order */
                for (int k = 1; k < n; ++k) {
                    fan[k].ordinal = ord0 + slot_extra++;
                    if (alive[k]) extras[nextra++] = fan[k];
                }
            } else if (s->draw_mode == 1) {
                if (process_line(&c, idx, &fan[0])) {
                    fan[0].ordinal = ord0 + (uint32_t)i;
                    rasterize(&c, &fan[0]);
                }
            } else {
                if (process_point(&c, idx, &fan[0])) {
                    fan[0].ordinal = ord0 + (uint32_t)i;
                    rasterize(&c, &fan[0]);
                }
            }
        }
        for (int64_t k = 0; k < nextra; ++k) rasterize(&c, &extras[k]);
    }
    free(extras);
    tex_free_mips(&c);
    return 0;
}

/* Rasterizer::drawTriangle on screen-space input (RasterizerTest.cpp:55-80):
 * verts = ntri * 3 * {x, y, z, w, a0, a1, a2}; ordinal of triangle t is t. */
int oracle_draw_raster_triangles(swr_scene *s, const float *verts, int64_t ntri)
{
    ctx c;
    int vs_save = s->vs_kind;
    s->vs_kind = SWR_VS_POS_COLOR;
    int rc = ctx_init(&c, s);
    s->vs_kind = vs_save;
    if (rc) return rc;
    for (int64_t t = 0; t < ntri; ++t) {
        prim p;
        memset(&p, 0, sizeof(p));
        p.n = 3;
        p.ordinal = (uint32_t)t;
        for (int k = 0; k < 3; ++k) {
            const float *f = verts + (t * 3 + k) * 7;
            p.v[k].x = f[0]; p.v[k].y = f[1]; p.v[k].z = f[2]; p.v[k].w = f[3];
            p.v[k].avar[0] = f[4]; p.v[k].avar[1] = f[5]; p.v[k].avar[2] = f[6];
        }
        c.ordinal = p.ordinal;
        s->primitives_out++;
        draw_triangle(&c, &p.v[0], &p.v[1], &p.v[2]);
    }
    return 0;
}

/* Rasterizer::drawPoint / drawLine / drawTriangle on screen-space input (Rasterizer.h:99-114):
 * verts = nprim * (mode + 1) * {x, y, z, w, a0, a1, a2}; ordinal of primitive t is t. */
int oracle_draw_raster_prims(swr_scene *s, int mode, const float *verts, int64_t nprim)
{
    if (mode == 2) return oracle_draw_raster_triangles(s, verts, nprim);
    if (mode != 0 && mode != 1) return -2;
    ctx c;
    int vs_save = s->vs_kind, dm_save = s->draw_mode;
    s->vs_kind = SWR_VS_POS_COLOR;
    s->draw_mode = mode;
    int rc = ctx_init(&c, s);
    s->vs_kind = vs_save;
    s->draw_mode = dm_save;
    if (rc) return rc;
    const int per = mode + 1;
    for (int64_t t = 0; t < nprim; ++t) {
        prim p;
        memset(&p, 0, sizeof(p));
        p.n = per;
        p.ordinal = (uint32_t)t;
        for (int k = 0; k < per; ++k) {
            const float *f = verts + (t * per + k) * 7;
            p.v[k].x = f[0]; p.v[k].y = f[1]; p.v[k].z = f[2]; p.v[k].w = f[3];
            p.v[k].avar[0] = f[4]; p.v[k].avar[1] = f[5]; p.v[k].avar[2] = f[6];
        }
        c.ordinal = p.ordinal;
        s->primitives_out++;
        if (mode == 0) draw_point(&c, &p.v[0]);
        else draw_line(&c, &p.v[0], &p.v[1]);
    }
    return 0;
}
