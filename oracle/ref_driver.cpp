// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin driver around the UNMODIFIED reference renderer.  The reference sources are
// not copied: oracle/Makefile compiles this file with -I/root/reference/src/renderer
// and links /root/reference/src/renderer/{VertexProcessor,PolyClipper,LineClipper}.cpp
// where they lie; the only output is oracle/_ref/libswr_ref.so (git-ignored).
//
// What is ours here: the stock shaders (CRTP classes on the reference's own
// VertexShaderBase / PixelShaderBase, same bodies as oracle/swr_oracle.c and
// softwarerenderer_b200/csrc/stock_shaders.cuh), a recording IRasterizer that stamps
// the emission ordinal before forwarding every primitive to the reference Rasterizer,
// and the flat C entry point  ref_draw(swr_scene*).
//
// Mandatory flags: -O2 -std=c++17 -ffp-contract=off (FMA contraction
// changes coverage).
#include "Renderer.h"      // reference: src/renderer/Renderer.h
#include "Random.h"        // reference: src/examples/Random.h
#include "vector_math.h"   // reference: src/examples/vector_math.h
#include "SDL.h"           // oracle/sdl_shim/SDL.h (stand-in for the SDL2 names Texture.h uses)
#include "Texture.h"       // reference: src/examples/Texture.h, unmodified

#include <cmath>
#include <cstring>
#include <cstdint>

#include "swr_scene.h"

using namespace swr;

namespace {

swr_scene *g_s = nullptr;       // scene being drawn (single-threaded)
uint32_t g_ordinal = 0;         // emission ordinal of the primitive being rasterized

inline uint32_t packRGB(const PixelData &p)
{
    // RasterizerTest.cpp:39-45
    int rint = (int)(p.avar[0] * 255);
    int gint = (int)(p.avar[1] * 255);
    int bint = (int)(p.avar[2] * 255);
    return (uint32_t)(rint << 16 | gint << 8 | bint);
}

// ---------------------------------------------------------------- vertex shaders
struct PosColorVertex { float x, y, z, r, g, b; };
struct ObjVertex { float px, py, pz, nx, ny, nz, u, v; };

struct VSPosColor : public VertexShaderBase<VSPosColor> {
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 0;
    static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const PosColorVertex *d = static_cast<const PosColorVertex *>(in[0]);
        out->x = d->x; out->y = d->y; out->z = d->z; out->w = 1.0f;
        out->avar[0] = d->r; out->avar[1] = d->g; out->avar[2] = d->b;
    }
};

inline void mvpTransform(const float *m, float x, float y, float z, VertexShaderOutput *out)
{
    const float w = 1.0f;
    out->x = m[0] * x + m[1] * y + m[2] * z + m[3] * w;
    out->y = m[4] * x + m[5] * y + m[6] * z + m[7] * w;
    out->z = m[8] * x + m[9] * y + m[10] * z + m[11] * w;
    out->w = m[12] * x + m[13] * y + m[14] * z + m[15] * w;
}

struct VSMvpColor : public VertexShaderBase<VSMvpColor> {
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 0;
    static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const PosColorVertex *d = static_cast<const PosColorVertex *>(in[0]);
        mvpTransform(g_s->mvp, d->x, d->y, d->z, out);
        out->avar[0] = d->r; out->avar[1] = d->g; out->avar[2] = d->b;
    }
};

struct VSMvpNormalUv : public VertexShaderBase<VSMvpNormalUv> {
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const ObjVertex *d = static_cast<const ObjVertex *>(in[0]);
        mvpTransform(g_s->mvp, d->px, d->py, d->pz, out);
        out->avar[0] = d->nx; out->avar[1] = d->ny; out->avar[2] = d->nz;
        out->pvar[0] = d->u; out->pvar[1] = d->v;
    }
};

// ---------------------------------------------------------------- pixel shaders
struct PSFlat : public PixelShaderBase<PSFlat> {
    static const int AVarCount = 3;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        g_s->color[p.x + g_s->width * p.y] = 1;
    }
};

struct PSCountId : public PixelShaderBase<PSCountId> {
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        int i = p.x + g_s->width * p.y;
        g_s->count[i]++;
        g_s->prim_id[i] = g_ordinal;
    }
};

struct PSGouraud : public PixelShaderBase<PSGouraud> {
    static const int AVarCount = 3;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        g_s->color[p.x + g_s->width * p.y] = packRGB(p);
    }
};

struct PSGouraudDepth : public PixelShaderBase<PSGouraudDepth> {
    static const bool InterpolateZ = true;
    static const int AVarCount = 3;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        int i = p.x + g_s->width * p.y;
        if (p.z < g_s->depth[i]) {
            g_s->depth[i] = p.z;
            g_s->color[i] = packRGB(p);
        }
    }
};

struct PSVaryDump : public PixelShaderBase<PSVaryDump> {
    static const bool InterpolateZ = true;
    static const bool InterpolateW = true;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        size_t n = (size_t)g_s->width * g_s->height;
        size_t i = (size_t)p.x + (size_t)g_s->width * p.y;
        float *v = g_s->vary;
        v[0 * n + i] = p.z;
        v[1 * n + i] = p.w;
        v[2 * n + i] = p.invw;
        v[3 * n + i] = p.avar[0];
        v[4 * n + i] = p.avar[1];
        v[5 * n + i] = p.avar[2];
        v[6 * n + i] = p.pvar[0];
        v[7 * n + i] = p.pvar[1];
        g_s->count[i]++;
    }
};

struct PSTextured : public PixelShaderBase<PSTextured> {
    static const bool InterpolateW = true;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        int tx = (int)floorf(p.pvar[0] * (float)g_s->tex_w) & (g_s->tex_w - 1);
        int ty = (int)floorf(p.pvar[1] * (float)g_s->tex_h) & (g_s->tex_h - 1);
        g_s->color[p.x + g_s->width * p.y] = g_s->texture[ty * g_s->tex_w + tx];
    }
};

// Box.cpp:39-62: perspective-correct UV derivatives + the reference's anisotropic / trilinear sampler.
Texture *g_texture = nullptr;

struct PSTexturedAniso : public PixelShaderBase<PSTexturedAniso> {
    static const bool InterpolateZ = false;
    static const bool InterpolateW = true;
    static const int AVarCount = 0;
    static const int PVarCount = 2;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        float dudx, dudy, dvdx, dvdy;
        p.computePerspectiveDerivatives(*p.equations, 0, dudx, dudy);
        p.computePerspectiveDerivatives(*p.equations, 1, dvdx, dvdy);
        Uint32 sampledColor;
        g_texture->sample(p.pvar[0], p.pvar[1], dudx, dvdx, dudy, dvdy, sampledColor);
        g_s->color[p.x + g_s->width * p.y] = sampledColor;
    }
};

// ---------------------------------------------------------------- recording rasterizer
// Forwards primitive by primitive to the reference Rasterizer (same loops as
// Rasterizer.h:116-141) and stamps the emission ordinal first.
class RecordingRasterizer : public IRasterizer {
public:
    Rasterizer inner;
    mutable uint32_t batch = 0;

    void record(uint32_t ordinal, int n, const RasterizerVertex *a, const RasterizerVertex *b, const RasterizerVertex *c) const
    {
        g_s->primitives_out++;
        if (g_s->stream) {
            if (g_s->stream_len < g_s->stream_cap) {
                float *r = g_s->stream + g_s->stream_len * SWR_STREAM_FLOATS;
                uint32_t un = (uint32_t)n;
                memset(r, 0, sizeof(float) * SWR_STREAM_FLOATS);
                memcpy(r + 0, &ordinal, 4);
                memcpy(r + 1, &un, 4);
                const RasterizerVertex *v[3] = { a, b, c };
                for (int k = 0; k < n; ++k) {
                    r[2 + 4 * k + 0] = v[k]->x; r[2 + 4 * k + 1] = v[k]->y;
                    r[2 + 4 * k + 2] = v[k]->z; r[2 + 4 * k + 3] = v[k]->w;
                }
            }
            g_s->stream_len++;
        }
    }

    void drawPointList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        for (size_t i = 0; i < indexCount; ++i) {
            if (indices[i] == -1) continue;
            g_ordinal = batch * SWR_ORDINAL_STRIDE + (uint32_t)i;
            record(g_ordinal, 1, &vertices[indices[i]], nullptr, nullptr);
            inner.drawPoint(vertices[indices[i]]);
        }
        batch++;
    }

    void drawLineList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        for (size_t i = 0; i + 2 <= indexCount; i += 2) {
            if (indices[i] == -1) continue;
            g_ordinal = batch * SWR_ORDINAL_STRIDE + (uint32_t)(i / 2);
            record(g_ordinal, 2, &vertices[indices[i]], &vertices[indices[i + 1]], nullptr);
            inner.drawLine(vertices[indices[i]], vertices[indices[i + 1]]);
        }
        batch++;
    }

    void drawTriangleList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        for (size_t i = 0; i + 3 <= indexCount; i += 3) {
            if (indices[i] == -1) continue;
            g_ordinal = batch * SWR_ORDINAL_STRIDE + (uint32_t)(i / 3);
            record(g_ordinal, 3, &vertices[indices[i]], &vertices[indices[i + 1]], &vertices[indices[i + 2]]);
            inner.drawTriangle(vertices[indices[i]], vertices[indices[i + 1]], vertices[indices[i + 2]]);
        }
        batch++;
    }
};

template <class PS>
void drawWithPS(RecordingRasterizer &r, VertexProcessor &v, swr_scene *s)
{
    r.inner.setPixelShader<PS>();
    v.drawElements((DrawMode)s->draw_mode, (size_t)s->index_count, const_cast<int *>(s->indices));
}

} // namespace

extern "C" {

// Draw one scene with the unmodified reference VertexProcessor + Rasterizer.
int ref_draw(swr_scene *s)
{
    g_s = s;
    g_ordinal = 0;
    s->fragments = 0;
    s->primitives_out = 0;
    s->stream_len = 0;

    RecordingRasterizer r;
    VertexProcessor v(&r);

    r.inner.setRasterMode((RasterMode)s->raster_mode);
    r.inner.setScissorRect(s->sc_x, s->sc_y, s->sc_w, s->sc_h);
    v.setViewport(s->vp_x, s->vp_y, s->vp_w, s->vp_h);
    v.setDepthRange(s->depth_n, s->depth_f);
    v.setCullMode((CullMode)s->cull_mode);
    v.setVertexAttribPointer(0, s->stride, s->vertices);

    switch (s->vs_kind) {
    case SWR_VS_POS_COLOR: v.setVertexShader<VSPosColor>(); break;
    case SWR_VS_MVP_COLOR: v.setVertexShader<VSMvpColor>(); break;
    case SWR_VS_MVP_NORMAL_UV: v.setVertexShader<VSMvpNormalUv>(); break;
    default: return -1;
    }

    switch (s->ps_kind) {
    case SWR_PS_FLAT: drawWithPS<PSFlat>(r, v, s); break;
    case SWR_PS_COUNT_ID: drawWithPS<PSCountId>(r, v, s); break;
    case SWR_PS_GOURAUD: drawWithPS<PSGouraud>(r, v, s); break;
    case SWR_PS_GOURAUD_DEPTH: drawWithPS<PSGouraudDepth>(r, v, s); break;
    case SWR_PS_VARY_DUMP: drawWithPS<PSVaryDump>(r, v, s); break;
    case SWR_PS_TEXTURED: drawWithPS<PSTextured>(r, v, s); break;
    case SWR_PS_TEXTURED_ANISO: {
        if (s->draw_mode != 2 || !s->texture) return -5;       // Box.cpp's shader dereferences p.equations
        SDL_Surface base = { s->tex_w, s->tex_h, s->tex_w * 4, const_cast<uint32_t *>(s->texture) };
        Texture tex(&base);                                    // builds the mip chain (Texture.h:220-293)
        g_texture = &tex;
        drawWithPS<PSTexturedAniso>(r, v, s);
        g_texture = nullptr;
        break;
    }
    default: return -2;
    }
    g_s = nullptr;
    return 0;
}

// Rasterizer::drawTriangle called directly on screen-space vertices (the
// RasterizerTest.cpp:55-80 entry).  verts: n triangles * 3 vertices * 7 floats
// {x, y, z, w, a0, a1, a2}; ordinal of triangle t is t.
int ref_draw_raster_triangles(swr_scene *s, const float *verts, int64_t ntri)
{
    g_s = s;
    s->fragments = 0;
    s->primitives_out = 0;
    s->stream_len = 0;
    Rasterizer r;
    r.setRasterMode((RasterMode)s->raster_mode);
    r.setScissorRect(s->sc_x, s->sc_y, s->sc_w, s->sc_h);
    switch (s->ps_kind) {
    case SWR_PS_FLAT: r.setPixelShader<PSFlat>(); break;
    case SWR_PS_COUNT_ID: r.setPixelShader<PSCountId>(); break;
    case SWR_PS_GOURAUD: r.setPixelShader<PSGouraud>(); break;
    case SWR_PS_GOURAUD_DEPTH: r.setPixelShader<PSGouraudDepth>(); break;
    default: return -2;
    }
    for (int64_t t = 0; t < ntri; ++t) {
        RasterizerVertex v[3];
        memset(v, 0, sizeof(v));
        for (int k = 0; k < 3; ++k) {
            const float *f = verts + (t * 3 + k) * 7;
            v[k].x = f[0]; v[k].y = f[1]; v[k].z = f[2]; v[k].w = f[3];
            v[k].avar[0] = f[4]; v[k].avar[1] = f[5]; v[k].avar[2] = f[6];
        }
        g_ordinal = (uint32_t)t;
        s->primitives_out++;
        r.drawTriangle(v[0], v[1], v[2]);
    }
    g_s = nullptr;
    return 0;
}

// Rasterizer::drawPoint / drawLine / drawTriangle on screen-space vertices (Rasterizer.h:99-114).
// verts: nprim * (mode + 1) vertices * 7 floats {x, y, z, w, a0, a1, a2}; ordinal of primitive t is t.
int ref_draw_raster_prims(swr_scene *s, int mode, const float *verts, int64_t nprim)
{
    if (mode == 2) return ref_draw_raster_triangles(s, verts, nprim);
    if (mode != 0 && mode != 1) return -2;
    g_s = s;
    s->fragments = 0;
    s->primitives_out = 0;
    s->stream_len = 0;
    Rasterizer r;
    r.setRasterMode((RasterMode)s->raster_mode);
    r.setScissorRect(s->sc_x, s->sc_y, s->sc_w, s->sc_h);
    switch (s->ps_kind) {
    case SWR_PS_FLAT: r.setPixelShader<PSFlat>(); break;
    case SWR_PS_COUNT_ID: r.setPixelShader<PSCountId>(); break;
    case SWR_PS_GOURAUD: r.setPixelShader<PSGouraud>(); break;
    case SWR_PS_GOURAUD_DEPTH: r.setPixelShader<PSGouraudDepth>(); break;
    default: return -2;
    }
    const int per = mode + 1;
    for (int64_t t = 0; t < nprim; ++t) {
        RasterizerVertex v[2];
        memset(v, 0, sizeof(v));
        for (int k = 0; k < per; ++k) {
            const float *f = verts + (t * per + k) * 7;
            v[k].x = f[0]; v[k].y = f[1]; v[k].z = f[2]; v[k].w = f[3];
            v[k].avar[0] = f[4]; v[k].avar[1] = f[5]; v[k].avar[2] = f[6];
        }
        g_ordinal = (uint32_t)t;
        s->primitives_out++;
        if (mode == 0) r.drawPoint(v[0]);
        else r.drawLine(v[0], v[1]);
    }
    g_s = nullptr;
    return 0;
}

// n doubles of the reference's Random(seed).NextDouble() stream (Random.cpp:45-50),
// used to check our own generator of Benchmark.cpp's vertex set.
void ref_random_doubles(int seed, int64_t n, double *out)
{
    Random rnd(seed);
    for (int64_t i = 0; i < n; ++i) out[i] = rnd.NextDouble();
}

// The Box.cpp:190-195 camera: perspective(60, 4/3, 0.1, 10) * lookat(eye, 0, +Y),
// written row-major (clip = M * (pos, 1)).
void ref_box_mvp(float ex, float ey, float ez, float fovy, float aspect, float zn, float zf, float *out16)
{
    typedef vmath::vec3<float> vec3f;
    typedef vmath::mat4<float> mat4f;
    mat4f look = vmath::lookat_matrix(vec3f(ex, ey, ez), vec3f(0.0f), vec3f(0.0f, 1.0f, 0.0f));
    mat4f persp = vmath::perspective_matrix(fovy, aspect, zn, zf);
    mat4f m = persp * look;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            out16[4 * r + c] = m.elem[r][c];
}

} // extern "C"
