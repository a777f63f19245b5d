// stock_shaders.cuh -- the stock shader pack compiled into libswr_b200.so.
//
// These are ordinary user shaders: CRTP classes on swr::VertexShaderBase / swr::PixelShaderBase,
// written exactly as one writes them against the reference (compare Benchmark.cpp:14-48,
// RasterizerTest.cpp:32-50, Box.cpp:39-87), with "static" uniforms replaced by swr::uniforms<>()
// and the framebuffer reached through swr::target<>().  The bodies are mirrored line for line by
// the CPU checkers (oracle/swr_oracle.c, oracle/ref_driver.cpp) -- this TU is built with
// -fmad=false so that the shader bodies round like the oracle's (-ffp-contract=off).
#pragma once

#include <cstddef>
#include <swr/VertexShaderBase.h>
#include <swr/PixelShaderBase.h>
#include <swr/detail/geometry.cuh>
#include <swr/detail/tile.cuh>
#include <swr/Texture.h>

namespace stock {

using namespace swr;

struct PosColorVertex { float x, y, z, r, g, b; };
struct ObjVertex { float px, py, pz, nx, ny, nz, u, v; };

__device__ __forceinline__ void mvpTransform(const float *m, float x, float y, float z, VertexShaderOutput *out)
{
    const float w = 1.0f;
    out->x = m[0] * x + m[1] * y + m[2] * z + m[3] * w;
    out->y = m[4] * x + m[5] * y + m[6] * z + m[7] * w;
    out->z = m[8] * x + m[9] * y + m[10] * z + m[11] * w;
    out->w = m[12] * x + m[13] * y + m[14] * z + m[15] * w;
}

// ---- vertex shaders ----------------------------------------------------------------------------
struct VSPosColor : public VertexShaderBase<VSPosColor> {      // Benchmark.cpp:28-48
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 0;
    __device__ static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const PosColorVertex *d = static_cast<const PosColorVertex *>(in[0]);
        out->x = d->x; out->y = d->y; out->z = d->z; out->w = 1.0f;
        out->avar[0] = d->r; out->avar[1] = d->g; out->avar[2] = d->b;
    }
};

struct VSMvpColor : public VertexShaderBase<VSMvpColor> {
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 0;
    __device__ static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const PosColorVertex *d = static_cast<const PosColorVertex *>(in[0]);
        mvpTransform(uniforms<swr_stock_uniforms>().mvp, d->x, d->y, d->z, out);
        out->avar[0] = d->r; out->avar[1] = d->g; out->avar[2] = d->b;
    }
};

struct VSMvpNormalUv : public VertexShaderBase<VSMvpNormalUv> { // Box.cpp:65-87 (+ the normal as 3 avars)
    static const int AttribCount = 1;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    __device__ static void processVertex(VertexShaderInput in, VertexShaderOutput *out)
    {
        const ObjVertex v = fetchAttrib<ObjVertex>(in[0]), *d = &v;      // two 128-bit loads when the buffer allows (-1 % on C5)
        mvpTransform(uniforms<swr_stock_uniforms>().mvp, d->px, d->py, d->pz, out);
        out->avar[0] = d->nx; out->avar[1] = d->ny; out->avar[2] = d->nz;
        out->pvar[0] = d->u; out->pvar[1] = d->v;
    }
};

// ---- pixel shaders -----------------------------------------------------------------------------
__device__ __forceinline__ unsigned packRGB(const PixelData &p)   // RasterizerTest.cpp:39-45
{
    int rint = (int)(p.avar[0] * 255);
    int gint = (int)(p.avar[1] * 255);
    int bint = (int)(p.avar[2] * 255);
    return (unsigned)(rint << 16 | gint << 8 | bint);
}

struct PSFlat : public PixelShaderBase<PSFlat> {               // Benchmark.cpp:14-26
    static const int AVarCount = 3;
    static const int RenderTargets = 1;
    __device__ static void drawPixel(const PixelData &p) { target<unsigned>(p, SWR_RT_COLOR) = 1u; }
};

struct PSCountId : public PixelShaderBase<PSCountId> {
    static const int RenderTargets = 4;
    __device__ static void drawPixel(const PixelData &p)
    {
        target<unsigned>(p, SWR_RT_COUNT) += 1u;
        target<unsigned>(p, SWR_RT_PRIM_ID) = p.primitiveOrdinal;
    }
};

struct PSGouraud : public PixelShaderBase<PSGouraud> {         // RasterizerTest.cpp:32-50
    static const int AVarCount = 3;
    static const int RenderTargets = 1;
    __device__ static void drawPixel(const PixelData &p) { target<unsigned>(p, SWR_RT_COLOR) = packRGB(p); }
};

struct PSGouraudDepth : public PixelShaderBase<PSGouraudDepth> {
    static const bool InterpolateZ = true;
    static const int AVarCount = 3;
    static const int RenderTargets = 2;
    __device__ static void drawPixel(const PixelData &p)
    {
        float &depth = target<float>(p, SWR_RT_DEPTH);
        if (p.z < depth) {
            depth = p.z;
            target<unsigned>(p, SWR_RT_COLOR) = packRGB(p);
        }
    }
};

struct PSVaryDump : public PixelShaderBase<PSVaryDump> {
    static const bool InterpolateZ = true;
    static const bool InterpolateW = true;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    static const int RenderTargets = 12;
    __device__ static void drawPixel(const PixelData &p)
    {
        target<float>(p, SWR_RT_VARY0 + 0) = p.z;
        target<float>(p, SWR_RT_VARY0 + 1) = p.w;
        target<float>(p, SWR_RT_VARY0 + 2) = p.invw;
        target<float>(p, SWR_RT_VARY0 + 3) = p.avar[0];
        target<float>(p, SWR_RT_VARY0 + 4) = p.avar[1];
        target<float>(p, SWR_RT_VARY0 + 5) = p.avar[2];
        target<float>(p, SWR_RT_VARY0 + 6) = p.pvar[0];
        target<float>(p, SWR_RT_VARY0 + 7) = p.pvar[1];
        target<unsigned>(p, SWR_RT_COUNT) += 1u;
    }
};

struct PSTextured : public PixelShaderBase<PSTextured> {       // Box.cpp:39-62, nearest fetch
    static const bool InterpolateW = true;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    static const int RenderTargets = 1;
    __device__ static void drawPixel(const PixelData &p)
    {
        const swr_stock_uniforms &u = uniforms<swr_stock_uniforms>();
        int tx = (int)floorf(p.pvar[0] * (float)u.tex_w) & (u.tex_w - 1);
        int ty = (int)floorf(p.pvar[1] * (float)u.tex_h) & (u.tex_h - 1);
        target<unsigned>(p, SWR_RT_COLOR) = __ldg(u.texture + ty * u.tex_w + tx);
    }
};

static_assert(SWR_MAX_MIP_LEVELS == swr::kMaxMipLevels, "uniform block and TextureView disagree");
static_assert(offsetof(swr_stock_uniforms, max_anisotropy) - offsetof(swr_stock_uniforms, mip) == offsetof(swr::TextureView, maxAnisotropy),
              "swr_stock_uniforms::mip.. must have the layout of swr::TextureView");

struct PSTexturedAniso : public PixelShaderBase<PSTexturedAniso> {   // Box.cpp:39-62, verbatim structure
    static const bool InterpolateZ = false;
    static const bool InterpolateW = true;  // Required for perspective correct texturing
    static const int AVarCount = 0;
    static const int PVarCount = 2;         // UV coordinates
    static const int RenderTargets = 1;
    __device__ static void drawPixel(const PixelData &p)
    {
        // Compute texture coordinate derivatives
        float dudx, dudy, dvdx, dvdy;
        p.computePerspectiveDerivatives(*p.equations, 0, dudx, dudy); // U derivatives
        p.computePerspectiveDerivatives(*p.equations, 1, dvdx, dvdy); // V derivatives
        const TextureView &tex = *reinterpret_cast<const TextureView *>(&uniforms<swr_stock_uniforms>().mip[0]);
        target<unsigned>(p, SWR_RT_COLOR) = textureSample(tex, p.pvar[0], p.pvar[1], dudx, dvdx, dudy, dvdy);
    }
};

} // namespace stock
