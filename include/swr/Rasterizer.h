// swr/Rasterizer.h -- the Rasterizer class, same public surface as the reference
// (src/renderer/Rasterizer.h:52-141): setRasterMode, setScissorRect, setPixelShader<PS>,
// drawPoint/Line/Triangle, draw{Point,Line,Triangle}List.  Defaults as Rasterizer.h:67-72:
// RasterMode::Span, scissor (0,0,0,0), NullPixelShader.
//
// The object owns a swr_context (one CUDA stream + device scratch).  setPixelShader<PS>()
// captures host launchers of the tile kernel instantiated with PS in the caller's translation
// unit (the reference captures three member-function-template pointers, Rasterizer.h:90-96).
// Additive: setRenderTarget, setUniforms, finish, context().
#pragma once

#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "IRasterizer.h"
#include "PixelShaderBase.h"
#include "detail/tile.cuh"

namespace swr {

/// Rasterizer mode (Rasterizer.h:45-49).
enum class RasterMode {
    Span,
    Block,
    Adaptive
};

namespace detail {
inline void check(int rc, const char *what)
{
    if (rc != 0) {
        std::fprintf(stderr, "swr: %s failed: %s\n", what, swr_last_error());
        std::abort();   // the reference API returns void everywhere; failures are fatal like its asserts
    }
}
} // namespace detail

/// Rasterizer main class.
class Rasterizer : public IRasterizer {
public:
    /// Constructor. `cudaDevice` is additive (default: the current device 0).
    explicit Rasterizer(int cudaDevice = 0) : m_ctx(nullptr)
    {
        detail::check(swr_create(&m_ctx, cudaDevice), "swr_create");
        setRasterMode(RasterMode::Span);
        setScissorRect(0, 0, 0, 0);
#if defined(__CUDACC__)
        setPixelShader<NullPixelShader>();
#endif
    }
    ~Rasterizer() { swr_destroy(m_ctx); }
    Rasterizer(const Rasterizer &) = delete;
    Rasterizer &operator=(const Rasterizer &) = delete;

    /// Set the raster mode. The default is RasterMode::Span.
    void setRasterMode(RasterMode mode) { detail::check(swr_set_raster_mode(m_ctx, (int)mode), "setRasterMode"); }

    /// Set the scissor rectangle.
    void setScissorRect(int x, int y, int width, int height)
    {
        detail::check(swr_set_scissor_rect(m_ctx, x, y, width, height), "setScissorRect");
    }

#if defined(__CUDACC__)
    /// Set the pixel shader.
    template <class PixelShader>
    void setPixelShader()
    {
        detail::check(swr_set_pixel_shader(m_ctx, detail::pixelShaderBinding<PixelShader>()), "setPixelShader");
    }
#endif

    /// Draw a single point / line / triangle given in screen space (Rasterizer.h:99-114).
    void drawPoint(const RasterizerVertex &v) const
    {
        const int idx[1] = { 0 };
        drawPointList(&v, idx, 1);
    }
    void drawLine(const RasterizerVertex &v0, const RasterizerVertex &v1) const
    {
        const RasterizerVertex v[2] = { v0, v1 };
        const int idx[2] = { 0, 1 };
        drawLineList(v, idx, 2);
    }
    void drawTriangle(const RasterizerVertex &v0, const RasterizerVertex &v1, const RasterizerVertex &v2) const
    {
        const RasterizerVertex v[3] = { v0, v1, v2 };
        const int idx[3] = { 0, 1, 2 };
        drawTriangleList(v, idx, 3);
    }

    void drawPointList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        drawList(SWR_DRAW_POINT, vertices, indices, indexCount);
    }
    void drawLineList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        drawList(SWR_DRAW_LINE, vertices, indices, indexCount);
    }
    void drawTriangleList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const override
    {
        drawList(SWR_DRAW_TRIANGLE, vertices, indices, indexCount);
    }

    // ---- additive ------------------------------------------------------------------------------
    /// Register a 32-bit-per-pixel device surface as render target `slot` (see swr::target).
    void setRenderTarget(int slot, void *devicePtr, int pitchBytes, int width, int height)
    {
        detail::check(swr_set_render_target(m_ctx, slot, devicePtr, pitchBytes, width, height), "setRenderTarget");
    }
    /// Bytes handed to swr::uniforms<T>() in the shaders of the next draw.
    void setUniforms(const void *data, size_t bytes) { detail::check(swr_set_uniforms(m_ctx, data, bytes), "setUniforms"); }
    /// Wait for all enqueued draws.
    void finish() const { detail::check(swr_finish(m_ctx), "finish"); }
    /// Multi-GPU, one process per GPU (sort-first): this context shades the screen tiles of `rank` out of `world`.
    void setTilePartition(int rank, int world) { detail::check(swr_set_tile_partition(m_ctx, rank, world), "setTilePartition"); }
    void setTileSplit(int groups) { detail::check(swr_set_tile_split(m_ctx, groups), "setTileSplit"); }      // heavy tiles shaded by four CTAs
    void setIndexNarrowing(int mode) { detail::check(swr_set_index_narrowing(m_ctx, mode), "setIndexNarrowing"); }   // big host index arrays as 16-bit offsets over PCIe
    /// ... and stores every finished tile of `slot` also into `count` peer surfaces (mapped with swr_ipc_open): the
    /// composite then needs no pass of its own, only a cross-rank barrier after the draws (see swr_b200.h).
    void setTileMirrors(int slot, int count, void *const *surfaces)
    {
        detail::check(swr_set_tile_mirrors(m_ctx, slot, count, surfaces), "setTileMirrors");
    }
    /// Enqueue on a caller-owned CUDA stream (nullptr: the context's own).
    void setStream(void *cudaStream) { detail::check(swr_set_stream(m_ctx, cudaStream), "setStream"); }
    swr_context *context() const { return m_ctx; }

private:
    void drawList(int mode, const RasterizerVertex *vertices, const int *indices, size_t indexCount) const
    {
        // the vertex array extent is not part of the reference signature: take max(index) + 1
        int maxIndex = -1;
        for (size_t i = 0; i < indexCount; ++i) maxIndex = indices[i] > maxIndex ? indices[i] : maxIndex;
        detail::check(swr_draw_raster_list(m_ctx, mode, vertices, (size_t)(maxIndex + 1), indices, indexCount), "draw*List");
    }

    swr_context *m_ctx;
};

} // namespace swr
