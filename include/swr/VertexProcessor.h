// swr/VertexProcessor.h -- the VertexProcessor class, same public surface as the reference
// (src/renderer/VertexProcessor.h:56-91): VertexProcessor(IRasterizer*), setRasterizer,
// setViewport, setDepthRange, setCullMode, setVertexShader<VS>, setVertexAttribPointer,
// drawElements.  Defaults as VertexProcessor.cpp:29-35: CullMode::CW, depth range (0,1),
// DummyVertexShader.
//
// The primitives never leave the device between the two stages, so the rasterizer handed in must
// be a swr::Rasterizer (the reference's IRasterizer seam with host arrays remains available on
// the Rasterizer itself through draw*List).
#pragma once

#include <cassert>

#include "Rasterizer.h"
#include "VertexShaderBase.h"
#include "detail/geometry.cuh"

namespace swr {

/// Primitive draw mode (VertexProcessor.h:42-46).
enum class DrawMode {
    Point,
    Line,
    Triangle
};

/// Triangle culling mode (VertexProcessor.h:49-53).
enum class CullMode {
    None,
    CCW,
    CW
};

/// Process vertices and pass them to a rasterizer.
class VertexProcessor {
public:
    /// Constructor.
    VertexProcessor(IRasterizer *rasterizer) : m_rasterizer(nullptr)
    {
        setRasterizer(rasterizer);
        setCullMode(CullMode::CW);
        setDepthRange(0.0f, 1.0f);
#if defined(__CUDACC__)
        setVertexShader<DummyVertexShader>();
#endif
    }

    /// Change the rasterizer where the primitives are sent.
    void setRasterizer(IRasterizer *rasterizer)
    {
        assert(rasterizer != nullptr);
        m_rasterizer = dynamic_cast<Rasterizer *>(rasterizer);
        assert(m_rasterizer != nullptr && "swr::VertexProcessor needs a swr::Rasterizer");
    }

    /// Set the viewport. Top-Left is (0, 0).
    void setViewport(int x, int y, int width, int height) { detail::check(swr_set_viewport(ctx(), x, y, width, height), "setViewport"); }

    /// Set the depth range. Default is (0, 1).
    void setDepthRange(float n, float f) { detail::check(swr_set_depth_range(ctx(), n, f), "setDepthRange"); }

    /// Set the cull mode. Default is CullMode::CW to cull clockwise triangles.
    void setCullMode(CullMode mode) { detail::check(swr_set_cull_mode(ctx(), (int)mode), "setCullMode"); }

#if defined(__CUDACC__)
    /// Set the vertex shader.
    template <class VertexShader>
    void setVertexShader()
    {
        assert(VertexShader::AttribCount <= MaxVertexAttribs);
        detail::check(swr_set_vertex_shader(ctx(), detail::vertexShaderBinding<VertexShader>()), "setVertexShader");
    }
#endif

    /// Set a vertex attrib pointer (device memory, or managed memory; host memory needs the
    /// sized overload because the reference signature carries no extent).
    void setVertexAttribPointer(int index, int stride, const void *buffer)
    {
        assert(index < MaxVertexAttribs);
        detail::check(swr_set_vertex_attrib_pointer(ctx(), index, stride, buffer, 0), "setVertexAttribPointer");
    }
    /// Additive sized overload: `bytes` readable bytes behind `buffer` (host or device).
    void setVertexAttribPointer(int index, int stride, const void *buffer, size_t bytes)
    {
        assert(index < MaxVertexAttribs);
        detail::check(swr_set_vertex_attrib_pointer(ctx(), index, stride, buffer, bytes), "setVertexAttribPointer");
    }

    /// Draw a number of points, lines or triangles.  Like the reference, results are complete on
    /// return (the call waits for the device).
    void drawElements(DrawMode mode, size_t count, int *indices) const
    {
        detail::check(swr_draw_elements(ctx(), (int)mode, count, indices), "drawElements");
        detail::check(swr_finish(ctx()), "drawElements");
    }
    /// Additive: enqueue only; pair with Rasterizer::finish().
    void drawElementsAsync(DrawMode mode, size_t count, const int *indices) const
    {
        detail::check(swr_draw_elements(ctx(), (int)mode, count, indices), "drawElements");
    }

private:
    swr_context *ctx() const { return m_rasterizer->context(); }
    Rasterizer *m_rasterizer;
};

} // namespace swr
