// swr/VertexProcessor.h -- the VertexProcessor class, same public surface as the reference
// (src/renderer/VertexProcessor.h:56-91): VertexProcessor(IRasterizer*), setRasterizer,
// setViewport, setDepthRange, setCullMode, setVertexShader<VS>, setVertexAttribPointer,
// drawElements.  Defaults as VertexProcessor.cpp:29-35: CullMode::CW, depth range (0,1),
// DummyVertexShader.
//
// With a swr::Rasterizer the primitives never leave the device between the two stages; any other
// IRasterizer gets the reference's host arrays batch by batch (swr_process_elements).  Uniforms for the
// vertex shader of such a stand-alone vertex stage: setUniforms below.
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
///
/// With a swr::Rasterizer the primitives never leave the device between the two stages.  Any other IRasterizer
/// (IRasterizer.h:56-71) is served the way the reference serves it (VertexProcessor.cpp:302-317): the vertex stage
/// runs on the device, and per batch of 1024 input primitives the rasterizer's draw*List is called with host arrays of
/// screen-space RasterizerVertex records and indices (-1 = dropped primitive, clipper fan triangles appended).
class VertexProcessor {
public:
    /// Constructor.
    VertexProcessor(IRasterizer *rasterizer) : m_rasterizer(nullptr), m_foreign(nullptr), m_own(nullptr), m_vs(nullptr)
    {
        for (int i = 0; i < MaxVertexAttribs; ++i) { m_attribs[i].buffer = nullptr; m_attribs[i].stride = 0; m_attribs[i].bytes = 0; m_attribs[i].set = false; }
        m_vp[0] = m_vp[1] = m_vp[2] = m_vp[3] = 0;
        m_haveViewport = false;
        m_depthN = 0.0f; m_depthF = 1.0f;
        m_cull = CullMode::CW;
        setRasterizer(rasterizer);
#if defined(__CUDACC__)
        setVertexShader<DummyVertexShader>();
#endif
    }
    ~VertexProcessor() { if (m_own) swr_destroy(m_own); }
    VertexProcessor(const VertexProcessor &) = delete;
    VertexProcessor &operator=(const VertexProcessor &) = delete;

    /// Change the rasterizer where the primitives are sent.
    void setRasterizer(IRasterizer *rasterizer)
    {
        assert(rasterizer != nullptr);
        m_rasterizer = dynamic_cast<Rasterizer *>(rasterizer);
        m_foreign = m_rasterizer ? nullptr : rasterizer;
        if (m_foreign && !m_own) detail::check(swr_create(&m_own, 0), "swr_create");   // a context for the vertex stage alone
        apply();
    }

    /// Set the viewport. Top-Left is (0, 0).
    void setViewport(int x, int y, int width, int height)
    {
        m_vp[0] = x; m_vp[1] = y; m_vp[2] = width; m_vp[3] = height;
        m_haveViewport = true;
        detail::check(swr_set_viewport(ctx(), x, y, width, height), "setViewport");
    }

    /// Set the depth range. Default is (0, 1).
    void setDepthRange(float n, float f)
    {
        m_depthN = n; m_depthF = f;
        detail::check(swr_set_depth_range(ctx(), n, f), "setDepthRange");
    }

    /// Set the cull mode. Default is CullMode::CW to cull clockwise triangles.
    void setCullMode(CullMode mode)
    {
        m_cull = mode;
        detail::check(swr_set_cull_mode(ctx(), (int)mode), "setCullMode");
    }

#if defined(__CUDACC__)
    /// Set the vertex shader.
    template <class VertexShader>
    void setVertexShader()
    {
        assert(VertexShader::AttribCount <= MaxVertexAttribs);
        m_vs = detail::vertexShaderBinding<VertexShader>();
        detail::check(swr_set_vertex_shader(ctx(), m_vs), "setVertexShader");
    }
#endif

    /// Set a vertex attrib pointer: device or managed memory is read in place; host memory is staged per draw up
    /// to stride * (largest index of the draw + 1), the reference signature carrying no extent.
    void setVertexAttribPointer(int index, int stride, const void *buffer)
    {
        setVertexAttribPointer(index, stride, buffer, 0);
    }
    /// Additive sized overload: `bytes` readable bytes behind `buffer` (host or device).
    void setVertexAttribPointer(int index, int stride, const void *buffer, size_t bytes)
    {
        assert(index < MaxVertexAttribs);
        m_attribs[index].buffer = buffer; m_attribs[index].stride = stride; m_attribs[index].bytes = bytes; m_attribs[index].set = true;
        detail::check(swr_set_vertex_attrib_pointer(ctx(), index, stride, buffer, bytes), "setVertexAttribPointer");
    }

    /// Draw a number of points, lines or triangles.  Like the reference, results are complete on
    /// return (the call waits for the device).
    void drawElements(DrawMode mode, size_t count, int *indices) const
    {
        if (m_foreign) {
            detail::check(swr_process_elements(ctx(), (int)mode, count, indices, &VertexProcessor::emit, m_foreign), "drawElements");
            return;
        }
        detail::check(swr_draw_elements(ctx(), (int)mode, count, indices), "drawElements");
        detail::check(swr_finish(ctx()), "drawElements");
    }
    /// Additive: enqueue only; pair with Rasterizer::finish().  (With a foreign IRasterizer the call is synchronous.)
    void drawElementsAsync(DrawMode mode, size_t count, const int *indices) const
    {
        if (m_foreign) {
            detail::check(swr_process_elements(ctx(), (int)mode, count, indices, &VertexProcessor::emit, m_foreign), "drawElements");
            return;
        }
        detail::check(swr_draw_elements(ctx(), (int)mode, count, indices), "drawElements");
    }

    /// Additive: bytes handed to swr::uniforms<T>() in the vertex shader (same block as Rasterizer::setUniforms when the
    /// rasterizer is a swr::Rasterizer).
    void setUniforms(const void *data, size_t bytes) { detail::check(swr_set_uniforms(ctx(), data, bytes), "setUniforms"); }

private:
    struct AttribState { const void *buffer; int stride; size_t bytes; bool set; };

    // IRasterizer::draw*List of the user's rasterizer, one call per batch (VertexProcessor.cpp:302-317)
    static void emit(void *user, int mode, const void *vertices, size_t, const int32_t *indices, size_t indexCount)
    {
        const IRasterizer *r = static_cast<const IRasterizer *>(user);
        const RasterizerVertex *v = static_cast<const RasterizerVertex *>(vertices);
        static_assert(sizeof(int) == sizeof(int32_t), "index type");
        const int *idx = reinterpret_cast<const int *>(indices);
        if (mode == SWR_DRAW_TRIANGLE) r->drawTriangleList(v, idx, indexCount);
        else if (mode == SWR_DRAW_LINE) r->drawLineList(v, idx, indexCount);
        else r->drawPointList(v, idx, indexCount);
    }

    // the vertex-stage state lives in this object (as in the reference) and follows it to another rasterizer
    void apply()
    {
        if (m_haveViewport) detail::check(swr_set_viewport(ctx(), m_vp[0], m_vp[1], m_vp[2], m_vp[3]), "setViewport");
        detail::check(swr_set_depth_range(ctx(), m_depthN, m_depthF), "setDepthRange");
        detail::check(swr_set_cull_mode(ctx(), (int)m_cull), "setCullMode");
        if (m_vs) detail::check(swr_set_vertex_shader(ctx(), m_vs), "setVertexShader");
        for (int i = 0; i < MaxVertexAttribs; ++i)
            if (m_attribs[i].set)
                detail::check(swr_set_vertex_attrib_pointer(ctx(), i, m_attribs[i].stride, m_attribs[i].buffer, m_attribs[i].bytes), "setVertexAttribPointer");
    }

    swr_context *ctx() const { return m_rasterizer ? m_rasterizer->context() : m_own; }
    Rasterizer *m_rasterizer;
    IRasterizer *m_foreign;
    swr_context *m_own;
    const swr_vertex_shader *m_vs;
    AttribState m_attribs[MaxVertexAttribs];
    int m_vp[4];
    bool m_haveViewport;
    float m_depthN, m_depthF;
    CullMode m_cull;
};

} // namespace swr
