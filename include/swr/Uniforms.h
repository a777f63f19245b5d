// swr/Uniforms.h -- the framework-provided uniform block (additive; see VertexShaderBase.h).
//
// Each translation unit that instantiates shaders gets its own SWR_MAX_UNIFORM_BYTES block of
// __constant__ memory; Rasterizer/VertexProcessor::setUniforms (swr_set_uniforms in the C ABI)
// copies the caller's bytes into it before the next draw.
#pragma once

#include "detail/common.h"

#if defined(__CUDACC__)
namespace swr {
namespace detail {
static __constant__ unsigned char g_uniformBlock[SWR_MAX_UNIFORM_BYTES];

// static: every translation unit must get its OWN copy bound to its own g_uniformBlock (an inline
// function would be merged across TUs by the linker and write a single block).
static int uploadUniforms(const void *data, size_t bytes, void *stream)
{
    if (bytes > SWR_MAX_UNIFORM_BYTES) return -1;
    return cudaMemcpyToSymbolAsync(g_uniformBlock, data, bytes, 0, cudaMemcpyHostToDevice, (cudaStream_t)stream) == cudaSuccess ? 0 : -2;
}
} // namespace detail

/// The current uniform block viewed as T (T must be trivially copyable and <= SWR_MAX_UNIFORM_BYTES).
template <class T>
SWR_D const T &uniforms()
{
    static_assert(sizeof(T) <= SWR_MAX_UNIFORM_BYTES, "uniform block too large");
    return *reinterpret_cast<const T *>(detail::g_uniformBlock);
}
} // namespace swr
#endif
