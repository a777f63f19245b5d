// swr/VertexShaderBase.h -- CRTP base for vertex shaders (reference: src/renderer/VertexShaderBase.h:33-53).
//
// Derive and redefine AttribCount / AVarCount / PVarCount and processVertex exactly as with the
// reference; the only source-level differences for GPU shaders are
//   * processVertex is a  static __device__  function,
//   * "uniforms" cannot be static data members (nvcc rejects memory-space qualifiers on data
//     members): read them with swr::uniforms<T>() (see Uniforms.h) or from your own
//     namespace-scope __constant__ / __device__ variables.
#pragma once

#include "VertexConfig.h"

namespace swr {

template <class Derived>
class VertexShaderBase {
public:
    /// Number of vertex attribute pointers this vertex shader uses.
    static const int AttribCount = 0;
    /// Number of affine output variables.
    static const int AVarCount = 0;
    /// Number of perspective correct output variables.
    static const int PVarCount = 0;

    /// Process a single vertex. Implement this in your own vertex shader.
    SWR_D static void processVertex(VertexShaderInput in, VertexShaderOutput *out) { (void)in; (void)out; }
};

class DummyVertexShader : public VertexShaderBase<DummyVertexShader> {};

} // namespace swr
