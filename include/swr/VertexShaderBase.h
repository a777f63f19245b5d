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

#if defined(__CUDACC__)
/// Optional helper for processVertex: reads one attribute struct with the widest loads its address allows
/// (128-bit when the struct size is a multiple of 16 and the address is 16-byte aligned, else 64-bit, else
/// member by member as `*static_cast<const T *>(p)` would).  Interleaved vertex structs of 24 / 32 bytes then
/// cost three / two load instructions instead of six / eight.
template <class T>
__device__ __forceinline__ T fetchAttrib(const void *p)
{
    T v;
    const unsigned long long a = (unsigned long long)p;
    if (sizeof(T) % 16 == 0 && (a & 15) == 0) {
#pragma unroll
        for (unsigned i = 0; i < sizeof(T) / 16; ++i) reinterpret_cast<uint4 *>(&v)[i] = static_cast<const uint4 *>(p)[i];
    } else if (sizeof(T) % 8 == 0 && (a & 7) == 0) {
#pragma unroll
        for (unsigned i = 0; i < sizeof(T) / 8; ++i) reinterpret_cast<uint2 *>(&v)[i] = static_cast<const uint2 *>(p)[i];
    } else {
        v = *static_cast<const T *>(p);
    }
    return v;
}
#endif

} // namespace swr
