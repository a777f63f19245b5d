// swr/VertexConfig.h -- vertex stage types (reference: src/renderer/VertexConfig.h:34-40).
#pragma once

#include "IRasterizer.h"

namespace swr {

/// Maximum supported number of vertex attributes.
const int MaxVertexAttribs = SWR_MAX_VERTEX_ATTRIBS;

/// Vertex shader output.
typedef RasterizerVertex VertexShaderOutput;

/// Vertex shader input is an array of vertex attribute pointers.
typedef const void *VertexShaderInput[MaxVertexAttribs];

} // namespace swr
