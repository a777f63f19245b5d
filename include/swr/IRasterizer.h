// swr/IRasterizer.h -- constants, the rasterizer vertex record and the IRasterizer interface.
// Same names, values and field order as the reference (src/renderer/IRasterizer.h:33-71).
#pragma once

#include <cstddef>
#include "detail/common.h"

namespace swr {

const int BlockSize = SWR_BLOCK_SIZE;      // IRasterizer.h:33

/// Maximum affine variables used for interpolation across the triangle.
const int MaxAVars = SWR_MAX_AVARS;        // IRasterizer.h:36

/// Maximum perspective variables used for interpolation across the triangle.
const int MaxPVars = SWR_MAX_PVARS;        // IRasterizer.h:39

/// Vertex input structure for the Rasterizer. Output from the VertexProcessor (IRasterizer.h:42-53).
struct RasterizerVertex {
    float x;
    float y;
    float z;
    float w;
    float avar[MaxAVars];
    float pvar[MaxPVars];
};

/// Interface for the rasterizer used by the VertexProcessor (IRasterizer.h:56-71).
/// Primitives whose first index is -1 are ignored.  Arrays may be host or device memory.
class IRasterizer {
public:
    virtual ~IRasterizer() {}
    virtual void drawPointList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const = 0;
    virtual void drawLineList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const = 0;
    virtual void drawTriangleList(const RasterizerVertex *vertices, const int *indices, size_t indexCount) const = 0;
};

} // namespace swr
