// swr/PixelShaderBase.h -- CRTP base for pixel shaders (reference: src/renderer/PixelShaderBase.h:41-137).
//
// Derive, redefine the traits and drawPixel exactly as with the reference.  GPU differences:
//   * drawPixel is a  static __device__  function (it is inlined into the tile kernel);
//   * the reference's drawBlock<TestEdges> / drawSpan walkers (PixelShaderBase.h:55-112) are not
//     members here: the tile kernel (detail/tile.cuh) generates the same fragments with the same
//     incremental fp32 values and calls drawPixel for each, in the reference's per-pixel order;
//   * RenderTargets (additive): how many registered render-target slots, starting at 0, the
//     shader reaches through swr::target<T>(p, slot).  They are staged in shared memory per
//     screen tile and written back with 128-bit stores.  A shader may still write raw global
//     pointers it got through uniforms, like the reference's shaders do.
#pragma once

#include "PixelData.h"

namespace swr {

template <class Derived>
class PixelShaderBase {
public:
    /// Tells the rasterizer to interpolate the z component.
    static const int InterpolateZ = false;
    /// Tells the rasterizer to interpolate the w component.
    static const int InterpolateW = false;
    /// Tells the rasterizer how many affine vars to interpolate.
    static const int AVarCount = 0;
    /// Tells the rasterizer how many perspective vars to interpolate.
    static const int PVarCount = 0;
    /// Additive: number of staged render-target slots (see swr::target).
    static const int RenderTargets = 0;

    /// This is called per pixel. Implement this in your derived class to display single pixels.
    SWR_D static void drawPixel(const PixelData &p) { (void)p; }
};

class NullPixelShader : public PixelShaderBase<NullPixelShader> {};

} // namespace swr
