// swr/PixelData.h -- the fragment record handed to PixelShader::drawPixel.
//
// Reference: src/renderer/PixelData.h:38-58.  Field names and meaning are the reference's:
// x, y (pixel), z, w, invw, avar[], pvar[], pvarTemp[], equations.  z is valid only if the shader
// sets InterpolateZ; w / invw only if InterpolateW or PVarCount > 0 (PixelData.h:64-71).  The
// values are bit-identical to the reference's incremental chains (PixelData.h:61-125): the tile
// kernel replays the same fp32 additions in the same order for every fragment.
//
// Additive members (not in the reference): primitiveOrdinal and the staged render-target view
// used by swr::target<T>().
#pragma once

#include "TriangleEquations.h"

namespace swr {

struct PixelData {
    int x; ///< The x coordinate.
    int y; ///< The y coordinate.

    float z;    ///< The interpolated z value.
    float w;    ///< The interpolated w value.
    float invw; ///< The interpolated 1 / w value.

    /// Affine variables.
    float avar[MaxAVars];
    /// Perspective variables.
    float pvar[MaxPVars];
    // Used internally.
    float pvarTemp[MaxPVars];

    /// Triangle equations needed for derivative computation (nullptr for lines and points).
    const TriangleEquations *equations;

    // ---- additive -------------------------------------------------------------------------
    /// Emission ordinal of the primitive: batch * SWR_ORDINAL_STRIDE + slot in the batch's output list.
    unsigned primitiveOrdinal;
    /// Staged render targets of the current tile (see swr::target).
    char *rtBase;
    int rtSlotStride;
    int rtOffset;

    /// Derivatives of a perspective-correct variable (PixelData.h:128-145).
    SWR_HD void computePerspectiveDerivatives(const TriangleEquations &eqn, int varIndex, float &ddx, float &ddy) const
    {
        using namespace detail;
        ParameterEquation pv = eqn.pvar[varIndex];
        float var = pv.evaluate(fadd(i2f(x), 0.5f), fadd(i2f(y), 0.5f));
        float curInvw = eqn.invw.evaluate(fadd(i2f(x), 0.5f), fadd(i2f(y), 0.5f));
        float dvar_dx = pv.a, dvar_dy = pv.b;
        float dinvw_dx = eqn.invw.a, dinvw_dy = eqn.invw.b;
        ddx = fdiv(fsub(fmul(curInvw, dvar_dx), fmul(var, dinvw_dx)), fmul(curInvw, curInvw));
        ddy = fdiv(fsub(fmul(curInvw, dvar_dy), fmul(var, dinvw_dy)), fmul(curInvw, curInvw));
    }
};

/// This fragment's pixel in registered render target `slot` (swr_set_render_target /
/// Rasterizer::setRenderTarget).  Slots below PixelShader::RenderTargets are staged in shared
/// memory for the duration of a tile; the reference resolves to the staged copy.  T must be a
/// 32-bit type.
template <class T>
SWR_D T &target(const PixelData &p, int slot)
{
    static_assert(sizeof(T) == 4, "render targets are 32 bits per pixel");
    return *reinterpret_cast<T *>(p.rtBase + (size_t)slot * (size_t)p.rtSlotStride + (size_t)p.rtOffset);
}

} // namespace swr
