// swr/TriangleEquations.h -- what PixelData::equations points at inside drawPixel.
//
// Reference: src/renderer/TriangleEquations.h:35-72.  Member names are the reference's
// (area2, e0..e2, z, invw, avar[i], pvar[i]); on the GPU the object is a per-fragment view
// whose avar / pvar arrays read the triangle's plane record in HBM on demand, so a shader that
// never touches p.equations pays nothing.  Setup itself (TriangleEquations' constructor) runs
// once per triangle in the geometry kernel (detail/geometry.cuh, setupTriangle).
#pragma once

#include "ParameterEquation.h"

namespace swr {

/// Read-only array view over consecutive planes stored as (a, b, c, 0).
struct ParameterEquationArray {
    const float *planes;
    SWR_HD ParameterEquation operator[](int i) const
    {
        ParameterEquation p;
        p.a = planes[4 * i + 0];
        p.b = planes[4 * i + 1];
        p.c = planes[4 * i + 2];
        return p;
    }
};

struct TriangleEquations {
    float area2;
    EdgeEquation e0;
    EdgeEquation e1;
    EdgeEquation e2;
    ParameterEquation z;
    ParameterEquation invw;
    ParameterEquationArray avar;
    ParameterEquationArray pvar;
};

} // namespace swr
