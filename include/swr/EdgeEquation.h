// swr/EdgeEquation.h -- reference: src/renderer/EdgeEquation.h:31-86 (P8, P10 of the parity ledger).
#pragma once

#include "IRasterizer.h"

namespace swr {

struct EdgeEquation {
    float a;
    float b;
    float c;
    bool tie;

    SWR_HD void init(float v0x, float v0y, float v1x, float v1y)
    {
        using namespace detail;
        a = fsub(v0y, v1y);
        b = fsub(v1x, v0x);
        c = fmul(-fadd(fmul(a, fadd(v0x, v1x)), fmul(b, fadd(v0y, v1y))), 0.5f);   // x / 2 == x * 0.5 to the bit (EdgeEquation.h:44 divides by 2)
        tie = a != 0 ? a > 0 : b > 0;
    }
    SWR_HD void init(const RasterizerVertex &v0, const RasterizerVertex &v1) { init(v0.x, v0.y, v1.x, v1.y); }

    /// Evaluate the edge equation for the given point: (a*x + b*y) + c, no contraction.
    SWR_HD float evaluate(float x, float y) const
    {
        using namespace detail;
        return fadd(fadd(fmul(a, x), fmul(b, y)), c);
    }
    SWR_HD bool test(float x, float y) const { return test(evaluate(x, y)); }
    /// Top-left rule on an evaluated value.
    SWR_HD bool test(float v) const { return (v > 0 || (v == 0 && tie)); }
    SWR_HD float stepX(float v) const { return detail::fadd(v, a); }
    SWR_HD float stepX(float v, float stepSize) const { return detail::fadd(v, detail::fmul(a, stepSize)); }
    SWR_HD float stepY(float v) const { return detail::fadd(v, b); }
    SWR_HD float stepY(float v, float stepSize) const { return detail::fadd(v, detail::fmul(b, stepSize)); }
};

} // namespace swr
