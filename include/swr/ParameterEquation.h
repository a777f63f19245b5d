// swr/ParameterEquation.h -- reference: src/renderer/ParameterEquation.h:31-79 (P9).
#pragma once

#include "EdgeEquation.h"

namespace swr {

struct ParameterEquation {
    float a;
    float b;
    float c;

    SWR_HD void init(float p0, float p1, float p2, const EdgeEquation &e0, const EdgeEquation &e1, const EdgeEquation &e2, float factor)
    {
        using namespace detail;
        a = fmul(factor, fadd(fadd(fmul(p0, e0.a), fmul(p1, e1.a)), fmul(p2, e2.a)));
        b = fmul(factor, fadd(fadd(fmul(p0, e0.b), fmul(p1, e1.b)), fmul(p2, e2.b)));
        c = fmul(factor, fadd(fadd(fmul(p0, e0.c), fmul(p1, e1.c)), fmul(p2, e2.c)));
    }
    SWR_HD float evaluate(float x, float y) const
    {
        using namespace detail;
        return fadd(fadd(fmul(a, x), fmul(b, y)), c);
    }
    SWR_HD float stepX(float v) const { return detail::fadd(v, a); }
    SWR_HD float stepX(float v, float stepSize) const { return detail::fadd(v, detail::fmul(a, stepSize)); }
    SWR_HD float stepY(float v) const { return detail::fadd(v, b); }
    SWR_HD float stepY(float v, float stepSize) const { return detail::fadd(v, detail::fmul(b, stepSize)); }
};

} // namespace swr
