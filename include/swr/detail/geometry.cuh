// swr/detail/geometry.cuh -- the geometry kernel (K1+K2+K3 of SURVEY.md 2.3), templated on the
// user's vertex shader so processVertex is inlined.
//
// One CTA = one reference batch of 1024 input primitives (VertexProcessor.cpp:110-116), one
// thread = one primitive at a time (4 rounds of 256).  Per primitive, as a pure function:
//   fetch indices -> VS::processVertex per corner (VertexProcessor.cpp:97-105,134-150)
//   -> clip mask (VertexProcessor.cpp:122-132) -> Sutherland-Hodgman against the flagged planes in
//   the fixed order +X,-X,+Y,-Y,+Z,-Z (PolyClipper.cpp:45-81) or the parametric line clip
//   (LineClipper.cpp:30-56) -> perspective divide + viewport + depth range
//   (VertexProcessor.cpp:347-377) -> cull / re-orient (VertexProcessor.cpp:319-345)
//   -> TriangleEquations setup (TriangleEquations.h:47-71) + footprint box -> record in HBM.
// The footprint of a Block-mode triangle is its CERTIFIED PIXEL BOUNDS (tightPixelBounds below): the
// smallest pixel rectangle this file can prove to contain every fragment the reference's block walk
// produces; binning and coverage work on it instead of the reference's 8-aligned block box.
// The reference's 16-entry VertexCache only avoids re-shading (results are identical without
// it); here the three corner fetches of neighbouring primitives hit L1/L2 (a
// warp-cooperative de-duplication exists as a measured build variant, SWR_GEOM_DEDUP).
// With several ranks the kernel has a second form (MULTI): the batches are sharded over the ranks and
// every surviving record is written into the scratch of the ranks that own the tiles it touches --
// locally or through a peer mapping (NVLink) -- see RecordSink in common.h.
// All parity-critical arithmetic goes through fmul/fadd/fsub/fdiv (never contracted).
#pragma once

#include "common.h"
#include "../VertexShaderBase.h"
#include "../ParameterEquation.h"
#include "../Uniforms.h"

namespace swr {
namespace detail {

template <int NA, int NP>
struct CVert {
    float x, y, z, w;
    float a[NA > 0 ? NA : 1];
    float p[NP > 0 ? NP : 1];
};

// VertexProcessor.cpp:122-132 -- strict '<' on w-x, x+w, ...
SWR_HD int outcode(float x, float y, float z, float w)
{
    int m = 0;
    if (fsub(w, x) < 0) m |= 0x01;
    if (fadd(x, w) < 0) m |= 0x02;
    if (fsub(w, y) < 0) m |= 0x04;
    if (fadd(y, w) < 0) m |= 0x08;
    if (fsub(w, z) < 0) m |= 0x10;
    if (fadd(z, w) < 0) m |= 0x20;
    return m;
}

// a*x + b*y + c*z + d*w, left to right, with the literal plane coefficients of
// VertexProcessor.cpp:187-192 / 237-242 (PolyClipper.cpp:56,62, LineClipper.cpp:35-36).
SWR_HD float planeDist(int plane, float x, float y, float z, float w)
{
    float a = 0.0f, b = 0.0f, c = 0.0f;
    const float d = 1.0f;
    switch (plane) {
    case 0: a = -1.0f; break;
    case 1: a = 1.0f; break;
    case 2: b = -1.0f; break;
    case 3: b = 1.0f; break;
    case 4: c = -1.0f; break;
    default: c = 1.0f; break;
    }
    return fadd(fadd(fadd(fmul(a, x), fmul(b, y)), fmul(c, z)), fmul(d, w));
}

// PolyClipper.h:34-48 -- v0*(1-t) + v1*t on x,y,z,w and the VERTEX shader's variable counts.
template <int NA, int NP>
SWR_HD void lerpVert(CVert<NA, NP> &o, const CVert<NA, NP> &v0, const CVert<NA, NP> &v1, float t)
{
    const float s = fsub(1.0f, t);
    o.x = fadd(fmul(v0.x, s), fmul(v1.x, t));
    o.y = fadd(fmul(v0.y, s), fmul(v1.y, t));
    o.z = fadd(fmul(v0.z, s), fmul(v1.z, t));
    o.w = fadd(fmul(v0.w, s), fmul(v1.w, t));
#pragma unroll
    for (int i = 0; i < NA; ++i) o.a[i] = fadd(fmul(v0.a[i], s), fmul(v1.a[i], t));
#pragma unroll
    for (int i = 0; i < NP; ++i) o.p[i] = fadd(fmul(v0.p[i], s), fmul(v1.p[i], t));
}

SWR_HD int sgn3(float v) { return (0.0f < v) - (v < 0.0f); }   // PolyClipper.h:91-94

// One Sutherland-Hodgman pass (PolyClipper.cpp:45-81).  Returns the new vertex count; -1 when
// the polygon would exceed kMaxPoly.
template <int NA, int NP>
SWR_HD int clipPolyPlane(const CVert<NA, NP> *in, int n, CVert<NA, NP> *out, int plane)
{
    int m = 0;
    int iprev = 0;
    float dprev = planeDist(plane, in[0].x, in[0].y, in[0].z, in[0].w);
    for (int i = 1; i <= n; ++i) {
        const int icur = (i == n) ? 0 : i;
        const float d = planeDist(plane, in[icur].x, in[icur].y, in[icur].z, in[icur].w);
        if (dprev >= 0) {
            if (m >= kMaxPoly) return -1;
            out[m++] = in[iprev];
        }
        if (sgn3(d) != sgn3(dprev)) {
            const float t = d < 0 ? fdiv(dprev, fsub(dprev, d)) : fdiv(-dprev, fsub(d, dprev));
            if (m >= kMaxPoly) return -1;
            lerpVert(out[m++], in[iprev], in[icur], t);
        }
        iprev = icur;
        dprev = d;
    }
    return m;
}

// Clip a triangle against the planes flagged in `mask`.  bufA holds the 3 input vertices; the
// result is in *res (bufA or bufB).  Returns the polygon size (0 when fully clipped, or on overflow of
// kMaxPoly vertices, which *overflow then reports).
template <int NA, int NP>
SWR_HD int clipTriangle(CVert<NA, NP> *bufA, CVert<NA, NP> *bufB, int mask, CVert<NA, NP> **res, bool *overflow = nullptr)
{
    CVert<NA, NP> *in = bufA, *out = bufB;
    int n = 3;
    for (int pl = 0; pl < 6; ++pl) {
        if (!(mask & (1 << pl))) continue;
        if (n < 3) break;                                   // PolyClipper.cpp:47-48
        n = clipPolyPlane(in, n, out, pl);
        if (n < 0) { n = 0; if (overflow) *overflow = true; break; }   // > kMaxPoly vertices: dropped and reported (swr_finish)
        CVert<NA, NP> *t = in; in = out; out = t;
    }
    *res = in;
    return n < 3 ? 0 : n;                                   // VertexProcessor.cpp:244-250
}

// VertexProcessor.cpp:347-377 -- perspective divide, viewport (y flipped), depth range.  w stays.
template <int NA, int NP>
SWR_HD void toScreen(const GeomArgs &g, CVert<NA, NP> &v)
{
    const float invW = frcp(v.w);
    v.x = fmul(v.x, invW);
    v.y = fmul(v.y, invW);
    v.z = fmul(v.z, invW);
    v.x = fadd(fmul(g.px, v.x), g.ox);
    v.y = fadd(fmul(g.py, -v.y), g.oy);
    v.z = fadd(fmul(fmul(0.5f, fsub(g.depthF, g.depthN)), v.z), fmul(0.5f, fadd(g.depthN, g.depthF)));
}

SWR_HD float min3f(float a, float b, float c) { float m = b < a ? b : a; return c < m ? c : m; }   // std::min nesting
SWR_HD float max3f(float a, float b, float c) { float m = a < b ? b : a; return m < c ? c : m; }   // std::max nesting
SWR_HD int imin(int a, int b) { return a < b ? a : b; }
SWR_HD int imax(int a, int b) { return a > b ? a : b; }
SWR_HD int16_t clamp16(int v) { return (int16_t)imin(imax(v, -32768), 32767); }

SWR_HD Box16 makeBox(int x0, int y0, int x1, int y1)
{
    // clamp to the addressable screen; empty after clamping => dead
    x0 = imax(x0, 0); y0 = imax(y0, 0);
    x1 = imin(x1, 32767); y1 = imin(y1, 32767);
    if (x0 > x1 || y0 > y1) return deadBox();
    Box16 b; b.x0 = (int16_t)x0; b.y0 = (int16_t)y0; b.x1 = (int16_t)x1; b.y1 = (int16_t)y1;
    return b;
}

// One scan-converted half of a triangle (Rasterizer.h:360-411): rows [y0, y1), per row
// x = (vx + inv * dy) + 0.5 with dy = (row - vy) + 0.5.
struct SpanHalf { float vx, vy, inv1, inv2; int y0, y1; };

SWR_HD float spanX(float vx, float vy, float inv, int row)
{
    const float dy = fadd(fsub(i2f(row), vy), 0.5f);
    return fadd(fadd(vx, fmul(inv, dy)), 0.5f);
}

// Rasterizer.h:318-358: sort by y, split at the middle vertex, left/right by x.
SWR_HD void spanSetup(float x0, float y0, float x1, float y1, float x2, float y2, SpanHalf &bot, SpanHalf &top)
{
    float tx = x0, ty = y0, mx = x1, my = y1, bx = x2, by = y2, f;
    if (ty > my) { f = tx; tx = mx; mx = f; f = ty; ty = my; my = f; }
    if (my > by) { f = mx; mx = bx; bx = f; f = my; my = by; by = f; }
    if (ty > my) { f = tx; tx = mx; mx = f; f = ty; ty = my; my = f; }
    const float dy = fsub(by, ty);
    const float iy = fsub(my, ty);
    bot.vx = bot.vy = bot.inv1 = bot.inv2 = 0.0f; bot.y0 = bot.y1 = 0;
    top = bot;
    float lx, ly, rx, ry;
    bool hasBot = true, hasTop = true;
    if (my == ty) {                      // flat top: l = m, r = t (Rasterizer.h:330-335)
        lx = mx; ly = my; rx = tx; ry = ty;
        hasBot = false;
    } else if (my == by) {               // flat bottom: l = m, r = b (Rasterizer.h:336-341)
        lx = mx; ly = my; rx = bx; ry = by;
        hasTop = false;
    } else {                             // general: split vertex v4 on the long edge (Rasterizer.h:344-350)
        lx = mx; ly = my;
        ry = my;
        rx = fadd(tx, fmul(fdiv(fsub(bx, tx), dy), iy));
    }
    if (lx > rx) { f = lx; lx = rx; rx = f; f = ly; ly = ry; ry = f; }
    if (hasBot) {                        // drawBottomFlatTriangle(eqn, t, l, r)  (Rasterizer.h:360-385)
        bot.vx = tx; bot.vy = ty;
        bot.inv1 = fdiv(fsub(lx, tx), fsub(ly, ty));
        bot.inv2 = fdiv(fsub(rx, tx), fsub(ry, ty));
        bot.y0 = f2i(fadd(ty, 0.5f));
        bot.y1 = f2i(fadd(ly, 0.5f));
    }
    if (hasTop) {                        // drawTopFlatTriangle(eqn, l, r, b)  (Rasterizer.h:387-411)
        top.vx = bx; top.vy = by;
        top.inv1 = fdiv(fsub(bx, lx), fsub(by, ly));
        top.inv2 = fdiv(fsub(bx, rx), fsub(by, ry));
        top.y0 = f2i(fsub(ly, 0.5f)) + 1;                   // rows int(v0.y-.5) < row <= int(v2.y-.5)
        top.y1 = f2i(fsub(by, 0.5f)) + 1;
    }
}

// Pixel range [xl, xr) of one row of one half, X-clamped to the scissor (Rasterizer.h:376-378).
SWR_HD void spanRow(const SpanHalf &h, int row, int scMinX, int scMaxX, int &xl, int &xr)
{
    xl = imax(scMinX, f2i(spanX(h.vx, h.vy, h.inv1, row)));
    xr = imin(scMaxX, f2i(spanX(h.vx, h.vy, h.inv2, row)));
}

// Exact footprint of the two halves: the row-wise x is monotone in the row (every fp32 step of
// spanX is monotone), so the extremes sit on the first / last row of each half.
SWR_HD Box16 spanBox(const SpanHalf &bot, const SpanHalf &top, int scMinX, int scMaxX)
{
    int x0 = 32767, x1 = -32768, y0 = 32767, y1 = -32768;
    const SpanHalf *hs[2] = { &bot, &top };
    for (int k = 0; k < 2; ++k) {
        const SpanHalf &h = *hs[k];
        if (h.y0 >= h.y1) continue;
        int xl, xr;
        spanRow(h, h.y0, scMinX, scMaxX, xl, xr);
        x0 = imin(x0, xl); x1 = imax(x1, xr - 1);
        spanRow(h, h.y1 - 1, scMinX, scMaxX, xl, xr);
        x0 = imin(x0, xl); x1 = imax(x1, xr - 1);
        y0 = imin(y0, h.y0); y1 = imax(y1, h.y1 - 1);
    }
    return makeBox(x0, y0, x1, y1);
}

SWR_HD float4 mkf4(float a, float b, float c, float d) { float4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }
SWR_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
SWR_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}

// Certified pixel bounds of a Block-mode triangle.  The reference visits every 8x8 block of the truncated,
// 8-aligned bounding box (Rasterizer.h:234-255) and tests each pixel with incremental fp32 chains
// (EdgeData.h:37-66); most of those pixels can be excluded up front without changing a single result.
// With u = 2^-24 (half an ulp, relative) and everything below in REAL arithmetic on the stored fp32 coefficients:
//
//   (1) every value the reference compares for edge k at a pixel -- the per-pixel chain value as well as the four
//       block-corner values of Rasterizer.h:272-275 -- is E_k(centre) = a_k x + b_k y + c_k plus at most 18
//       roundings of values bounded by M_k = |a_k| X + |b_k| Y + |c_k| (X, Y = the box's far pixel edge), so it
//       differs from E_k by less than 32 u M_k.  A pixel (or block corner) that passes test k has E_k > -32 u M_k.
//   (2) the stored line misses the two vertices p, q it was built from (EdgeEquation.h:38-46) by at most
//       |E_k(p)|, |E_k(q)| <= 2u (|a||b| + |a|(|p.x|+|q.x|) + |b|(|p.y|+|q.y|))   (a, b: one rounding each;
//       c: three roundings of the two products and their sum).
//   (3) so every fragment lies in the triangle spanned by the lines E_k = -32 u M_k, whose corner at the edges
//       i, j is displaced from the vertex the two edges share by a vector s with |n_i.s| <= D_i, |n_j.s| <= D_j,
//       D_k = (1) + (2); Cramer's rule bounds |s.x| <= (D_i |b_j| + D_j |b_i|) / det, |s.y| likewise, with
//       det = a_i b_j - a_j b_i > 0 taken at its fp32 lower bound.
// The bounds are the vertices' bounding box grown by the largest displacement and 0.02 pixel (which covers the
// fp32 roundings of this function itself).  Pixels outside can pass no edge test of the reference and no block they
// lie in can be classified "fully covered", so restricting binning and coverage to the bounds is exact.  Returns
// false when nothing can be certified (degenerate / non-finite coefficients): the caller keeps the block box.
SWR_HD float fabs32(float v) { return v < 0 ? -v : v; }

SWR_HD bool tightPixelBounds(const EdgeEquation &e0, const EdgeEquation &e1, const EdgeEquation &e2,
                             float v0x, float v0y, float v1x, float v1y, float v2x, float v2y,
                             float fminX, float fminY, float fmaxX, float fmaxY,
                             int bx0, int by0, int bx1, int by1, int &x0, int &y0, int &x1, int &y1)
{
    const float X = i2f(bx1 + 1), Y = i2f(by1 + 1);
    const float a[3] = { fabs32(e0.a), fabs32(e1.a), fabs32(e2.a) };
    const float b[3] = { fabs32(e0.b), fabs32(e1.b), fabs32(e2.b) };
    const float c[3] = { fabs32(e0.c), fabs32(e1.c), fabs32(e2.c) };
    // end points of the edges: e0 = (v1, v2), e1 = (v2, v0), e2 = (v0, v1)   (emitScreenTriangle)
    const float sx[3] = { fabs32(v1x) + fabs32(v2x), fabs32(v2x) + fabs32(v0x), fabs32(v0x) + fabs32(v1x) };
    const float sy[3] = { fabs32(v1y) + fabs32(v2y), fabs32(v2y) + fabs32(v0y), fabs32(v0y) + fabs32(v1y) };
    const float kEval = 1.0f / 524288.0f;        // 32 u = 2^-19
    const float kLine = 1.0f / 8388608.0f;       // 2 u = 2^-23
    const float kUp = 1.0f + 1.0f / 65536.0f;    // absorbs the roundings of the bound's own arithmetic
    float D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float M = a[k] * X + b[k] * Y + c[k];
        const float L = a[k] * b[k] + a[k] * sx[k] + b[k] * sy[k];
        D[k] = (M * kEval + L * kLine) * kUp + 1e-30f;
    }
    const float sa[3] = { e0.a, e1.a, e2.a }, sb[3] = { e0.b, e1.b, e2.b };
    float dx = 0.0f, dy = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = i == 2 ? 0 : i + 1;
        const float p1 = sa[i] * sb[j], p2 = sa[j] * sb[i];
        const float det = (p1 - p2) - (fabs32(p1) + fabs32(p2)) * (1.0f / 4194304.0f);     // lower bound: - 4u (|p1| + |p2|)
        if (!(det > 1e-30f)) return false;
        const float r = kUp / det;
        const float ddx = (D[i] * b[j] + D[j] * b[i]) * r, ddy = (D[i] * a[j] + D[j] * a[i]) * r;
        dx = ddx > dx ? ddx : dx;
        dy = ddy > dy ? ddy : dy;
    }
    if (!(dx < 64.0f) || !(dy < 64.0f)) return false;            // (also catches NaN) nothing worth certifying
    // pixel p has its centre at p + 0.5: it can hold a fragment only if  min - d <= p + 0.5 <= max + d
    const float lx0 = fminX - dx - 0.52f, ly0 = fminY - dy - 0.52f, lx1 = fmaxX + dx - 0.48f, ly1 = fmaxY + dy - 0.48f;
    if (!(lx0 > -1e9f) || !(ly0 > -1e9f) || !(lx1 < 1e9f) || !(ly1 < 1e9f)) return false;
    // ceil / floor, clamped to the reference's block box [bx0, bx1] x [by0, by1]
    const float fx0 = i2f(bx0), fy0 = i2f(by0), fx1 = i2f(bx1), fy1 = i2f(by1);
    int px0 = bx0, py0 = by0, px1 = bx1, py1 = by1;
    if (lx0 > fx0) { if (lx0 > fx1) px0 = bx1 + 1; else { const int t = f2i(lx0); px0 = i2f(t) < lx0 ? t + 1 : t; } }
    if (ly0 > fy0) { if (ly0 > fy1) py0 = by1 + 1; else { const int t = f2i(ly0); py0 = i2f(t) < ly0 ? t + 1 : t; } }
    if (lx1 < fx1) { if (lx1 < fx0) px1 = bx0 - 1; else px1 = f2i(lx1); }
    if (ly1 < fy1) { if (ly1 < fy0) py1 = by0 - 1; else py1 = f2i(ly1); }
    x0 = px0; y0 = py0; x1 = px1; y1 = py1;
    return true;
}

// ---- records ---------------------------------------------------------------------------------------
// A primitive is set up in registers first (its footprint decides whether, and to which ranks, a record is
// written at all) and stored once its position in the destination's compacted record range is known.

constexpr int kMaxPlanes = 2;        // z, 1/w in front of the avar / pvar planes

// TriangleEquations (TriangleEquations.h:47-71) of one screen-space triangle plus its footprint.
template <int NA, int NP>
struct TriRecord {
    float4 h0, h1, h2;                               // edges, flags, ordinal, area2 (layout of `head`)
    float pa[kMaxPlanes + NA + NP], pb[kMaxPlanes + NA + NP], pc[kMaxPlanes + NA + NP];   // planes: [0] z, [1] 1/w, avar, pvar
    SpanHalf bot, top;                               // Span / Adaptive only
    bool span;
};

// Screen-space triangle: cull / re-orient (VertexProcessor.cpp:319-345), setup (TriangleEquations.h:47-71),
// footprint (the certified pixel bounds of the reference's block walk, Rasterizer.h:234-255, or the span halves).
// Returns the footprint box (dead when the triangle is dropped anywhere on the way).
template <int NA, int NP, bool SPAN = true>
SWR_HD Box16 setupScreenTriangle(const GeomArgs &g, uint32_t ordinal, bool doCull,
                                 const CVert<NA, NP> &s0, const CVert<NA, NP> &s1, const CVert<NA, NP> &s2, TriRecord<NA, NP> &R)
{
    const CVert<NA, NP> *v0 = &s0, *v1 = &s1, *v2 = &s2;
    if (doCull) {
        const float facing = fsub(fmul(fsub(s0.x, s1.x), fsub(s2.y, s1.y)), fmul(fsub(s2.x, s1.x), fsub(s0.y, s1.y)));
        if (facing < 0) {
            if (g.cullMode == SWR_CULL_CW) return deadBox();
        } else {
            if (g.cullMode == SWR_CULL_CCW) return deadBox();
            v0 = &s2; v2 = &s0;                             // std::swap(idx0, idx2)
        }
    }
    EdgeEquation e0 = {}, e1 = {}, e2 = {};
    e0.init(v1->x, v1->y, v2->x, v2->y);
    e1.init(v2->x, v2->y, v0->x, v0->y);
    e2.init(v0->x, v0->y, v1->x, v1->y);
    const float area2 = fadd(fadd(e0.c, e1.c), e2.c);
    if (area2 <= 0) return deadBox();                       // Rasterizer.h:231,315

    // raster mode of this triangle (Rasterizer.h:413-445; Adaptive's literals are doubles)
    bool span = SPAN && g.rasterMode == SWR_RASTER_SPAN;     // (SPAN == false: the caller knows the mode is Block)
    const float fminX = min3f(v0->x, v1->x, v2->x), fmaxX = max3f(v0->x, v1->x, v2->x);
    const float fminY = min3f(v0->y, v1->y, v2->y), fmaxY = max3f(v0->y, v1->y, v2->y);
    if (SPAN && g.rasterMode == SWR_RASTER_ADAPTIVE) {
        const float orient = fdiv(fsub(fmaxX, fminX), fsub(fmaxY, fminY));
        span = !((double)orient > 0.4 && (double)orient < 1.6);
    }

    Box16 box;
    R.span = span;
    if (span) {
        spanSetup(v0->x, v0->y, v1->x, v1->y, v2->x, v2->y, R.bot, R.top);
        box = spanBox(R.bot, R.top, g.scMinX, g.scMaxX);
        if (box.x0 > box.x1) return box;
    } else {
        int minX = f2i(fminX), maxX = f2i(fmaxX), minY = f2i(fminY), maxY = f2i(fmaxY);
        minX = imax(minX, g.scMinX); maxX = imin(maxX, g.scMaxX);      // max stays the exclusive edge (P11)
        minY = imax(minY, g.scMinY); maxY = imin(maxY, g.scMaxY);
        minX &= ~7; maxX &= ~7; minY &= ~7; maxY &= ~7;
        const int stepsX = (maxX - minX) / 8 + 1, stepsY = (maxY - minY) / 8 + 1;
        if (stepsX <= 0 || stepsY <= 0) return deadBox();            // reference loop does not run (P23 aside)
        int tx0 = minX, ty0 = minY, tx1 = maxX + 7, ty1 = maxY + 7;
        if (!g.noTightBox && minX >= 0 && minY >= 0 && maxX + 7 <= 32767 && maxY + 7 <= 32767)
            tightPixelBounds(e0, e1, e2, v0->x, v0->y, v1->x, v1->y, v2->x, v2->y, fminX, fminY, fmaxX, fmaxY,
                             minX, minY, maxX + 7, maxY + 7, tx0, ty0, tx1, ty1);
        box = makeBox(tx0, ty0, tx1, ty1);                            // a sub-pixel triangle that misses every pixel centre dies here
        if (box.x0 > box.x1) return box;
    }

    const uint32_t flags = (e0.tie ? kTie0 : 0u) | (e1.tie ? kTie1 : 0u) | (e2.tie ? kTie2 : 0u) | (span ? kModeSpan : 0u);
    R.h0 = mkf4(e0.a, e0.b, e0.c, e1.a);
    R.h1 = mkf4(e1.b, e1.c, e2.a, e2.b);
    R.h2 = mkf4(e2.c, u2f(flags), u2f(ordinal), area2);

    // interpolation planes (TriangleEquations.h:59-70)
    const float factor = frcp(area2);
    ParameterEquation pe;
    if (g.useZ) {
        pe.init(v0->z, v1->z, v2->z, e0, e1, e2, factor);
        R.pa[0] = pe.a; R.pb[0] = pe.b; R.pc[0] = pe.c;
    }
    float iw0 = 0.0f, iw1 = 0.0f, iw2 = 0.0f;
    if (g.useW || g.nP > 0) {
        iw0 = frcp(v0->w); iw1 = frcp(v1->w); iw2 = frcp(v2->w);
        pe.init(iw0, iw1, iw2, e0, e1, e2, factor);
        R.pa[1] = pe.a; R.pb[1] = pe.b; R.pc[1] = pe.c;
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (i < g.nA) {
            pe.init(v0->a[i], v1->a[i], v2->a[i], e0, e1, e2, factor);
            R.pa[kMaxPlanes + i] = pe.a; R.pb[kMaxPlanes + i] = pe.b; R.pc[kMaxPlanes + i] = pe.c;
        }
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        if (i < g.nP) {
            pe.init(fmul(v0->p[i], iw0), fmul(v1->p[i], iw1), fmul(v2->p[i], iw2), e0, e1, e2, factor);
            R.pa[kMaxPlanes + NA + i] = pe.a; R.pb[kMaxPlanes + NA + i] = pe.b; R.pc[kMaxPlanes + NA + i] = pe.c;
        }
    }
    return box;
}

// Write a set-up triangle as record `rec` of `sink`: head, planes in the order z?, invw?, avar[nA], pvar[nP]
// (one float4 (a, b, c, 0) each), span halves.
template <int NA, int NP>
SWR_HD void storeTriangle(const GeomArgs &g, const RecordSink &sink, uint32_t rec, const TriRecord<NA, NP> &R, float4 *spanDst = nullptr)
{
    float4 *hd = sink.head + (size_t)rec * 3;
    hd[0] = R.h0; hd[1] = R.h1; hd[2] = R.h2;
    float4 *pp4 = reinterpret_cast<float4 *>(sink.params + (size_t)rec * g.paramStride);
    if (g.useZ) *pp4++ = mkf4(R.pa[0], R.pb[0], R.pc[0], 0.0f);
    if (g.useW || g.nP > 0) *pp4++ = mkf4(R.pa[1], R.pb[1], R.pc[1], 0.0f);
#pragma unroll
    for (int i = 0; i < NA; ++i)
        if (i < g.nA) *pp4++ = mkf4(R.pa[kMaxPlanes + i], R.pb[kMaxPlanes + i], R.pc[kMaxPlanes + i], 0.0f);
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i < g.nP) *pp4++ = mkf4(R.pa[kMaxPlanes + NA + i], R.pb[kMaxPlanes + NA + i], R.pc[kMaxPlanes + NA + i], 0.0f);
    if (R.span) {
        float4 *sp = spanDst ? spanDst : sink.span + (size_t)rec * 3;
        sp[0] = mkf4(R.bot.vx, R.bot.vy, R.bot.inv1, R.bot.inv2);
        sp[1] = mkf4(R.top.vx, R.top.vy, R.top.inv1, R.top.inv2);
        sp[2] = mkf4(u2f((uint32_t)R.bot.y0), u2f((uint32_t)R.bot.y1), u2f((uint32_t)R.top.y0), u2f((uint32_t)R.top.y1));
    }
}

// Clip-space fan triangle -> screen -> set-up record.
template <int NA, int NP, bool SPAN = true>
SWR_HD Box16 setupClipTriangle(const GeomArgs &g, uint32_t ordinal, CVert<NA, NP> a, CVert<NA, NP> b, CVert<NA, NP> c, TriRecord<NA, NP> &R)
{
    toScreen(g, a);
    toScreen(g, b);
    toScreen(g, c);
    return setupScreenTriangle<NA, NP, SPAN>(g, ordinal, true, a, b, c, R);
}

// One-call forms (sequential hosts: tests/hostcheck, the raster-list kernel): set up and store as record `rec`.
template <int NA, int NP>
SWR_HD Box16 emitScreenTriangle(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal, bool doCull,
                                const CVert<NA, NP> &s0, const CVert<NA, NP> &s1, const CVert<NA, NP> &s2)
{
    TriRecord<NA, NP> R;
    const Box16 box = setupScreenTriangle(g, ordinal, doCull, s0, s1, s2, R);
    if (box.x0 <= box.x1) storeTriangle(g, sink, rec, R);
    return box;
}
template <int NA, int NP>
SWR_HD Box16 emitClipTriangle(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal,
                              const CVert<NA, NP> &a, const CVert<NA, NP> &b, const CVert<NA, NP> &c)
{
    TriRecord<NA, NP> R;
    const Box16 box = setupClipTriangle(g, ordinal, a, b, c, R);
    if (box.x0 <= box.x1) storeTriangle(g, sink, rec, R);
    return box;
}

// Rasterizer.h:144-147
SWR_HD bool scissorTest(int minX, int minY, int maxX, int maxY, float x, float y)
{
    return x >= i2f(minX) && x < i2f(maxX) && y >= i2f(minY) && y < i2f(maxY);
}

constexpr int kMaxLineSteps = 1 << 17;   // longer DDA walks cannot touch a <= 32767-pixel screen meaningfully

// Screen-space line: footprint (Rasterizer.h:175-222).  steps = 0 with a dead box when nothing is drawn;
// *tooLong is set for walks beyond kMaxLineSteps (dropped, reported by swr_finish).
template <int NA, int NP>
SWR_HD Box16 setupScreenLine(const GeomArgs &g, const CVert<NA, NP> &v0, const CVert<NA, NP> &v1, int &steps, bool &tooLong)
{
    const int ix0 = f2i(v0.x), iy0 = f2i(v0.y), ix1 = f2i(v1.x), iy1 = f2i(v1.y);
    const int adx = ix1 > ix0 ? ix1 - ix0 : ix0 - ix1;
    const int ady = iy1 > iy0 ? iy1 - iy0 : iy0 - iy1;
    steps = imax(adx, ady);
    tooLong = false;
    if (steps <= 0) return deadBox();
    if (steps > kMaxLineSteps) { tooLong = true; return deadBox(); }
    // Fragments must pass the float scissor test, so the footprint is inside the scissor.  The walk adds the rounded
    // step `steps` times: each addition is off by at most half an ulp of the running coordinate, and the step itself by
    // half an ulp of its own value times the number of additions, so position k lies within
    // (steps + 1) * ulp(max |coordinate|) of the straight line between the end points -- the margin below.
    float mag = fabs32(v0.x);
    mag = fabs32(v0.y) > mag ? fabs32(v0.y) : mag;
    mag = fabs32(v1.x) > mag ? fabs32(v1.x) : mag;
    mag = fabs32(v1.y) > mag ? fabs32(v1.y) : mag;
    const float drift = i2f(steps + 1) * mag * (1.0f / 8388608.0f);        // (steps + 1) * 2^-23 * max|coordinate| >= (steps + 1) ulps
    const int margin = drift < 30000.0f ? f2i(drift) + 2 : 32767;
    return makeBox(imax(imin(ix0, ix1) - margin, g.scMinX), imax(imin(iy0, iy1) - margin, g.scMinY),
                   imin(imax(ix0, ix1) + margin, g.scMaxX - 1), imin(imax(iy0, iy1) + margin, g.scMaxY - 1));
}

// head0 = {x0, y0, stepx, stepy}, head1 = {steps, ordinal}; params = (start, step) pairs of z?, w?, avar[], pvar[].
template <int NA, int NP>
SWR_HD void storeLine(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal, int steps,
                      const CVert<NA, NP> &v0, const CVert<NA, NP> &v1)
{
    const float fs = i2f(steps);
    float4 *hd = sink.head + (size_t)rec * 3;
    hd[0] = mkf4(v0.x, v0.y, fdiv(fsub(v1.x, v0.x), fs), fdiv(fsub(v1.y, v0.y), fs));
    hd[1] = mkf4(u2f((uint32_t)steps), u2f(ordinal), 0.0f, 0.0f);
    float *pp = sink.params + (size_t)rec * g.paramStride;
    if (g.useZ) { pp[0] = v0.z; pp[1] = fdiv(fsub(v1.z, v0.z), fs); pp += 2; }
    if (g.useW) { pp[0] = v0.w; pp[1] = fdiv(fsub(v1.w, v0.w), fs); pp += 2; }
#pragma unroll
    for (int i = 0; i < NA; ++i)
        if (i < g.nA) { pp[0] = v0.a[i]; pp[1] = fdiv(fsub(v1.a[i], v0.a[i]), fs); pp += 2; }
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i < g.nP) { pp[0] = v0.p[i]; pp[1] = fdiv(fsub(v1.p[i], v0.p[i]), fs); pp += 2; }
}

// Screen-space point: footprint (Rasterizer.h:149-173).
template <int NA, int NP>
SWR_HD Box16 setupScreenPoint(const GeomArgs &g, const CVert<NA, NP> &v)
{
    if (!scissorTest(g.scMinX, g.scMinY, g.scMaxX, g.scMaxY, v.x, v.y)) return deadBox();
    const int ix = f2i(v.x), iy = f2i(v.y);
    return makeBox(ix, iy, ix, iy);
}

template <int NA, int NP>
SWR_HD void storePoint(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal, const CVert<NA, NP> &v)
{
    float4 *hd = sink.head + (size_t)rec * 3;
    hd[0] = mkf4(v.x, v.y, 0.0f, 0.0f);
    hd[1] = mkf4(0.0f, u2f(ordinal), 0.0f, 0.0f);
    float *pp = sink.params + (size_t)rec * g.paramStride;
    if (g.useZ) { pp[0] = v.z; pp += 1; }
    if (g.useW) { pp[0] = v.w; pp += 1; }
#pragma unroll
    for (int i = 0; i < NA; ++i)
        if (i < g.nA) { pp[0] = v.a[i]; pp += 1; }
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i < g.nP) { pp[0] = v.p[i]; pp += 1; }
}

// One input line in clip space: VertexProcessor.cpp:167-215 + LineClipper.cpp:30-56.  Returns false when the line is
// clipped away; else a, b are the screen-space end points.
template <int NA, int NP>
SWR_HD bool clipLineToScreen(const GeomArgs &g, const CVert<NA, NP> &c0, const CVert<NA, NP> &c1, CVert<NA, NP> &a, CVert<NA, NP> &b)
{
    const int m0 = outcode(c0.x, c0.y, c0.z, c0.w), m1 = outcode(c1.x, c1.y, c1.z, c1.w);
    const int mask = m0 | m1;
    float t0 = 0.0f, t1 = 1.0f;
    for (int pl = 0; pl < 6; ++pl) {
        if (!(mask & (1 << pl))) continue;
        const float d0 = planeDist(pl, c0.x, c0.y, c0.z, c0.w);
        const float d1 = planeDist(pl, c1.x, c1.y, c1.z, c1.w);
        const bool n0 = d0 < 0, n1 = d1 < 0;
        if (n0 && n1) return false;
        if (n0) {
            const float t = fdiv(-d0, fsub(d1, d0));
            t0 = t0 < t ? t : t0;                           // std::max(t0, t)
        } else {
            const float t = fdiv(d0, fsub(d0, d1));
            t1 = t < t1 ? t : t1;                           // std::min(t1, t)
        }
    }
    a = c0; b = c1;
    if (m0) lerpVert(a, c0, c1, t0);
    if (m1) lerpVert(b, c0, c1, t1);
    toScreen(g, a);
    toScreen(g, b);
    return true;
}

// One-call forms for sequential hosts.
template <int NA, int NP>
SWR_HD Box16 emitScreenLine(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal, const CVert<NA, NP> &v0, const CVert<NA, NP> &v1)
{
    int steps;
    bool tooLong;
    const Box16 box = setupScreenLine(g, v0, v1, steps, tooLong);
    if (tooLong) { sink.errorFlag[0] |= 2u; sink.errorFlag[1] |= 2u; }
    if (box.x0 <= box.x1) storeLine(g, sink, rec, ordinal, steps, v0, v1);
    return box;
}
template <int NA, int NP>
SWR_HD Box16 emitClipLine(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal, const CVert<NA, NP> &c0, const CVert<NA, NP> &c1)
{
    CVert<NA, NP> a, b;
    if (!clipLineToScreen(g, c0, c1, a, b)) return deadBox();
    return emitScreenLine(g, sink, rec, ordinal, a, b);
}
template <int NA, int NP>
SWR_HD Box16 emitScreenPoint(const GeomArgs &g, const RecordSink &sink, uint32_t rec, uint32_t ordinal, const CVert<NA, NP> &v)
{
    const Box16 box = setupScreenPoint(g, v);
    if (box.x0 <= box.x1) storePoint(g, sink, rec, ordinal, v);
    return box;
}

#if defined(__CUDACC__)

template <class VS>
SWR_D void shadeVertex(const GeomArgs &g, int index, CVert<VS::AVarCount, VS::PVarCount> &o)
{
    VertexShaderInput in;
#pragma unroll
    for (int i = 0; i < MaxVertexAttribs; ++i)
        in[i] = (i < VS::AttribCount) ? (const void *)((const char *)g.attribPtr[i] + (size_t)g.attribStride[i] * (size_t)index)
                                      : nullptr;                     // VertexProcessor.cpp:134-150
    VertexShaderOutput out;
    VS::processVertex(in, &out);
    o.x = out.x; o.y = out.y; o.z = out.z; o.w = out.w;
#pragma unroll
    for (int i = 0; i < VS::AVarCount; ++i) o.a[i] = out.avar[i];
#pragma unroll
    for (int i = 0; i < VS::PVarCount; ++i) o.p[i] = out.pvar[i];
}

// Mark every screen tile of `rank`'s partition that the group box (x0, y0)-(x1, y1) touches in that rank's
// tile x chunk bitmap (called by a whole warp: the lanes share the tiles).  The groups of one batch mostly touch
// the same few tiles, so a small per-CTA cache of (rank, chunk parity, tile) keys already marked by this batch
// filters the repeats -- what remains is a fire-and-forget reduction (no NVLink round trip when the bitmap is a
// peer's).  A cache miss of a repeat only costs a redundant reduction.
constexpr int kMarkCache = 256;
SWR_D void markTiles(const GeomArgs &g, const RecordSink &sink, int rank, uint32_t *cache, uint32_t chunk, int x0, int y0, int x1, int y1)
{
    if (x0 > x1) return;
    const int tx0 = x0 >> g.tileShift, ty0 = y0 >> g.tileShift;
    const int tx1 = min(x1 >> g.tileShift, g.tilesX - 1), ty1 = min(y1 >> g.tileShift, g.tilesY - 1);
    if (tx0 > tx1 || ty0 > ty1) return;
    const int lane = threadIdx.x & 31;
    const int nx = tx1 - tx0 + 1, nt = nx * (ty1 - ty0 + 1);
    const uint32_t bit = 1u << (chunk & 31);
    const bool oneRow = ty0 == ty1, oneCol = tx0 == tx1;     // the usual shapes: no integer division for them
    for (int i = lane; i < nt; i += 32) {
        const int ty = oneRow ? ty0 : oneCol ? ty0 + i : ty0 + i / nx;
        const int tx = oneRow ? tx0 + i : oneCol ? tx0 : tx0 + i % nx;
        if (!tileOwned(tx, ty, rank, g.world)) continue;
        const uint32_t tile = (uint32_t)(ty * g.tilesX + tx);
        const uint32_t key = ((uint32_t)rank << 28) | ((chunk & 1u) << 27) | tile;
        uint32_t *slot = cache + ((tile * 2654435761u + (uint32_t)rank * 40503u + (chunk & 1u)) >> 24);
        if (atomicExch(slot, key) == key) continue;          // (one shared-memory atomic: check and claim without a race)
        atomicOr(sink.tilemap + (size_t)tile * g.chunkWords + (chunk >> 5), bit);   // result unused: compiles to RED
    }
}

#ifndef SWR_GEOM_MINB
#define SWR_GEOM_MINB 4      // one rank: <= 64 registers, four 256-thread CTAs per SM (measured: 5 or 6 CTAs with spills are slower)
#endif
#ifndef SWR_GEOM_MINB_SHARDED
#define SWR_GEOM_MINB_SHARDED 3   // sharded: <= 80 registers (the record a thread holds until its slot is known would spill at 64)
#endif

// Batch run by CTA `block` of a launch: all batches in turn, or -- sharded -- the batches of this rank.
SWR_D int batchOfBlock(const GeomArgs &g, int block)
{
    if (!g.shard) return block;
    return ((block / kShardBatches) * g.world + g.rank) * kShardBatches + block % kShardBatches;
}

// Dynamic shared memory of the geometry kernel: the record staging area, or 0 when the shader's records are too
// large for it (then the records are stored directly).
constexpr int kGeomThreadsSharded = 256;    // threads per batch of the sharded form (measured on C3 at N = 2: 1024 threads, i.e. one
                                            // CTA per SM and the whole batch in one round, is 15 % slower: no second CTA fills its barriers)

template <int NA, int NP>
struct GeomStage {
    static constexpr size_t want = (size_t)(3 + kMaxPlanes + NA + NP) * 32 * 16 * (kGeomThreadsSharded / 32);
    static constexpr size_t bytes = want <= 200 * 1024 ? want : 0;
};

// Warp-cooperative vertex de-duplication (the reference's VertexCache, VertexCache.h:29-63, as a build variant):
// the 96 corner indices of a warp's 32 triangles go through a small per-warp hash table in shared memory, every
// distinct vertex is shaded once by one lane, and the triangles read the results back.  Results are identical with
// and without it (the cache only avoids re-shading).  Measured on B200: with the
// stock vertex shaders (a 4x4 transform, ~45 instructions) it LOSES -- a warp of a grid mesh holds 34 distinct
// vertices, i.e. two shading rounds instead of three, and the table costs more than the saved round -- so the
// product build leaves it off; -DSWR_GEOM_DEDUP=1 turns it on (heavier vertex shaders).
#ifndef SWR_GEOM_DEDUP
#define SWR_GEOM_DEDUP 0
#endif
constexpr int kDedupSlots = 128, kDedupVerts = 96;
template <int NA, int NP>
struct GeomDedup {
    static constexpr int VF = 4 + NA + NP;
    static constexpr size_t perWarp = kDedupSlots * 4 + kDedupSlots + kDedupVerts * 4 + (size_t)kDedupVerts * VF * 4;
    static constexpr size_t bytes = SWR_GEOM_DEDUP ? ((perWarp + 15) / 16 * 16) * (kGeomThreads / 32) : 0;
};

// MODE = draw mode, SPAN = false when the raster mode is known to be Block (no span state is carried), MULTI = sharded
// geometry (records compacted per destination rank and pushed to the tile owners; otherwise every record stays at
// its own slot of this rank's scratch): launch-time constants, and as template parameters they keep the registers
// and the instructions down to what the launch really needs.
template <class VS, int MODE, bool SPAN, bool MULTI>
__global__ void __launch_bounds__(MULTI ? kGeomThreadsSharded : kGeomThreads, MULTI ? SWR_GEOM_MINB_SHARDED : SWR_GEOM_MINB) geometryKernel(const GeomArgs g)
{
    constexpr int kThreads = MULTI ? kGeomThreadsSharded : kGeomThreads;
    constexpr int kSlots = kBatch / kThreads;            // slots per thread in the block-wide scans
    constexpr int NA = VS::AVarCount, NP = VS::PVarCount;
    typedef CVert<NA, NP> V;
    constexpr int kMaxExtraGroups = kBatch * (kMaxFan - 1) / kGroup;
    constexpr int kWarps = kThreads / 32;

    __shared__ uint16_t sExtraCnt[kBatch];     // fan extras per primitive of this batch
    __shared__ uint16_t sExtraOfs[kBatch];     // exclusive prefix in primitive order
    __shared__ uint32_t sWarpSum[kWarps];
    __shared__ uint32_t sExtraBase, sExtraTotal, sClipTotal;
    __shared__ uint16_t sClipSlot[kBatch];     // slots of the primitives that have fan extras, in order
    __shared__ int sGx0[kMaxExtraGroups], sGy0[kMaxExtraGroups], sGx1[kMaxExtraGroups], sGy1[kMaxExtraGroups];
    __shared__ Box16 sGb[kMaxRanks][kBatch / kGroup];               // group boxes / record counts of the batch per rank:
    __shared__ uint8_t sGc[kMaxRanks][kBatch / kGroup];             //   written out in one piece per rank at the end
    __shared__ uint32_t sMark[kMarkCache];                          // markTiles' filter

    // per-warp staging of one group's records (head: 3 quads, params: up to kStageQuads - 3 quads per record)
    constexpr int kStageQuads = 3 + kMaxPlanes + NA + NP;
    constexpr bool kStage = MULTI && GeomStage<NA, NP>::bytes > 0;
    extern __shared__ float4 sStageAll[];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float4 *stage = sStageAll + (kStage ? wid * 32 * kStageQuads : 0);
    const int batch = batchOfBlock(g, blockIdx.x);
    const int primBase = batch * kBatch;
    if (primBase >= g.numPrims) return;
    const int cnt = min(kBatch, g.numPrims - primBase);
    const uint32_t ord0 = (g.firstBatch + (uint32_t)batch) * SWR_ORDINAL_STRIDE;
    constexpr int per = MODE + 1;
    const int world = g.world;
    // replicated geometry: only the records of this rank's tiles are kept
    const uint32_t keepMask = g.shard ? (1u << world) - 1u : (1u << g.rank);
    bool anyExtra = false;
    for (int i = tid; i < kMarkCache; i += kThreads) sMark[i] = 0xffffffffu;
    __syncthreads();

    // the next round's indices are fetched while the current round computes (the vertex fetch
    // depends on them, so this takes one DRAM round trip off every round's critical path)
    constexpr int kRounds = kBatch / kThreads;
    int32_t cur[3] = { 0, 0, 0 }, nxt[3] = { 0, 0, 0 };
    auto fetch = [&](int r, int32_t *dst) {
        const int slot = r * kThreads + tid;
        if (slot < cnt) {
            const int32_t *ip = g.indices + (size_t)(primBase + slot) * per;
            dst[0] = ip[0];
            if (per > 1) dst[1] = ip[1];
            if (per > 2) dst[2] = ip[2];
        }
    };
    fetch(0, cur);

#pragma unroll 1
    for (int r = 0; r < kRounds; ++r) {
        if (r + 1 < kRounds) fetch(r + 1, nxt);
        const int slot = r * kThreads + tid;
        const uint32_t group = (uint32_t)(primBase + slot) >> 5;      // the 32 slots this warp handles in this round
        Box16 box = deadBox();
        int extras = 0;
        const uint32_t ordinal = ord0 + (uint32_t)slot;
        TriRecord<NA, NP> R;
        V la, lb;                                  // line end points / the point, in screen space
        int steps = 0;
        R.span = false;
        V dv[3];
        if (SWR_GEOM_DEDUP && !MULTI && MODE == SWR_DRAW_TRIANGLE) {
            constexpr int VF = GeomDedup<NA, NP>::VF;
            char *wbase = reinterpret_cast<char *>(sStageAll) + (size_t)wid * ((GeomDedup<NA, NP>::perWarp + 15) / 16 * 16);
            int *keys = reinterpret_cast<int *>(wbase);
            int *idxOf = keys + kDedupSlots;
            float *verts = reinterpret_cast<float *>(idxOf + kDedupVerts);
            uint8_t *uidOf = reinterpret_cast<uint8_t *>(verts + kDedupVerts * VF);
            __syncwarp();
            for (int i = lane; i < kDedupSlots; i += 32) keys[i] = -1;
            __syncwarp();
            const bool act = slot < cnt;
            int hs[3] = { 0, 0, 0 };
            bool win[3] = { false, false, false };
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (act) {
                    const int idx = cur[k];
                    uint32_t h = ((uint32_t)idx * 2654435761u) >> 25;
                    while (true) {
                        const int old = atomicCAS(&keys[h], -1, idx);
                        if (old == -1) { win[k] = true; break; }
                        if (old == idx) break;
                        h = (h + 1) & (kDedupSlots - 1);
                    }
                    hs[k] = (int)h;
                }
            }
            __syncwarp();
            uint32_t nUnique = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t b = __ballot_sync(0xffffffffu, win[k]);
                if (win[k]) {
                    const uint32_t uid = nUnique + (uint32_t)__popc(b & ((1u << lane) - 1u));
                    uidOf[hs[k]] = (uint8_t)uid;
                    idxOf[uid] = cur[k];
                }
                nUnique += (uint32_t)__popc(b);
            }
            __syncwarp();
            for (uint32_t u = lane; u < nUnique; u += 32) {
                V t;
                shadeVertex<VS>(g, idxOf[u], t);
                float *o = verts + u * VF;
                o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
#pragma unroll
                for (int i = 0; i < NA; ++i) o[4 + i] = t.a[i];
#pragma unroll
                for (int i = 0; i < NP; ++i) o[4 + NA + i] = t.p[i];
            }
            __syncwarp();
            if (act) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float *o = verts + (uint32_t)uidOf[hs[k]] * VF;
                    dv[k].x = o[0]; dv[k].y = o[1]; dv[k].z = o[2]; dv[k].w = o[3];
#pragma unroll
                    for (int i = 0; i < NA; ++i) dv[k].a[i] = o[4 + i];
#pragma unroll
                    for (int i = 0; i < NP; ++i) dv[k].p[i] = o[4 + NA + i];
                }
            }
        }
        if (slot < cnt) {
            const int32_t ip[3] = { cur[0], cur[1], cur[2] };
            if (MODE == SWR_DRAW_TRIANGLE) {
                V v0, v1, v2;
                if (SWR_GEOM_DEDUP && !MULTI) {
                    v0 = dv[0]; v1 = dv[1]; v2 = dv[2];
                } else {
                    shadeVertex<VS>(g, ip[0], v0);
                    shadeVertex<VS>(g, ip[1], v1);
                    shadeVertex<VS>(g, ip[2], v2);
                }
                const int m0 = outcode(v0.x, v0.y, v0.z, v0.w), m1 = outcode(v1.x, v1.y, v1.z, v1.w),
                          m2 = outcode(v2.x, v2.y, v2.z, v2.w);
                const int mask = m0 | m1 | m2;
                if (mask == 0) {
                    box = setupClipTriangle<NA, NP, SPAN>(g, ordinal, v0, v1, v2, R);
                } else if ((mask & -mask) & (m0 & m1 & m2)) {
                    // Trivial reject, exactly as the reference computes it: the FIRST plane it clips
                    // against (lowest flagged bit, VertexProcessor.cpp:237-242) has all three original
                    // vertices outside (its distance a*x+b*y+c*z+d*w equals the outcode's w-x, x+w, ...
                    // bit for bit on finite inputs), so that pass emits nothing (PolyClipper.cpp:64-74)
                    // and the triangle is fully clipped.  No polygon is built.
                } else {
                    V a[kMaxPoly], b[kMaxPoly], *poly;      // rare path: polygons live in local memory
                    a[0] = v0; a[1] = v1; a[2] = v2;
                    bool overflow = false;
                    const int n = clipTriangle<NA, NP>(a, b, mask, &poly, &overflow);
                    if (overflow) { atomicOr(g.sink[g.rank].errorFlag, 4u); atomicOr(g.sink[g.rank].errorFlag + 1, 4u); }
                    if (n >= 3) {
                        box = setupClipTriangle<NA, NP, SPAN>(g, ordinal, poly[0], poly[1], poly[2], R);
                        extras = n - 3;
                    }
                }
            } else if (MODE == SWR_DRAW_LINE) {
                V c0, c1;
                shadeVertex<VS>(g, ip[0], c0);
                shadeVertex<VS>(g, ip[1], c1);
                if (clipLineToScreen<NA, NP>(g, c0, c1, la, lb)) {
                    bool tooLong;
                    box = setupScreenLine<NA, NP>(g, la, lb, steps, tooLong);
                    if (tooLong) { atomicOr(g.sink[g.rank].errorFlag, 2u); atomicOr(g.sink[g.rank].errorFlag + 1, 2u); }
                }
            } else {
                shadeVertex<VS>(g, ip[0], la);
                if (outcode(la.x, la.y, la.z, la.w) == 0) {     // VertexProcessor.cpp:152-165
                    toScreen(g, la);
                    box = setupScreenPoint<NA, NP>(g, la);
                }
            }
        }
        sExtraCnt[slot] = (uint16_t)extras;
        anyExtra |= extras > 0;

        const int gl = slot >> 5;                             // group within the batch
        if (!MULTI) {
            // ---- one destination (this rank); dropped primitives -- and, with several ranks and replicated geometry,
            // primitives that touch none of this rank's tiles -- leave no record
            const RecordSink &sk = g.sink[g.rank];
            if (world > 1 && !((boxOwnerMask(box, g.tileShift, g.tilesX, g.tilesY, world) >> g.rank) & 1u)) box = deadBox();
            // The surviving records of the warp's group move to the front of its 32 slots, in submission order (ballot +
            // popcount, as in the sharded form below): dead boxes are neither written nor read again (on C3 two thirds
            // of the slots: -55 MB of box writes here, -55 MB of box reads in the tile kernel; time unchanged).
            const bool live = box.x0 <= box.x1;
            const uint32_t liveMask = __ballot_sync(0xffffffffu, live);
            const uint32_t rec = (group << 5) + (uint32_t)__popc(liveMask & ((1u << lane) - 1u));
            if (live) {
                sk.bbox[rec] = box;
                if (MODE == SWR_DRAW_TRIANGLE) storeTriangle<NA, NP>(g, sk, rec, R);
                else if (MODE == SWR_DRAW_LINE) storeLine<NA, NP>(g, sk, rec, ordinal, steps, la, lb);
                else storePoint<NA, NP>(g, sk, rec, ordinal, la);
            }
            const int x0 = __reduce_min_sync(0xffffffffu, (int)box.x0), y0 = __reduce_min_sync(0xffffffffu, (int)box.y0);
            const int x1 = __reduce_max_sync(0xffffffffu, (int)box.x1), y1 = __reduce_max_sync(0xffffffffu, (int)box.y1);
            if (lane == 0) {
                Box16 u; u.x0 = (int16_t)x0; u.y0 = (int16_t)y0; u.x1 = (int16_t)x1; u.y1 = (int16_t)y1;
                sGb[0][gl] = u;
                sGc[0][gl] = (uint8_t)__popc(liveMask);
            }
            markTiles(g, sk, g.rank, sMark, 2u * (uint32_t)batch, x0, y0, x1, y1);
            cur[0] = nxt[0]; cur[1] = nxt[1]; cur[2] = nxt[2];
            continue;
        }
        // ---- per destination rank: the surviving records of the group, compacted to the front of the group's 32
        // slots in submission order (ballot + popcount: the warps never wait for each other), their count, the
        // union of their boxes and the tiles it touches.  Slots behind the count are never written, nor read.
        const uint32_t dm = boxOwnerMask(box, g.tileShift, g.tilesX, g.tilesY, world) & keepMask;
        const uint32_t anyD = __reduce_or_sync(0xffffffffu, dm);
        if (lane < world && !((anyD >> lane) & 1u)) {
            sGb[lane][gl] = deadBox();                        // nothing of this group concerns rank `lane`
            sGc[lane][gl] = 0;
        }
        for (uint32_t m = anyD; m; m &= m - 1) {
            const int d = __ffs((int)m) - 1;
            const bool mine = (dm >> d) & 1u;
            const uint32_t b = __ballot_sync(0xffffffffu, mine);
            const int x0 = __reduce_min_sync(0xffffffffu, mine ? (int)box.x0 : 32767), y0 = __reduce_min_sync(0xffffffffu, mine ? (int)box.y0 : 32767);
            const int x1 = __reduce_max_sync(0xffffffffu, mine ? (int)box.x1 : -32768), y1 = __reduce_max_sync(0xffffffffu, mine ? (int)box.y1 : -32768);
            const RecordSink &sk = g.sink[d];
            const uint32_t pos = (uint32_t)__popc(b & ((1u << lane) - 1u)), n = (uint32_t)__popc(b);
            if (kStage) {
                // The records go through shared memory so that every store instruction of the warp writes one
                // contiguous run (32 lanes x 16 bytes) instead of 32 pieces 48 bytes apart: what matters when the
                // destination is a peer (every store instruction becomes its own train of NVLink write packets).
                if (mine) {
                    RecordSink st = sk;
                    st.head = stage;
                    st.params = reinterpret_cast<float *>(stage + 32 * 3);
                    if (MODE == SWR_DRAW_TRIANGLE) storeTriangle<NA, NP>(g, st, pos, R, sk.span ? sk.span + (size_t)((group << 5) + pos) * 3 : nullptr);
                    else if (MODE == SWR_DRAW_LINE) storeLine<NA, NP>(g, st, pos, ordinal, steps, la, lb);
                    else storePoint<NA, NP>(g, st, pos, ordinal, la);
                    sk.bbox[(group << 5) + pos] = box;
                }
                __syncwarp();
                const uint32_t pq = (uint32_t)g.paramStride >> 2;
                float4 *dh = sk.head + (size_t)(group << 5) * 3;
                float4 *dp = reinterpret_cast<float4 *>(sk.params + (size_t)(group << 5) * g.paramStride);
                for (uint32_t i = lane; i < n * 3u; i += 32) dh[i] = stage[i];
                for (uint32_t i = lane; i < n * pq; i += 32) dp[i] = stage[32 * 3 + i];
                __syncwarp();
            } else if (mine) {
                const uint32_t rec = (group << 5) + pos;
                sk.bbox[rec] = box;
                if (MODE == SWR_DRAW_TRIANGLE) storeTriangle<NA, NP>(g, sk, rec, R);
                else if (MODE == SWR_DRAW_LINE) storeLine<NA, NP>(g, sk, rec, ordinal, steps, la, lb);
                else storePoint<NA, NP>(g, sk, rec, ordinal, la);
            }
            if (lane == 0) {
                Box16 u; u.x0 = (int16_t)x0; u.y0 = (int16_t)y0; u.x1 = (int16_t)x1; u.y1 = (int16_t)y1;
                sGb[d][gl] = u;
                sGc[d][gl] = (uint8_t)n;
            }
            markTiles(g, sk, d, sMark, 2u * (uint32_t)batch, x0, y0, x1, y1);
        }
        cur[0] = nxt[0]; cur[1] = nxt[1]; cur[2] = nxt[2];
    }
    // groups of a short last batch that no warp visited: the binning reads all ceil(cnt / 32) groups of a chunk
    // (all 32 warps-rounds ran above, so every group < 32 was written; nothing to do)

    const bool haveExtras = __syncthreads_or(anyExtra);
    if (tid < world && ((keepMask >> tid) & 1u) && (MODE != SWR_DRAW_TRIANGLE || !haveExtras)) g.sink[tid].extra[batch] = make_uint2(0u, 0u);
    {   // the batch's 32 group boxes and counts, one contiguous run per rank
        static_assert(kThreads >= kMaxRanks * (kBatch / kGroup), "one thread per (rank, group)");
        const int d = tid >> 5;
        if (MULTI ? (d < world) : (d == 0)) {
            const RecordSink &sk = g.sink[MULTI ? d : g.rank];
            sk.gbox[((uint32_t)primBase >> 5) + lane] = sGb[d][lane];
            sk.gcnt[((uint32_t)primBase >> 5) + lane] = sGc[d][lane];
        }
    }
    if (MODE != SWR_DRAW_TRIANGLE || !haveExtras) return;

    // ---- clipper fan extras: appended behind the batch's original slots, in primitive order.  They keep one slot
    // each (no compaction): every rank gets the whole range of boxes, dead where the triangle is not its business.
    {   // exclusive scan of sExtraCnt over the 1024 slots: thread t owns the kSlots consecutive slots from kSlots * t
        // (packed: fan extras in the low 16 bits, clipped primitives that have any in the high 16 bits)
        uint32_t cs[kSlots], sum = 0;
#pragma unroll
        for (int k = 0; k < kSlots; ++k) { cs[k] = sExtraCnt[kSlots * tid + k]; sum += cs[k] + (cs[k] ? 0x10000u : 0u); }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) sWarpSum[wid] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < wid; ++w) base += sWarpSum[w];
        uint32_t ex = base + incl - sum;
        {
            // offsets of every slot's extras, and the list of the slots that have any: the second phase below takes
            // its clipped primitives from that list, so its warps are full (in slot order one lane in nine is busy)
            uint32_t run = ex & 0xffffu, nclip = ex >> 16;
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                sExtraOfs[kSlots * tid + k] = (uint16_t)run;
                run += cs[k];
                if (cs[k]) sClipSlot[nclip++] = (uint16_t)(kSlots * tid + k);
            }
        }
        if (tid == kThreads - 1) {
            sClipTotal = (ex + sum) >> 16;
            const uint32_t total = (ex + sum) & 0xffffu;
            const uint32_t padded = (total + kGroup - 1) & ~(uint32_t)(kGroup - 1);
            uint32_t b0 = atomicAdd(g.extraAlloc, padded);
            if (b0 + padded > g.extrasEnd) {             // scratch exhausted: the draw is void on every rank
                for (int d = 0; d < world; ++d)
                    if ((keepMask >> d) & 1u) { atomicOr(g.sink[d].errorFlag, 1u); atomicOr(g.sink[d].errorFlag + 1, 1u); }
                b0 = 0xffffffffu;
            }
            sExtraBase = b0;
            sExtraTotal = total;
        }
        for (int i = tid; i < kMaxExtraGroups; i += kThreads) { sGx0[i] = 32767; sGy0[i] = 32767; sGx1[i] = -32768; sGy1[i] = -32768; }
        __syncthreads();
    }
    const uint32_t ebase = sExtraBase, etotal = sExtraTotal;
    if (tid < world && ((keepMask >> tid) & 1u))
        g.sink[tid].extra[batch] = (ebase == 0xffffffffu) ? make_uint2(0u, 0u) : make_uint2(ebase, etotal);
    if (ebase == 0xffffffffu) return;                       // flagged, draw is void

    const uint32_t nClipped = sClipTotal;
    for (uint32_t ci = tid; ci < nClipped; ci += kThreads) {
        const int slot = sClipSlot[ci];
        const uint32_t ofs = sExtraOfs[slot];
        const int32_t *ip = g.indices + (size_t)(primBase + slot) * 3;
        V a[kMaxPoly], b[kMaxPoly], *poly;
        shadeVertex<VS>(g, ip[0], a[0]);
        shadeVertex<VS>(g, ip[1], a[1]);
        shadeVertex<VS>(g, ip[2], a[2]);
        const int mask = outcode(a[0].x, a[0].y, a[0].z, a[0].w) | outcode(a[1].x, a[1].y, a[1].z, a[1].w) |
                         outcode(a[2].x, a[2].y, a[2].z, a[2].w);
        const int n = clipTriangle<NA, NP>(a, b, mask, &poly, nullptr);
        for (int k = 1; k + 2 < n; ++k) {                   // fan (p0, p[k+1], p[k+2]), VertexProcessor.cpp:257-261
            const uint32_t e = ofs + (uint32_t)(k - 1);
            const uint32_t rec = ebase + e;
            TriRecord<NA, NP> R;
            const Box16 box = setupClipTriangle<NA, NP, SPAN>(g, ord0 + (uint32_t)cnt + e, poly[0], poly[k + 1], poly[k + 2], R);
            const uint32_t dm = boxOwnerMask(box, g.tileShift, g.tilesX, g.tilesY, world) & keepMask;
            for (int d = 0; d < world; ++d) {
                if (!((keepMask >> d) & 1u)) continue;
                const bool mine = (dm >> d) & 1u;
                g.sink[d].bbox[rec] = mine ? box : deadBox();
                if (mine) storeTriangle<NA, NP>(g, g.sink[d], rec, R);
            }
            if (dm) {
                atomicMin(&sGx0[e >> 5], (int)box.x0); atomicMin(&sGy0[e >> 5], (int)box.y0);
                atomicMax(&sGx1[e >> 5], (int)box.x1); atomicMax(&sGy1[e >> 5], (int)box.y1);
            }
        }
    }
    __syncthreads();
    const uint32_t padded = (etotal + kGroup - 1) & ~(uint32_t)(kGroup - 1);
    // group boxes of the extras (one union box per group, shared by all ranks: each rank marks only its own tiles)
    const int ngroups = (int)(padded >> 5);
    for (int i = wid; i < ngroups * world; i += kWarps) {
        const int gi = i / world, d = i - gi * world;
        if (!((keepMask >> d) & 1u)) continue;
        const RecordSink &sk = g.sink[d];
        if (lane == 0) {
            Box16 u; u.x0 = (int16_t)sGx0[gi]; u.y0 = (int16_t)sGy0[gi]; u.x1 = (int16_t)sGx1[gi]; u.y1 = (int16_t)sGy1[gi];
            sk.gbox[(ebase >> 5) + (uint32_t)gi] = u;
            sk.gcnt[(ebase >> 5) + (uint32_t)gi] = (uint8_t)min(32u, etotal - (uint32_t)gi * 32u);
        }
        markTiles(g, sk, d, sMark, 2u * (uint32_t)batch + 1u, sGx0[gi], sGy0[gi], sGx1[gi], sGy1[gi]);
    }
}

template <class VS>
void launchGeometry(const void *args, void *stream)
{
    const GeomArgs *g = static_cast<const GeomArgs *>(args);
    const int batches = (g->numPrims + kBatch - 1) / kBatch;
    int blocks = batches;
    if (g->shard) {
        // runs of kShardBatches batches go round-robin to the ranks: CTAs for every run of this rank (the kernel
        // drops the batches past the end of the pass)
        const int runs = (batches + kShardBatches - 1) / kShardBatches;
        const int mine = runs > g->rank ? (runs - g->rank + g->world - 1) / g->world : 0;
        blocks = mine * kShardBatches;
    }
    // numPrims == 0: nothing is launched, but the kernel the launch would use is resolved (loaded onto the device).  The
    // runtime asks for this before it enqueues a cross-GPU barrier, so that the launch after the barrier cannot stall
    // on a lazy module load.
    cudaStream_t st = (cudaStream_t)stream;
    constexpr size_t stageBytes = GeomStage<VS::AVarCount, VS::PVarCount>::bytes;
    auto launch = [&](auto kernel, size_t smem) {
        if (smem > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (blocks > 0) kernel<<<blocks, g->shard ? kGeomThreadsSharded : kGeomThreads, smem, st>>>(*g);
        else { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kernel); }
    };
    if (g->shard) {
        if (g->drawMode == SWR_DRAW_TRIANGLE) {
            if (g->rasterMode == SWR_RASTER_BLOCK) launch(geometryKernel<VS, SWR_DRAW_TRIANGLE, false, true>, stageBytes);
            else launch(geometryKernel<VS, SWR_DRAW_TRIANGLE, true, true>, stageBytes);
        } else if (g->drawMode == SWR_DRAW_LINE) {
            launch(geometryKernel<VS, SWR_DRAW_LINE, false, true>, stageBytes);
        } else {
            launch(geometryKernel<VS, SWR_DRAW_POINT, false, true>, stageBytes);
        }
    } else {
        if (g->drawMode == SWR_DRAW_TRIANGLE) {
            constexpr size_t dedupBytes = GeomDedup<VS::AVarCount, VS::PVarCount>::bytes;
            if (g->rasterMode == SWR_RASTER_BLOCK) launch(geometryKernel<VS, SWR_DRAW_TRIANGLE, false, false>, dedupBytes);
            else launch(geometryKernel<VS, SWR_DRAW_TRIANGLE, true, false>, dedupBytes);
        } else if (g->drawMode == SWR_DRAW_LINE) {
            launch(geometryKernel<VS, SWR_DRAW_LINE, false, false>, 0);
        } else {
            launch(geometryKernel<VS, SWR_DRAW_POINT, false, false>, 0);
        }
    }
}

// ---- stream-out: the vertex stage alone, for a VertexProcessor whose IRasterizer is not ours ----------------
// One CTA per batch of 1024 input primitives (VertexProcessor.cpp:110-116) writes what the reference hands to
// IRasterizer::draw{Point,Line,Triangle}List (VertexProcessor.cpp:302-317): screen-space RasterizerVertex records and
// the index list, with -1 for primitives dropped by clipping / culling (VertexProcessor.cpp:152-165, 196-200,
// 244-250, 331-340), culled-in triangles re-oriented by swapping their first and last index (VertexProcessor.cpp:342),
// and the clipper's fan triangles appended behind the batch's own primitives in primitive order
// (VertexProcessor.cpp:252-261).  The vertex array is not de-duplicated (every primitive owns `per` consecutive
// vertices): the reference's VertexCache only affects which array positions the indices name, never the primitives
// a rasterizer sees.  Layout of batch b: vertices [b * soStride, ...), indices likewise, soStride = per * 1024 +
// 3 * soExtraCap; soCounts[b] = fan triangles appended.
struct SoVertex { float v[4 + SWR_MAX_AVARS + SWR_MAX_PVARS]; };

template <int NA, int NP>
SWR_D void soWrite(SoVertex *dst, const CVert<NA, NP> &c)
{
    float4 *q = reinterpret_cast<float4 *>(dst);
    float f[36];
#pragma unroll
    for (int i = 0; i < 36; ++i) f[i] = 0.0f;
    f[0] = c.x; f[1] = c.y; f[2] = c.z; f[3] = c.w;
#pragma unroll
    for (int i = 0; i < NA; ++i) f[4 + i] = c.a[i];
#pragma unroll
    for (int i = 0; i < NP; ++i) f[4 + SWR_MAX_AVARS + i] = c.p[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) q[i] = mkf4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
}

// Screen-space triangle into the stream: cull (drop / keep / swap first and last index), three vertices, three indices.
template <int NA, int NP>
SWR_D void soTriangle(const GeomArgs &g, SoVertex *verts, int32_t *idx, int at, CVert<NA, NP> a, CVert<NA, NP> b, CVert<NA, NP> c)
{
    toScreen(g, a); toScreen(g, b); toScreen(g, c);
    const float facing = fsub(fmul(fsub(a.x, b.x), fsub(c.y, b.y)), fmul(fsub(c.x, b.x), fsub(a.y, b.y)));
    bool drop = false, swap = false;
    if (facing < 0) drop = g.cullMode == SWR_CULL_CW;
    else { drop = g.cullMode == SWR_CULL_CCW; swap = !drop; }
    if (drop) { idx[at] = idx[at + 1] = idx[at + 2] = -1; return; }
    soWrite(verts + at, a); soWrite(verts + at + 1, b); soWrite(verts + at + 2, c);
    idx[at] = swap ? at + 2 : at; idx[at + 1] = at + 1; idx[at + 2] = swap ? at : at + 2;
}

template <class VS>
__global__ void __launch_bounds__(kGeomThreads) streamOutKernel(const GeomArgs g)
{
    constexpr int NA = VS::AVarCount, NP = VS::PVarCount;
    typedef CVert<NA, NP> V;
    __shared__ uint16_t sExtraCnt[kBatch], sExtraOfs[kBatch];
    __shared__ uint32_t sWarpSum[kGeomThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int batch = blockIdx.x;
    const int primBase = batch * kBatch;
    const int cnt = min(kBatch, g.numPrims - primBase);
    const int per = g.drawMode + 1;
    const size_t stride = (size_t)per * kBatch + 3 * (size_t)g.soExtraCap;
    SoVertex *verts = static_cast<SoVertex *>(g.soVerts) + (size_t)batch * stride;
    int32_t *idx = g.soIndices + (size_t)batch * stride;
    bool anyExtra = false;
    for (int r = 0; r < kBatch / kGeomThreads; ++r) {
        const int slot = r * kGeomThreads + tid;
        int extras = 0;
        if (slot < cnt) {
            const int32_t *ip = g.indices + (size_t)(primBase + slot) * per;
            const int at = slot * per;
            if (g.drawMode == SWR_DRAW_TRIANGLE) {
                V v0, v1, v2;
                shadeVertex<VS>(g, ip[0], v0); shadeVertex<VS>(g, ip[1], v1); shadeVertex<VS>(g, ip[2], v2);
                const int mask = outcode(v0.x, v0.y, v0.z, v0.w) | outcode(v1.x, v1.y, v1.z, v1.w) | outcode(v2.x, v2.y, v2.z, v2.w);
                if (mask == 0) {
                    soTriangle<NA, NP>(g, verts, idx, at, v0, v1, v2);
                } else {
                    V a[kMaxPoly], b[kMaxPoly], *poly;
                    a[0] = v0; a[1] = v1; a[2] = v2;
                    bool overflow = false;
                    const int n = clipTriangle<NA, NP>(a, b, mask, &poly, &overflow);
                    if (overflow) { atomicOr(g.sink[g.rank].errorFlag, 4u); atomicOr(g.sink[g.rank].errorFlag + 1, 4u); }
                    if (n >= 3) { soTriangle<NA, NP>(g, verts, idx, at, poly[0], poly[1], poly[2]); extras = n - 3; }
                    else idx[at] = idx[at + 1] = idx[at + 2] = -1;
                }
            } else if (g.drawMode == SWR_DRAW_LINE) {
                V c0, c1, a, b;
                shadeVertex<VS>(g, ip[0], c0); shadeVertex<VS>(g, ip[1], c1);
                if (clipLineToScreen<NA, NP>(g, c0, c1, a, b)) {
                    soWrite(verts + at, a); soWrite(verts + at + 1, b);
                    idx[at] = at; idx[at + 1] = at + 1;
                } else idx[at] = idx[at + 1] = -1;
            } else {
                V c0;
                shadeVertex<VS>(g, ip[0], c0);
                if (outcode(c0.x, c0.y, c0.z, c0.w) == 0) { toScreen(g, c0); soWrite(verts + at, c0); idx[at] = at; }
                else idx[at] = -1;
            }
        }
        sExtraCnt[slot] = (uint16_t)extras;
        anyExtra |= extras > 0;
    }
    if (!__syncthreads_or(anyExtra)) {
        if (tid == 0) g.soCounts[batch] = 0;
        return;
    }
    uint32_t c0 = sExtraCnt[4 * tid], c1 = sExtraCnt[4 * tid + 1], c2 = sExtraCnt[4 * tid + 2], c3 = sExtraCnt[4 * tid + 3];
    uint32_t sum = c0 + c1 + c2 + c3, incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) sWarpSum[wid] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < wid; ++w) base += sWarpSum[w];
    const uint32_t ex = base + incl - sum;
    sExtraOfs[4 * tid] = (uint16_t)ex; sExtraOfs[4 * tid + 1] = (uint16_t)(ex + c0);
    sExtraOfs[4 * tid + 2] = (uint16_t)(ex + c0 + c1); sExtraOfs[4 * tid + 3] = (uint16_t)(ex + c0 + c1 + c2);
    if (tid == kGeomThreads - 1) g.soCounts[batch] = ex + sum;
    __syncthreads();
    for (int r = 0; r < kBatch / kGeomThreads; ++r) {
        const int slot = r * kGeomThreads + tid;
        const int extras = sExtraCnt[slot];
        if (extras == 0) continue;
        const int32_t *ip = g.indices + (size_t)(primBase + slot) * 3;
        V a[kMaxPoly], b[kMaxPoly], *poly;
        shadeVertex<VS>(g, ip[0], a[0]); shadeVertex<VS>(g, ip[1], a[1]); shadeVertex<VS>(g, ip[2], a[2]);
        const int mask = outcode(a[0].x, a[0].y, a[0].z, a[0].w) | outcode(a[1].x, a[1].y, a[1].z, a[1].w) | outcode(a[2].x, a[2].y, a[2].z, a[2].w);
        const int n = clipTriangle<NA, NP>(a, b, mask, &poly, nullptr);
        for (int k = 1; k + 2 < n; ++k)
            soTriangle<NA, NP>(g, verts, idx, 3 * cnt + 3 * ((int)sExtraOfs[slot] + k - 1), poly[0], poly[k + 1], poly[k + 2]);
    }
}

template <class VS>
void launchStreamOut(const void *args, void *stream)
{
    const GeomArgs *g = static_cast<const GeomArgs *>(args);
    const int batches = (g->numPrims + kBatch - 1) / kBatch;
    if (batches > 0) streamOutKernel<VS><<<batches, kGeomThreads, 0, (cudaStream_t)stream>>>(*g);
}

template <class VS>
const swr_vertex_shader *vertexShaderBinding(const char *name = "user")
{
    static const swr_vertex_shader d = { &launchGeometry<VS>, &launchStreamOut<VS>, &uploadUniforms, VS::AttribCount, VS::AVarCount, VS::PVarCount, name, SWR_ARGS_LAYOUT };
    return &d;
}

#endif // __CUDACC__

} // namespace detail
} // namespace swr
