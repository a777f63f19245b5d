// hostcheck.cu -- TEST INFRASTRUCTURE ONLY (never part of the product, never shipped).
//
// Runs the product's parity-critical DEVICE MATH on the host: the __host__ __device__ functions
// of include/swr/detail/{geometry,tile}.cuh (clip, transform, cull, setup, footprint, coverage
// masks, fragment chains) are called from a plain sequential loop that stands in for the two
// kernels' thread / warp orchestration.  The result is diffed against the oracle by
// tests/test_hostcheck.py without a GPU, so that a GPU run only has to prove the orchestration
// (binning order, scans, shared-memory staging).  Built by nvcc as host code with
// -Xcompiler -ffp-contract=off; no kernel is launched.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <swr/VertexShaderBase.h>
#include <swr/PixelShaderBase.h>
#include <swr/detail/geometry.cuh>
#include <swr/detail/tile.cuh>
#include <swr/Texture.h>

#include "../../oracle/swr_scene.h"

using namespace swr;
using namespace swr::detail;

namespace {

swr_scene *g_s = nullptr;

struct PosColorVertex { float x, y, z, r, g, b; };
struct ObjVertex { float px, py, pz, nx, ny, nz, u, v; };

void mvpTransform(const float *m, float x, float y, float z, VertexShaderOutput *out)
{
    const float w = 1.0f;
    out->x = m[0] * x + m[1] * y + m[2] * z + m[3] * w;
    out->y = m[4] * x + m[5] * y + m[6] * z + m[7] * w;
    out->z = m[8] * x + m[9] * y + m[10] * z + m[11] * w;
    out->w = m[12] * x + m[13] * y + m[14] * z + m[15] * w;
}

template <int NP_>
struct HostVS {
    static const int AVarCount = 3;
    static const int PVarCount = NP_;
    static void processVertex(const void *in, VertexShaderOutput *out)
    {
        if (g_s->vs_kind == SWR_VS_POS_COLOR) {
            const PosColorVertex *d = static_cast<const PosColorVertex *>(in);
            out->x = d->x; out->y = d->y; out->z = d->z; out->w = 1.0f;
            out->avar[0] = d->r; out->avar[1] = d->g; out->avar[2] = d->b;
        } else if (g_s->vs_kind == SWR_VS_MVP_COLOR) {
            const PosColorVertex *d = static_cast<const PosColorVertex *>(in);
            mvpTransform(g_s->mvp, d->x, d->y, d->z, out);
            out->avar[0] = d->r; out->avar[1] = d->g; out->avar[2] = d->b;
        } else {
            const ObjVertex *d = static_cast<const ObjVertex *>(in);
            mvpTransform(g_s->mvp, d->px, d->py, d->pz, out);
            out->avar[0] = d->nx; out->avar[1] = d->ny; out->avar[2] = d->nz;
            out->pvar[0] = d->u; out->pvar[1] = d->v;
        }
    }
};

unsigned packRGB(const PixelData &p)
{
    int rint = (int)(p.avar[0] * 255);
    int gint = (int)(p.avar[1] * 255);
    int bint = (int)(p.avar[2] * 255);
    return (unsigned)(rint << 16 | gint << 8 | bint);
}

struct HPSCountId : PixelShaderBase<HPSCountId> {
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        int i = p.x + g_s->width * p.y;
        g_s->count[i]++;
        g_s->prim_id[i] = p.primitiveOrdinal;
    }
};
struct HPSGouraud : PixelShaderBase<HPSGouraud> {
    static const int AVarCount = 3;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        g_s->color[p.x + g_s->width * p.y] = packRGB(p);
    }
};
struct HPSGouraudDepth : PixelShaderBase<HPSGouraudDepth> {
    static const bool InterpolateZ = true;
    static const int AVarCount = 3;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        int i = p.x + g_s->width * p.y;
        if (p.z < g_s->depth[i]) { g_s->depth[i] = p.z; g_s->color[i] = packRGB(p); }
    }
};
struct HPSVaryDump : PixelShaderBase<HPSVaryDump> {
    static const bool InterpolateZ = true;
    static const bool InterpolateW = true;
    static const int AVarCount = 3;
    static const int PVarCount = 2;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        size_t n = (size_t)g_s->width * g_s->height, i = (size_t)p.x + (size_t)g_s->width * p.y;
        float *v = g_s->vary;
        v[0 * n + i] = p.z; v[1 * n + i] = p.w; v[2 * n + i] = p.invw;
        v[3 * n + i] = p.avar[0]; v[4 * n + i] = p.avar[1]; v[5 * n + i] = p.avar[2];
        v[6 * n + i] = p.pvar[0]; v[7 * n + i] = p.pvar[1];
        g_s->count[i]++;
    }
};

TextureView g_tex;

struct HPSTexturedAniso : PixelShaderBase<HPSTexturedAniso> {
    static const bool InterpolateW = true;
    static const int PVarCount = 2;
    static void drawPixel(const PixelData &p)
    {
        g_s->fragments++;
        float dudx, dudy, dvdx, dvdy;
        p.computePerspectiveDerivatives(*p.equations, 0, dudx, dudy);
        p.computePerspectiveDerivatives(*p.equations, 1, dvdx, dvdy);
        g_s->color[p.x + g_s->width * p.y] = textureSample(g_tex, p.pvar[0], p.pvar[1], dudx, dvdx, dudy, dvdy);
    }
};

template <class VS>
void shadeHost(const swr_scene *s, int index, CVert<VS::AVarCount, VS::PVarCount> &o)
{
    VertexShaderOutput out;
    VS::processVertex((const char *)s->vertices + (size_t)s->stride * (size_t)index, &out);
    o.x = out.x; o.y = out.y; o.z = out.z; o.w = out.w;
    for (int i = 0; i < VS::AVarCount; ++i) o.a[i] = out.avar[i];
    for (int i = 0; i < VS::PVarCount; ++i) o.p[i] = out.pvar[i];
}

struct Store {
    std::vector<Box16> bbox;
    std::vector<float4> head, span;
    std::vector<float> params;
    std::vector<uint32_t> order;     // record ids in emission order
};

template <class VS, class PS>
int run(swr_scene *s)
{
    constexpr int NA = VS::AVarCount, NP = VS::PVarCount;
    typedef CVert<NA, NP> V;
    typedef PsTraits<PS> TR;
    const int per = s->draw_mode + 1;
    const int64_t nprim = s->index_count / per;

    GeomArgs g;
    memset(&g, 0, sizeof(g));
    g.drawMode = s->draw_mode;
    g.px = s->vp_w / 2.0f; g.py = s->vp_h / 2.0f;
    g.ox = (s->vp_x + g.px); g.oy = (s->vp_y + g.py);
    g.depthN = s->depth_n; g.depthF = s->depth_f;
    g.cullMode = s->cull_mode; g.rasterMode = s->raster_mode;
    g.scMinX = s->sc_x; g.scMinY = s->sc_y; g.scMaxX = s->sc_x + s->sc_w; g.scMaxY = s->sc_y + s->sc_h;
    g.nA = TR::NA; g.nP = TR::NP; g.useZ = TR::Z; g.useW = TR::W;
    g.paramStride = paramFloats(s->draw_mode, g.nA, g.nP, g.useZ, g.useW);
    uint32_t errorFlags[2] = { 0, 0 };
    g.world = 1;
    g.noTightBox = getenv("HOSTCHECK_NO_TIGHT_BOX") ? 1 : 0;

    const size_t cap = (size_t)nprim * (s->draw_mode == 2 ? kMaxFan : 1) + 1;
    Store st;
    st.bbox.assign(cap, deadBox());
    st.head.resize(cap * 3);
    st.span.resize(cap * 3);
    st.params.resize(cap * (size_t)g.paramStride + 4);
    RecordSink &sink = g.sink[0];
    sink.bbox = st.bbox.data(); sink.head = st.head.data(); sink.span = st.span.data(); sink.params = st.params.data();
    sink.errorFlag = errorFlags;

    // ---- geometry stage, in emission order: per batch the original slots, then the fan extras
    uint32_t nextRec = 0;
    for (int64_t base = 0, batch = 0; base < nprim; base += kBatch, ++batch) {
        const int cnt = (int)std::min<int64_t>(kBatch, nprim - base);
        const uint32_t ord0 = (uint32_t)batch * SWR_ORDINAL_STRIDE;
        std::vector<uint32_t> extras;
        uint32_t slotExtra = (uint32_t)cnt;
        for (int i = 0; i < cnt; ++i) {
            const int32_t *ip = s->indices + (base + i) * per;
            const uint32_t rec = nextRec++;
            Box16 box = deadBox();
            if (s->draw_mode == 2) {
                V a[kMaxPoly], b[kMaxPoly], *poly = a;
                shadeHost<VS>(s, ip[0], a[0]); shadeHost<VS>(s, ip[1], a[1]); shadeHost<VS>(s, ip[2], a[2]);
                const int mask = outcode(a[0].x, a[0].y, a[0].z, a[0].w) | outcode(a[1].x, a[1].y, a[1].z, a[1].w) |
                                 outcode(a[2].x, a[2].y, a[2].z, a[2].w);
                int n = 3;
                if (mask) n = clipTriangle<NA, NP>(a, b, mask, &poly);
                if (n >= 3) {
                    box = emitClipTriangle<NA, NP>(g, sink, rec, ord0 + (uint32_t)i, poly[0], poly[1], poly[2]);
                    for (int k = 1; k + 2 < n; ++k) {
                        const uint32_t er = nextRec++;
                        st.bbox[er] = emitClipTriangle<NA, NP>(g, sink, er, ord0 + slotExtra++, poly[0], poly[k + 1], poly[k + 2]);
                        extras.push_back(er);
                    }
                }
            } else if (s->draw_mode == 1) {
                V c0, c1;
                shadeHost<VS>(s, ip[0], c0); shadeHost<VS>(s, ip[1], c1);
                box = emitClipLine<NA, NP>(g, sink, rec, ord0 + (uint32_t)i, c0, c1);
            } else {
                V c0;
                shadeHost<VS>(s, ip[0], c0);
                if (outcode(c0.x, c0.y, c0.z, c0.w) == 0) {
                    toScreen(g, c0);
                    box = emitScreenPoint<NA, NP>(g, sink, rec, ord0 + (uint32_t)i, c0);
                }
            }
            st.bbox[rec] = box;
            st.order.push_back(rec);
        }
        st.order.insert(st.order.end(), extras.begin(), extras.end());
    }

    // ---- raster stage: every record in emission order, every 8x8 block of its footprint
    TileArgs t;
    memset(&t, 0, sizeof(t));
    t.bbox = sink.bbox; t.head = sink.head; t.params = sink.params; t.span = sink.span; t.paramStride = g.paramStride;
    t.rtWidth = s->width; t.rtHeight = s->height;
    t.scMinX = g.scMinX; t.scMinY = g.scMinY; t.scMaxX = g.scMaxX; t.scMaxY = g.scMaxY;
    PixelData p;
    memset(&p, 0, sizeof(p));
    long statPruned = 0, pruneViolations = 0;
    long statLive = 0, statSingle = 0, statSingleZero = 0, statItems = 0, statZeroItems = 0, statZeroPrims = 0;
    for (uint32_t rec : st.order) {
        const Box16 bb = st.bbox[rec];
        if (bb.x0 > bb.x1) continue;
        s->primitives_out++;
        const bool single = (bb.x0 >> 3) == (bb.x1 >> 3) && (bb.y0 >> 3) == (bb.y1 >> 3);
        const uint64_t fragsBefore = s->fragments;
        statLive++;
        statSingle += single;
        for (int gy = bb.y0 & ~7; gy <= bb.y1; gy += 8) {
            for (int gx = bb.x0 & ~7; gx <= bb.x1; gx += 8) {
                uint64_t m;
                const float4 h0 = t.head[(size_t)rec * 3], h1 = t.head[(size_t)rec * 3 + 1], h2 = t.head[(size_t)rec * 3 + 2];
                if (s->draw_mode == 2) {
                    if (f2u(h2.y) & kModeSpan) m = coverSpan(t.span[(size_t)rec * 3], t.span[(size_t)rec * 3 + 1], t.span[(size_t)rec * 3 + 2], gx, gy, t.scMinX, t.scMaxX);
                    else {
                        // as the tile kernel calls it: only the part of the certified pixel bounds inside this block
                        m = coverBlock(h0, h1, h2, gx, gy, std::max((int)bb.x0 - gx, 0), std::max((int)bb.y0 - gy, 0),
                                       std::min((int)bb.x1 - gx, 7), std::min((int)bb.y1 - gy, 7));
                        if (!blockMayBeCovered(h0, h1, h2, gx, gy)) { statPruned++; if (m != 0) pruneViolations++; }
                    }
                } else if (s->draw_mode == 1) {
                    m = coverLine(h0, h1, gx, gy, t);
                } else {
                    const int lx = f2i(h0.x) - gx, ly = f2i(h0.y) - gy;
                    m = ((unsigned)lx < 8u && (unsigned)ly < 8u) ? 1ull << (ly * 8 + lx) : 0ull;
                }
                statItems++;
                statZeroItems += m == 0;
                for (int bit = 0; bit < 64; ++bit) {
                    if (!((m >> bit) & 1)) continue;
                    const int xx = bit & 7, yy = bit >> 3;
                    if (gx + xx >= s->width || gy + yy >= s->height) continue;
                    if (s->draw_mode == 2) shadeTriangleFragment<PS>(t, rec, gx, gy, xx, yy, p);
                    else if (s->draw_mode == 1) shadeLineFragments<PS>(t, rec, gx + xx, gy + yy, p);
                    else shadePointFragment<PS>(t, rec, gx + xx, gy + yy, p);
                }
            }
        }
        if (s->fragments == fragsBefore) { statZeroPrims++; statSingleZero += single; }
    }
    if (getenv("HOSTCHECK_STATS"))
        fprintf(stderr, "hostcheck stats: live %ld single-block %ld (%.1f%%) zero-fragment prims %ld (%.1f%%) of which single-block %ld; items %ld zero items %ld (%.1f%%)\n",
                statLive, statSingle, 100.0 * statSingle / statLive, statZeroPrims, 100.0 * statZeroPrims / statLive, statSingleZero,
                statItems, statZeroItems, 100.0 * statZeroItems / statItems);
    if (getenv("HOSTCHECK_STATS")) fprintf(stderr, "hostcheck stats: items pruned by blockMayBeCovered %ld, violations %ld\n", statPruned, pruneViolations);
    if (pruneViolations) return -77;      // the emptiness pre-test must never drop a covered block
    return (int)errorFlags[1];
}

template <class PS>
int runVS(swr_scene *s)
{
    if (s->vs_kind == SWR_VS_MVP_NORMAL_UV) return run<HostVS<2>, PS>(s);
    if (PS::PVarCount > 0) return -3;
    return run<HostVS<0>, PS>(s);
}

} // namespace

// Note: unlike the oracle, primitives_out counts records with a live footprint (the tile kernel's
// notion), and within one primitive the fragment order is block-major -- neither is observable in
// the per-pixel results the test compares.
extern "C" int hostcheck_draw(swr_scene *s)
{
    g_s = s;
    s->fragments = 0;
    s->primitives_out = 0;
    s->stream_len = 0;
    switch (s->ps_kind) {
    case SWR_PS_COUNT_ID: return runVS<HPSCountId>(s);
    case SWR_PS_GOURAUD: return runVS<HPSGouraud>(s);
    case SWR_PS_GOURAUD_DEPTH: return runVS<HPSGouraudDepth>(s);
    case SWR_PS_VARY_DUMP: return runVS<HPSVaryDump>(s);
    case 6: {   // SWR_PS_TEXTURED_ANISO
        if (s->draw_mode != 2 || !s->texture) return -5;
        std::vector<uint32_t> chain((size_t)s->tex_w * s->tex_h * 2 + 16);
        int32_t w[kMaxMipLevels], h[kMaxMipLevels];
        int64_t off[kMaxMipLevels];
        g_tex.levels = buildMipChain(s->texture, s->tex_w, s->tex_h, chain.data(), w, h, off);
        for (int i = 0; i < g_tex.levels; ++i) { g_tex.level[i] = chain.data() + off[i]; g_tex.w[i] = w[i]; g_tex.h[i] = h[i]; }
        g_tex.maxAnisotropy = 8;
        return runVS<HPSTexturedAniso>(s);
    }
    default: return -2;
    }
}


// ---- certified pixel bounds (geometry.cuh: tightPixelBounds) against the reference's full block walk ----------
// For `n` pseudo-random screen-space triangles of class `kind` the record is set up twice, with the certified
// bounds and with the reference's 8-aligned block box; the coverage of every block of the reference box is
// evaluated in full (coverBlock over the whole block) and (a) every covered pixel must lie inside the certified
// bounds, (b) the rectangle-limited coverBlock must return exactly the full mask.  Returns the number of
// violations; *checked = triangles that reached the rasterizer, *tightened = pixels excluded by the bounds.
namespace {
uint64_t rngState = 0;
double rnd01()
{
    rngState = rngState * 6364136223846793005ull + 1442695040888963407ull;
    return (double)((rngState >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53);
}
}
extern "C" long hostcheck_tight_box_fuzz(int kind, long n, unsigned long long seed, int width, int height, long *checked, long *tightened)
{
    typedef CVert<3, 0> V;
    rngState = seed * 2654435761ull + 12345u;
    GeomArgs g;
    memset(&g, 0, sizeof(g));
    g.drawMode = SWR_DRAW_TRIANGLE;
    g.cullMode = SWR_CULL_NONE; g.rasterMode = SWR_RASTER_BLOCK;
    g.scMinX = 0; g.scMinY = 0; g.scMaxX = width; g.scMaxY = height;
    g.nA = 0; g.nP = 0; g.useZ = 0; g.useW = 0;
    g.paramStride = 4;
    uint32_t errorFlags[2] = { 0, 0 };
    g.world = 1;
    std::vector<float4> head(6);
    std::vector<float> params(16);
    std::vector<float4> span(6);
    RecordSink &sink = g.sink[0];
    sink.head = head.data(); sink.params = params.data(); sink.span = span.data(); sink.errorFlag = errorFlags;
    long bad = 0;
    *checked = 0; *tightened = 0;
    for (long it = 0; it < n; ++it) {
        V v[3];
        memset(v, 0, sizeof(v));
        const double cx = rnd01() * width, cy = rnd01() * height;
        for (int k = 0; k < 3; ++k) {
            double x, y;
            switch (kind) {
            case 0: x = cx + (rnd01() - 0.5) * 3.0; y = cy + (rnd01() - 0.5) * 3.0; break;                     // tiny
            case 1: x = rnd01() * width; y = rnd01() * height; break;                                             // screen sized
            case 2: x = cx + (rnd01() - 0.5) * 40.0; y = cy + (rnd01() - 0.5) * 40.0; break;                      // a few blocks
            case 3: x = cx + (k == 2 ? 1e-3 * (rnd01() - 0.5) : 0.0) + (k == 1 ? 3000.0 * (rnd01() - 0.5) : 0.0); // slivers
                    y = cy + (k == 1 ? 3000.0 * (rnd01() - 0.5) : 0.0) * 0.37 + (k == 2 ? 2.0 * (rnd01() - 0.5) : 0.0); break;
            case 4: x = (double)(int)(cx + (rnd01() - 0.5) * 12.0) + 0.5; y = (double)(int)(cy + (rnd01() - 0.5) * 12.0) + 0.5; break;   // vertices on pixel centres
            case 5: x = (double)(int)(cx + (rnd01() - 0.5) * 20.0); y = (double)(int)(cy + (rnd01() - 0.5) * 20.0); break;             // vertices on pixel corners
            case 6: x = cx + (rnd01() - 0.5) * 0.05; y = cy + (rnd01() - 0.5) * 0.05; break;                      // far below a pixel
            default: x = (rnd01() - 0.25) * 2.0 * width; y = (rnd01() - 0.25) * 2.0 * height; break;              // partly off screen (guard band)
            }
            v[k].x = (float)x; v[k].y = (float)y; v[k].z = 0.5f; v[k].w = 1.0f;
        }
        g.noTightBox = 0;
        const Box16 tight = emitScreenTriangle<3, 0>(g, sink, 0, 0, true, v[0], v[1], v[2]);
        g.noTightBox = 1;
        const Box16 ref = emitScreenTriangle<3, 0>(g, sink, 1, 0, true, v[0], v[1], v[2]);
        if (ref.x0 > ref.x1) { if (tight.x0 <= tight.x1) ++bad; continue; }
        ++*checked;
        const float4 h0 = head[3], h1 = head[4], h2 = head[5];
        if (tight.x0 <= tight.x1 && (memcmp(&head[0], &head[3], 48) != 0)) ++bad;      // same record either way
        for (int gy = ref.y0 & ~7; gy <= ref.y1; gy += 8)
            for (int gx = ref.x0 & ~7; gx <= ref.x1; gx += 8) {
                const uint64_t full = coverBlock(h0, h1, h2, gx, gy);
                uint64_t lim = 0;
                const bool touches = tight.x0 <= tight.x1 && tight.x0 <= gx + 7 && tight.x1 >= gx && tight.y0 <= gy + 7 && tight.y1 >= gy;
                if (touches)
                    lim = coverBlock(h0, h1, h2, gx, gy, std::max((int)tight.x0 - gx, 0), std::max((int)tight.y0 - gy, 0),
                                     std::min((int)tight.x1 - gx, 7), std::min((int)tight.y1 - gy, 7));
                if (lim != full) ++bad;
                for (int bit = 0; bit < 64; ++bit) {
                    const int x = gx + (bit & 7), y = gy + (bit >> 3);
                    const bool inTight = tight.x0 <= tight.x1 && x >= tight.x0 && x <= tight.x1 && y >= tight.y0 && y <= tight.y1;
                    if (((full >> bit) & 1) && !inTight) ++bad;
                    if (!inTight) ++*tightened;
                }
            }
    }
    return bad;
}

extern "C" int hostcheck_draw_raster_triangles(swr_scene *, const float *, int64_t) { return -1; }
