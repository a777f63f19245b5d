// swr/detail/tile.cuh -- the tile kernel (K4+K5+K6 of SURVEY.md 2.3): order-preserving binning,
// coverage and shading of one screen tile per CTA, templated on the user's pixel shader so
// drawPixel is inlined.
//
// Per tile (T x T pixels, T = 32 or 64, so every reference 8x8 block lies in exactly one tile):
//   F1  read the tile's row of the tile x chunk bitmap      -> chunks that may touch the tile,
//   F2  test their 32-record group boxes                     -> groups,
//   F3  test the groups' record boxes                        -> queue of primitives,
//       each step is a block-wide ordered compaction (ballot-free prefix sums), so the queue is in
//       ascending (chunk, record) order = the reference's emission order, with no atomics and no sort;
//   A   one thread per (primitive, 8x8 block) item: exact coverage as a 64-bit mask
//       Block:  Rasterizer.h:257-305 corner classification (incl. the "special case" skip) and
//               the per-pixel add chains of PixelShaderBase.h:55-94 / EdgeData.h:45-66, walked only
//               inside the primitive's certified pixel bounds (geometry.cuh) with warp-uniform loops
//       Span:   per-row [xl, xr) of Rasterizer.h:360-411 (no edge tests)
//       lines:  the DDA walk of Rasterizer.h:175-207; points: Rasterizer.h:149-158
//   B   each 8x8 block is owned by one warp for the whole tile; it walks the block's items in
//       queue order, 32 fragments at a time (one lane per fragment, taken from a per-warp list the
//       item lanes write their (item, pixel) pairs into); fragments of different
//       primitives that hit the same pixel are serialised in emission order.  Every lane
//       replays the reference's incremental fp32 chain for its own pixel (PixelData.h:61-125),
//       so z / w / varyings are bit-identical, then calls PixelShader::drawPixel.
// Registered render targets are staged in shared memory (block-linear) for the whole tile and
// moved with 128-bit loads / stores.
#pragma once

#include "common.h"
#include "geometry.cuh"
#include "bin.cuh"
#include "../PixelShaderBase.h"
#include "../Uniforms.h"

#if defined(__CUDACC__)

namespace swr {
namespace detail {

#ifndef SWR_DENSE_MIN
#define SWR_DENSE_MIN 24      // items with at least this many covered pixels are shaded lane <-> pixel (measured 16 / 24 / 32 / 48:
                              // C3 tile phase 0.588 / 0.562 / 0.557 / 0.558 ms, fill 0.140 / 0.140 / 0.166 / 0.166, C0 at 4K 11.3 / 11.7 / 12.3 / 13.7)
#endif
#ifndef SWR_DENSE_PATH
#define SWR_DENSE_PATH 1
#endif

#ifndef SWR_QUEUE
#define SWR_QUEUE 1024
#endif
#ifndef SWR_ITEMS
#define SWR_ITEMS 4096
#endif
constexpr int kQueue = SWR_QUEUE;    // primitives per flush
constexpr int kItems = SWR_ITEMS;    // (primitive, block) items per flush
#ifndef SWR_CHUNK_LIST
#define SWR_CHUNK_LIST 1024
#endif
#ifndef SWR_GROUP_LIST
#define SWR_GROUP_LIST 2048
#endif
#ifndef SWR_FRAG_LIST
#define SWR_FRAG_LIST 256
#endif
constexpr int kChunkList = SWR_CHUNK_LIST;
constexpr int kGroupList = SWR_GROUP_LIST;    // >= the per-tile list capacity of the binning pass (runtime.cu)
constexpr int kFragList = SWR_FRAG_LIST;      // fragments of one sparse run (per warp)
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kPruneMax = 16;        // primitives with at most this many (primitive, block) items get the emptiness pre-test
static_assert(2 * kTileThreads >= kQueue, "the item re-indexing scan handles two queue entries per thread");
static_assert(4 * kTileThreads < (1 << 12) && 4 * 64 * kTileThreads < (1 << 20), "packing of the 32-bit record scan (count << 20 | items)");

template <int TLOG, int NRT>
struct TileSmem {
    static constexpr int T = 1 << TLOG;
    static constexpr int BPR = T / 8;                 // blocks per tile row
    static constexpr int NB = BPR * BPR;              // blocks per tile
    static constexpr int QW = kQueue / 32;            // bitmap words per block
    static constexpr size_t rtBytes = (size_t)NRT * T * T * 4;
    static constexpr size_t offMasks = (rtBytes + 15) & ~(size_t)15;
    static constexpr size_t offBlockmap = offMasks + (size_t)kItems * 8;
    static constexpr size_t offQRec = offBlockmap + (size_t)NB * QW * 4;
    static constexpr size_t offQRange = offQRec + (size_t)kQueue * 4;
    static constexpr size_t offQItem = offQRange + (size_t)kQueue * 4;
    static constexpr size_t offQValid = offQItem + (size_t)(kQueue + 1) * 4 + 12;
    static constexpr size_t offChunkGroup = offQValid + (size_t)kQueue * 4;
    static constexpr size_t offChunkPair = offChunkGroup + (size_t)kChunkList * 4;
    static constexpr size_t offGroupList = offChunkPair + (size_t)(kChunkList + 1) * 4 + 12;
    static constexpr size_t offFragList = offGroupList + (size_t)kGroupList * 4;
    static constexpr size_t offScan = offFragList + (size_t)kFragList * 2 * (kTileThreads / 32);
    static constexpr size_t offCtl = offScan + 2 * (kTileWarps + 1) * 8;
    static constexpr size_t bytes = offCtl + 64;
};

struct Ctl { uint32_t accQ, accItems; int nextBlock; };

// Exclusive block scan of a packed (hi: count, lo: sum) pair.  ONE barrier: every warp publishes its
// total, then each warp scans the (at most 32) warp totals for itself.  The scratch is double buffered,
// so back-to-back calls need no trailing barrier (a warp can be at most one call behind the others).
template <class T>
SWR_D T blockScanT(T v, T &total, uint64_t *scratch, int &phase)
{
    static_assert(kTileWarps <= 32, "one lane per warp total");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T *s = reinterpret_cast<T *>(scratch + phase * (kTileWarps + 1));
    phase ^= 1;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s[wid] = incl;
    __syncthreads();
    const T w = lane < kTileWarps ? s[lane] : 0;
    T wi = w;
#pragma unroll
    for (int o = 1; o < kTileWarps; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += n;
    }
    total = __shfl_sync(0xffffffffu, wi, kTileWarps - 1);
    return incl - v + __shfl_sync(0xffffffffu, wi - w, wid);
}
SWR_D uint64_t blockScan(uint64_t v, uint64_t &total, uint64_t *scratch, int &phase) { return blockScanT<uint64_t>(v, total, scratch, phase); }
// 32-bit flavour (half the shuffles) for the steps whose packed pair fits: hi 12 bits count, lo 20 bits sum
SWR_D uint32_t blockScan32(uint32_t v, uint32_t &total, uint64_t *scratch, int &phase) { return blockScanT<uint32_t>(v, total, scratch, phase); }

#ifndef SWR_PREFETCH
#define SWR_PREFETCH 7
#endif
#ifndef SWR_TILE_STATS
#define SWR_TILE_STATS 0          // per-tile / per-phase clocks (tools/tile_stats.py builds its own library variant)
#endif
SWR_D void prefetchL2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// qRange: a queued primitive's pixel bounds inside the tile, 6 bits each (tiles are at most 64 pixels wide)
SWR_HD uint32_t packRange(int lx0, int ly0, int lx1, int ly1) { return (uint32_t)lx0 | ((uint32_t)ly0 << 6) | ((uint32_t)lx1 << 12) | ((uint32_t)ly1 << 18); }
SWR_HD int rangeBx0(uint32_t rg) { return (int)((rg >> 3) & 7u); }
SWR_HD int rangeBy0(uint32_t rg) { return (int)((rg >> 9) & 7u); }
SWR_HD int rangeBx1(uint32_t rg) { return (int)((rg >> 15) & 7u); }
SWR_HD int rangeBy1(uint32_t rg) { return (int)((rg >> 21) & 7u); }

SWR_D bool boxOverlaps(const Box16 b, int X0, int Y0, int X1, int Y1)
{
    return b.x0 <= b.x1 && b.x0 <= X1 && b.x1 >= X0 && b.y0 <= Y1 && b.y1 >= Y0;
}

// li / nx for 0 <= li < 64, 1 <= nx <= 8 (item index -> block row of a primitive's block range) without the
// ~25-instruction integer division: floor(li * ceil(256 / nx) / 256), exact on that range (checked exhaustively)
static __constant__ uint16_t kCeil256Over[9] = { 0, 256, 128, 86, 64, 52, 43, 37, 32 };
SWR_D int divSmall(int li, int nx) { return (li * (int)kCeil256Over[nx]) >> 8; }

// position of the n-th (0-based) set bit: popcount binary search (the __fns intrinsic is a slow loop)
SWR_D int nthSetBit32(uint32_t w, int n)
{
    int base = 0, c;
    c = __popc(w & 0xffffu); if (n >= c) { n -= c; w >>= 16; base = 16; }
    c = __popc(w & 0xffu);   if (n >= c) { n -= c; w >>= 8;  base += 8; }
    c = __popc(w & 0xfu);    if (n >= c) { n -= c; w >>= 4;  base += 4; }
    c = __popc(w & 0x3u);    if (n >= c) { n -= c; w >>= 2;  base += 2; }
    c = (int)(w & 1u);       if (n >= c) base += 1;
    return base;
}
SWR_D int nthSetBit64(uint64_t m, int n)
{
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    const int cl = __popc(lo);
    return n < cl ? nthSetBit32(lo, n) : 32 + nthSetBit32(hi, n - cl);
}

// ---- coverage of one 8x8 block ------------------------------------------------------------------
// Block mode: Rasterizer.h:257-305 + PixelShaderBase.h:55-94.  (gx, gy) = block origin.
//
// Inside test (EdgeEquation.h:58-61)  v > 0 || (v == 0 && tie)  as ONE compare:  v > thr  with
// thr = 0 for tie == false and thr = -denorm_min for tie == true (no float lies between the two,
// -0 == 0, NaN fails both forms).
//
// Row skipping is exact, not conservative-approximate: along a row the reference's value chain
// v, v+a, (v+a)+a, ... is monotone (an fp32 add of a constant is monotone), so its maximum is the
// first value when a <= 0 and the last one when a > 0; if that maximum fails an edge's test no
// pixel of the row can pass, and the 8 x 3 per-pixel tests are skipped.
struct BlockEdges { float ea[3], eb[3], thr[3], e00[3]; };
constexpr int kNarrowCols = 3;      // coverRect: rectangles of at most this many columns (warp-wide) walk them in a loop

// Corner classification of one block (Rasterizer.h:272-303): 0 = skipped ("all out" with the reference's special
// case), 1 = fully covered (drawBlock<false>: all 64 pixels, no edge tests), 2 = partially covered (drawBlock<true>).
SWR_HD int classifyBlock(const float4 h0, const float4 h1, const float4 h2, int gx, int gy, BlockEdges &E)
{
    E.ea[0] = h0.x; E.ea[1] = h0.w; E.ea[2] = h1.z;
    E.eb[0] = h0.y; E.eb[1] = h1.x; E.eb[2] = h1.w;
    const float ec[3] = { h0.z, h1.y, h2.x };
    const uint32_t flags = f2u(h2.y);
    const float negTiny = u2f(0x80000001u);
    E.thr[0] = (flags & kTie0) ? negTiny : 0.0f; E.thr[1] = (flags & kTie1) ? negTiny : 0.0f; E.thr[2] = (flags & kTie2) ? negTiny : 0.0f;
    const float xf = fadd(i2f(gx), 0.5f), yf = fadd(i2f(gy), 0.5f);
    const float s = 7.0f;
    bool in[4][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a7 = fmul(E.ea[k], s);
        E.e00[k] = fadd(fadd(fmul(E.ea[k], xf), fmul(E.eb[k], yf)), ec[k]);   // EdgeData.h:37-42
        const float e01 = fadd(E.e00[k], fmul(E.eb[k], s));                   // stepY(s)
        const float e10 = fadd(E.e00[k], a7);                                 // stepX(s)
        const float e11 = fadd(e01, a7);
        in[0][k] = E.e00[k] > E.thr[k]; in[1][k] = e01 > E.thr[k];
        in[2][k] = e10 > E.thr[k]; in[3][k] = e11 > E.thr[k];
    }
    int all = 0;
    bool same = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        all += (in[j][0] && in[j][1] && in[j][2]) ? 1 : 0;
        same = same && ((in[j][0] == in[j][1]) == in[j][2]);               // C++ chained '==' (Rasterizer.h:287-290)
    }
    if (all == 4) return 1;
    if (all == 0 && same) return 0;
    return 2;
}

// Per-pixel tests of a partially covered block inside the pixel rectangle (rx0, ry0)-(rx1, ry1) (the part of the
// primitive's certified bounds in this block, geometry.cuh: tightPixelBounds), as R rows x C columns starting at
// (rx0, ry0).  R and C are the same for every lane of a warp (the warp's maxima), so the loops never diverge;
// rows / columns past a lane's own rectangle are masked off.  (The reference walks all 64 pixels of such a block; the
// rows it adds nothing in are exactly the ones outside the certified bounds.)  The values are the reference's: the rows above and
// the columns left of the rectangle still take their adds (the chains are replayed, not re-derived).
SWR_HD uint64_t coverRect(const BlockEdges &E, int rx0, int ry0, int rx1, int ry1, int R, int C)
{
    float r0 = E.e00[0], r1 = E.e00[1], r2 = E.e00[2];
#pragma unroll
    for (int i = 0; i < 7; ++i)
        if (i < ry0) { r0 = fadd(r0, E.eb[0]); r1 = fadd(r1, E.eb[1]); r2 = fadd(r2, E.eb[2]); }
    const uint32_t colMask = rx0 <= rx1 ? ((0xffu >> (7 - rx1)) & (0xffu << rx0)) : 0u;
    uint64_t mask = 0;
#pragma unroll 1
    for (int rr = 0; rr < R; ++rr) {
        float v0 = r0, v1 = r1, v2 = r2;
        uint32_t rowMask = 0;
        if (C > kNarrowCols) {
            // all 8 pixels of the row, unrolled (4 instructions per test, no loop, no lead-in)
#pragma unroll
            for (int xx = 0; xx < 8; ++xx) {
#if defined(__CUDA_ARCH__)
                // the three compares chained into one predicate + a predicated OR (nvcc materialises them with SELs otherwise)
                asm("{\n\t.reg .pred p;\n\t"
                    "setp.gt.f32 p, %1, %4;\n\t"
                    "setp.gt.and.f32 p, %2, %5, p;\n\t"
                    "setp.gt.and.f32 p, %3, %6, p;\n\t"
                    "@p or.b32 %0, %0, %7;\n\t}"
                    : "+r"(rowMask)
                    : "f"(v0), "f"(v1), "f"(v2), "f"(E.thr[0]), "f"(E.thr[1]), "f"(E.thr[2]), "r"(1u << xx));
#else
                if (v0 > E.thr[0] && v1 > E.thr[1] && v2 > E.thr[2]) rowMask |= 1u << xx;
#endif
                if (xx < 7) { v0 = fadd(v0, E.ea[0]); v1 = fadd(v1, E.ea[1]); v2 = fadd(v2, E.ea[2]); }
            }
        } else {
            // a few columns: lead-in adds up to the rectangle, then C tests
#pragma unroll
            for (int i = 0; i < 7; ++i)
                if (i < rx0) { v0 = fadd(v0, E.ea[0]); v1 = fadd(v1, E.ea[1]); v2 = fadd(v2, E.ea[2]); }
            uint32_t bit = 1u << rx0;
#pragma unroll 1
            for (int cc = 0; cc < C; ++cc) {
#if defined(__CUDA_ARCH__)
                asm("{\n\t.reg .pred p;\n\t"
                    "setp.gt.f32 p, %1, %4;\n\t"
                    "setp.gt.and.f32 p, %2, %5, p;\n\t"
                    "setp.gt.and.f32 p, %3, %6, p;\n\t"
                    "@p or.b32 %0, %0, %7;\n\t}"
                    : "+r"(rowMask)
                    : "f"(v0), "f"(v1), "f"(v2), "f"(E.thr[0]), "f"(E.thr[1]), "f"(E.thr[2]), "r"(bit));
#else
                if (v0 > E.thr[0] && v1 > E.thr[1] && v2 > E.thr[2]) rowMask |= bit;
#endif
                bit <<= 1;
                v0 = fadd(v0, E.ea[0]); v1 = fadd(v1, E.ea[1]); v2 = fadd(v2, E.ea[2]);
            }
        }
        const int yy = ry0 + rr;
        if (yy <= ry1) mask |= (uint64_t)(rowMask & colMask) << (yy * 8);
        r0 = fadd(r0, E.eb[0]); r1 = fadd(r1, E.eb[1]); r2 = fadd(r2, E.eb[2]);
    }
    return mask;
}

// Scalar form (host checks, and the reference for the warp-cooperative use in the tile kernel).
SWR_HD uint64_t coverBlock(const float4 h0, const float4 h1, const float4 h2, int gx, int gy, int rx0, int ry0, int rx1, int ry1)
{
    BlockEdges E;
    const int cls = classifyBlock(h0, h1, h2, gx, gy, E);
    if (cls == 1) return ~0ull;                                            // drawBlock<false>
    if (cls == 0) return 0ull;                                             // "special case": block skipped
    return coverRect(E, rx0, ry0, rx1, ry1, ry1 - ry0 + 1, rx1 - rx0 + 1);
}
SWR_HD uint64_t coverBlock(const float4 h0, const float4 h1, const float4 h2, int gx, int gy)
{
    return coverBlock(h0, h1, h2, gx, gy, 0, 0, 7, 7);
}

// Exact emptiness test of one 8x8 block (Block mode): for every edge the largest of the 64 per-pixel
// chain values sits at pixel (a > 0 ? 7 : 0, b > 0 ? 7 : 0) -- the row-start chain is monotone in the
// row, and a row's chain is monotone in the column and in its start value -- so if that one value,
// computed by the reference's own additions, fails the edge's inside test, no pixel of the block can
// pass: coverBlock would return 0.  ~60 instructions instead of a full coverBlock.
SWR_HD bool blockMayBeCovered(const float4 h0, const float4 h1, const float4 h2, int gx, int gy)
{
    const float ea[3] = { h0.x, h0.w, h1.z }, eb[3] = { h0.y, h1.x, h1.w }, ec[3] = { h0.z, h1.y, h2.x };
    const uint32_t flags = f2u(h2.y);
    const float negTiny = u2f(0x80000001u);
    const float thr[3] = { (flags & kTie0) ? negTiny : 0.0f, (flags & kTie1) ? negTiny : 0.0f, (flags & kTie2) ? negTiny : 0.0f };
    const float xf = fadd(i2f(gx), 0.5f), yf = fadd(i2f(gy), 0.5f);
    bool may = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = fadd(fadd(fmul(ea[k], xf), fmul(eb[k], yf)), ec[k]);
        const bool up = eb[k] > 0, right = ea[k] > 0;
#pragma unroll
        for (int s = 0; s < 7; ++s) if (up) v = fadd(v, eb[k]);
#pragma unroll
        for (int s = 0; s < 7; ++s) if (right) v = fadd(v, ea[k]);
        const bool ordered = (ea[k] <= 0 || right) && (eb[k] <= 0 || up);     // false only for NaN coefficients
        if (ordered && !(v > thr[k])) may = false;
    }
    return may;
}

SWR_HD SpanHalf loadHalf(const float4 v, uint32_t y0, uint32_t y1)
{
    SpanHalf h; h.vx = v.x; h.vy = v.y; h.inv1 = v.z; h.inv2 = v.w; h.y0 = (int)y0; h.y1 = (int)y1;
    return h;
}

// Span mode: rows of the two halves, [xl, xr) per row (Rasterizer.h:360-411).
SWR_HD uint64_t coverSpan(const float4 s0, const float4 s1, const float4 s2, int gx, int gy, int scMinX, int scMaxX)
{
    const SpanHalf bot = loadHalf(s0, f2u(s2.x), f2u(s2.y)), top = loadHalf(s1, f2u(s2.z), f2u(s2.w));
    uint64_t mask = 0;
#pragma unroll
    for (int yy = 0; yy < 8; ++yy) {
        const int row = gy + yy;
        const bool inBot = row >= bot.y0 && row < bot.y1, inTop = row >= top.y0 && row < top.y1;
        if (!inBot && !inTop) continue;
        int xl, xr;
        spanRow(inBot ? bot : top, row, scMinX, scMaxX, xl, xr);
        const int a = imax(xl - gx, 0), b = imin(xr - gx, 8);
        if (a < b) mask |= (uint64_t)((0xffu >> (8 - (b - a))) << a) << (yy * 8);
    }
    return mask;
}

// Lines: pixels of the DDA walk that fall into this block and pass the float scissor test.
SWR_HD uint64_t coverLine(const float4 h0, const float4 h1, int gx, int gy, const TileArgs &t)
{
    float x = h0.x, y = h0.y;
    const int steps = (int)f2u(h1.x);
    uint64_t mask = 0;
    for (int k = 0; k < steps; ++k) {
        if (scissorTest(t.scMinX, t.scMinY, t.scMaxX, t.scMaxY, x, y)) {
            const int lx = f2i(x) - gx, ly = f2i(y) - gy;
            if ((unsigned)lx < 8u && (unsigned)ly < 8u) mask |= 1ull << (ly * 8 + lx);
        }
        x = fadd(x, h0.z);
        y = fadd(y, h0.w);
    }
    return mask;
}

// ---- shading --------------------------------------------------------------------------------------
template <class PS>
struct PsTraits {
    static constexpr int NA = PS::AVarCount, NP = PS::PVarCount;
    static constexpr bool Z = PS::InterpolateZ != 0, W = PS::InterpolateW != 0;
    static constexpr bool WP = W || NP > 0;               // PixelData.h:67,89,107
    static constexpr int NRT = PS::RenderTargets;
};

// One fragment of a triangle.  Block mode: start at the block origin, `yy` row steps, then `xx`
// column steps (PixelShaderBase.h:61-92).  Span mode: start at (xl + .5, y + .5), x - xl column
// steps (PixelShaderBase.h:96-112).  All lanes run the same instruction stream (the 7+7 / span
// steps are predicated adds), so there is no divergence between fragments of a round.
#pragma nv_exec_check_disable
template <class PS>
SWR_HD void shadeTriangleFragment(const TileArgs &t, uint32_t rec, int gx, int gy, int xx, int yy, PixelData &p)
{
    typedef PsTraits<PS> TR;
    const float4 h2 = t.head[(size_t)rec * 3 + 2];
    const uint32_t flags = f2u(h2.y);
    const float *pl = t.params + (size_t)rec * t.paramStride;
    p.primitiveOrdinal = f2u(h2.z);
    p.x = gx + xx;
    p.y = gy + yy;

    float xf, yf;
    int nrow, ncol;
    if (flags & kModeSpan) {
        const float4 *sp = t.span + (size_t)rec * 3;
        const float4 s2 = sp[2];
        const SpanHalf bot = loadHalf(sp[0], f2u(s2.x), f2u(s2.y)), top = loadHalf(sp[1], f2u(s2.z), f2u(s2.w));
        const bool inBot = p.y >= bot.y0 && p.y < bot.y1;
        const int xl = imax(t.scMinX, f2i(spanX(inBot ? bot.vx : top.vx, inBot ? bot.vy : top.vy, inBot ? bot.inv1 : top.inv1, p.y)));
        xf = fadd(i2f(xl), 0.5f);
        yf = fadd(i2f(p.y), 0.5f);
        nrow = 0;
        ncol = p.x - xl;
    } else {
        xf = fadd(i2f(gx), 0.5f);
        yf = fadd(i2f(gy), 0.5f);
        nrow = yy;
        ncol = xx;
    }

    TriangleEquations eq;
    {
        const float4 h0 = t.head[(size_t)rec * 3], h1 = t.head[(size_t)rec * 3 + 1];
        eq.area2 = h2.w;
        eq.e0.a = h0.x; eq.e0.b = h0.y; eq.e0.c = h0.z; eq.e0.tie = (flags & kTie0) != 0;
        eq.e1.a = h0.w; eq.e1.b = h1.x; eq.e1.c = h1.y; eq.e1.tie = (flags & kTie1) != 0;
        eq.e2.a = h1.z; eq.e2.b = h1.w; eq.e2.c = h2.x; eq.e2.tie = (flags & kTie2) != 0;
    }

    // All planes walk the same steps, so the steps are the outer loop and up to kChainGroup planes the inner
    // one: one step predicate serves the whole group (per plane on its own it is recomputed for every plane:
    // 14 compares each).  Per plane the additions and their order are untouched.
    constexpr int NPL = (TR::Z ? 1 : 0) + (TR::WP ? 1 : 0) + TR::NA + TR::NP;
    constexpr int kChainGroup = 6;
    const float4 *pl4 = reinterpret_cast<const float4 *>(pl);      // one 128-bit load per plane
    float val[NPL > 0 ? NPL : 1];
#pragma unroll
    for (int g0 = 0; g0 < NPL; g0 += kChainGroup) {
        float ca[kChainGroup], cb[kChainGroup], cv[kChainGroup];
#pragma unroll
        for (int i = 0; i < kChainGroup; ++i) {
            if (g0 + i < NPL) {
                const float4 q = pl4[g0 + i];
                ca[i] = q.x; cb[i] = q.y;
                cv[i] = fadd(fadd(fmul(q.x, xf), fmul(q.y, yf)), q.z);      // ParameterEquation::evaluate
            }
        }
        if (flags & kModeSpan) {
            for (int s = 0; s < ncol; ++s) {
#pragma unroll
                for (int i = 0; i < kChainGroup; ++i)
                    if (g0 + i < NPL) cv[i] = fadd(cv[i], ca[i]);
            }
        } else {
#pragma unroll
            for (int s = 0; s < 7; ++s) {                                    // stepY
                const bool on = s < nrow;
#pragma unroll
                for (int i = 0; i < kChainGroup; ++i)
                    if (g0 + i < NPL && on) cv[i] = fadd(cv[i], cb[i]);
            }
#pragma unroll
            for (int s = 0; s < 7; ++s) {                                    // stepX
                const bool on = s < ncol;
#pragma unroll
                for (int i = 0; i < kChainGroup; ++i)
                    if (g0 + i < NPL && on) cv[i] = fadd(cv[i], ca[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < kChainGroup; ++i)
            if (g0 + i < NPL) val[g0 + i] = cv[i];
    }

    int o = 0;
    if (TR::Z) {
        const float4 q = pl4[o];
        eq.z.a = q.x; eq.z.b = q.y; eq.z.c = q.z;
        p.z = val[o++];
    }
    if (TR::WP) {
        const float4 q = pl4[o];
        eq.invw.a = q.x; eq.invw.b = q.y; eq.invw.c = q.z;
        p.invw = val[o++];
        p.w = frcp(p.invw);
    }
    eq.avar.planes = pl + 4 * o;
#pragma unroll
    for (int i = 0; i < TR::NA; ++i) p.avar[i] = val[o + i];
    o += TR::NA;
    eq.pvar.planes = pl + 4 * o;
#pragma unroll
    for (int i = 0; i < TR::NP; ++i) {
        p.pvarTemp[i] = val[o + i];
        p.pvar[i] = fmul(p.pvarTemp[i], p.w);
    }
    p.equations = &eq;
    PS::drawPixel(p);
}

// All hits of pixel (px, py) by one line, in step order (Rasterizer.h:175-207, 160-173).
#pragma nv_exec_check_disable
template <class PS>
SWR_HD int shadeLineFragments(const TileArgs &t, uint32_t rec, int px, int py, PixelData &p)
{
    typedef PsTraits<PS> TR;
    const float4 h0 = t.head[(size_t)rec * 3], h1 = t.head[(size_t)rec * 3 + 1];
    const float *pl = t.params + (size_t)rec * t.paramStride;
    const int steps = (int)f2u(h1.x);
    p.primitiveOrdinal = f2u(h1.y);
    p.equations = nullptr;
    float x = h0.x, y = h0.y;
    int o = 0;
    float z = 0.0f, dz = 0.0f, w = 0.0f, dw = 0.0f;
    float av[TR::NA > 0 ? TR::NA : 1], da[TR::NA > 0 ? TR::NA : 1], pv[TR::NP > 0 ? TR::NP : 1], dp[TR::NP > 0 ? TR::NP : 1];
    if (TR::Z) { z = pl[o]; dz = pl[o + 1]; o += 2; }
    if (TR::W) { w = pl[o]; dw = pl[o + 1]; o += 2; }
#pragma unroll
    for (int i = 0; i < TR::NA; ++i) { av[i] = pl[o]; da[i] = pl[o + 1]; o += 2; }
#pragma unroll
    for (int i = 0; i < TR::NP; ++i) { pv[i] = pl[o]; dp[i] = pl[o + 1]; o += 2; }
    int drawn = 0;
    for (int k = 0; k < steps; ++k) {
        if (f2i(x) == px && f2i(y) == py && scissorTest(t.scMinX, t.scMinY, t.scMaxX, t.scMaxY, x, y)) {
            p.x = px;
            p.y = py;
            if (TR::Z) p.z = z;
            if (TR::W) { p.w = w; p.invw = frcp(w); }
#pragma unroll
            for (int i = 0; i < TR::NA; ++i) p.avar[i] = av[i];
#pragma unroll
            for (int i = 0; i < TR::NP; ++i) p.pvar[i] = pv[i];
            PS::drawPixel(p);
            ++drawn;
        }
        x = fadd(x, h0.z);
        y = fadd(y, h0.w);
        if (TR::Z) z = fadd(z, dz);
        if (TR::W) w = fadd(w, dw);
#pragma unroll
        for (int i = 0; i < TR::NA; ++i) av[i] = fadd(av[i], da[i]);
#pragma unroll
        for (int i = 0; i < TR::NP; ++i) pv[i] = fadd(pv[i], dp[i]);
    }
    return drawn;
}

#pragma nv_exec_check_disable
template <class PS>
SWR_HD void shadePointFragment(const TileArgs &t, uint32_t rec, int px, int py, PixelData &p)
{
    typedef PsTraits<PS> TR;
    const float4 h1 = t.head[(size_t)rec * 3 + 1];
    const float *pl = t.params + (size_t)rec * t.paramStride;
    p.primitiveOrdinal = f2u(h1.y);
    p.equations = nullptr;
    p.x = px;
    p.y = py;
    int o = 0;
    if (TR::Z) { p.z = pl[o]; o += 1; }
    if (TR::W) { p.w = pl[o]; p.invw = frcp(p.w); o += 1; }
#pragma unroll
    for (int i = 0; i < TR::NA; ++i) p.avar[i] = pl[o + i];
    o += TR::NA;
#pragma unroll
    for (int i = 0; i < TR::NP; ++i) p.pvar[i] = pl[o + i];
    PS::drawPixel(p);
}

// ---- staged render targets ------------------------------------------------------------------------
// Shared-memory layout is block-linear: pixel (lx, ly) of the tile lives at word
// ((ly/8)*BPR + lx/8)*64 + (ly%8)*8 + lx%8, so one 8x8 block is 256 contiguous bytes and a full
// block round of 32 lanes touches 32 distinct banks.
// `sub` selects what is moved: 0 = the whole tile, 1 + q = quadrant q (bit 0: right half, bit 1: lower half) of it
// (heavy-tile split: the CTA only owns that quadrant's pixels).
template <int TLOG, bool STORE>
SWR_D void moveSlot(const TileArgs &t, char *g, int pitch, uint32_t *sm, int X0, int Y0, int sub)
{
    constexpr int T = 1 << TLOG, BPR = T / 8;
    const int rlog = sub ? TLOG - 1 : TLOG, R = 1 << rlog;                 // region size
    const int ox = (sub && ((sub - 1) & 1)) ? T / 2 : 0, oy = (sub && ((sub - 1) & 2)) ? T / 2 : 0;
    const bool vec = ((((uintptr_t)g) | (uintptr_t)pitch) & 15) == 0 && (t.rtWidth & 3) == 0;
    if (vec) {
        for (int i = threadIdx.x; i < (R * R) / 4; i += kTileThreads) {
            const int ly = oy + (i >> (rlog - 2)), lx = ox + (i & (R / 4 - 1)) * 4;
            const int x = X0 + lx, y = Y0 + ly;
            if (x < t.rtWidth && y < t.rtHeight) {
                uint4 *gp = (uint4 *)(g + (size_t)y * pitch + (size_t)x * 4);
                uint4 *sp = (uint4 *)(sm + (((ly >> 3) * BPR + (lx >> 3)) * 64 + (ly & 7) * 8 + (lx & 7)));
                if (STORE) *gp = *sp; else *sp = *gp;
            }
        }
    } else {
        for (int i = threadIdx.x; i < R * R; i += kTileThreads) {
            const int ly = oy + (i >> rlog), lx = ox + (i & (R - 1));
            const int x = X0 + lx, y = Y0 + ly;
            if (x < t.rtWidth && y < t.rtHeight) {
                uint32_t *gp = (uint32_t *)(g + (size_t)y * pitch + (size_t)x * 4);
                uint32_t *sp = sm + (((ly >> 3) * BPR + (lx >> 3)) * 64 + (ly & 7) * 8 + (lx & 7));
                if (STORE) *gp = *sp; else *sp = *gp;
            }
        }
    }
}

template <int TLOG, int NRT, bool STORE>
SWR_D void moveTile(const TileArgs &t, char *rtSmem, int X0, int Y0, int sub)
{
    constexpr int T = 1 << TLOG;
#pragma unroll 1
    for (int s = 0; s < NRT; ++s)
        moveSlot<TLOG, STORE>(t, (char *)t.rt[s].ptr, t.rt[s].pitch, (uint32_t *)(rtSmem + (size_t)s * T * T * 4), X0, Y0, sub);
    // Sort-first composite fused into the store: the finished tile of one slot also goes to the peers'
    // surfaces (plain stores through NVLink peer mappings; they overlap the tiles still being shaded).
    if (STORE && t.mirrorCount > 0 && t.mirrorSlot < NRT) {
#pragma unroll 1
        for (int m = 0; m < t.mirrorCount; ++m)
            moveSlot<TLOG, true>(t, (char *)t.mirror[m], t.rt[t.mirrorSlot].pitch, (uint32_t *)(rtSmem + (size_t)t.mirrorSlot * T * T * 4), X0, Y0, sub);
    }
}

// ---- the kernel -----------------------------------------------------------------------------------
template <class PS, int MODE, int TLOG>
__global__ void __launch_bounds__(kTileThreads, SWR_TILE_MINB) tileKernel(const TileArgs t)
{
    typedef PsTraits<PS> TR;
    typedef TileSmem<TLOG, TR::NRT> SM;
    constexpr int T = SM::T, BPR = SM::BPR, NB = SM::NB, QW = SM::QW;

    // Which pixels, and whose lists: normally one CTA per tile of the grid.  Heavy-tile split (common.h): the first
    // 4 * splitCap CTAs of the launch (they are scheduled first) each take one quadrant of a tile the binning pass
    // flagged as heavy -- same shared-memory tile, same bitmap row and group list, but the record boxes are clipped
    // to the quadrant and only its pixels are loaded, shaded and stored -- and the tile's own CTA exits.
    const uint32_t nSplit = 4u * (uint32_t)t.splitCap;
    int listTile, sub = 0;
    if (blockIdx.x < nSplit) {
        const uint32_t hi = blockIdx.x >> 2;
        if (hi >= min(t.heavyList[0], (uint32_t)t.splitCap)) return;
        listTile = (int)t.heavyList[1 + hi];
        sub = 1 + (int)(blockIdx.x & 3u);
    } else {
        listTile = (int)(blockIdx.x - nSplit);
        if (!tileOwned(listTile % t.tilesX, listTile / t.tilesX, t.rank, t.world)) return;
        if (nSplit && t.heavyFlag[listTile]) return;
    }
    const int X0 = (listTile % t.tilesX) << TLOG, Y0 = (listTile / t.tilesX) << TLOG;      // origin of the shared-memory tile
    // the pixels this CTA owns (records are clipped to them)
    const int CX0 = X0 + ((sub && ((sub - 1) & 1)) ? (1 << TLOG) / 2 : 0), CY0 = Y0 + ((sub && ((sub - 1) & 2)) ? (1 << TLOG) / 2 : 0);
    const int X1 = CX0 + (sub ? (1 << TLOG) / 2 : (1 << TLOG)) - 1, Y1 = CY0 + (sub ? (1 << TLOG) / 2 : (1 << TLOG)) - 1;
    if (CX0 >= t.rtWidth || CY0 >= t.rtHeight) return;
    if (*t.errorFlag & 1u) return;                           // geometry scratch exhausted: the draw is void

    extern __shared__ __align__(16) char smem[];
    char *rtSmem = smem;
    uint64_t *sMasks = (uint64_t *)(smem + SM::offMasks);
    uint32_t *sBlockmap = (uint32_t *)(smem + SM::offBlockmap);
    uint32_t *qRec = (uint32_t *)(smem + SM::offQRec);
    uint32_t *qRange = (uint32_t *)(smem + SM::offQRange);
    uint32_t *qItem = (uint32_t *)(smem + SM::offQItem);
    uint32_t *qValid = (uint32_t *)(smem + SM::offQValid);
    uint32_t *cGroup = (uint32_t *)(smem + SM::offChunkGroup);     // first group of each listed chunk
    uint32_t *cPair = (uint32_t *)(smem + SM::offChunkPair);       // exclusive prefix of group counts
    uint32_t *gList = (uint32_t *)(smem + SM::offGroupList);
    uint16_t *sFrag = (uint16_t *)(smem + SM::offFragList) + (threadIdx.x >> 5) * kFragList;   // this warp's fragment list
    uint64_t *sScan = (uint64_t *)(smem + SM::offScan);
    Ctl *ctl = (Ctl *)(smem + SM::offCtl);

    const int tid = threadIdx.x, lane = tid & 31;
    int phase = 0;
    bool loaded = false;
    uint32_t nGroup = 0, nQ = 0, nItems = 0, primsSeen = 0;
    unsigned long long frags = 0;
    unsigned long long tStart = 0;
    long long cycA0 = 0, cycA = 0, cycB = 0, cycMark = 0;      // debug: per-phase clocks of thread 0
    long long cycPre = 0, cycF3 = 0, cycF12 = 0, cycIn = 0;
    uint32_t dbgPairs = 0, dbgGroups = 0;
    uint32_t dbgFlush = 0;
    const bool timing = SWR_TILE_STATS && t.tileStats != nullptr && tid == 0;
    if (timing) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tStart));

    // ---- flush: coverage (A) + shading (B) of the queued primitives -----------------------------
    auto flushQueue = [&]() __attribute__((always_inline)) {
        if (nQ == 0) return;
        primsSeen += nQ;
        if (timing) { cycMark = clock64(); cycIn = cycMark; ++dbgFlush; }
        if (tid == 0) ctl->nextBlock = 0;
        for (int i = tid; i < NB * QW; i += kTileThreads) sBlockmap[i] = 0;
        if (!loaded) {
            moveTile<TLOG, TR::NRT, false>(t, rtSmem, X0, Y0, sub);
            loaded = true;
        }
        if (timing) { const long long c = clock64(); cycPre += c - cycMark; cycMark = c; }

        // A0: drop the (primitive, block) items whose block provably holds no covered pixel (about
        // half of them on tiny-triangle meshes: the reference visits every block of the truncated
        // bounding box).  qValid = bitmap over the primitive's items, row-major in its block range.
        // Worth its barrier + scan only when primitives span several blocks each (big triangles).
        const bool prune = MODE == SWR_DRAW_TRIANGLE && nItems >= 3u * nQ;
        for (uint32_t q = tid; q < nQ; q += kTileThreads) {
            const uint32_t rg = qRange[q];
            const int bx0 = rangeBx0(rg), by0 = rangeBy0(rg), nx = rangeBx1(rg) - bx0 + 1;
            const int n = nx * (rangeBy1(rg) - by0 + 1);
            uint32_t valid = n >= 32 ? 0xffffffffu : (1u << n) - 1u;
            if (prune && n <= kPruneMax) {
                const uint32_t rec = qRec[q];
                const float4 h2 = t.head[(size_t)rec * 3 + 2];
                if (!(f2u(h2.y) & kModeSpan)) {
                    const float4 h0 = t.head[(size_t)rec * 3], h1 = t.head[(size_t)rec * 3 + 1];
                    valid = 0;
                    for (int li = 0; li < n; ++li) {
                        const int ly = divSmall(li, nx), lx = li - ly * nx;
                        if (blockMayBeCovered(h0, h1, h2, X0 + (bx0 + lx) * 8, Y0 + (by0 + ly) * 8)) valid |= 1u << li;
                    }
                }
            }
            qValid[q] = valid;
        }
        if (!prune && tid == 0) qItem[nQ] = nItems;       // qItem already holds the prefix F3 computed
        __syncthreads();
        if (prune) {   // re-index the surviving items: qItem = exclusive prefix of the per-primitive item counts
            auto itemCount = [&](uint32_t q) __attribute__((always_inline)) -> uint32_t {
                if (q >= nQ) return 0u;
                const uint32_t rg = qRange[q];
                const int n = (rangeBx1(rg) - rangeBx0(rg) + 1) * (rangeBy1(rg) - rangeBy0(rg) + 1);
                return n <= kPruneMax ? (uint32_t)__popc(qValid[q]) : (uint32_t)n;
            };
            const uint32_t c0 = itemCount(2 * tid), c1 = itemCount(2 * tid + 1);
            uint32_t total;
            const uint32_t ex = blockScan32(c0 + c1, total, sScan, phase);
            if (2u * tid < nQ) qItem[2 * tid] = ex;
            if (2u * tid + 1 < nQ) qItem[2 * tid + 1] = ex + c0;
            nItems = (uint32_t)total;
            if (tid == 0) qItem[nQ] = nItems;
            __syncthreads();
        }

        if (timing) { const long long c = clock64(); cycA0 += c - cycMark; cycMark = c; }
        // A: one thread per (primitive, block) item.  The loop is warp-uniform (every lane takes part in the warp's
        // reductions of coverRect's loop bounds; lanes without an item contribute nothing).
        for (uint32_t itBase = 0; itBase < nItems; itBase += kTileThreads) {
            const uint32_t it = itBase + tid;
            const bool ivalid = it < nItems;
            uint32_t lo = 0;                                 // largest q < nQ with qItem[q] <= it (qItem[0] = 0)
#pragma unroll
            for (uint32_t step = kQueue / 2; step > 0; step >>= 1) {
                const uint32_t mid = lo + step;
                if (mid < nQ && qItem[mid] <= it) lo = mid;
            }
            const uint32_t q = lo, rec = qRec[q], rg = qRange[q];
            const int bx0 = rangeBx0(rg), by0 = rangeBy0(rg), nx = rangeBx1(rg) - bx0 + 1;
            const int nAll = nx * (rangeBy1(rg) - by0 + 1);
            const int k = ivalid ? (int)(it - qItem[q]) : 0;
            const int li = nAll <= kPruneMax ? nthSetBit32(qValid[q], k) : k;     // k-th surviving item -> block
            const int liy = divSmall(li, nx);
            const int bx = bx0 + li - liy * nx, by = by0 + liy;
            const int gx = X0 + bx * 8, gy = Y0 + by * 8;
            uint64_t m = 0;
            if (MODE == SWR_DRAW_TRIANGLE) {
                const float4 h2 = t.head[(size_t)rec * 3 + 2];
                const bool spanMode = (f2u(h2.y) & kModeSpan) != 0;
                // the primitive's pixel bounds, clipped to this block
                const int rx0 = max((int)(rg & 63u) - bx * 8, 0), ry0 = max((int)((rg >> 6) & 63u) - by * 8, 0);
                const int rx1 = min((int)((rg >> 12) & 63u) - bx * 8, 7), ry1 = min((int)((rg >> 18) & 63u) - by * 8, 7);
                BlockEdges E;
                int cls = 0;
                if (ivalid && !spanMode) cls = classifyBlock(t.head[(size_t)rec * 3], t.head[(size_t)rec * 3 + 1], h2, gx, gy, E);
                else {
#pragma unroll
                    for (int e = 0; e < 3; ++e) { E.ea[e] = 0.0f; E.eb[e] = 0.0f; E.thr[e] = 0.0f; E.e00[e] = 0.0f; }
                }
                if (cls == 1) m = ~0ull;
                const bool partial = cls == 2;
                const int R = __reduce_max_sync(0xffffffffu, partial ? ry1 - ry0 + 1 : 0);
                const int Cw = __reduce_max_sync(0xffffffffu, partial ? rx1 - rx0 + 1 : 0);
                if (R > 0) {
                    const uint64_t pm = coverRect(E, partial ? rx0 : 0, partial ? ry0 : 0, partial ? rx1 : -1, partial ? ry1 : -1, R, Cw);
                    if (partial) m = pm;
                }
                if (ivalid && spanMode) {
                    const float4 *sp = t.span + (size_t)rec * 3;
                    m = coverSpan(sp[0], sp[1], sp[2], gx, gy, t.scMinX, t.scMaxX);
                }
            } else if (MODE == SWR_DRAW_LINE) {
                if (ivalid) m = coverLine(t.head[(size_t)rec * 3], t.head[(size_t)rec * 3 + 1], gx, gy, t);
            } else if (ivalid) {
                const float4 h0 = t.head[(size_t)rec * 3];
                const int lx = f2i(h0.x) - gx, ly = f2i(h0.y) - gy;
                m = ((unsigned)lx < 8u && (unsigned)ly < 8u) ? 1ull << (ly * 8 + lx) : 0ull;
            }
            if (ivalid) sMasks[it] = m;
            if ((SWR_PREFETCH & 2) && ivalid && m) {
                const char *pp = (const char *)(t.params + (size_t)rec * t.paramStride);
                prefetchL2(pp);
                prefetchL2(pp + t.paramStride * 4 - 4);
            }
            if (ivalid && m) atomicOr(&sBlockmap[(by * BPR + bx) * QW + (q >> 5)], 1u << (q & 31));
        }
        __syncthreads();

        if (timing) { const long long c = clock64(); cycA += c - cycMark; cycMark = c; }
        // B: per block, in queue order, 32 fragments per round
        PixelData p;
        p.rtBase = rtSmem;
        p.rtSlotStride = T * T * 4;
        // a block belongs to ONE warp for the whole flush (per-pixel order); which warp takes which
        // block is dynamic so that heavy blocks do not pile up on one warp
        while (true) {
            int b = 0;
            if (lane == 0) b = atomicAdd(&ctl->nextBlock, 1);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b >= (sub ? NB / 4 : NB)) break;
            if (sub) b = ((((sub - 1) >> 1) * (BPR / 2) + b / (BPR / 2)) * BPR) + ((sub - 1) & 1) * (BPR / 2) + b % (BPR / 2);   // block of the quadrant
            const int bx = b % BPR, by = b / BPR;
            const int gx = X0 + bx * 8, gy = Y0 + by * 8;
            // the block's bitmap over the queue, 32 words (1024 queue entries) at a time, in order
#pragma unroll 1
            for (int wc = 0; wc < QW; wc += 32) {
            const uint32_t bm = wc + lane < QW ? sBlockmap[b * QW + wc + lane] : 0u;
            uint32_t wincl = __popc(bm);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, wincl, o);
                if (lane >= o) wincl += n;
            }
            const uint32_t wex = wincl - __popc(bm);
            const uint32_t totalItems = __shfl_sync(0xffffffffu, wincl, 31);
            for (uint32_t ibase = 0; ibase < totalItems; ibase += 32) {
                // lane -> the (ibase + lane)-th item of this block
                const uint32_t j = ibase + lane;
                const bool ivalid = j < totalItems;
                int wl = 0;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {           // largest word wl with wex[wl] <= j
                    const int c = wl + s;
                    const uint32_t e = __shfl_sync(0xffffffffu, wex, c & 31);
                    if (wc + c < QW && e <= j) wl = c;
                }
                const uint32_t wbm = __shfl_sync(0xffffffffu, bm, wl);
                const uint32_t wbase = __shfl_sync(0xffffffffu, wex, wl);
                uint32_t q = 0, rec = 0;
                uint64_t m = 0;
                if (ivalid) {
                    q = (uint32_t)(wc + wl) * 32u + (uint32_t)nthSetBit32(wbm, (int)(j - wbase));
                    rec = qRec[q];
                    const uint32_t rg = qRange[q];
                    const int bx0 = rangeBx0(rg), by0 = rangeBy0(rg), nx = rangeBx1(rg) - bx0 + 1;
                    const int nAll = nx * (rangeBy1(rg) - by0 + 1);
                    const int li = (by - by0) * nx + (bx - bx0);
                    const int k = nAll <= kPruneMax ? __popc(qValid[q] & ((1u << li) - 1u)) : li;
                    m = sMasks[qItem[q] + (uint32_t)k];
                }
                // One fragment of primitive `frec` at bit `fbit` of this block.
                auto shadeOne = [&](uint32_t frec, int fbit) __attribute__((always_inline)) {
                    const int xx = fbit & 7, yy = fbit >> 3;
                    if (gx + xx >= t.rtWidth || gy + yy >= t.rtHeight) return;      // block sticks out of the surface
                    p.rtOffset = (b * 64 + fbit) * 4;
                    if (MODE == SWR_DRAW_TRIANGLE) {
                        shadeTriangleFragment<PS>(t, frec, gx, gy, xx, yy, p);
                        ++frags;
                    } else if (MODE == SWR_DRAW_LINE) {
                        frags += shadeLineFragments<PS>(t, frec, gx + xx, gy + yy, p);
                    } else {
                        shadePointFragment<PS>(t, frec, gx + xx, gy + yy, p);
                        ++frags;
                    }
                };

                // Items are consumed in lane (= queue) order.  A DENSE item (>= SWR_DENSE_MIN covered pixels) is
                // shaded on its own with lane <-> pixel fixed: no search, no conflict detection, and a
                // pixel always meets the same lane.  A run of SPARSE items (tiny triangles) is packed,
                // 32 fragments per round, one lane per fragment.
                const uint32_t validMask = __ballot_sync(0xffffffffu, ivalid);
                const uint32_t denseMask = SWR_DENSE_PATH ? __ballot_sync(0xffffffffu, __popcll(m) >= SWR_DENSE_MIN) : 0u;
                const int nvalid = __popc(validMask);
                int pos = 0;
                while (pos < nvalid) {
                    if ((denseMask >> pos) & 1u) {
                        const uint64_t dm = __shfl_sync(0xffffffffu, m, pos);
                        const uint32_t drec = __shfl_sync(0xffffffffu, rec, pos);
                        if ((dm >> lane) & 1ull) shadeOne(drec, lane);
                        if ((dm >> (lane + 32)) & 1ull) shadeOne(drec, lane + 32);
                        __syncwarp();
                        ++pos;
                        continue;
                    }
                    // sparse run [pos, runEnd): as many of the next sparse items as hold at most kFragList fragments.
                    // Every item lane writes its fragments -- (item lane, pixel) -- into the warp's list at its prefix
                    // position; the rounds then read "fragment f" straight from the list (no search for the owner, no
                    // n-th-set-bit search for the pixel).
                    const uint32_t after = denseMask >> pos;
                    int runEnd = after ? min(nvalid, pos + (__ffs((int)after) - 1)) : nvalid;
                    const uint64_t mr0 = (lane >= pos && lane < runEnd) ? m : 0ull;
                    uint32_t fincl = __popcll(mr0);
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        uint32_t n = __shfl_up_sync(0xffffffffu, fincl, o);
                        if (lane >= o) fincl += n;
                    }
                    runEnd = pos + __popc(__ballot_sync(0xffffffffu, lane >= pos && lane < runEnd && fincl <= (uint32_t)kFragList));
                    const uint64_t mr = lane < runEnd ? mr0 : 0ull;
                    const uint32_t totalF = __shfl_sync(0xffffffffu, fincl, runEnd - 1);
                    {
                        uint32_t w = fincl - __popcll(mr0);
                        uint32_t lo = (uint32_t)mr, hi = (uint32_t)(mr >> 32);
                        while (lo) { sFrag[w++] = (uint16_t)((lane << 6) | (__ffs((int)lo) - 1)); lo &= lo - 1; }
                        while (hi) { sFrag[w++] = (uint16_t)((lane << 6) | (__ffs((int)hi) + 31)); hi &= hi - 1; }
                    }
                    __syncwarp();
                    for (uint32_t fbase = 0; fbase < totalF; fbase += 32) {
                        const uint32_t f = fbase + lane;
                        const bool fvalid = f < totalF;
                        const uint32_t fe = fvalid ? sFrag[f] : 0u;
                        const int ol = (int)(fe >> 6), bit = (int)(fe & 63u);
                        const uint32_t orec = __shfl_sync(0xffffffffu, rec, ol);
                        // fragments of different primitives on the same pixel run in emission order
                        // (skipped when the round holds one primitive only, or -- the usual case on a mesh -- when the
                        // round's pixels are provably all different: the OR of the lanes' one-hot pixel masks has as
                        // many bits as there are fragments)
                        int rank = 0, maxRank = 0;
                        const int first = __shfl_sync(0xffffffffu, ol, 0);
                        bool clash = !__all_sync(0xffffffffu, !fvalid || ol == first);
                        if (clash) {
                            const uint32_t oneLo = (fvalid && bit < 32) ? 1u << bit : 0u, oneHi = (fvalid && bit >= 32) ? 1u << (bit - 32) : 0u;
                            const int distinct = __popc(__reduce_or_sync(0xffffffffu, oneLo)) + __popc(__reduce_or_sync(0xffffffffu, oneHi));
                            clash = distinct != (int)min(32u, totalF - fbase);
                        }
                        if (clash) {
                            const uint32_t peers = __match_any_sync(0xffffffffu, fvalid ? bit : 64 + lane);
                            rank = __popc(peers & ((1u << lane) - 1u));
                            maxRank = __reduce_max_sync(0xffffffffu, rank);
                        }
                        for (int r = 0; r <= maxRank; ++r) {
                            if (fvalid && rank == r) shadeOne(orec, bit);
                            __syncwarp();
                        }
                    }
                    __syncwarp();                                // the list is rewritten by the next run
                    pos = runEnd;
                }
            }
            }   // word chunks
        }
        __syncthreads();
        if (timing) { const long long c = clock64(); cycB += c - cycMark; cycF3 -= c - cycIn; }
        nQ = 0;
        nItems = 0;
    };

    // ---- F3: records of the listed groups -> queue ------------------------------------------------
    // Four consecutive records per thread and scan step (one 32-byte read of their boxes): a quarter of
    // the barriers, four loads in flight per thread.
    auto drainGroups = [&]() __attribute__((always_inline)) {
        __syncthreads();                                     // gList writes of the caller are visible
        const uint32_t npairs = nGroup * 32u;
        const long long f3In = timing ? clock64() : 0;
        if (SWR_TILE_STATS) dbgPairs += npairs;
        for (uint32_t pb = 0; pb < npairs; pb += 4u * kTileThreads) {
            const uint32_t pr0 = pb + 4u * tid;
            uint32_t pend = 0, rec0 = 0;                     // bit k: record rec0 + k still has to be tested / queued
            if (pr0 < npairs) {
                const uint32_t ge = gList[pr0 >> 5];
                const uint32_t n = groupCount(ge), q0 = pr0 & 31u;       // the group's records sit in its first n slots
                rec0 = groupOf(ge) * 32u + q0;
                pend = q0 >= n ? 0u : (n - q0 >= 4u ? 0xfu : (1u << (n - q0)) - 1u);
            }
            while (true) {
                // (re)derive the block ranges of the pending records; after a flush this re-reads the
                // boxes (L1 hits), so only `pend` and `rec0` live across flushQueue()
                uint32_t range[4] = { 0, 0, 0, 0 }, items[4] = { 0, 0, 0, 0 };
                uint32_t cnt = 0, sum = 0;
                if (pend) {
                    const uint4 *bp = reinterpret_cast<const uint4 *>(t.bbox + rec0);
                    const uint4 b01 = bp[0], b23 = bp[1];
                    const uint32_t w[8] = { b01.x, b01.y, b01.z, b01.w, b23.x, b23.y, b23.z, b23.w };
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        Box16 bb;
                        bb.x0 = (int16_t)(w[2 * k] & 0xffffu); bb.y0 = (int16_t)(w[2 * k] >> 16);
                        bb.x1 = (int16_t)(w[2 * k + 1] & 0xffffu); bb.y1 = (int16_t)(w[2 * k + 1] >> 16);
                        if (((pend >> k) & 1u) && boxOverlaps(bb, CX0, CY0, X1, Y1)) {
                            const int lx0 = max((int)bb.x0, CX0) - X0, ly0 = max((int)bb.y0, CY0) - Y0;
                            const int lx1 = min((int)bb.x1, X1) - X0, ly1 = min((int)bb.y1, Y1) - Y0;
                            range[k] = packRange(lx0, ly0, lx1, ly1);
                            items[k] = (uint32_t)(((lx1 >> 3) - (lx0 >> 3) + 1) * ((ly1 >> 3) - (ly0 >> 3) + 1));
                            ++cnt;
                            sum += items[k];
                        } else {
                            pend &= ~(1u << k);
                        }
                    }
                }
                uint32_t total;                              // packed: count (<= 2048 per step) << 20 | items (<= 2^17 per step)
                const uint32_t ex = blockScan32((cnt << 20) | sum, total, sScan, phase);
                const uint32_t totHit = total >> 20, totItems = total & 0xfffffu;
                if (totHit == 0) break;
                uint32_t eh = ex >> 20, ei = ex & 0xfffffu;
                const bool fitsAll = nQ + totHit <= kQueue && nItems + totItems <= kItems;
                uint32_t nAcc = 0, accItems = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((pend >> k) & 1u) {
                        if (nQ + eh < kQueue && nItems + ei + items[k] <= kItems) {
                            qRec[nQ + eh] = rec0 + k;
                            qRange[nQ + eh] = range[k];
                            qItem[nQ + eh] = nItems + ei;
                            pend &= ~(1u << k);
                            ++nAcc;
                            accItems += items[k];
                        }
                        ++eh;
                        ei += items[k];
                    }
                }
                if (fitsAll) {
                    nQ += totHit;
                    nItems += totItems;
                    break;
                }
                // queue full: count what was accepted, flush, retry the rest
                if (tid == 0) { ctl->accQ = 0; ctl->accItems = 0; }
                __syncthreads();
                if (nAcc) { atomicAdd(&ctl->accQ, nAcc); atomicAdd(&ctl->accItems, accItems); }
                __syncthreads();
                nQ += ctl->accQ;
                nItems += ctl->accItems;
                __syncthreads();
                flushQueue();
            }
        }
        __syncthreads();
        nGroup = 0;
        if (timing) { const long long c = clock64(); cycF3 += c - f3In; cycF12 -= c - f3In; }
    };

    // ---- F1 + F2 ----------------------------------------------------------------------------------
    const uint32_t *row = t.tilemap + (size_t)listTile * t.chunkWords;
    const long long f12In = timing ? clock64() : 0;
    // Normally binKernel (bin.cuh) has already listed this tile's groups; the in-kernel F1 + F2 below only
    // run for tiles whose list overflowed its capacity, or when the binning pass is switched off.
    const uint32_t listed = t.groupCount ? t.groupCount[listTile] : 0xffffffffu;
    if (listed <= (uint32_t)kGroupList) {                    // (the overflow mark is 0xffffffff)
        const uint32_t *lst = t.groupList + (size_t)listTile * t.groupCap;
        for (uint32_t i = tid; i < listed; i += kTileThreads) gList[i] = lst[i];
        nGroup = listed;                                     // drained by the common drainGroups() below
        if (SWR_TILE_STATS) dbgGroups += listed;
    } else
    for (int wb = 0; wb < t.chunkWords; wb += kTileThreads) {
        uint32_t pending = (wb + tid < t.chunkWords) ? row[wb + tid] : 0u;
        while (true) {
            // F1: this step's set bits -> chunk list {first group, group count}
            uint32_t myGroups = 0;
            {
                uint32_t bits = pending;
                while (bits) {
                    const int bpos = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const uint32_t c = (uint32_t)(wb + tid) * 32u + (uint32_t)bpos;
                    uint32_t cnt;
                    if (c & 1u) cnt = t.extra[c >> 1].y;
                    else cnt = (uint32_t)min(kBatch, t.numPrims - (int)(c >> 1) * kBatch);
                    myGroups += (cnt + 31u) >> 5;
                }
            }
            uint64_t total;
            const uint64_t ex = blockScan(((uint64_t)__popc(pending) << 32) | myGroups, total, sScan, phase);
            const uint32_t totChunks = (uint32_t)(total >> 32);
            if (totChunks == 0) break;
            uint32_t ci = (uint32_t)(ex >> 32), pairBase = (uint32_t)ex;
            while (pending && ci < kChunkList) {
                const int bpos = __ffs(pending) - 1;
                pending &= pending - 1;
                const uint32_t c = (uint32_t)(wb + tid) * 32u + (uint32_t)bpos;
                uint32_t cnt, g0;
                if (c & 1u) { const uint2 e = t.extra[c >> 1]; g0 = e.x >> 5; cnt = e.y; }
                else { g0 = (c >> 1) * (kBatch / 32); cnt = (uint32_t)min(kBatch, t.numPrims - (int)(c >> 1) * kBatch); }
                cGroup[ci] = g0;
                cPair[ci] = pairBase;
                pairBase += (cnt + 31u) >> 5;
                ++ci;
                if (ci == min(totChunks, (uint32_t)kChunkList)) cPair[ci] = pairBase;   // sentinel by the last writer
            }
            __syncthreads();
            const uint32_t nChunk = min(totChunks, (uint32_t)kChunkList);
            const uint32_t npairs = cPair[nChunk];
            if (SWR_TILE_STATS) dbgGroups += npairs;

            // F2: group boxes of the listed chunks -> group list (four consecutive groups per thread and scan step)
            for (uint32_t pb = 0; pb < npairs; pb += 4u * kTileThreads) {
                if (nGroup + 4u * kTileThreads > kGroupList) drainGroups();
                const uint32_t pr0 = pb + 4u * tid;
                uint32_t grp[4] = { 0, 0, 0, 0 }, hits = 0;
                if (pr0 < npairs) {
                    uint32_t lo = 0, hi = nChunk;            // largest chunk entry with cPair[entry] <= pr0
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (cPair[mid] <= pr0) lo = mid; else hi = mid;
                    }
                    uint32_t first = cPair[lo], next = cPair[lo + 1];
                    Box16 gb[4];
                    uint32_t valid = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t pr = pr0 + k;
                        if (pr < npairs) {
                            while (pr >= next) { ++lo; first = next; next = cPair[lo + 1]; }
                            grp[k] = cGroup[lo] + (pr - first);
                            gb[k] = t.gbox[grp[k]];
                            valid |= 1u << k;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (((valid >> k) & 1u) && boxOverlaps(gb[k], CX0, CY0, X1, Y1)) hits |= 1u << k;
                }
                uint32_t tot2;
                uint32_t e2 = nGroup + blockScan32((uint32_t)__popc(hits), tot2, sScan, phase);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((hits >> k) & 1u) gList[e2++] = packGroup(grp[k], t.gcnt[grp[k]]);
                nGroup += (uint32_t)tot2;
            }
            __syncthreads();
            if (totChunks <= kChunkList) break;
        }
    }
    drainGroups();
    if (timing) cycF12 += clock64() - f12In;
    const long long tailIn = timing ? clock64() : 0;
    flushQueue();
    if (timing) cycF3 += clock64() - tailIn;     // flushQueue subtracted its own time from cycF3

    if (loaded) moveTile<TLOG, TR::NRT, true>(t, rtSmem, X0, Y0, sub);

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) frags += __shfl_xor_sync(0xffffffffu, frags, o);
    if (lane == 0 && frags) atomicAdd(t.fragCounter, frags);
    if (SWR_TILE_STATS && t.tileStats) {
        if (lane == 0 && frags) atomicAdd(&t.tileStats[listTile * 16 + 3], (uint32_t)frags);
        if (tid == 0) {
            unsigned long long tEnd;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tEnd));
            t.tileStats[listTile * 16 + 0] = (uint32_t)tStart;
            t.tileStats[listTile * 16 + 1] = (uint32_t)(tEnd - tStart);
            t.tileStats[listTile * 16 + 2] = primsSeen;
            t.tileStats[listTile * 16 + 4] = (uint32_t)(cycA0 >> 4);
            t.tileStats[listTile * 16 + 5] = (uint32_t)(cycA >> 4);
            t.tileStats[listTile * 16 + 6] = (uint32_t)(cycB >> 4);
            t.tileStats[listTile * 16 + 7] = dbgFlush;
            t.tileStats[listTile * 16 + 8] = (uint32_t)(cycPre >> 4);
            t.tileStats[listTile * 16 + 9] = (uint32_t)(cycF3 >> 4);
            t.tileStats[listTile * 16 + 10] = (uint32_t)(cycF12 >> 4);
            t.tileStats[listTile * 16 + 11] = dbgPairs;
            t.tileStats[listTile * 16 + 12] = dbgGroups;
        }
    }
}

template <class PS, int MODE, int TLOG>
void launchTiles(const void *args, void *stream)
{
    typedef TileSmem<TLOG, PsTraits<PS>::NRT> SM;
    const TileArgs *t = static_cast<const TileArgs *>(args);
    cudaFuncSetAttribute(tileKernel<PS, MODE, TLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::bytes);
    const int grid = 4 * t->splitCap + t->tilesX * t->tilesY;
    tileKernel<PS, MODE, TLOG><<<grid, kTileThreads, SM::bytes, (cudaStream_t)stream>>>(*t);
}

template <class PS>
const swr_pixel_shader *pixelShaderBinding(const char *name = "user")
{
    static const swr_pixel_shader d = {
        { { &launchTiles<PS, 0, 5>, &launchTiles<PS, 0, 6> },
          { &launchTiles<PS, 1, 5>, &launchTiles<PS, 1, 6> },
          { &launchTiles<PS, 2, 5>, &launchTiles<PS, 2, 6> } },
        &uploadUniforms,
        PS::AVarCount, PS::PVarCount, PS::InterpolateZ ? 1 : 0, PS::InterpolateW ? 1 : 0,
        PS::RenderTargets, name, SWR_ARGS_LAYOUT };
    return &d;
}

} // namespace detail
} // namespace swr

#endif // __CUDACC__
