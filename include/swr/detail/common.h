// swr/detail/common.h -- record layouts, kernel argument blocks and exact-fp32 helpers shared by
// the geometry kernel (geometry.cuh), the tile kernel (tile.cuh) and the runtime (runtime.cu).
//
// HBM layout of one pass (up to `maxPrims` input primitives, R = record index):
//   bbox   [R]  4 x int16  inclusive pixel bounds of the primitive's fragment footprint
//                         (min > max  <=>  dead: clipped away, culled, zero area, off-screen)
//   gbox   [R/32] 4 x int16 union of the records of one group (one warp of the geometry kernel)
//   head   [R]  3 x float4 edge equations / line start+step / point position, flags, ordinal
//   params [R]  paramStride floats: interpolation planes, one float4 (a,b,c,0) each, in the order z?, invw?, avar[], pvar[]
//               (lines: (start, step) pairs; points: values)
//   span   [R]  3 x float4 (Span / Adaptive only) the two scan-converted halves
//   tilemap[tile][chunk/32] one bit per (screen tile, chunk): chunk may touch the tile
//   gcnt   [R/32] records of the group (they sit at the front of its 32 slots)
// Records [0, maxPrims) are the "original slots": group G (32 consecutive primitives of the pass, one warp of the
// geometry kernel) owns the slots [32 G, 32 G + 32) and fills the first gcnt[G] of them with its surviving primitives
// IN SUBMISSION ORDER (culled / clipped-away / off-screen primitives and, with several ranks, primitives that touch
// no tile of the rank leave no record; the slots behind the count are neither written nor read).  Clipper fan extras
// of batch b live in a bump-allocated, 32-aligned range extra[b] = {base, count} behind the original slots.  Chunk 2b = original slots of batch b,
// chunk 2b+1 = its extras: ascending chunk id, then ascending record, is exactly the reference's
// emission order (VertexProcessor.cpp:252-261, Rasterizer.h:134-141).
//
// Several ranks (sort-first, one GPU each): every rank owns the screen tiles t with (tx + 3 ty) % world == rank and
// keeps its own copy of the arrays above, holding only the records that touch its tiles.  The geometry stage is
// sharded by batches: the rank that runs batch b writes each surviving record straight into the scratch of the
// owners of the tiles it touches -- its own or, through peer mappings over NVLink, another GPU's (RecordSink).
#pragma once

#include <stdint.h>
#include "../../swr_b200.h"

#if defined(__CUDACC__)
#define SWR_HD __host__ __device__ __forceinline__
#define SWR_D __device__ __forceinline__
#else
#define SWR_HD inline
#define SWR_D inline
#endif

namespace swr {
namespace detail {

constexpr int kBatch = SWR_BATCH_PRIMS;          // 1024 input primitives (VertexProcessor.cpp:110)
constexpr int kGroup = 32;                       // records per group AABB
constexpr int kMaxPoly = SWR_MAX_POLY;
constexpr int kMaxFan = kMaxPoly - 2;            // triangles per clipped input triangle
constexpr int kGeomThreads = 256;
#ifndef SWR_TILE_THREADS
#define SWR_TILE_THREADS 512
#endif
#ifndef SWR_TILE_MINB
#define SWR_TILE_MINB 2
#endif
constexpr int kTileThreads = SWR_TILE_THREADS;

// ---- exact fp32: never contracted into FMA, whatever flags the including TU is built with ----
#if defined(__CUDA_ARCH__)
SWR_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
SWR_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
SWR_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
SWR_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SWR_HD float frcp(float a) { return __frcp_rn(a); }           // == 1.0f / a, correctly rounded: same bits as fdiv(1.0f, a)
SWR_HD int f2i(float a) { return __float2int_rz(a); }
SWR_HD float i2f(int a) { return __int2float_rn(a); }
#else
// Host build (tests/hostcheck only): compile with -ffp-contract=off.
SWR_HD float fmul(float a, float b) { return a * b; }
SWR_HD float fadd(float a, float b) { return a + b; }
SWR_HD float fsub(float a, float b) { return a - b; }
SWR_HD float fdiv(float a, float b) { return a / b; }
SWR_HD float frcp(float a) { return 1.0f / a; }
SWR_HD int f2i(float a) { return (int)a; }
SWR_HD float i2f(int a) { return (float)a; }
#endif

SWR_HD int ffs32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v);
#else
    return __builtin_ffs((int)v);
#endif
}

struct alignas(8) Box16 { int16_t x0, y0, x1, y1; };        // inclusive; dead when x0 > x1
SWR_HD Box16 deadBox() { Box16 b; b.x0 = 32767; b.y0 = 32767; b.x1 = -32768; b.y1 = -32768; return b; }

// head flags
enum : uint32_t {
    kTie0 = 1u, kTie1 = 2u, kTie2 = 4u,          // EdgeEquation::tie of e0, e1, e2
    kModeSpan = 8u,                               // this triangle is scan-converted (Span, or Adaptive's choice)
};

struct RenderTargetDesc { void *ptr; int32_t pitch; int32_t pad; };

constexpr int kMaxRanks = SWR_MAX_RANKS;
constexpr int kShardBatches = 16;    // sharded geometry: runs of this many batches go to the ranks round-robin

// One rank's per-pass scratch as a geometry kernel sees it: its own, or a peer's mapped over NVLink.
struct RecordSink {
    Box16 *bbox;
    Box16 *gbox;
    float4 *head;
    float *params;
    float4 *span;
    uint32_t *tilemap;
    uint8_t *gcnt;                   // per group: records at the front of its 32 slots
    uint2 *extra;                    // per batch {base, count} of the fan extras
    uint32_t *errorFlag;             // [0]: per draw (bit 0 voids the draw on that rank), [1]: sticky copy
};

// Arguments of the geometry kernel (one launch = one pass).
struct GeomArgs {
    // input
    const int32_t *indices;          // first index of the pass
    int32_t numPrims;                // primitives in this pass
    uint32_t firstBatch;             // global batch index of the pass's first batch (ordinals)
    int32_t drawMode;
    const void *attribPtr[SWR_MAX_VERTEX_ATTRIBS];
    int32_t attribStride[SWR_MAX_VERTEX_ATTRIBS];
    // raster-list entry (IRasterizer::draw*List): when rasterVerts != nullptr the vertex shader,
    // clipping, transform and culling are skipped and 144-byte RasterizerVertex records are read
    const void *rasterVerts;
    // fixed-function state
    float px, py, ox, oy;            // VertexProcessor.cpp:50-53
    float depthN, depthF;
    int32_t cullMode, rasterMode;
    int32_t scMinX, scMinY, scMaxX, scMaxY;   // Rasterizer.h:81-87 (max exclusive)
    // what the pixel shader interpolates (TriangleEquations uses the PIXEL shader's counts)
    int32_t nA, nP, useZ, useW;
    // output: sink[r] = scratch of rank r.  Replicated geometry (shard == 0): only sink[rank] is set and records
    // for other ranks are dropped; sharded geometry: all `world` sinks are set.
    RecordSink sink[kMaxRanks];
    int32_t paramStride;
    int32_t chunkWords;              // 32-bit words per tilemap row
    int32_t tileShift;               // log2(tile size in pixels)
    int32_t tilesX, tilesY;
    uint32_t *extraAlloc;            // bump allocator cursor of this rank's share of the extras range (record index)
    uint32_t extrasEnd;              // ... and its end
    int32_t rank, world;             // sort-first tile ownership
    int32_t shard;                   // 1: this launch runs only the batches of `rank` (kShardBatches-cyclic) and pushes to the owners
    int32_t noTightBox;              // debug: keep the reference's 8-aligned block box instead of the certified pixel bounds
    // optional stream-out (VertexProcessor with a foreign IRasterizer): per input primitive 3 RasterizerVertex-sized
    // records + 3 indices in the reference's batch layout; see geometry.cuh streamOutKernel
    void *soVerts;
    int32_t *soIndices;
    uint32_t *soCounts;
    uint32_t soExtraCap;
};

// Arguments of the tile kernel.
struct TileArgs {
    const Box16 *bbox;
    const Box16 *gbox;
    const float4 *head;
    const float *params;
    const float4 *span;
    int32_t paramStride;
    const uint32_t *tilemap;
    int32_t chunkWords;
    int32_t numChunks;               // 2 * batches in the pass
    int32_t numPrims;
    const uint2 *extra;
    const uint8_t *gcnt;             // per group: records at the front of its 32 slots (geometry.cuh)
    int32_t tilesX, tilesY;
    int32_t rank, world;             // sort-first ownership: (tx + 3*ty) % world == rank
    int32_t rtWidth, rtHeight;
    int32_t numRT;
    RenderTargetDesc rt[SWR_MAX_RENDER_TARGETS];
    int32_t scMinX, scMinY, scMaxX, scMaxY;
    unsigned long long *fragCounter;
    const uint32_t *errorFlag;
    uint32_t *groupCount;            // per tile: groups listed by binKernel (bin.cuh), 0xffffffff = list overflowed; nullptr = no binning pass
    uint32_t *groupList;             // per tile groupCap group ids, ascending
    uint32_t groupCap;
    // heavy-tile split: a tile that lists more than splitThreshold groups (or whose list overflowed) is flagged by the
    // binning pass (heavyFlag[tile] = 1, heavyList[1 + atomicAdd(heavyList[0])] = tile, at most splitCap of them) and
    // shaded by four CTAs, one per quadrant, which are the first 4 * splitCap CTAs of the tile launch (tile.cuh).
    // splitCap = 0: no split.  heavyList[0] is cleared with the tile bitmap.
    uint8_t *heavyFlag;
    uint32_t *heavyList;
    uint32_t splitThreshold;
    int32_t splitCap;
    int32_t mirrorSlot, mirrorCount; // finished tiles of render target `mirrorSlot` are also stored to mirror[0..mirrorCount)
    void *mirror[SWR_MAX_TILE_MIRRORS];   // surfaces of the same pitch / size, typically the peers' framebuffers (NVLink stores)
    uint32_t *tileStats;             // optional debug: 16 words per tile {globaltimer ns start, ns duration, primitives, fragments, A0/A/B clocks >> 4 of thread 0, flushes, flush prologue / F3 / F1+F2 clocks >> 4, records tested, groups tested, 0...}
};

// The kernels of a user's shader TU are compiled against these argument structs, the library fills them: a TU
// built with other headers than the library is refused when its shader is set (swr_set_*_shader).
#define SWR_ARGS_LAYOUT ((uint32_t)(sizeof(::swr::detail::GeomArgs) * 65599u + sizeof(::swr::detail::TileArgs) * 31u + 2u))

// number of floats of one params record
SWR_HD int paramFloats(int drawMode, int nA, int nP, int useZ, int useW)
{
    if (drawMode == SWR_DRAW_TRIANGLE) {
        int planes = (useZ ? 1 : 0) + ((useW || nP > 0) ? 1 : 0) + nA + nP;
        return planes * 4;                           // one float4 (a, b, c, 0) per plane: a single 128-bit access each
    }
    int vars = (useZ ? 1 : 0) + (useW ? 1 : 0) + nA + nP;
    int n = drawMode == SWR_DRAW_LINE ? vars * 2 : vars;
    return (n + 3) & ~3;
}

SWR_HD bool tileOwned(int tx, int ty, int rank, int world) { return world <= 1 || ((tx + 3 * ty) % world) == rank; }

// Owners of the tiles a pixel box touches, as a bit mask over the ranks (owner = (tx + 3 ty) mod world: any `world`
// consecutive tiles of a row contain every owner, so only narrow boxes need the loop).
SWR_HD uint32_t boxOwnerMask(const Box16 b, int tileShift, int tilesX, int tilesY, int world)
{
    if (b.x0 > b.x1) return 0u;
    int tx0 = b.x0 >> tileShift, ty0 = b.y0 >> tileShift, tx1 = b.x1 >> tileShift, ty1 = b.y1 >> tileShift;
    if (tx1 > tilesX - 1) tx1 = tilesX - 1;
    if (ty1 > tilesY - 1) ty1 = tilesY - 1;
    if (tx0 > tx1 || ty0 > ty1) return 0u;
    if (world <= 1) return 1u;
    const uint32_t all = (1u << world) - 1u;
    if (tx1 - tx0 + 1 >= world) return all;
    uint32_t m = 0;
    int o0 = (tx0 + 3 * ty0) % world;
    for (int ty = ty0; ty <= ty1 && m != all; ++ty) {
        int o = o0;
        for (int tx = tx0; tx <= tx1; ++tx) {
            m |= 1u << o;
            if (++o == world) o = 0;
        }
        o0 += 3;
        while (o0 >= world) o0 -= world;
    }
    return m;
}

} // namespace detail
} // namespace swr
