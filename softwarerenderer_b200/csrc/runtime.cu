// runtime.cu -- the shader-agnostic runtime behind the C ABI (include/swr_b200.h): contexts,
// device scratch, host-pointer staging, pass planning, kernel orchestration, measurement, the
// raster-list entry (IRasterizer::draw*List), the stream-out entry for a foreign IRasterizer
// (swr_process_elements), and the multi-GPU plumbing: shared scratch arena + geometry shards, the
// flag barrier between the GPUs, tile mirrors and the tile pack / unpack kernels of the collective
// composite.  Shader-templated kernels live in include/swr/detail/{geometry,tile}.cuh and reach this
// file only as launcher thunks (swr_vertex_shader / swr_pixel_shader).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <swr_b200.h>
#include <swr/detail/common.h>
#include <swr/detail/geometry.cuh>
#include <swr/detail/bin.cuh>

using namespace swr::detail;

namespace {

thread_local std::string g_lastError;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return fail(-100, "%s: %s", #expr, cudaGetErrorString(e__));       \
    } while (0)

// hostpack.cpp (g++ -mavx2): the host side of the narrowed index upload
extern "C" {
void *swr_hostpack_create(int workers);
void swr_hostpack_destroy(void *pool);
int swr_hostpack_pack16(void *pool, const int32_t *src, size_t count, uint16_t *dst16, int32_t *bases);
}
constexpr size_t kIdxBlock = 4096;         // indices per base of the narrowed form (hostpack.cpp: kBlock)

struct HostBuf {                           // page-locked host memory, grow-only
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need)
    {
        if (need <= bytes) return 0;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        bytes = 0;
        const cudaError_t e = cudaMallocHost(&ptr, need);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(-102, "cudaMallocHost(%zu bytes): %s", need, cudaGetErrorString(e));
        }
        bytes = need;
        return 0;
    }
    void release()
    {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need)
    {
        if (need <= bytes) return 0;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        size_t want = need + need / 8;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&ptr, need);
            want = need;
        }
        if (e != cudaSuccess) return fail(-101, "cudaMalloc(%zu bytes): %s", need, cudaGetErrorString(e));
        bytes = want;
        return 0;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

struct Attrib {
    const void *ptr = nullptr;
    int stride = 0;
    size_t bytes = 0;
};

bool isDevicePointer(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

} // namespace

// One set of per-pass scratch: a single allocation ("blob") cut into the arrays of common.h by layoutOf(), plus a
// small counter block.  Two sets alternate when the geometry kernel of pass k+1 (auxiliary stream) overlaps the tile
// kernel of pass k (swr_set_pipeline).  With a shared scratch arena (swr_shared_scratch_create: sharded geometry,
// the peers write into this memory) there is one set and both live inside the arena, at offsets that are the same
// on every rank.
struct SetLayout {
    size_t bbox, gbox, gcnt, head, params, span, tilemap, extra, groupList, groupCount, heavyList, heavyFlag, bytes;
};

struct Counters {               // one small device block
    uint32_t extraAlloc;
    uint32_t errorFlag;         // per draw: bit 0 voids the draw (tile kernel exits)
    uint32_t stickyFlag;        // same bits, accumulated until swr_finish reports them
    uint32_t pad;
    unsigned long long fragments;
};

constexpr size_t kArenaHeader = 4096;        // Counters at 0, barrier flags (kMaxRanks words) at kArenaFlags
constexpr size_t kArenaFlags = 512;

struct ScratchSet {
    DevBuf blob, counters;
    char *base = nullptr;               // blob.ptr, or the arena's record area
    Counters *ctr = nullptr;
    cudaEvent_t geomDone = nullptr, tileDone = nullptr;
    bool tilePending = false, countersInit = false;
    size_t bytes() const { return blob.bytes + counters.bytes; }
    void release() { blob.release(); counters.release(); base = nullptr; ctr = nullptr; countersInit = false; }
};

static size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

static SetLayout layoutOf(size_t recCap, int paramStride, bool needSpan, size_t tiles, int chunkWords, size_t passBatches, uint32_t groupCap, bool binPass, int splitCap)
{
    SetLayout L;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = alignUp(o + bytes, 256); return at; };
    L.bbox = take(recCap * sizeof(Box16) + 64);
    L.gbox = take((recCap / kGroup + 1) * sizeof(Box16));
    L.gcnt = take(recCap / kGroup + 1);
    L.head = take(recCap * 48);
    L.params = take(recCap * (size_t)paramStride * 4 + 16);
    L.span = take(needSpan ? recCap * 48 : 0);
    L.tilemap = take(tiles * (size_t)chunkWords * 4);
    L.heavyList = take(4 + (size_t)splitCap * 4);        // right behind the bitmap: cleared with it (clearSet)
    L.extra = take(passBatches * sizeof(uint2));
    L.groupList = take(binPass ? tiles * (size_t)groupCap * 4 : 0);
    L.groupCount = take(binPass ? tiles * 4 : 0);
    L.heavyFlag = take(splitCap ? tiles : 0);
    L.bytes = o;
    return L;
}

static RecordSink sinkAt(char *base, const SetLayout &L, Counters *ctr)
{
    RecordSink k;
    k.bbox = reinterpret_cast<Box16 *>(base + L.bbox);
    k.gbox = reinterpret_cast<Box16 *>(base + L.gbox);
    k.head = reinterpret_cast<float4 *>(base + L.head);
    k.params = reinterpret_cast<float *>(base + L.params);
    k.span = reinterpret_cast<float4 *>(base + L.span);
    k.tilemap = reinterpret_cast<uint32_t *>(base + L.tilemap);
    k.gcnt = reinterpret_cast<uint8_t *>(base + L.gcnt);
    k.extra = reinterpret_cast<uint2 *>(base + L.extra);
    k.errorFlag = &ctr->errorFlag;
    return k;
}

struct swr_context {
    int device = 0;
    cudaStream_t stream = nullptr;      // main stream: tile kernels, everything the caller orders against
    cudaStream_t ownStream = nullptr;   // created by swr_create
    cudaStream_t aux = nullptr;         // geometry kernels and their input staging
    cudaStream_t copy = nullptr;        // host -> device staging of big draws, pass by pass (created on first use)
    std::vector<cudaEvent_t> idxReady;  // one per pass of the current draw
    cudaEvent_t evGeom0 = nullptr, evGeom1 = nullptr, evTile0 = nullptr, evTile1 = nullptr, evTimer0 = nullptr, evTimer1 = nullptr,
                evDrawStart = nullptr;
    ScratchSet sets[2];
    uint64_t passSeq = 0;
    int lastSet = 0;
    bool pipeline = false;              // geometry on the auxiliary stream (two scratch sets); see swr_set_pipeline
    bool overlapDraws = false;          // ... and across draws (caller promises not to touch the inputs in between)
    int passesHint = 0;
    bool haveDrawEvents = false;

    // VertexProcessor state (defaults VertexProcessor.cpp:29-35)
    int vpX = 0, vpY = 0, vpW = 0, vpH = 0;
    float px = 0, py = 0, ox = 0, oy = 0;
    float depthN = 0.0f, depthF = 1.0f;
    int cullMode = SWR_CULL_CW;
    Attrib attribs[SWR_MAX_VERTEX_ATTRIBS];
    const swr_vertex_shader *vs = nullptr;
    // Rasterizer state (defaults Rasterizer.h:67-72)
    int rasterMode = SWR_RASTER_SPAN;
    int scMinX = 0, scMinY = 0, scMaxX = 0, scMaxY = 0;
    const swr_pixel_shader *ps = nullptr;
    // additive state
    RenderTargetDesc rt[SWR_MAX_RENDER_TARGETS] = {};
    int rtW = 0, rtH = 0;
    unsigned char uniforms[SWR_MAX_UNIFORM_BYTES] = {};
    size_t uniformBytes = 0;
    int tileSizeReq = 0, rank = 0, world = 1;
    int tileSplitReq = -1;              // heavy-tile split: -1 automatic, 0 off, > 0 listed groups above which a tile is split
    size_t scratchLimit = (size_t)48 << 30;   // per-pass scratch (worst case 10 records per triangle): C5's 50M triangles then take 2 passes (5 at 16 GB: +4 %)

    // scratch
    DevBuf stageIdx, l2flush, ownedIdx, tileStats, arena, soVerts, soIndices, soCounts, maxIndexBuf;
    bool debugTileStats = false;
    // sharded geometry (swr_shared_scratch_create / swr_set_geometry_shards)
    bool shardGeometry = false;
    char *peerArena[kMaxRanks] = {};
    uint32_t barrierEpoch = 0;
    uint64_t clearedKey = 0;            // layout the set's tile bitmap / counters are currently cleared for (0: dirty)
    bool fencedSinceClear = false;      // ... and a peer barrier ran after that clear
    int pinnedTileShift = 0;            // world > 1: the tile size of the first draw is kept (tile ownership depends on it)
    int mirrorSlot = 0, mirrorCount = 0;
    void *mirror[SWR_MAX_TILE_MIRRORS] = {};
    std::vector<std::pair<void *, void *>> ipcOpened;   // {pointer handed out, mapped base}
    int lastTiles = 0;
    int ownedKey[5] = { 0, 0, 0, 0, 0 };   // {tile size, rank, world, width, height} of ownedIdx
    DevBuf stageAttrib[SWR_MAX_VERTEX_ATTRIBS];
    // second staging set: streamed draws alternate, so the uploads of draw k+1 can run under the kernels of draw k
    DevBuf stageIdxB, stageAttribB[SWR_MAX_VERTEX_ATTRIBS];
    cudaEvent_t stageFree[2] = { nullptr, nullptr };   // recorded after the last kernel of the draw that last read the set
    bool stageUsed[2] = { false, false };
    int stageSet = 0;
    // narrowed index upload of streamed draws (hostpack.cpp): per staging set the packed form in page-locked host
    // memory and its device copy; layout of both: 16-bit offsets of the whole draw, then one int32 base per block
    void *packPool = nullptr;
    HostBuf pack16[2];
    DevBuf dev16[2];
    int narrowMode = -1;                // -1 automatic (by the measured packing rate), 0 off, 1 always try
    double packGBps = 0.0;              // best packing rate of a draw so far (source bytes / host time)
    int packDraws = 0;                  // draws that measured it
    uint32_t *hostFlags = nullptr;   // pinned: error flag read-back
    int lastDrawMode = 0;

    swr_stats stats = {};
};

namespace {

int setDevice(swr_context *c)
{
    CUDA_TRY(cudaSetDevice(c->device));
    return 0;
}

size_t scratchBytes(const swr_context *c)
{
    size_t n = c->sets[0].bytes() + c->sets[1].bytes() + c->arena.bytes + c->stageIdx.bytes + c->l2flush.bytes + c->ownedIdx.bytes +
               c->soVerts.bytes + c->soIndices.bytes + c->soCounts.bytes;
    for (int i = 0; i < SWR_MAX_VERTEX_ATTRIBS; ++i) n += c->stageAttrib[i].bytes + c->stageAttribB[i].bytes;
    n += c->stageIdxB.bytes + c->dev16[0].bytes + c->dev16[1].bytes;
    return n;
}

// ---- raster-list entry: IRasterizer::draw{Point,Line,Triangle}List on screen-space vertices -------
// Rasterizer.h:116-141: primitives whose FIRST index is -1 are skipped; no clipping, transform or
// culling happens here (that was the VertexProcessor's job).
__global__ void __launch_bounds__(kGeomThreads) rasterListKernel(const GeomArgs g)
{
    typedef CVert<SWR_MAX_AVARS, SWR_MAX_PVARS> V;
    __shared__ uint32_t sMark[kMarkCache];
    for (int i = threadIdx.x; i < kMarkCache; i += kGeomThreads) sMark[i] = 0xffffffffu;
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31;
    const int batch = blockIdx.x;
    const int primBase = batch * kBatch;
    const int cnt = min(kBatch, g.numPrims - primBase);
    const uint32_t ord0 = (g.firstBatch + (uint32_t)batch) * SWR_ORDINAL_STRIDE;
    const int per = g.drawMode + 1;
    const V *verts = static_cast<const V *>(g.rasterVerts);     // RasterizerVertex has exactly this layout
    const RecordSink &sk = g.sink[g.rank];                      // every rank walks the whole list and keeps what touches its tiles
    for (int r = 0; r < kBatch / kGeomThreads; ++r) {
        const int slot = r * kGeomThreads + tid;
        const uint32_t rec = (uint32_t)(primBase + slot);
        Box16 box = deadBox();
        if (slot < cnt) {
            const int32_t *ip = g.indices + (size_t)rec * per;
            const uint32_t ordinal = ord0 + (uint32_t)slot;
            if (ip[0] != -1) {
                if (g.drawMode == SWR_DRAW_TRIANGLE)
                    box = emitScreenTriangle<SWR_MAX_AVARS, SWR_MAX_PVARS>(g, sk, rec, ordinal, false, verts[ip[0]], verts[ip[1]], verts[ip[2]]);
                else if (g.drawMode == SWR_DRAW_LINE)
                    box = emitScreenLine<SWR_MAX_AVARS, SWR_MAX_PVARS>(g, sk, rec, ordinal, verts[ip[0]], verts[ip[1]]);
                else
                    box = emitScreenPoint<SWR_MAX_AVARS, SWR_MAX_PVARS>(g, sk, rec, ordinal, verts[ip[0]]);
                if (!((boxOwnerMask(box, g.tileShift, g.tilesX, g.tilesY, g.world) >> g.rank) & 1u)) box = deadBox();
            }
        }
        sk.bbox[rec] = box;
        // union of the warp's boxes -> group box, tiles of the union -> bitmap
        const int x0 = __reduce_min_sync(0xffffffffu, (int)box.x0), y0 = __reduce_min_sync(0xffffffffu, (int)box.y0);
        const int x1 = __reduce_max_sync(0xffffffffu, (int)box.x1), y1 = __reduce_max_sync(0xffffffffu, (int)box.y1);
        if (lane == 0) {
            Box16 u; u.x0 = (int16_t)x0; u.y0 = (int16_t)y0; u.x1 = (int16_t)x1; u.y1 = (int16_t)y1;
            sk.gbox[rec >> 5] = u;
            sk.gcnt[rec >> 5] = 32;                             // the list is not compacted: skipped entries are dead boxes
        }
        markTiles(g, sk, g.rank, sMark, 2u * (uint32_t)batch, x0, y0, x1, y1);
    }
    if (tid == 0) sk.extra[batch] = make_uint2(0u, 0u);
}

void launchRasterList(const void *args, void *stream)
{
    const GeomArgs *g = static_cast<const GeomArgs *>(args);
    const int batches = (g->numPrims + kBatch - 1) / kBatch;
    if (batches > 0) rasterListKernel<<<batches, kGeomThreads, 0, (cudaStream_t)stream>>>(*g);
}

// ---- multi-GPU composite helpers ---------------------------------------------------------------------
// Tile-major exchange buffer: the owned tiles of one rank in increasing tile id, T*T words each,
// row-major inside a tile.  One CTA per owned tile, 128-bit moves.
template <bool PACK>
__global__ void tileExchangeKernel(char *surface, int pitch, int width, int height, int tileShift, int tilesX, int tilesY,
                                   int rank, int world, const int *ownedIndex, uint32_t *buf)
{
    const int tile = blockIdx.x;
    const int tx = tile % tilesX, ty = tile / tilesX;
    if (!tileOwned(tx, ty, rank, world)) return;
    const int T = 1 << tileShift;
    uint32_t *tb = buf + (size_t)ownedIndex[tile] * T * T;
    const bool vec = ((((uintptr_t)surface) | (uintptr_t)pitch) & 15) == 0 && (width & 3) == 0;
    if (vec) {
        for (int i = threadIdx.x; i < T * T / 4; i += blockDim.x) {
            const int ly = i / (T / 4), lx = (i % (T / 4)) * 4;
            const int x = (tx << tileShift) + lx, y = (ty << tileShift) + ly;
            uint4 *bp = (uint4 *)(tb + ly * T + lx);
            if (x < width && y < height) {
                uint4 *gp = (uint4 *)(surface + (size_t)y * pitch + (size_t)x * 4);
                if (PACK) *bp = *gp; else *gp = *bp;
            } else if (PACK) {
                *bp = make_uint4(0, 0, 0, 0);
            }
        }
    } else {
        for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
            const int ly = i / T, lx = i % T;
            const int x = (tx << tileShift) + lx, y = (ty << tileShift) + ly;
            if (x < width && y < height) {
                uint32_t *gp = (uint32_t *)(surface + (size_t)y * pitch + (size_t)x * 4);
                if (PACK) tb[ly * T + lx] = *gp; else *gp = tb[ly * T + lx];
            } else if (PACK) {
                tb[ly * T + lx] = 0;
            }
        }
    }
}

__global__ void fill32Kernel(uint32_t *dst, uint32_t value, size_t count)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = value;
}

int chooseTileShift(const swr_context *c, int renderTargets, size_t primitives)
{
    auto fits64 = [&]() { return (size_t)renderTargets * 64 * 64 * 4 + 96 * 1024 <= (size_t)227 * 1024; };
    int req = c->tileSizeReq;
    if (req == 0) {
        const char *env = getenv("SWR_TILE_SIZE");
        if (env) req = atoi(env);
    }
    if (req == 64 && fits64()) return 6;
    if (req == 32) return 5;
    // 64-pixel tiles amortise the binning scans better, 32-pixel tiles balance better and give the shading
    // phase four times as many CTAs: take 64 only for meshes of small triangles (fewer than 10 surface pixels per
    // primitive of one pass) and when this rank still gets >= 768 tiles (2.5 waves of the 296 CTA slots; C3 on two
    // ranks, 1020 tiles each: 0.664 / 0.596 ms per frame, on four ranks, 510 each: 0.414 / 0.549).  Measured on B200 (ms, 32 / 64):
    // 10M tiny triangles at 4K 1.28 / 1.13, 5 x 10M triangles at 8K 12.9 / 10.9, 1M-triangle grid at 1080p
    // (510 tiles of 64) 0.31 / 0.35, Benchmark.cpp's 40960 large triangles at 4K 13.0 / 22.8.
    const long tiles64 = (long)((c->rtW + 63) / 64) * ((c->rtH + 63) / 64);
    const bool tiny = (double)primitives * 10.0 >= (double)c->rtW * (double)c->rtH;
    return (tiny && tiles64 / (c->world > 0 ? c->world : 1) >= 768 && fits64()) ? 6 : 5;
}

// ---- cross-GPU barrier of a sort-first partition ---------------------------------------------------
// Every rank owns kMaxRanks flag words in its arena header; word s is written by rank s only.  One warp: signal the
// peers (release store of the epoch over NVLink), then wait until every peer's word has reached the epoch.
// Stream order makes this a barrier between whole kernels: everything this rank enqueued before it -- e.g. the
// geometry kernel's record pushes into the peers' scratch -- is visible at system scope before the signal, and
// what follows (bin / tile kernels reading the own scratch) starts only after every peer has signalled.
struct PeerFlags { uint32_t *flags[kMaxRanks]; };

__global__ void peerBarrierKernel(PeerFlags pf, int rank, int world, uint32_t epoch, uint32_t *stickyFlag)
{
    const int lane = threadIdx.x;
    __threadfence_system();
    if (lane < world && lane != rank) {
        uint32_t *dst = pf.flags[lane] + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
        const uint32_t *src = pf.flags[rank] + lane;
        const long long t0 = clock64();
        while (true) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (clock64() - t0 > 6000000000ll) { atomicOr(stickyFlag, 8u); break; }     // ~3 s: a peer is gone; reported by swr_finish
            __nanosleep(64);
        }
    }
    __threadfence_system();
}

int enqueuePeerBarrier(swr_context *c)
{
    if (c->world <= 1) return 0;
    if (!c->shardGeometry || !c->arena.ptr) return fail(-40, "no shared scratch / geometry shards set (swr_shared_scratch_create, swr_set_geometry_shards)");
    PeerFlags pf;
    for (int r = 0; r < kMaxRanks; ++r) pf.flags[r] = r < c->world ? reinterpret_cast<uint32_t *>(c->peerArena[r] + kArenaFlags) : nullptr;
    Counters *ctr = reinterpret_cast<Counters *>(c->arena.ptr);
    peerBarrierKernel<<<1, 32, 0, c->stream>>>(pf, c->rank, c->world, ++c->barrierEpoch, &ctr->stickyFlag);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    if (c->clearedKey) c->fencedSinceClear = true;
    return 0;
}

// Narrowed index upload: out[i] = bases[i / 4096] + offsets[i] (see hostpack.cpp).  Eight indices per thread: one
// 16-byte read, two 16-byte writes; `offsets` and `out` are 32-byte aligned (slices start at multiples of 1024 primitives).
__global__ void widenIndicesKernel(const uint16_t *offsets, const int32_t *bases, int32_t *out, size_t count)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n8 = count / 8;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n8; t += stride) {
        const uint4 v = reinterpret_cast<const uint4 *>(offsets)[t];
        const int32_t b = bases[(t * 8) / kIdxBlock];
        int4 lo, hi;
        lo.x = b + (int)(v.x & 0xffffu); lo.y = b + (int)(v.x >> 16); lo.z = b + (int)(v.y & 0xffffu); lo.w = b + (int)(v.y >> 16);
        hi.x = b + (int)(v.z & 0xffffu); hi.y = b + (int)(v.z >> 16); hi.z = b + (int)(v.w & 0xffffu); hi.w = b + (int)(v.w >> 16);
        reinterpret_cast<int4 *>(out)[2 * t] = lo;
        reinterpret_cast<int4 *>(out)[2 * t + 1] = hi;
    }
    if (blockIdx.x == 0 && threadIdx.x < (count & 7)) {
        const size_t i = n8 * 8 + threadIdx.x;
        out[i] = bases[i / kIdxBlock] + (int)offsets[i];
    }
}

__global__ void maxIndexKernel(const int32_t *idx, size_t n, int *out)
{
    int m = -1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = max(m, idx[i]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// Largest index of a draw (host or device array): the extent of a host attribute array the reference API gives no size for.
int maxIndexOf(swr_context *c, const int32_t *indices, size_t count, bool onDevice, int *out)
{
    if (!onDevice) {
        int m = -1;
        for (size_t i = 0; i < count; ++i) m = indices[i] > m ? indices[i] : m;
        *out = m;
        return 0;
    }
    if (int rc = c->maxIndexBuf.reserve(sizeof(int))) return rc;
    const int init = -1;
    CUDA_TRY(cudaMemcpyAsync(c->maxIndexBuf.ptr, &init, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    maxIndexKernel<<<148 * 4, 256, 0, c->stream>>>(indices, count, static_cast<int *>(c->maxIndexBuf.ptr));
    c->stats.kernel_launches++;
    CUDA_TRY(cudaMemcpyAsync(out, c->maxIndexBuf.ptr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// One draw = one or more passes of {geometry kernel, [peer barrier,] bin kernel, tile kernel}.
int drawCommon(swr_context *c, int drawMode, size_t count, const int32_t *indices, const void *rasterVerts, size_t rasterVertCount)
{
    if (int rc = setDevice(c)) return rc;
    if (drawMode < 0 || drawMode > 2) return fail(-2, "bad draw mode %d", drawMode);
    const swr_pixel_shader *ps = c->ps;
    const swr_vertex_shader *vs = c->vs;
    if (!ps) return fail(-3, "no pixel shader set");
    if (!rasterVerts && !vs) return fail(-3, "no vertex shader set");
    const int per = drawMode + 1;
    const size_t nprims = count / per;
    c->stats.draws++;
    c->stats.primitives_in += nprims;
    if (nprims == 0) return 0;
    if (nprims > (size_t)0x7fffffff / 4) return fail(-4, "too many primitives in one draw (%zu)", nprims);
    if (c->rtW <= 0 || c->rtH <= 0) return fail(-5, "no render target registered (swr_set_render_target)");
    if (c->rtW > 32767 || c->rtH > 32767) return fail(-5, "render target too large");
    for (int s = 0; s < ps->render_targets; ++s)
        if (!c->rt[s].ptr) return fail(-5, "pixel shader '%s' stages render target slots [0,%d) but slot %d is not registered", ps->name, ps->render_targets, s);
    if (c->scMinX < 0 || c->scMinY < 0) return fail(-6, "scissor origin must be >= 0");
    if (!rasterVerts) {
        if (ps->avar_count > vs->avar_count || ps->pvar_count > vs->pvar_count)
            return fail(-7, "pixel shader '%s' interpolates more variables than vertex shader '%s' outputs", ps->name, vs->name);
        for (int i = 0; i < vs->attrib_count; ++i)
            if (!c->attribs[i].ptr) return fail(-8, "vertex attribute %d not set", i);
    }
    const bool shard = c->shardGeometry && c->world > 1;
    const bool pipeline = c->pipeline && !shard;             // (the peers write into ONE scratch set)

    // ---- streams: geometry-side work goes to the auxiliary stream.  Unless the caller opted into
    // overlapping whole draws, it first waits for everything already enqueued on the main stream
    // (so buffers the caller filled on that stream are visible to the vertex stage).
    cudaStream_t gs = pipeline ? c->aux : c->stream;
    if (pipeline && !c->overlapDraws) {
        CUDA_TRY(cudaEventRecord(c->evDrawStart, c->stream));
        CUDA_TRY(cudaStreamWaitEvent(gs, c->evDrawStart, 0));
    }

    // ---- inputs: device memory in place, host memory staged
    GeomArgs g;
    memset(&g, 0, sizeof(g));
    const int32_t *devIndices = indices;
    const bool idxOnDevice = isDevicePointer(indices);
    // Host indices of a big draw are streamed: the staging copies run on their own stream, attributes
    // first, then the indices pass by pass, and pass k's kernels only wait for their own slice -- the
    // PCIe transfer of pass k+1 runs under the kernels of pass k.
    const size_t idxBytes = nprims * per * sizeof(int32_t);
    const bool streamIdx = !pipeline && !shard && !rasterVerts && !idxOnDevice && idxBytes >= ((size_t)32 << 20) &&
                           !getenv("SWR_NO_INDEX_STREAMING");
    cudaStream_t hs = gs;                                    // stream of the staging copies
    if (streamIdx) {
        if (!c->copy) CUDA_TRY(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
        hs = c->copy;
        c->stageSet ^= 1;                                                 // the other set than the previous streamed draw
    }
    const int S = streamIdx ? c->stageSet : 0;
    DevBuf &stageIdx = S ? c->stageIdxB : c->stageIdx;
    DevBuf *stageAttrib = S ? c->stageAttribB : c->stageAttrib;
    bool staged = false;
    if (streamIdx && c->stageUsed[S]) CUDA_TRY(cudaStreamWaitEvent(hs, c->stageFree[S], 0));   // kernels of the draw that last read this set
    if (!idxOnDevice) {
        if (int rc = stageIdx.reserve(count * sizeof(int32_t))) return rc;
        if (!streamIdx) {
            CUDA_TRY(cudaMemcpyAsync(stageIdx.ptr, indices, idxBytes, cudaMemcpyHostToDevice, gs));
            c->stats.h2d_bytes += idxBytes;
        }
        devIndices = static_cast<const int32_t *>(stageIdx.ptr);
        staged = true;
    }
    if (rasterVerts) {
        const void *dv = rasterVerts;
        if (!isDevicePointer(rasterVerts)) {
            const size_t bytes = rasterVertCount * 144;
            if (int rc = stageAttrib[0].reserve(bytes)) return rc;
            CUDA_TRY(cudaMemcpyAsync(stageAttrib[0].ptr, rasterVerts, bytes, cudaMemcpyHostToDevice, gs));
            c->stats.h2d_bytes += bytes;
            dv = stageAttrib[0].ptr;
            staged = true;
        }
        g.rasterVerts = dv;
    } else {
        int maxIndex = -2;                                   // computed at most once per draw
        for (int i = 0; i < vs->attrib_count; ++i) {
            const Attrib &a = c->attribs[i];
            const void *dp = a.ptr;
            if (!isDevicePointer(a.ptr)) {
                size_t bytes = a.bytes;
                if (bytes == 0) {
                    // The reference signature setVertexAttribPointer(index, stride, buffer) (VertexProcessor.h:88) carries no
                    // extent: the vertices a draw can touch end at stride * (largest index + 1).
                    if (a.stride <= 0) return fail(-9, "vertex attribute %d is host memory with stride %d: its extent is required (bytes > 0)", i, a.stride);
                    if (maxIndex == -2)
                        if (int rc = maxIndexOf(c, indices, nprims * per, idxOnDevice, &maxIndex)) return rc;
                    if (maxIndex < 0) return fail(-9, "vertex attribute %d is host memory and the draw has no valid index", i);
                    bytes = (size_t)a.stride * ((size_t)maxIndex + 1);
                }
                if (int rc = stageAttrib[i].reserve(bytes)) return rc;
                CUDA_TRY(cudaMemcpyAsync(stageAttrib[i].ptr, a.ptr, bytes, cudaMemcpyHostToDevice, hs));
                c->stats.h2d_bytes += bytes;
                dp = stageAttrib[i].ptr;
                staged = true;
            }
            g.attribPtr[i] = dp;
            g.attribStride[i] = a.stride;
        }
    }

    // ---- uniforms into the shader TUs' constant blocks
    if (c->uniformBytes) {
        // A vertex and a pixel shader of the same translation unit share one uniform block: it is then
        // written once, on the main stream, and the geometry stream waits for it.
        const bool shared = !rasterVerts && ps->set_uniforms == vs->set_uniforms;
        if (!rasterVerts && !shared && vs->set_uniforms && vs->set_uniforms(c->uniforms, c->uniformBytes, gs) != 0)
            return fail(-10, "uniform upload failed (vertex shader '%s')", vs->name);
        if (ps->set_uniforms && ps->set_uniforms(c->uniforms, c->uniformBytes, c->stream) != 0)
            return fail(-10, "uniform upload failed (pixel shader '%s')", ps->name);
        if (shared && gs != c->stream) {
            CUDA_TRY(cudaStreamSynchronize(gs));             // no geometry kernel may still read the old block
            CUDA_TRY(cudaEventRecord(c->evDrawStart, c->stream));
            CUDA_TRY(cudaStreamWaitEvent(gs, c->evDrawStart, 0));
        }
    }

    // ---- pass plan
    const int nA = ps->avar_count, nP = ps->pvar_count, useZ = ps->interpolate_z, useW = ps->interpolate_w;
    const int paramStride = paramFloats(drawMode, nA, nP, useZ, useW);
    const bool needSpan = drawMode == SWR_DRAW_TRIANGLE && c->rasterMode != SWR_RASTER_BLOCK;
    const bool tri = drawMode == SWR_DRAW_TRIANGLE && !rasterVerts;
    const size_t recBytes = sizeof(Box16) + 48 + (size_t)paramStride * 4 + (needSpan ? 48 : 0);
    const size_t perPrimWorst = recBytes * (tri ? (size_t)(1 + kMaxFan - 1) : 1) + 16;
    const size_t budget = shard ? (c->arena.bytes > kArenaHeader ? c->arena.bytes - kArenaHeader : 0) * 9 / 10 : c->scratchLimit / (pipeline ? 2 : 1);
    size_t passPrims = budget / perPrimWorst;                           // pipelining alternates two scratch sets
    passPrims = std::max<size_t>(kBatch, passPrims / kBatch * kBatch);
    passPrims = std::min(passPrims, (nprims + kBatch - 1) / kBatch * kBatch);
    if (pipeline) {
        // big draws are cut into a few passes so that geometry(k+1) runs under tiles(k)
        int want = c->passesHint;
        if (want <= 0) {
            const char *env = getenv("SWR_PASSES");
            want = env ? atoi(env) : 1;   // measured: one pass per draw is fastest (every tile pass has its own tail)
        }
        want = std::max(1, std::min(want, 64));
        const size_t per_pass = ((nprims + want - 1) / want + kBatch - 1) / kBatch * kBatch;
        passPrims = std::min(passPrims, std::max<size_t>(kBatch, per_pass));
    }
    if (streamIdx) {
        // 2..8 passes of >= 32 MB of indices each: enough slices to hide the kernels, few enough tile passes
        const char *env = getenv("SWR_STREAM_SLICE_MB");
        const size_t slice = (size_t)std::max(1, env ? atoi(env) : 32) << 20;
        const size_t want = std::max<size_t>(2, std::min<size_t>(8, idxBytes / slice));
        const size_t per_pass = ((nprims + want - 1) / want + kBatch - 1) / kBatch * kBatch;
        passPrims = std::min(passPrims, std::max<size_t>(kBatch, per_pass));
    }
    // Tile ownership of a partition depends on the tile size, so with several ranks the size chosen for the first
    // draw stays (every draw of a frame, and the composite, must agree on who owns which pixels).
    int tileShift = c->world > 1 ? c->pinnedTileShift : 0;
    if (tileShift == 0) {
        tileShift = chooseTileShift(c, ps->render_targets, std::min(nprims, passPrims));   // density of ONE pass
        if (c->world > 1) c->pinnedTileShift = tileShift;
    } else if ((size_t)ps->render_targets * (1u << (2 * tileShift)) * 4 + 96 * 1024 > (size_t)227 * 1024) {
        return fail(-2, "pixel shader '%s' stages %d render targets: the partition's %d-pixel tiles do not fit in shared memory (swr_set_tile_size(32) before the first draw)",
                    ps->name, ps->render_targets, 1 << tileShift);
    }
    const int T = 1 << tileShift;
    const int tilesX = (c->rtW + T - 1) / T, tilesY = (c->rtH + T - 1) / T;
    const size_t firstsCap = passPrims;                                  // multiple of kBatch
    const size_t passBatches = passPrims / kBatch;
    // fan extras: worst case 9 per primitive.  Sharded: every rank appends the extras of its own batches to its own
    // share of the range (the cursor is local, no NVLink round trip), so the range is `world` worst-case shares.
    size_t extraShare = 0, extrasCap = 0;
    if (tri) {
        if (shard) {
            const size_t runs = (passBatches + kShardBatches - 1) / kShardBatches;
            const size_t myBatchesMax = (runs + c->world - 1) / c->world * kShardBatches;
            extraShare = myBatchesMax * kBatch * (size_t)(kMaxFan - 1);
            extrasCap = extraShare * c->world;
        } else {
            extraShare = extrasCap = passPrims * (size_t)(kMaxFan - 1);
        }
    }
    const size_t recCap = firstsCap + extrasCap;
    if (recCap > 0xfffffff0u) return fail(-4, "pass too large");
    const int chunkWords = (int)((2 * passBatches + 31) / 32);

    // chunk + group binning as a pass of its own (bin.cuh); the per-tile list capacity is a tuning knob, tiles
    // beyond it bin themselves inside the tile kernel
    const bool binPass = !getenv("SWR_NO_BIN_PASS");
    const uint32_t groupCap = tileShift == 6 ? 2048u : 1024u;          // <= kGroupList of tile.cuh
    // heavy-tile split (tile.cuh, swr_set_tile_split): tiles that list more than `splitThreshold` groups are shaded by four
    // CTAs (quadrants) that start first.  Measured on C3 (tools/rank_emul.py, profiles/r02_tile_split.txt): every quadrant
    // repeats the tile's record tests, which costs more than the shorter tail gives back at every GPU count (two ranks,
    // 64-pixel tiles: 0.324 -> 0.367 ms; eight ranks: 0.152 ms against 0.137 ms with 32-pixel tiles), so "automatic" is off.
    int splitReq = c->tileSplitReq;
    if (const char *env = getenv("SWR_TILE_SPLIT")) splitReq = atoi(env);
    if (splitReq < 0) splitReq = 0;
    const int splitCap = (binPass && splitReq > 0) ? 256 : 0;
    const SetLayout L = layoutOf(recCap, paramStride, needSpan, (size_t)tilesX * tilesY, chunkWords, passBatches, groupCap, binPass, splitCap);
    if (shard) {
        if (kArenaHeader + L.bytes > c->arena.bytes)
            return fail(-41, "shared scratch of %zu bytes is too small for this draw (needs %zu)", c->arena.bytes, kArenaHeader + L.bytes);
        ScratchSet &ss = c->sets[0];
        ss.base = static_cast<char *>(c->arena.ptr) + kArenaHeader;
        ss.ctr = reinterpret_cast<Counters *>(c->arena.ptr);
        ss.countersInit = true;                                          // zeroed when the arena was created
    } else {
        for (ScratchSet &ss : c->sets) {
            if (int rc = ss.blob.reserve(L.bytes)) return rc;
            if (int rc = ss.counters.reserve(sizeof(Counters))) return rc;
            ss.base = static_cast<char *>(ss.blob.ptr);
            ss.ctr = static_cast<Counters *>(ss.counters.ptr);
            if (!ss.countersInit) {
                CUDA_TRY(cudaMemset(ss.counters.ptr, 0, sizeof(Counters)));
                ss.countersInit = true;
            }
            if (!pipeline) break;                                        // without overlap only set 0 is used
        }
    }
    if (!c->hostFlags) {
        CUDA_TRY(cudaMallocHost((void **)&c->hostFlags, 64));
        c->hostFlags[0] = c->hostFlags[1] = 0;
    }

    g.drawMode = drawMode;
    g.px = c->px; g.py = c->py; g.ox = c->ox; g.oy = c->oy;
    g.depthN = c->depthN; g.depthF = c->depthF;
    g.cullMode = c->cullMode; g.rasterMode = c->rasterMode;
    g.scMinX = c->scMinX; g.scMinY = c->scMinY; g.scMaxX = c->scMaxX; g.scMaxY = c->scMaxY;
    g.nA = nA; g.nP = nP; g.useZ = useZ; g.useW = useW;
    g.paramStride = paramStride;
    g.chunkWords = chunkWords;
    g.tileShift = tileShift;
    g.tilesX = tilesX; g.tilesY = tilesY;
    g.rank = c->rank;
    g.world = c->world;
    g.shard = shard && !rasterVerts ? 1 : 0;                  // raster lists: every rank walks the whole list

    TileArgs t;
    memset(&t, 0, sizeof(t));
    t.paramStride = paramStride;
    t.chunkWords = chunkWords;
    t.tilesX = tilesX; t.tilesY = tilesY;
    t.rank = c->rank; t.world = c->world;
    t.mirrorSlot = c->mirrorSlot;
    t.mirrorCount = c->mirrorCount;
    for (int m = 0; m < SWR_MAX_TILE_MIRRORS; ++m) t.mirror[m] = m < c->mirrorCount ? c->mirror[m] : nullptr;
    t.rtWidth = c->rtW; t.rtHeight = c->rtH;
    t.numRT = ps->render_targets;
    for (int s = 0; s < SWR_MAX_RENDER_TARGETS; ++s) t.rt[s] = c->rt[s];
    t.scMinX = c->scMinX; t.scMinY = c->scMinY; t.scMaxX = c->scMaxX; t.scMaxY = c->scMaxY;
    if (c->debugTileStats) {
        if (int rc = c->tileStats.reserve((size_t)tilesX * tilesY * 64)) return rc;
        CUDA_TRY(cudaMemsetAsync(c->tileStats.ptr, 0, (size_t)tilesX * tilesY * 64, c->stream));
        t.tileStats = static_cast<uint32_t *>(c->tileStats.ptr);
        c->lastTiles = tilesX * tilesY;
    }

    swr_launch_fn geomLaunch = rasterVerts ? &launchRasterList : vs->launch_geometry;
    swr_launch_fn tileLaunch = ps->launch_tiles[drawMode][tileShift - 5];
    // this rank's share of the extras range
    const uint32_t extrasBegin = (uint32_t)(firstsCap + (g.shard ? extraShare * (size_t)c->rank : 0));
    g.extrasEnd = (uint32_t)(extrasBegin + extraShare);
    // what the scratch was cleared for: the bitmap's place and size (the peers rely on the clear, see below)
    const uint64_t layoutKey = ((uint64_t)L.tilemap * 1000003ull) ^ ((uint64_t)tilesX * tilesY * (uint64_t)chunkWords * 2654435761ull) ^ ((uint64_t)extrasBegin << 1) | 1ull;

    if (streamIdx) {
        const size_t npass = (nprims + passPrims - 1) / passPrims;
        while (c->idxReady.size() < npass) {
            cudaEvent_t ev;
            CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->idxReady.push_back(ev);
        }
        // Narrowed upload (hostpack.cpp): a slice whose blocks of 4096 indices each span at most 65535 vertices crosses
        // PCIe as 16-bit offsets + one base per block and is widened on the device; the host threads pack slice k + 1
        // while slice k (and, before the first one, the vertex attributes) are on the link.  Automatic mode keeps it
        // only where the host packs faster than the link would have carried the saved bytes.
        int narrow = c->narrowMode;
        if (const char *env = getenv("SWR_INDEX_NARROWING")) narrow = atoi(env);
        if (narrow < 0) {
            const char *env = getenv("SWR_NARROW_MIN_GBPS");
            narrow = (c->packDraws < 3 || c->packGBps >= (env ? atof(env) : 32.0)) ? 1 : 0;   // judged by the best of the first draws
        }
        if (narrow && !__builtin_cpu_supports("avx2")) narrow = 0;
        uint16_t *h16 = nullptr, *d16 = nullptr;
        int32_t *hBase = nullptr, *dBase = nullptr;
        if (narrow) {
            const size_t nIdx = nprims * per, nBlocksMax = nIdx / kIdxBlock + npass + 1;
            const size_t baseOfs = (nIdx * 2 + 255) / 256 * 256;
            if (int rc = c->pack16[S].reserve(baseOfs + nBlocksMax * 4)) return rc;
            if (int rc = c->dev16[S].reserve(baseOfs + nBlocksMax * 4)) return rc;
            if (!c->packPool) c->packPool = swr_hostpack_create(-1);
            // the copies that last read this set's host buffer belong to a draw whose kernels stageFree[S] follows
            if (c->stageUsed[S]) CUDA_TRY(cudaEventSynchronize(c->stageFree[S]));
            h16 = static_cast<uint16_t *>(c->pack16[S].ptr);
            d16 = static_cast<uint16_t *>(c->dev16[S].ptr);
            hBase = reinterpret_cast<int32_t *>(static_cast<char *>(c->pack16[S].ptr) + baseOfs);
            dBase = reinterpret_cast<int32_t *>(static_cast<char *>(c->dev16[S].ptr) + baseOfs);
        }
        size_t k = 0, blockAt = 0;
        double packSec = 0.0, packBytes = 0.0;
        for (size_t first = 0; first < nprims; first += passPrims, ++k) {
            const size_t n = std::min(passPrims, nprims - first);
            const size_t at = first * per, cnt = n * per;
            bool narrowed = false;
            if (narrow) {
                const size_t nb = (cnt + kIdxBlock - 1) / kIdxBlock;
                const auto t0 = std::chrono::steady_clock::now();
                narrowed = swr_hostpack_pack16(c->packPool, indices + at, cnt, h16 + at, hBase + blockAt) != 0;
                const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                packSec += sec;
                if (narrowed) {
                    packBytes += (double)cnt * 4.0;
                    CUDA_TRY(cudaMemcpyAsync(d16 + at, h16 + at, cnt * 2, cudaMemcpyHostToDevice, hs));
                    CUDA_TRY(cudaMemcpyAsync(dBase + blockAt, hBase + blockAt, nb * 4, cudaMemcpyHostToDevice, hs));
                    widenIndicesKernel<<<148 * 4, 256, 0, hs>>>(d16 + at, dBase + blockAt, static_cast<int32_t *>(stageIdx.ptr) + at, cnt);
                    c->stats.kernel_launches++;
                    c->stats.h2d_bytes += cnt * 2 + nb * 4;
                    blockAt += nb;
                }
            }
            if (!narrowed) {
                CUDA_TRY(cudaMemcpyAsync(static_cast<int32_t *>(stageIdx.ptr) + at, indices + at, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, hs));
                c->stats.h2d_bytes += cnt * sizeof(int32_t);
            }
            CUDA_TRY(cudaEventRecord(c->idxReady[k], hs));
        }
        if (narrow && packSec > 0.0) {                     // source bytes packed per second of this draw (failed slices count as time only)
            c->packGBps = std::max(c->packGBps, packBytes / packSec * 1e-9);
            c->packDraws++;
        }
    }
    CUDA_TRY(cudaEventRecord(c->evGeom0, gs));
    bool firstPass = true;
    size_t passIndex = 0;
    c->stats.last_geometry_ms = 0.0f;
    for (size_t first = 0; first < nprims; first += passPrims, ++passIndex) {
        const size_t n = std::min(passPrims, nprims - first);
        if (streamIdx) CUDA_TRY(cudaStreamWaitEvent(gs, c->idxReady[passIndex], 0));
        ScratchSet &ss = c->sets[pipeline ? (c->passSeq & 1) : 0];
        Counters *dc = ss.ctr;
        g.indices = devIndices + first * per;
        g.numPrims = (int)n;
        g.firstBatch = (uint32_t)(first / kBatch);
        g.extraAlloc = &dc->extraAlloc;
        for (int r = 0; r < kMaxRanks; ++r) memset(&g.sink[r], 0, sizeof(RecordSink));
        const RecordSink own = sinkAt(ss.base, L, dc);
        if (shard) {
            for (int r = 0; r < c->world; ++r)
                g.sink[r] = r == c->rank ? own : sinkAt(c->peerArena[r] + kArenaHeader, L, reinterpret_cast<Counters *>(c->peerArena[r]));
            if (getenv("SWR_DEBUG_FAKE_PEERS")) {
                // measurement aid (results are wrong): the records for the peers go to a local buffer instead of over NVLink
                if (int rc = c->l2flush.reserve(c->arena.bytes)) return rc;
                for (int r = 0; r < c->world; ++r)
                    if (r != c->rank) g.sink[r] = sinkAt(static_cast<char *>(c->l2flush.ptr) + kArenaHeader, L, reinterpret_cast<Counters *>(c->l2flush.ptr));
            }
        } else {
            g.sink[c->rank] = own;
        }
        t.bbox = own.bbox; t.gbox = own.gbox; t.head = own.head; t.params = own.params; t.span = own.span;
        t.tilemap = own.tilemap;
        t.gcnt = own.gcnt;
        t.groupList = binPass ? reinterpret_cast<uint32_t *>(ss.base + L.groupList) : nullptr;
        t.groupCount = binPass ? reinterpret_cast<uint32_t *>(ss.base + L.groupCount) : nullptr;
        t.groupCap = groupCap;
        t.splitCap = splitCap;
        t.splitThreshold = (uint32_t)splitReq;
        t.heavyList = reinterpret_cast<uint32_t *>(ss.base + L.heavyList);
        t.heavyFlag = splitCap ? reinterpret_cast<uint8_t *>(ss.base + L.heavyFlag) : nullptr;
        t.extra = own.extra;
        t.fragCounter = &dc->fragments;
        t.errorFlag = &dc->errorFlag;
        t.numPrims = (int)n;
        t.numChunks = (int)(2 * ((n + kBatch - 1) / kBatch));

        // geometry stream: wait until the tile kernel that last read this set is done, then refill it
        if (gs != c->stream && ss.tilePending) CUDA_TRY(cudaStreamWaitEvent(gs, ss.tileDone, 0));
        const uint32_t init[2] = { extrasBegin, 0u };                    // extraAlloc, errorFlag
        auto clearSet = [&]() -> int {
            CUDA_TRY(cudaMemsetAsync(own.tilemap, 0, L.heavyList + 4 - L.tilemap, gs));           // the bitmap and the heavy-tile counter behind it
            CUDA_TRY(cudaMemcpyAsync(&dc->extraAlloc, init, sizeof(init), cudaMemcpyHostToDevice, gs));
            return 0;
        };
        if (shard) {
            // The peers' geometry kernels write into this rank's bitmap and records, so (1) the bitmap must be clear
            // and the previous pass's tile kernel done on EVERY rank before any geometry kernel of this pass starts,
            // and (2) every geometry kernel must be done before any tile kernel reads.  (1) is a clear followed by a
            // barrier; the clear is issued right after the tile kernel of the previous pass, so that a barrier the
            // host enqueues anyway at the end of a frame (swr_peer_barrier) also serves as (1) of the next draw.
            {   // resolve the geometry kernel before anything waits on a peer (see launchGeometry)
                GeomArgs pre = g;
                pre.numPrims = 0;
                geomLaunch(&pre, gs);
            }
            if (c->clearedKey != layoutKey) {
                if (int rc = clearSet()) return rc;
                c->clearedKey = layoutKey;
                c->fencedSinceClear = false;
            }
            if (!c->fencedSinceClear)
                if (int rc = enqueuePeerBarrier(c)) return rc;
            c->clearedKey = 0;                                           // the geometry kernels are about to dirty it
            c->fencedSinceClear = false;
        } else {
            if (int rc = clearSet()) return rc;
        }
        geomLaunch(&g, gs);
        if (first + passPrims >= nprims) CUDA_TRY(cudaEventRecord(c->evGeom1, gs));
        // main stream: tiles of this pass after its geometry
        if (gs != c->stream) {
            CUDA_TRY(cudaEventRecord(ss.geomDone, gs));
            CUDA_TRY(cudaStreamWaitEvent(c->stream, ss.geomDone, 0));
        }
        if (shard)
            if (int rc = enqueuePeerBarrier(c)) return rc;               // (2)
        if (firstPass) CUDA_TRY(cudaEventRecord(c->evTile0, c->stream));
        firstPass = false;
        if (binPass) {
            launchBin(t, tileShift, c->stream);
            c->stats.kernel_launches++;
        }
        tileLaunch(&t, c->stream);
        if (gs != c->stream) {
            CUDA_TRY(cudaEventRecord(ss.tileDone, c->stream));
            ss.tilePending = true;
        }
        if (shard) {
            if (int rc = clearSet()) return rc;                          // ready for the next pass / draw of the same shape
            c->clearedKey = layoutKey;
            c->fencedSinceClear = false;
        }
        c->stats.kernel_launches += 2;
        c->stats.passes++;
        c->passSeq++;
    }
    CUDA_TRY(cudaEventRecord(c->evTile1, c->stream));
    if (staged) {
        CUDA_TRY(cudaEventRecord(c->stageFree[S], c->stream));
        c->stageUsed[S] = true;
    }
    CUDA_TRY(cudaGetLastError());
    c->haveDrawEvents = true;
    c->lastDrawMode = drawMode;
    c->stats.last_tile_size = T;
    return 0;
}

} // namespace

// =============================================================================================== C ABI
extern "C" {

int swr_abi_version(void) { return 1; }
const char *swr_last_error(void) { return g_lastError.c_str(); }

int swr_create(swr_context **out, int cuda_device)
{
    if (!out) return fail(-1, "null out pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(-20, "no CUDA device: this library has no CPU fallback");
    }
    if (cuda_device < 0 || cuda_device >= ndev) return fail(-21, "bad CUDA device %d (have %d)", cuda_device, ndev);
    swr_context *c = new swr_context();
    c->device = cuda_device;
    cudaError_t e = cudaSetDevice(cuda_device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking);
    c->stream = c->ownStream;
    cudaEvent_t *evs[] = { &c->evGeom0, &c->evGeom1, &c->evTile0, &c->evTile1, &c->evTimer0, &c->evTimer1 };
    for (cudaEvent_t *ev : evs)
        if (e == cudaSuccess) e = cudaEventCreate(ev);
    cudaEvent_t *sync[] = { &c->evDrawStart, &c->sets[0].geomDone, &c->sets[0].tileDone, &c->sets[1].geomDone, &c->sets[1].tileDone,
                            &c->stageFree[0], &c->stageFree[1] };
    for (cudaEvent_t *ev : sync)
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (const char *env = getenv("SWR_PIPELINE")) c->pipeline = atoi(env) != 0;
    if (const char *env = getenv("SWR_SCRATCH_LIMIT_GB")) c->scratchLimit = (size_t)std::max(1, atoi(env)) << 30;
    if (e != cudaSuccess) {
        delete c;
        return fail(-22, "context creation: %s", cudaGetErrorString(e));
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cuda_device) == cudaSuccess && prop.major < 10) {
        swr_destroy(c);
        return fail(-23, "device %d is sm_%d%d; this library is built for sm_100a only", cuda_device, prop.major, prop.minor);
    }
    *out = c;
    return 0;
}

void swr_destroy(swr_context *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->aux) cudaStreamSynchronize(c->aux);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = { &c->stageIdx, &c->stageIdxB, &c->l2flush, &c->ownedIdx, &c->tileStats, &c->arena, &c->soVerts, &c->soIndices, &c->soCounts, &c->maxIndexBuf };
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < SWR_MAX_VERTEX_ATTRIBS; ++i) c->stageAttribB[i].release();
    for (ScratchSet &ss : c->sets) ss.release();
    for (int i = 0; i < SWR_MAX_VERTEX_ATTRIBS; ++i) c->stageAttrib[i].release();
    for (auto &o : c->ipcOpened) cudaIpcCloseMemHandle(o.second);
    if (c->hostFlags) cudaFreeHost(c->hostFlags);
    if (c->packPool) swr_hostpack_destroy(c->packPool);
    for (int i = 0; i < 2; ++i) { c->pack16[i].release(); c->dev16[i].release(); }
    cudaEvent_t evs[] = { c->evGeom0, c->evGeom1, c->evTile0, c->evTile1, c->evTimer0, c->evTimer1, c->evDrawStart,
                          c->sets[0].geomDone, c->sets[0].tileDone, c->sets[1].geomDone, c->sets[1].tileDone, c->stageFree[0], c->stageFree[1] };
    for (cudaEvent_t ev : evs)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->idxReady) cudaEventDestroy(ev);
    if (c->copy) { cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy); }
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
}

int swr_set_viewport(swr_context *c, int x, int y, int width, int height)
{
    if (!c) return fail(-1, "null context");
    c->vpX = x; c->vpY = y; c->vpW = width; c->vpH = height;
    c->px = width / 2.0f;                                   // VertexProcessor.cpp:50-53
    c->py = height / 2.0f;
    c->ox = (x + c->px);
    c->oy = (y + c->py);
    return 0;
}

int swr_set_depth_range(swr_context *c, float n, float f)
{
    if (!c) return fail(-1, "null context");
    c->depthN = n; c->depthF = f;
    return 0;
}

int swr_set_cull_mode(swr_context *c, int mode)
{
    if (!c) return fail(-1, "null context");
    if (mode < 0 || mode > 2) return fail(-2, "bad cull mode %d", mode);
    c->cullMode = mode;
    return 0;
}

int swr_set_vertex_attrib_pointer(swr_context *c, int index, int stride, const void *buffer, size_t bytes)
{
    if (!c) return fail(-1, "null context");
    if (index < 0 || index >= SWR_MAX_VERTEX_ATTRIBS) return fail(-2, "attribute index %d out of range", index);   // assert at VertexProcessor.cpp:69
    c->attribs[index].ptr = buffer;
    c->attribs[index].stride = stride;
    c->attribs[index].bytes = bytes;
    return 0;
}

int swr_set_vertex_shader(swr_context *c, const swr_vertex_shader *vs)
{
    if (!c || !vs) return fail(-1, "null argument");
    if (vs->attrib_count > SWR_MAX_VERTEX_ATTRIBS) return fail(-2, "AttribCount %d > %d", vs->attrib_count, SWR_MAX_VERTEX_ATTRIBS);  // VertexProcessor.h:80
    if (vs->args_layout != SWR_ARGS_LAYOUT) return fail(-24, "vertex shader '%s' was compiled against other swr headers than this library: rebuild it", vs->name);
    c->vs = vs;
    return 0;
}

int swr_set_raster_mode(swr_context *c, int mode)
{
    if (!c) return fail(-1, "null context");
    if (mode < 0 || mode > 2) return fail(-2, "bad raster mode %d", mode);
    c->rasterMode = mode;
    return 0;
}

int swr_set_scissor_rect(swr_context *c, int x, int y, int width, int height)
{
    if (!c) return fail(-1, "null context");
    c->scMinX = x; c->scMinY = y;                           // Rasterizer.h:81-87
    c->scMaxX = x + width; c->scMaxY = y + height;
    return 0;
}

int swr_set_pixel_shader(swr_context *c, const swr_pixel_shader *ps)
{
    if (!c || !ps) return fail(-1, "null argument");
    if (ps->render_targets > SWR_MAX_RENDER_TARGETS) return fail(-2, "RenderTargets %d > %d", ps->render_targets, SWR_MAX_RENDER_TARGETS);
    if (ps->args_layout != SWR_ARGS_LAYOUT) return fail(-24, "pixel shader '%s' was compiled against other swr headers than this library: rebuild it", ps->name);
    c->ps = ps;
    return 0;
}

int swr_set_render_target(swr_context *c, int slot, void *device_ptr, int pitch_bytes, int width, int height)
{
    if (!c) return fail(-1, "null context");
    if (slot < 0 || slot >= SWR_MAX_RENDER_TARGETS) return fail(-2, "render target slot %d out of range", slot);
    if (device_ptr) {
        if (!isDevicePointer(device_ptr)) return fail(-2, "render target %d is not device memory", slot);
        if (width <= 0 || height <= 0 || pitch_bytes < width * 4 || (pitch_bytes & 3)) return fail(-2, "bad render target geometry");
        c->rtW = width;
        c->rtH = height;
    }
    c->rt[slot].ptr = device_ptr;
    c->rt[slot].pitch = pitch_bytes;
    return 0;
}

int swr_set_uniforms(swr_context *c, const void *data, size_t bytes)
{
    if (!c) return fail(-1, "null context");
    if (bytes > SWR_MAX_UNIFORM_BYTES) return fail(-2, "uniform block of %zu bytes exceeds %d", bytes, SWR_MAX_UNIFORM_BYTES);
    if (bytes) memcpy(c->uniforms, data, bytes);
    c->uniformBytes = bytes;
    return 0;
}

int swr_set_tile_size(swr_context *c, int tile_size)
{
    if (!c) return fail(-1, "null context");
    if (tile_size != 0 && tile_size != 32 && tile_size != 64) return fail(-2, "tile size must be 0, 32 or 64");
    if (tile_size != c->tileSizeReq) c->pinnedTileShift = 0;
    c->tileSizeReq = tile_size;
    return 0;
}

int swr_set_index_narrowing(swr_context *c, int mode)
{
    if (!c) return fail(-1, "null context");
    if (mode < -1 || mode > 1) return fail(-2, "index narrowing mode must be -1 (automatic), 0 (off) or 1 (always try)");
    c->narrowMode = mode;
    return 0;
}

int swr_debug_pack_indices16(const int32_t *indices, size_t count, uint16_t *offsets, int32_t *bases)
{
    if (!indices || !offsets || !bases) return fail(-1, "null argument");
    if (!__builtin_cpu_supports("avx2")) return -1;
    void *pool = swr_hostpack_create(3);
    const int ok = swr_hostpack_pack16(pool, indices, count, offsets, bases);
    swr_hostpack_destroy(pool);
    return ok;
}

int swr_set_tile_split(swr_context *c, int groups)
{
    if (!c) return fail(-1, "null context");
    if (groups < -1) return fail(-2, "tile split threshold must be -1 (automatic), 0 (off) or a number of 32-record groups");
    c->tileSplitReq = groups;
    return 0;
}

int swr_set_tile_partition(swr_context *c, int rank, int world)
{
    if (!c) return fail(-1, "null context");
    if (world < 1 || rank < 0 || rank >= world) return fail(-2, "bad partition %d/%d", rank, world);
    if (rank != c->rank || world != c->world) c->pinnedTileShift = 0;
    c->rank = rank;
    c->world = world;
    if (world == 1) c->shardGeometry = false;
    return 0;
}

int swr_shared_scratch_create(swr_context *c, size_t bytes, void **base)
{
    if (!c || !base) return fail(-1, "null argument");
    if (int rc = setDevice(c)) return rc;
    if (bytes < kArenaHeader + ((size_t)1 << 20)) return fail(-2, "shared scratch must be at least %zu bytes", kArenaHeader + ((size_t)1 << 20));
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->arena.release();
    c->shardGeometry = false;
    c->sets[0].release();
    c->sets[1].release();
    c->arena.bytes = 0;
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);                   // exact size: the peers compute the same layout from it
    if (e != cudaSuccess) return fail(-101, "cudaMalloc(%zu bytes) for the shared scratch: %s", bytes, cudaGetErrorString(e));
    c->arena.ptr = p;
    c->arena.bytes = bytes;
    CUDA_TRY(cudaMemset(p, 0, kArenaHeader));
    c->barrierEpoch = 0;
    c->clearedKey = 0;
    c->fencedSinceClear = false;
    *base = p;
    return 0;
}

int swr_set_geometry_shards(swr_context *c, int rank, int world, void *const *arenas)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(-2, "bad partition %d/%d (at most %d ranks)", rank, world, kMaxRanks);
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (world > 1) {
        if (!arenas) return fail(-1, "null arena list");
        if (!c->arena.ptr) return fail(-40, "swr_shared_scratch_create first");
        for (int r = 0; r < world; ++r) {
            if (r != rank && !arenas[r]) return fail(-2, "shared scratch of rank %d is null", r);
            c->peerArena[r] = r == rank ? static_cast<char *>(c->arena.ptr) : static_cast<char *>(arenas[r]);
        }
    }
    if (rank != c->rank || world != c->world) c->pinnedTileShift = 0;
    c->rank = rank;
    c->world = world;
    c->shardGeometry = world > 1;
    c->clearedKey = 0;
    c->fencedSinceClear = false;
    return 0;
}

int swr_peer_barrier(swr_context *c)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    return enqueuePeerBarrier(c);
}

int swr_set_stream(swr_context *c, void *cuda_stream)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (ScratchSet &ss : c->sets) ss.tilePending = false;
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->ownStream;
    return 0;
}

int swr_set_pipeline(swr_context *c, int enable, int passes_hint, int overlap_draws)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (ScratchSet &ss : c->sets) ss.tilePending = false;
    c->pipeline = enable != 0;
    c->passesHint = passes_hint;
    c->overlapDraws = enable != 0 && overlap_draws != 0;
    return 0;
}

int swr_set_scratch_limit(swr_context *c, size_t bytes)
{
    if (!c) return fail(-1, "null context");
    c->scratchLimit = std::max<size_t>(bytes, (size_t)1 << 20);
    return 0;
}

int swr_draw_elements(swr_context *c, int draw_mode, size_t count, const int32_t *indices)
{
    if (!c) return fail(-1, "null context");
    if (count && !indices) return fail(-1, "null indices");
    return drawCommon(c, draw_mode, count, indices, nullptr, 0);
}

int swr_draw_raster_list(swr_context *c, int draw_mode, const void *vertices, size_t vertex_count, const int32_t *indices, size_t index_count)
{
    if (!c) return fail(-1, "null context");
    if (index_count && (!indices || !vertices)) return fail(-1, "null vertices / indices");
    if (index_count == 0) return 0;
    return drawCommon(c, draw_mode, index_count, indices, vertices, vertex_count);
}

int swr_process_elements(swr_context *c, int drawMode, size_t count, const int32_t *indices, swr_stream_out_fn emit, void *user)
{
    if (!c || !emit) return fail(-1, "null argument");
    if (count && !indices) return fail(-1, "null indices");
    if (int rc = setDevice(c)) return rc;
    if (drawMode < 0 || drawMode > 2) return fail(-2, "bad draw mode %d", drawMode);
    const swr_vertex_shader *vs = c->vs;
    if (!vs) return fail(-3, "no vertex shader set");
    if (!vs->launch_stream_out) return fail(-3, "vertex shader '%s' has no stream-out launcher", vs->name);
    for (int i = 0; i < vs->attrib_count; ++i)
        if (!c->attribs[i].ptr) return fail(-8, "vertex attribute %d not set", i);
    const int per = drawMode + 1;
    const size_t nprims = count / per;
    if (nprims == 0) return 0;
    if (nprims > (size_t)0x7fffffff / 4) return fail(-4, "too many primitives in one draw (%zu)", nprims);
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    cudaStream_t st = c->stream;

    GeomArgs g;
    memset(&g, 0, sizeof(g));
    const int32_t *devIndices = indices;
    const bool idxOnDevice = isDevicePointer(indices);
    if (!idxOnDevice) {
        if (int rc = c->stageIdx.reserve(count * sizeof(int32_t))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->stageIdx.ptr, indices, nprims * per * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        devIndices = static_cast<const int32_t *>(c->stageIdx.ptr);
    }
    int maxIndex = -2;
    for (int i = 0; i < vs->attrib_count; ++i) {
        const Attrib &a = c->attribs[i];
        const void *dp = a.ptr;
        if (!isDevicePointer(a.ptr)) {
            size_t bytes = a.bytes;
            if (bytes == 0) {
                if (a.stride <= 0) return fail(-9, "vertex attribute %d is host memory with stride %d: its extent is required (bytes > 0)", i, a.stride);
                if (maxIndex == -2)
                    if (int rc = maxIndexOf(c, indices, nprims * per, idxOnDevice, &maxIndex)) return rc;
                if (maxIndex < 0) return fail(-9, "vertex attribute %d is host memory and the draw has no valid index", i);
                bytes = (size_t)a.stride * ((size_t)maxIndex + 1);
            }
            if (int rc = c->stageAttrib[i].reserve(bytes)) return rc;
            CUDA_TRY(cudaMemcpyAsync(c->stageAttrib[i].ptr, a.ptr, bytes, cudaMemcpyHostToDevice, st));
            dp = c->stageAttrib[i].ptr;
        }
        g.attribPtr[i] = dp;
        g.attribStride[i] = a.stride;
    }
    if (c->uniformBytes && vs->set_uniforms && vs->set_uniforms(c->uniforms, c->uniformBytes, st) != 0)
        return fail(-10, "uniform upload failed (vertex shader '%s')", vs->name);

    ScratchSet &ss = c->sets[0];
    if (!ss.ctr) {
        if (int rc = ss.counters.reserve(sizeof(Counters))) return rc;
        CUDA_TRY(cudaMemset(ss.counters.ptr, 0, sizeof(Counters)));
        ss.ctr = static_cast<Counters *>(ss.counters.ptr);
        ss.countersInit = true;
    }
    const size_t passBatches = 32;
    const uint32_t extraCap = drawMode == SWR_DRAW_TRIANGLE ? (uint32_t)(kBatch * (kMaxFan - 1)) : 0u;
    const size_t stride = (size_t)per * kBatch + 3 * (size_t)extraCap;
    if (int rc = c->soVerts.reserve(passBatches * stride * 144)) return rc;
    if (int rc = c->soIndices.reserve(passBatches * stride * sizeof(int32_t))) return rc;
    if (int rc = c->soCounts.reserve(passBatches * sizeof(uint32_t))) return rc;

    g.drawMode = drawMode;
    g.px = c->px; g.py = c->py; g.ox = c->ox; g.oy = c->oy;
    g.depthN = c->depthN; g.depthF = c->depthF;
    g.cullMode = c->cullMode; g.rasterMode = c->rasterMode;
    g.rank = 0; g.world = 1;
    g.sink[0].errorFlag = &ss.ctr->errorFlag;
    g.soVerts = c->soVerts.ptr;
    g.soIndices = static_cast<int32_t *>(c->soIndices.ptr);
    g.soCounts = static_cast<uint32_t *>(c->soCounts.ptr);
    g.soExtraCap = extraCap;

    std::vector<uint32_t> counts(passBatches);
    std::vector<unsigned char> hv;
    std::vector<int32_t> hi;
    const size_t passPrims = passBatches * kBatch;
    c->stats.draws++;
    c->stats.primitives_in += nprims;
    for (size_t first = 0; first < nprims; first += passPrims) {
        const size_t n = std::min(passPrims, nprims - first);
        const size_t batches = (n + kBatch - 1) / kBatch;
        g.indices = devIndices + first * per;
        g.numPrims = (int)n;
        g.firstBatch = (uint32_t)(first / kBatch);
        vs->launch_stream_out(&g, st);
        c->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(counts.data(), c->soCounts.ptr, batches * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (size_t b = 0; b < batches; ++b) {
            const size_t cnt = std::min<size_t>(kBatch, n - b * kBatch);
            const size_t items = per * cnt + 3 * (size_t)counts[b];       // vertices == indices of the batch
            hv.resize(items * 144);
            hi.resize(items);
            CUDA_TRY(cudaMemcpy(hv.data(), static_cast<char *>(c->soVerts.ptr) + b * stride * 144, items * 144, cudaMemcpyDeviceToHost));
            CUDA_TRY(cudaMemcpy(hi.data(), static_cast<int32_t *>(c->soIndices.ptr) + b * stride, items * sizeof(int32_t), cudaMemcpyDeviceToHost));
            emit(user, drawMode, hv.data(), items, hi.data(), items);
        }
        c->stats.passes++;
    }
    return 0;
}

int swr_finish(swr_context *c)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    for (int k = 0; k < 2; ++k) {
        ScratchSet &ss = c->sets[k];
        if (!ss.countersInit) continue;
        Counters *dc = ss.ctr;
        CUDA_TRY(cudaMemcpyAsync(c->hostFlags + k, &dc->stickyFlag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemsetAsync(&dc->stickyFlag, 0, sizeof(uint32_t), c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (ScratchSet &ss : c->sets) ss.tilePending = false;
    if (c->hostFlags && (c->hostFlags[0] | c->hostFlags[1])) {
        const uint32_t f = c->hostFlags[0] | c->hostFlags[1];
        c->hostFlags[0] = c->hostFlags[1] = 0;
        if (f & 1u) return fail(-30, "geometry scratch exhausted: the last draw produced nothing (raise swr_set_scratch_limit)");
        if (f & 2u) return fail(-31, "a line longer than %d DDA steps was dropped", kMaxLineSteps);
        if (f & 4u) return fail(-32, "a clipped polygon of more than %d vertices was dropped", kMaxPoly);
        if (f & 8u) return fail(-33, "peer barrier timed out: a rank of the partition did not reach it");
    }
    return 0;
}

int swr_get_stats(swr_context *c, swr_stats *out)
{
    if (!c || !out) return fail(-1, "null argument");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->stats.fragments = 0;
    for (ScratchSet &ss : c->sets)
        if (ss.countersInit) {
            Counters h;
            CUDA_TRY(cudaMemcpy(&h, ss.ctr, sizeof(h), cudaMemcpyDeviceToHost));
            c->stats.fragments += h.fragments;
        }
    if (c->haveDrawEvents) {
        // geometry: first launch to last completion on its stream; tiles: first launch to the end of the
        // draw on the main stream.  With pipelining on, the two intervals overlap.
        cudaEventElapsedTime(&c->stats.last_geometry_ms, c->evGeom0, c->evGeom1);
        cudaEventElapsedTime(&c->stats.last_tile_ms, c->evTile0, c->evTile1);
    }
    c->stats.scratch_bytes = scratchBytes(c);
    *out = c->stats;
    return 0;
}

int swr_reset_stats(swr_context *c)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (ScratchSet &ss : c->sets)
        if (ss.countersInit) CUDA_TRY(cudaMemset(&ss.ctr->fragments, 0, sizeof(unsigned long long)));
    const uint64_t scratch = c->stats.scratch_bytes;
    const int tile = c->stats.last_tile_size;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.scratch_bytes = scratch;
    c->stats.last_tile_size = tile;
    return 0;
}

int swr_timer_begin(swr_context *c)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaEventRecord(c->evTimer0, c->stream));
    return 0;
}

int swr_timer_end(swr_context *c, float *ms)
{
    if (!c || !ms) return fail(-1, "null argument");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaEventRecord(c->evTimer1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->evTimer1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->evTimer0, c->evTimer1));
    return 0;
}

void *swr_device_alloc(swr_context *c, size_t bytes)
{
    if (!c) return nullptr;
    cudaSetDevice(c->device);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(-101, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

int swr_device_free(swr_context *c, void *ptr)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->aux));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaFree(ptr));
    return 0;
}

void *swr_host_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(-101, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}

int swr_host_free_pinned(void *ptr)
{
    CUDA_TRY(cudaFreeHost(ptr));
    return 0;
}

int swr_memcpy_h2d(swr_context *c, void *dst, const void *src, size_t bytes)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int swr_memcpy_d2h(swr_context *c, void *dst, const void *src, size_t bytes)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

int swr_memset32(swr_context *c, void *dst, uint32_t value, size_t count)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    if (count == 0) return 0;
    const int blocks = (int)std::min<size_t>((count + 1023) / 1024, 148 * 8);
    fill32Kernel<<<blocks, 256, 0, c->stream>>>(static_cast<uint32_t *>(dst), value, count);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int swr_flush_l2(swr_context *c)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    const size_t bytes = (size_t)256 << 20;
    if (int rc = c->l2flush.reserve(bytes)) return rc;
    fill32Kernel<<<148 * 8, 256, 0, c->stream>>>(static_cast<uint32_t *>(c->l2flush.ptr), 0u, bytes / 4);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int64_t swr_owned_tile_count(int width, int height, int tile_size, int rank, int world)
{
    const int tilesX = (width + tile_size - 1) / tile_size, tilesY = (height + tile_size - 1) / tile_size;
    int64_t n = 0;
    for (int ty = 0; ty < tilesY; ++ty)
        for (int tx = 0; tx < tilesX; ++tx) n += tileOwned(tx, ty, rank, world) ? 1 : 0;
    return n;
}

static int exchangeTiles(swr_context *c, int slot, int rank, int world, int tile_size, void *buf, bool pack)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    if (slot < 0 || slot >= SWR_MAX_RENDER_TARGETS || !c->rt[slot].ptr) return fail(-2, "render target slot %d not registered", slot);
    if (tile_size != 32 && tile_size != 64) return fail(-2, "tile size must be 32 or 64");
    const int shift = tile_size == 64 ? 6 : 5;
    const int tilesX = (c->rtW + tile_size - 1) / tile_size, tilesY = (c->rtH + tile_size - 1) / tile_size;
    // ownedIdx[r][tile] = position of `tile` in rank r's exchange buffer; built once per geometry
    const int ntiles = tilesX * tilesY;
    const int key[5] = { tile_size, 0, world, c->rtW, c->rtH };
    if (memcmp(key, c->ownedKey, sizeof(key)) != 0 || !c->ownedIdx.ptr) {
        std::vector<int> owned((size_t)world * ntiles, -1);
        for (int r = 0; r < world; ++r) {
            int n = 0;
            for (int t = 0; t < ntiles; ++t)
                if (tileOwned(t % tilesX, t / tilesX, r, world)) owned[(size_t)r * ntiles + t] = n++;
        }
        if (int rc = c->ownedIdx.reserve(owned.size() * sizeof(int))) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->ownedIdx.ptr, owned.data(), owned.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));   // `owned` is a temporary
        memcpy(c->ownedKey, key, sizeof(key));
    }
    const int *ownedIndex = (const int *)c->ownedIdx.ptr + (size_t)rank * ntiles;
    if (pack)
        tileExchangeKernel<true><<<tilesX * tilesY, 256, 0, c->stream>>>((char *)c->rt[slot].ptr, c->rt[slot].pitch, c->rtW, c->rtH, shift,
                                                                       tilesX, tilesY, rank, world, ownedIndex, (uint32_t *)buf);
    else
        tileExchangeKernel<false><<<tilesX * tilesY, 256, 0, c->stream>>>((char *)c->rt[slot].ptr, c->rt[slot].pitch, c->rtW, c->rtH, shift,
                                                                        tilesX, tilesY, rank, world, ownedIndex, (uint32_t *)buf);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int swr_pack_tiles(swr_context *c, int slot, int rank, int world, int tile_size, void *dst_device)
{
    return exchangeTiles(c, slot, rank, world, tile_size, dst_device, true);
}

int swr_unpack_tiles(swr_context *c, int slot, int rank, int world, int tile_size, const void *src_device)
{
    return exchangeTiles(c, slot, rank, world, tile_size, const_cast<void *>(src_device), false);
}

int swr_set_tile_mirrors(swr_context *c, int slot, int count, void *const *surfaces)
{
    if (!c) return fail(-1, "null context");
    if (count < 0 || count > SWR_MAX_TILE_MIRRORS) return fail(-2, "at most %d mirrors", SWR_MAX_TILE_MIRRORS);
    if (count > 0 && (slot < 0 || slot >= SWR_MAX_RENDER_TARGETS || !surfaces)) return fail(-2, "bad mirror slot / surfaces");
    for (int m = 0; m < count; ++m)
        if (!surfaces[m]) return fail(-2, "mirror surface %d is null", m);
    c->mirrorSlot = slot;
    c->mirrorCount = count;
    for (int m = 0; m < count; ++m) c->mirror[m] = surfaces[m];
    return 0;
}

// The allocation that holds `p`: cuMemGetAddressRange through the runtime's driver entry point (no libcuda link).
static int allocationBase(const void *p, void **base)
{
    typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
    static GetRange fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym)
            return fail(-100, "cuMemGetAddressRange is not available");
        fn = (GetRange)sym;
    }
    unsigned long long b = 0;
    size_t sz = 0;
    if (fn(&b, &sz, (unsigned long long)(uintptr_t)p) != 0) return fail(-2, "not a device allocation");
    *base = (void *)(uintptr_t)b;
    return 0;
}

int swr_ipc_get_handle(const void *device_ptr, void *handle64, int64_t *offset)
{
    if (!device_ptr || !handle64 || !offset) return fail(-1, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    void *base = nullptr;
    if (int rc = allocationBase(device_ptr, &base)) return rc;
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, base));
    memcpy(handle64, &h, 64);
    *offset = (int64_t)((const char *)device_ptr - (const char *)base);
    return 0;
}

int swr_ipc_open(swr_context *c, const void *handle64, int64_t offset, void **device_ptr)
{
    if (!c || !handle64 || !device_ptr) return fail(-1, "null argument");
    if (int rc = setDevice(c)) return rc;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *base = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *device_ptr = (char *)base + offset;
    c->ipcOpened.push_back({ *device_ptr, base });
    return 0;
}

int swr_ipc_close(swr_context *c, void *device_ptr)
{
    if (!c) return fail(-1, "null context");
    if (int rc = setDevice(c)) return rc;
    for (size_t i = 0; i < c->ipcOpened.size(); ++i)
        if (c->ipcOpened[i].first == device_ptr) {
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            CUDA_TRY(cudaIpcCloseMemHandle(c->ipcOpened[i].second));
            c->ipcOpened.erase(c->ipcOpened.begin() + (long)i);
            return 0;
        }
    return fail(-2, "pointer was not opened with swr_ipc_open");
}

int swr_debug_enable_tile_stats(swr_context *c, int enable)
{
    if (!c) return fail(-1, "null context");
#if defined(SWR_TILE_STATS) && SWR_TILE_STATS
    c->debugTileStats = enable != 0;
    return 0;
#else
    (void)enable;
    return fail(-7, "per-tile statistics are compiled out of this build (make VARIANT=_stats EXTRA=-DSWR_TILE_STATS=1)");
#endif
}

int64_t swr_debug_read_tile_stats(swr_context *c, uint32_t *out, int64_t cap_tiles)
{
    if (!c || !out) return fail(-1, "null argument");
    if (int rc = setDevice(c)) return rc;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail(-100, "sync failed");
    if (!c->tileStats.ptr) return 0;
    const int64_t n = std::min<int64_t>(cap_tiles, c->lastTiles);
    if (cudaMemcpy(out, c->tileStats.ptr, (size_t)n * 64, cudaMemcpyDeviceToHost) != cudaSuccess) return fail(-100, "copy failed");
    return c->lastTiles;
}

} // extern "C"
