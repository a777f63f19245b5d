// swr/detail/bin.cuh -- chunk and group binning as a pass of its own (between the geometry and the tile kernel).
//
// One small CTA per screen tile reads the tile's row of the tile x chunk bitmap, tests the 32-record group
// boxes of the flagged chunks and writes the hit groups, in ascending order (= the reference's emission
// order, see tile.cuh), to a per-tile list in global memory.  The tile kernel then starts directly with the
// record boxes of those groups.  Why a separate pass: these two steps are a chain of dependent global reads
// with almost no arithmetic; inside the tile kernel (two fat CTAs per SM) nothing hides their latency and with
// 32-pixel tiles they were 19-25 % of its time, here 8 CTAs per SM overlap them.  A tile whose list would
// exceed the per-tile capacity is flagged and the tile kernel falls back to doing the two steps itself, so
// the capacity is a tuning knob, never a limit.
#pragma once

#include "common.h"

#if defined(__CUDACC__)

namespace swr {
namespace detail {

constexpr int kBinThreads = 256;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinChunkList = 1024;
constexpr uint32_t kBinOverflow = 0xffffffffu;

// group list entry: group id (27 bits) | records of the group - 1 (5 bits); only groups with records are listed
SWR_HD uint32_t packGroup(uint32_t group, uint32_t count) { return group | ((count - 1u) << 27); }
SWR_HD uint32_t groupOf(uint32_t entry) { return entry & 0x07ffffffu; }
SWR_HD uint32_t groupCount(uint32_t entry) { return (entry >> 27) + 1u; }

// exclusive block scan of a packed (hi: count, lo: sum) pair, one barrier, double-buffered scratch
SWR_D uint64_t binScan(uint64_t v, uint64_t &total, uint64_t *scratch, int &phase)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t *s = scratch + phase * kBinWarps;
    phase ^= 1;
    uint64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s[wid] = incl;
    __syncthreads();
    const uint64_t w = lane < kBinWarps ? s[lane] : 0;
    uint64_t wi = w;
#pragma unroll
    for (int o = 1; o < kBinWarps; o <<= 1) {
        uint64_t n = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += n;
    }
    total = __shfl_sync(0xffffffffu, wi, kBinWarps - 1);
    return incl - v + __shfl_sync(0xffffffffu, wi - w, wid);
}

template <int TLOG>
__global__ void __launch_bounds__(kBinThreads) binKernel(const TileArgs t)
{
    constexpr int T = 1 << TLOG;
    const int tx = blockIdx.x % t.tilesX, ty = blockIdx.x / t.tilesX;
    if (!tileOwned(tx, ty, t.rank, t.world)) return;
    if (*t.errorFlag & 1u) return;

    __shared__ uint32_t cGroup[kBinChunkList];          // first group of each listed chunk
    __shared__ uint32_t cPair[kBinChunkList + 4];       // exclusive prefix of the group counts
    __shared__ uint64_t sScan[2 * kBinWarps];

    const int tid = threadIdx.x;
    const int X0 = tx << TLOG, Y0 = ty << TLOG, X1 = X0 + T - 1, Y1 = Y0 + T - 1;
    const uint32_t cap = t.groupCap;
    uint32_t *out = t.groupList + (size_t)blockIdx.x * cap;
    uint32_t nOut = 0;
    int phase = 0;

    const uint32_t *row = t.tilemap + (size_t)blockIdx.x * t.chunkWords;
    for (int wb = 0; wb < t.chunkWords && nOut <= cap; wb += kBinThreads) {
        uint32_t pending = (wb + tid < t.chunkWords) ? row[wb + tid] : 0u;
        while (true) {
            // this step's set bits -> chunk list {first group, group count}
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
            const uint64_t ex = binScan(((uint64_t)__popc(pending) << 32) | myGroups, total, sScan, phase);
            const uint32_t totChunks = (uint32_t)(total >> 32);
            if (totChunks == 0) break;
            uint32_t ci = (uint32_t)(ex >> 32), pairBase = (uint32_t)ex;
            while (pending && ci < kBinChunkList) {
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
                if (ci == min(totChunks, (uint32_t)kBinChunkList)) cPair[ci] = pairBase;   // sentinel by the last writer
            }
            __syncthreads();
            const uint32_t nChunk = min(totChunks, (uint32_t)kBinChunkList);
            const uint32_t npairs = cPair[nChunk];

            // group boxes of the listed chunks -> the tile's list (four consecutive groups per thread and scan step)
            for (uint32_t pb = 0; pb < npairs && nOut <= cap; pb += 4u * kBinThreads) {
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
                        if (((valid >> k) & 1u) && gb[k].x0 <= gb[k].x1 && gb[k].x0 <= X1 && gb[k].x1 >= X0 && gb[k].y0 <= Y1 && gb[k].y1 >= Y0)
                            hits |= 1u << k;
                }
                uint64_t tot2;
                uint32_t e2 = nOut + (uint32_t)binScan((uint64_t)__popc(hits), tot2, sScan, phase);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((hits >> k) & 1u) {
                        if (e2 < cap) out[e2] = packGroup(grp[k], t.gcnt[grp[k]]);
                        ++e2;
                    }
                nOut += (uint32_t)tot2;
            }
            __syncthreads();                                  // cGroup / cPair are rewritten by the next step
            if (totChunks <= kBinChunkList || nOut > cap) break;
        }
    }
    if (tid == 0) {
        t.groupCount[blockIdx.x] = nOut <= cap ? nOut : kBinOverflow;
        if (t.splitCap > 0) {                                // heavy tiles are shaded quadrant by quadrant (common.h)
            bool heavy = false;
            if (nOut > t.splitThreshold) {
                const uint32_t pos = atomicAdd(&t.heavyList[0], 1u);
                if (pos < (uint32_t)t.splitCap) { t.heavyList[1 + pos] = blockIdx.x; heavy = true; }
            }
            t.heavyFlag[blockIdx.x] = heavy ? 1 : 0;
        }
    }
}

inline void launchBin(const TileArgs &t, int tileShift, cudaStream_t stream)
{
    if (tileShift == 6) binKernel<6><<<t.tilesX * t.tilesY, kBinThreads, 0, stream>>>(t);
    else binKernel<5><<<t.tilesX * t.tilesY, kBinThreads, 0, stream>>>(t);
}

} // namespace detail
} // namespace swr

#endif // __CUDACC__
