// hostpack.cpp -- host side of the narrowed index upload (runtime.cu: streamed host indices).
//
// drawElements takes 32-bit indices in host memory (VertexProcessor.h:90, Benchmark.cpp:99-101), and on a big draw
// their PCIe transfer is what a frame costs end to end.  Indices of a mesh are local: the 4096 consecutive indices
// of a block rarely span more than 65535 vertices.  A slice of the index array whose blocks all do is sent as
// one int32 base per block plus 16-bit offsets -- half the bytes -- and widened again on the device
// (widenIndicesKernel); any other slice goes as it is.  Lossless either way.
//
// Compiled by g++ with -mavx2 (the callers check __builtin_cpu_supports("avx2") first); a small persistent thread
// pool packs the blocks of a slice in parallel, the caller's thread included.
#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace swr {
namespace hostpack {

constexpr size_t kBlock = 4096;          // indices per block (one base each); runtime.cu's widenIndicesKernel agrees

class Pool {
public:
    explicit Pool(int workers)
    {
        for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
    }
    int workers() const { return (int)threads_.size(); }

    // fn(task) for task in [0, n), on the workers and on the calling thread; returns when all are done
    void run(size_t n, const std::function<void(size_t)> &fn)
    {
        if (n == 0) return;
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn;
            total_ = n;
            next_.store(0);
            done_.store(0);
            ++generation_;
        }
        cv_.notify_all();
        work(fn);
        std::unique_lock<std::mutex> l(m_);
        doneCv_.wait(l, [this] { return done_.load() == total_ && active_ == 0; });
        fn_ = nullptr;
    }

private:
    void work(const std::function<void(size_t)> &fn)
    {
        size_t finished = 0;
        for (;;) {
            const size_t t = next_.fetch_add(1);
            if (t >= total_) break;
            fn(t);
            ++finished;
        }
        if (finished && done_.fetch_add(finished) + finished == total_) {
            std::lock_guard<std::mutex> l(m_);
            doneCv_.notify_all();
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(size_t)> *fn;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                fn = fn_;
                if (!fn) continue;               // woke up after that run had already finished without this worker
                ++active_;                       // run() does not return (and *fn stays alive) while a worker is active
            }
            work(*fn);
            {
                std::lock_guard<std::mutex> l(m_);
                --active_;
                doneCv_.notify_all();
            }
        }
    }

    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, doneCv_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t total_ = 0;
    std::atomic<size_t> next_{ 0 }, done_{ 0 };
    uint64_t generation_ = 0;
    int active_ = 0;
    bool stop_ = false;
};

// One block: base = smallest index; false (nothing written that matters) when the block spans more than 16 bits.
static bool packBlock(const int32_t *src, size_t n, uint16_t *dst, int32_t *base)
{
    size_t i = 0;
    __m256i vmin = _mm256_set1_epi32(INT32_MAX), vmax = _mm256_set1_epi32(INT32_MIN);
    for (; i + 8 <= n; i += 8) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        vmin = _mm256_min_epi32(vmin, a);
        vmax = _mm256_max_epi32(vmax, a);
    }
    alignas(32) int32_t lo[8], hi[8];
    _mm256_store_si256(reinterpret_cast<__m256i *>(lo), vmin);
    _mm256_store_si256(reinterpret_cast<__m256i *>(hi), vmax);
    int32_t mn = INT32_MAX, mx = INT32_MIN;
    for (int k = 0; k < 8; ++k) { mn = lo[k] < mn ? lo[k] : mn; mx = hi[k] > mx ? hi[k] : mx; }
    for (; i < n; ++i) { mn = src[i] < mn ? src[i] : mn; mx = src[i] > mx ? src[i] : mx; }
    if ((int64_t)mx - (int64_t)mn > 65535) return false;
    *base = mn;
    const __m256i vb = _mm256_set1_epi32(mn);
    // the packed form is written once and next read by the DMA engine: streaming stores (no read-for-ownership of the
    // destination lines, a quarter of the memory traffic of this loop) when the destination is aligned for them
    const bool stream = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    for (i = 0; i + 16 <= n; i += 16) {
        const __m256i a0 = _mm256_sub_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), vb);
        const __m256i a1 = _mm256_sub_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 8)), vb);
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packus_epi32(a0, a1), 0xD8);
        if (stream) _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), p);
        else _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + i), p);
    }
    for (; i < n; ++i) dst[i] = (uint16_t)(src[i] - mn);
    if (stream) _mm_sfence();
    return true;
}

} // namespace hostpack
} // namespace swr

extern "C" {

// Opaque pool handle; workers < 0: as many as make sense on this machine -- half the hardware threads, at most 8
// threads with the caller's.  Measured on the 16-vCPU host of a B200 (C3 end to end, ms per frame, SWR_PACK_THREADS =
// 6 / 9 / 12 / 16: 3.69 / 3.72 / 3.86 / 3.93): a handful of threads already packs faster than PCIe takes the result,
// more only compete with the DMA engine for host memory bandwidth.
void *swr_hostpack_create(int workers)
{
    if (workers < 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        workers = (int)(hw / 2);
        if (workers > 7) workers = 7;
        if (workers < 1) workers = hw > 1 ? 1 : 0;
        if (const char *env = std::getenv("SWR_PACK_THREADS")) workers = std::atoi(env) > 0 ? std::atoi(env) - 1 : 0;
    }
    return new swr::hostpack::Pool(workers);
}

void swr_hostpack_destroy(void *pool) { delete static_cast<swr::hostpack::Pool *>(pool); }

// Packs `count` indices into 16-bit offsets (dst16[count]) and one base per block of 4096 (bases[ceil(count / 4096)]).
// Returns 1 when every block fitted, 0 when the slice has to be sent as it is (dst16 / bases are then undefined).
int swr_hostpack_pack16(void *pool, const int32_t *src, size_t count, uint16_t *dst16, int32_t *bases)
{
    using namespace swr::hostpack;
    const size_t nblocks = (count + kBlock - 1) / kBlock;
    std::atomic<int> ok{ 1 };
    // a task = 16 blocks (256 KB of source): coarse enough for the task counter, fine enough to balance
    constexpr size_t kPerTask = 16;
    const size_t ntasks = (nblocks + kPerTask - 1) / kPerTask;
    static_cast<Pool *>(pool)->run(ntasks, [&](size_t t) {
        if (!ok.load(std::memory_order_relaxed)) return;
        const size_t b1 = (t + 1) * kPerTask < nblocks ? (t + 1) * kPerTask : nblocks;
        for (size_t b = t * kPerTask; b < b1; ++b) {
            const size_t first = b * kBlock, n = first + kBlock <= count ? kBlock : count - first;
            if (!packBlock(src + first, n, dst16 + first, bases + b)) {
                ok.store(0, std::memory_order_relaxed);
                return;
            }
        }
    });
    return ok.load();
}

} // extern "C"
