// s2g_api.cu — the C ABI of libsphtogrid_cuda.so: context management, host<->device staging, dispatch.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "s2g_common.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void s2g_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* s2g_last_error(void) { return g_err; }
extern "C" const char* s2g_version(void) { return "sphtogrid_cuda 0.1.0 (sm_100a; reference SPHtoGrid.jl v0.5.3)"; }

extern "C" int s2g_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_init(int device, s2g_ctx** out)
{
    S2G_CHECK(out != nullptr, S2G_EINVAL, "s2g_init: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        s2g_set_error("s2g_init: no CUDA device available (%s) - this library has no CPU fallback",
                      e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return S2G_ECUDA;
    }
    S2G_CHECK(device >= 0 && device < ndev, S2G_EINVAL, "s2g_init: device %d out of range [0,%d)", device, ndev);
    S2G_CUDA(cudaSetDevice(device));
    s2g_ctx* ctx = new (std::nothrow) s2g_ctx();
    S2G_CHECK(ctx != nullptr, S2G_ENOMEM, "s2g_init: out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    S2G_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    S2G_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    for (auto& ev : ctx->ev) S2G_CUDA(cudaEventCreate(&ev));
    S2G_CUDA(cudaMalloc(&ctx->d_counters, CNT_N * sizeof(unsigned long long)));
    S2G_CUDA(cudaMallocHost(&ctx->h_counters, CNT_N * sizeof(unsigned long long)));
    S2G_CUDA(cudaMallocHost(&ctx->h_small, 256));
    S2G_CUDA(cudaMemset(ctx->d_counters, 0, CNT_N * sizeof(unsigned long long)));
    *out = ctx;
    return S2G_OK;
}

static void stager_destroy(s2g_ctx* ctx);

extern "C" int s2g_shutdown(s2g_ctx* ctx)
{
    if (!ctx) return S2G_OK;
    cudaSetDevice(ctx->device);
    stager_destroy(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->pool)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& t : ctx->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return S2G_OK;
}

#define CTX_ENTER(ctx)                                                          \
    S2G_CHECK((ctx) != nullptr, S2G_EINVAL, "%s: ctx is NULL", __func__);       \
    S2G_CUDA(cudaSetDevice((ctx)->device))

extern "C" int s2g_sync(s2g_ctx* ctx)
{
    CTX_ENTER(ctx);
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}

extern "C" int s2g_set_stream(s2g_ctx* ctx, void* cuda_stream)
{
    CTX_ENTER(ctx);
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return S2G_OK;
}

extern "C" int s2g_set_strategy(s2g_ctx* ctx, int strategy)
{
    S2G_CHECK(ctx != nullptr, S2G_EINVAL, "s2g_set_strategy: ctx is NULL");
    S2G_CHECK(strategy >= S2G_STRATEGY_AUTO && strategy <= S2G_STRATEGY_GATHER, S2G_EINVAL,
              "s2g_set_strategy: unknown strategy %d", strategy);
    ctx->strategy = strategy;
    return S2G_OK;
}

extern "C" int s2g_set_exact_norm(s2g_ctx* ctx, int on)
{
    S2G_CHECK(ctx != nullptr, S2G_EINVAL, "s2g_set_exact_norm: ctx is NULL");
    ctx->exact_norm = on ? 1 : 0;
    return S2G_OK;
}

extern "C" int s2g_set_accumulate_mode(s2g_ctx* ctx, int mode)
{
    S2G_CHECK(ctx != nullptr, S2G_EINVAL, "s2g_set_accumulate_mode: ctx is NULL");
    S2G_CHECK(mode == S2G_ACCUM_F64 || mode == S2G_ACCUM_F32, S2G_EINVAL,
              "s2g_set_accumulate_mode: mode must be 0 (FP64) or 1 (FP32 partial sums)");
    ctx->accum_f32 = mode == S2G_ACCUM_F32 ? 1 : 0;
    return S2G_OK;
}

int s2g_phase_begin(s2g_ctx* ctx, int phase)
{
    if (ctx->timers_used == ctx->timers.size()) {
        s2g_ctx::timer t;
        t.phase = phase;
        if (cudaEventCreate(&t.a) != cudaSuccess || cudaEventCreate(&t.b) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        ctx->timers.push_back(t);
    }
    const int h = (int)ctx->timers_used++;
    ctx->timers[h].phase = phase;
    cudaEventRecord(ctx->timers[h].a, ctx->stream);
    return h;
}

void s2g_phase_end(s2g_ctx* ctx, int handle)
{
    if (handle < 0) return;
    cudaEventRecord(ctx->timers[handle].b, ctx->stream);
}

int s2g_scratch(s2g_ctx* ctx, const char* name, size_t bytes, void** out)
{
    s2g_buffer& b = ctx->pool[name];
    if (b.bytes < bytes) {
        if (b.ptr) {
            S2G_CUDA(cudaStreamSynchronize(ctx->stream));
            S2G_CUDA(cudaFree(b.ptr));
            b.ptr = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&b.ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&b.ptr, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            b.ptr = nullptr;
            s2g_set_error("device allocation of %zu bytes for '%s' failed: %s", bytes, name, cudaGetErrorString(e));
            return S2G_ENOMEM;
        }
        b.bytes = want;
    }
    *out = b.ptr;
    return S2G_OK;
}

extern "C" int s2g_host_alloc(void** out, uint64_t bytes)
{
    S2G_CHECK(out != nullptr, S2G_EINVAL, "s2g_host_alloc: out is NULL");
    S2G_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return S2G_OK;
}
extern "C" int s2g_host_free(void* p)
{
    if (p) S2G_CUDA(cudaFreeHost(p));
    return S2G_OK;
}
extern "C" int s2g_dev_alloc(s2g_ctx* ctx, void** out, uint64_t bytes)
{
    CTX_ENTER(ctx);
    S2G_CHECK(out != nullptr, S2G_EINVAL, "s2g_dev_alloc: out is NULL");
    S2G_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return S2G_OK;
}
extern "C" int s2g_dev_free(s2g_ctx* ctx, void* p)
{
    CTX_ENTER(ctx);
    if (p) S2G_CUDA(cudaFree(p));
    return S2G_OK;
}
extern "C" int s2g_memcpy_h2d(s2g_ctx* ctx, void* dst, const void* src, uint64_t bytes)
{
    CTX_ENTER(ctx);
    S2G_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}
extern "C" int s2g_memcpy_d2h(s2g_ctx* ctx, void* dst, const void* src, uint64_t bytes)
{
    CTX_ENTER(ctx);
    S2G_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}
extern "C" int s2g_memset_dev(s2g_ctx* ctx, void* dst, int value, uint64_t bytes)
{
    CTX_ENTER(ctx);
    S2G_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------------------
static int stats_begin(s2g_ctx* ctx, long long n_in)
{
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.n_in = n_in;
    ctx->host_pairs = 0;
    ctx->timers_used = 0;
    ctx->launches = 0;
    S2G_CUDA(cudaMemsetAsync(ctx->d_counters, 0, CNT_N * sizeof(unsigned long long), ctx->stream));
    return S2G_OK;
}

// copies the device counters back (synchronises the stream)
int s2g_stats_collect(s2g_ctx* ctx);
static int stager_finish(s2g_ctx* ctx);
static int stats_collect(s2g_ctx* ctx)
{
    S2G_TRY(stager_finish(ctx));   // the helper thread of an overlapped staging has issued all its copies
    S2G_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, CNT_N * sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.n_mapped = (int64_t)ctx->h_counters[CNT_MAPPED];
    ctx->stats.footprint_pixels = (int64_t)ctx->h_counters[CNT_FOOTPRINT];
    ctx->stats.touched_pixels = (int64_t)ctx->h_counters[CNT_TOUCHED];
    ctx->stats.n_fallback = (int64_t)ctx->h_counters[CNT_FALLBACK];
    ctx->stats.n_pairs = (int64_t)ctx->host_pairs;
    ctx->stats.n_scatter = (int64_t)ctx->h_counters[CNT_SCATTER];
    ctx->stats.n_gather = (int64_t)ctx->h_counters[CNT_GATHER];
    double ph[PH_N] = {0, 0, 0, 0, 0};
    for (size_t i = 0; i < ctx->timers_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->timers[i].a, ctx->timers[i].b) == cudaSuccess)
            ph[ctx->timers[i].phase] += ms;
        else
            cudaGetLastError();
    }
    ctx->stats.ms_prep = ph[PH_PREP];
    ctx->stats.ms_sort = ph[PH_SORT];
    ctx->stats.ms_norm = ph[PH_NORM];
    if (ph[PH_DEPOSIT] > 0) ctx->stats.ms_deposit = ph[PH_DEPOSIT];
    if (ph[PH_EPILOGUE] > 0) ctx->stats.ms_epilogue = ph[PH_EPILOGUE];
    ctx->stats.n_launches = ctx->launches;
    return S2G_OK;
}

int s2g_stats_collect(s2g_ctx* ctx) { return stats_collect(ctx); }
int s2g_stats_begin(s2g_ctx* ctx, long long n_in) { return stats_begin(ctx, n_in); }

static float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

extern "C" int s2g_get_stats(s2g_ctx* ctx, s2g_stats* out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(out != nullptr, S2G_EINVAL, "s2g_get_stats: out is NULL");
    S2G_TRY(stats_collect(ctx));
    *out = ctx->stats;
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// argument helpers
// ------------------------------------------------------------------------------------------------
static size_t esize(int dtype) { return dtype == S2G_F64 ? 8 : 4; }

static int check_common(const char* fn, const void* pos, const void* hsml, const void* m, const void* rho,
                        const void* binq, const void* w, int64_t n, int in_dtype, double len2pix, int64_t npix,
                        int kernel)
{
    S2G_CHECK(n >= 0, S2G_EINVAL, "%s: n < 0", fn);
    S2G_CHECK(n < 2147483647LL, S2G_EINVAL, "%s: n = %lld exceeds the per-call limit 2^31-1 (batch the particles)", fn,
              (long long)n);
    S2G_CHECK(n == 0 || (pos && hsml && m && rho && binq && w), S2G_EINVAL, "%s: NULL particle array", fn);
    S2G_CHECK(in_dtype == S2G_F32 || in_dtype == S2G_F64, S2G_EINVAL, "%s: in_dtype must be 0 (f32) or 1 (f64)", fn);
    S2G_CHECK(npix > 0 && npix <= 1000000, S2G_EINVAL, "%s: npix out of range", fn);
    S2G_CHECK(len2pix == len2pix, S2G_EINVAL, "%s: len2pix is NaN", fn);
    S2G_CHECK(kernel >= S2G_KERNEL_CUBIC && kernel <= S2G_KERNEL_WENDLAND_C8, S2G_EINVAL, "%s: unknown kernel id %d",
              fn, kernel);
    return S2G_OK;
}

static s2g_geom make_geom(double len2pix, int64_t npix, int n_images, int calc_mean)
{
    s2g_geom G;
    G.len2pix = len2pix;
    volatile double l2 = len2pix * len2pix;  // host arithmetic, individually rounded (no contraction possible)
    volatile double l3 = l2 * len2pix;
    G.l3 = l3;
    G.inv_l3 = 1.0 / l3;
    G.half_n = 0.5 * (double)npix;
    G.npix = npix;
    G.n_images = n_images;
    G.calc_mean = calc_mean;
    return G;
}

// ------------------------------------------------------------------------------------------------
// overlapped staging (see s2g_common.cuh): the Julia caller's arrays are pageable, and a pageable cudaMemcpyAsync
// blocks its host thread while the driver bounces the data through its own pinned buffers (~11 GB/s measured: 98 ms
// for the 1.07 GB of BASELINE config 2).  A helper thread does those copies on a non-blocking stream while the calling
// thread already runs the deposit of the particles that have arrived.
// ------------------------------------------------------------------------------------------------
#include <condition_variable>
#include <mutex>
#include <thread>

struct s2g_stager {
    // K helper threads (S2G_STAGE_THREADS, default 4): thread k stages the chunks k, k + K, ... on its own copy stream
    // through its own pinned bounce buffers.  One thread's memcpy() into the bounce buffer runs at ~10 GB/s — a third
    // of what the PCIe link takes — so the threads fill the link together.
    static constexpr int MAXT = 8;
    static constexpr int NBOUNCE = 3;
    static constexpr size_t BOUNCE_BYTES = 4u << 20;
    int nthreads = 0;
    std::thread th[MAXT];
    std::mutex mu;
    std::condition_variable cv;
    cudaStream_t copy_stream[MAXT] = {};
    cudaEvent_t start_ev = nullptr;
    // own pinned bounce buffers: a helper thread memcpy()s a piece of the caller's (pageable) array into one of them
    // and issues a truly asynchronous copy from there.  A pageable cudaMemcpyAsync instead holds the driver while it
    // stages the data, which delayed the kernel launches of the calling thread (measured: +45 ms per C2 map).
    void* bounce[MAXT][NBOUNCE] = {};
    cudaEvent_t bounce_ev[MAXT][NBOUNCE] = {};
    std::vector<cudaEvent_t> ev;   // one per chunk, recorded on the stream of the thread that staged it
    std::vector<char> done;        // chunk's event has been recorded
    long long chunk = 1 << 20;     // particles per chunk
    long long n = 0;
    int recorded = 0;              // length of the prefix of chunks whose events have been recorded
    int error = 0;                 // first cudaError_t of a helper thread
    bool active = false;
};

static void stager_join(s2g_ctx* ctx)
{
    s2g_stager* s = ctx->stager;
    if (!s) return;
    for (int k = 0; k < s2g_stager::MAXT; ++k)
        if (s->th[k].joinable()) s->th[k].join();
    s->active = false;
}

// joins the helper thread and reports its error, if any
static int stager_finish(s2g_ctx* ctx)
{
    s2g_stager* s = ctx->stager;
    if (!s || !s->active) return S2G_OK;
    stager_join(ctx);
    if (s->error != 0) {
        s2g_set_error("host->device staging failed: %s", cudaGetErrorString((cudaError_t)s->error));
        return S2G_ECUDA;
    }
    return S2G_OK;
}

// joins the helper thread on every exit path of an entry point (the caller's arrays must outlive the copies)
struct StageGuard {
    s2g_ctx* c;
    ~StageGuard() { stager_join(c); }
};

static void stager_destroy(s2g_ctx* ctx)
{
    s2g_stager* s = ctx->stager;
    if (!s) return;
    stager_join(ctx);
    for (auto e : s->ev) cudaEventDestroy(e);
    if (s->start_ev) cudaEventDestroy(s->start_ev);
    for (int k = 0; k < s2g_stager::MAXT; ++k) {
        for (int b = 0; b < s2g_stager::NBOUNCE; ++b) {
            if (s->bounce[k][b]) cudaFreeHost(s->bounce[k][b]);
            if (s->bounce_ev[k][b]) cudaEventDestroy(s->bounce_ev[k][b]);
        }
        if (s->copy_stream[k]) cudaStreamDestroy(s->copy_stream[k]);
    }
    delete s;
    ctx->stager = nullptr;
}

int s2g_stage_wait(s2g_ctx* ctx, long long upto)
{
    s2g_stager* s = ctx->stager;
    if (!s || !s->active || s->n <= 0) return S2G_OK;
    if (upto > s->n) upto = s->n;
    if (upto <= 0) return S2G_OK;
    const int c = (int)((upto - 1) / s->chunk);
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->recorded > c || s->error != 0; });
        if (s->error != 0) {
            s2g_set_error("host->device staging failed: %s", cudaGetErrorString((cudaError_t)s->error));
            return S2G_ECUDA;
        }
    }
    // the last chunk <= c of every helper thread (a thread's stream is in order: it covers the thread's earlier chunks)
    for (int k = 0; k < s->nthreads; ++k) {
        const int ck = c - ((c - k) % s->nthreads + s->nthreads) % s->nthreads;
        if (ck >= 0) S2G_CUDA(cudaStreamWaitEvent(ctx->stream, s->ev[(size_t)ck], 0));
    }
    return S2G_OK;
}

long long s2g_stage_first_slice(const s2g_ctx* ctx)
{
    return (ctx->stager && ctx->stager->active) ? ctx->stager->chunk : 0;
}

// Result map back to the caller's (pageable) array.  A pageable device->host cudaMemcpyAsync goes through the driver's
// staging at ~11 GB/s (the 1.07-GB grid of BASELINE config 3: ~95 ms); here the helper threads of the stager copy
// 4-MB pieces into their pinned bounce buffers on their own streams (up to three in flight per thread) and memcpy()
// them out, several at a time.  Used for results of at least S2G_UNSTAGE_MIN bytes (default 32 MB) when the context
// has a stager with bounce buffers (i.e. a large input came in through it); returns when `dst` is complete.
static int unstage_output(s2g_ctx* ctx, void* dst, const void* src_dev, size_t bytes)
{
    size_t min_bytes = (size_t)32 << 20;
    if (const char* e = getenv("S2G_UNSTAGE_MIN")) min_bytes = (size_t)atoll(e);
    s2g_stager* s = ctx->stager;
    bool plain = !s || s->nthreads <= 0 || !s->bounce[0][0] || bytes < min_bytes || bytes == 0;
    if (!plain) {   // a pinned (or registered) destination takes the direct copy at link speed
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, dst) == cudaSuccess) plain = at.type != cudaMemoryTypeUnregistered;
        else cudaGetLastError();
    }
    if (plain) {
        if (bytes > 0) S2G_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return S2G_OK;
    }
    S2G_TRY(stager_finish(ctx));   // the input copies are long done; the threads and buffers are free
    S2G_CUDA(cudaEventRecord(s->start_ev, ctx->stream));
    for (int k = 0; k < s->nthreads; ++k) S2G_CUDA(cudaStreamWaitEvent(s->copy_stream[k], s->start_ev, 0));
    const size_t piece = s2g_stager::BOUNCE_BYTES;
    const size_t npieces = (bytes + piece - 1) / piece;
    const int K = s->nthreads, device = ctx->device;
    int errs[s2g_stager::MAXT] = {};
    int started = 0;
    try {
        for (int k = 0; k < K; ++k) {
            s->th[k] = std::thread([=, &errs]() {
                cudaSetDevice(device);
                cudaStream_t cs = s->copy_stream[k];
                size_t off_of[s2g_stager::NBOUNCE] = {}, len_of[s2g_stager::NBOUNCE] = {};
                bool used[s2g_stager::NBOUNCE] = {};
                cudaError_t e = cudaSuccess;
                unsigned long long turn = 0;
                auto drain = [&](int b) {
                    if (!used[b] || e != cudaSuccess) return;
                    e = cudaEventSynchronize(s->bounce_ev[k][b]);
                    if (e == cudaSuccess) memcpy((char*)dst + off_of[b], s->bounce[k][b], len_of[b]);
                    used[b] = false;
                };
                for (size_t i = (size_t)k; i < npieces && e == cudaSuccess; i += (size_t)K) {
                    const int b = (int)(turn++ % s2g_stager::NBOUNCE);
                    drain(b);
                    if (e != cudaSuccess) break;
                    const size_t off = i * piece, len = std::min(piece, bytes - off);
                    e = cudaMemcpyAsync(s->bounce[k][b], (const char*)src_dev + off, len, cudaMemcpyDeviceToHost, cs);
                    if (e == cudaSuccess) e = cudaEventRecord(s->bounce_ev[k][b], cs);
                    off_of[b] = off; len_of[b] = len; used[b] = (e == cudaSuccess);
                }
                for (int j = 0; j < s2g_stager::NBOUNCE; ++j) drain((int)((turn + j) % s2g_stager::NBOUNCE));   // oldest first
                errs[k] = (int)e;
            });
            ++started;
        }
    } catch (...) {
    }
    for (int k = 0; k < s2g_stager::MAXT; ++k)
        if (s->th[k].joinable()) s->th[k].join();
    if (started < K) {   // not every piece had a thread: redo the whole copy the plain way
        S2G_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return S2G_OK;
    }
    for (int k = 0; k < K; ++k)
        if (errs[k] != 0) {
            s2g_set_error("device->host copy of the result failed: %s", cudaGetErrorString((cudaError_t)errs[k]));
            return S2G_ECUDA;
        }
    return S2G_OK;
}

// copies the six host arrays into scratch device buffers
static int stage_particles(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                           const void* binq, const void* w, int64_t n, int n_images, int in_dtype, s2g_particles& P)
{
    const size_t es = esize(in_dtype);
    void *dpos, *dh, *dm, *dr, *dq, *dw;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    S2G_TRY(s2g_scratch(ctx, "in_pos", 3 * nn * es, &dpos));
    S2G_TRY(s2g_scratch(ctx, "in_hsml", nn * es, &dh));
    S2G_TRY(s2g_scratch(ctx, "in_m", nn * es, &dm));
    S2G_TRY(s2g_scratch(ctx, "in_rho", nn * es, &dr));
    S2G_TRY(s2g_scratch(ctx, "in_q", nn * es * (size_t)n_images, &dq));
    S2G_TRY(s2g_scratch(ctx, "in_w", nn * es, &dw));
    const bool overlap_off = getenv("S2G_STAGE_OVERLAP") && atoi(getenv("S2G_STAGE_OVERLAP")) == 0;
    bool threaded = false;
    long long stage_min = 2 << 20;   // particles from which the helper thread pays off (S2G_STAGE_MIN: tests lower it)
    if (const char* e = getenv("S2G_STAGE_MIN")) stage_min = atoll(e);
    if (n > stage_min && !overlap_off) {
        // helper threads: chunks of 1 Mi particles, all six arrays of a chunk, then the chunk's event
        if (!ctx->stager) {
            ctx->stager = new s2g_stager();
            s2g_stager* s0 = ctx->stager;
            if (const char* e = getenv("S2G_STAGE_CHUNK")) {
                const long long c = atoll(e);
                if (c >= 1024) s0->chunk = c;
            }
            int nt = 4;
            if (const char* e = getenv("S2G_STAGE_THREADS")) nt = atoi(e);
            s0->nthreads = std::min(std::max(nt, 1), (int)s2g_stager::MAXT);
            S2G_CUDA(cudaEventCreateWithFlags(&s0->start_ev, cudaEventDisableTiming));
            const bool bounce_off = getenv("S2G_STAGE_BOUNCE") && atoi(getenv("S2G_STAGE_BOUNCE")) == 0;
            for (int k = 0; k < s0->nthreads; ++k) {
                S2G_CUDA(cudaStreamCreateWithFlags(&s0->copy_stream[k], cudaStreamNonBlocking));
                for (int b = 0; b < s2g_stager::NBOUNCE && !bounce_off; ++b) {
                    S2G_CUDA(cudaMallocHost(&s0->bounce[k][b], s2g_stager::BOUNCE_BYTES));
                    S2G_CUDA(cudaEventCreateWithFlags(&s0->bounce_ev[k][b], cudaEventDisableTiming));
                }
            }
        }
        s2g_stager* s = ctx->stager;
        stager_join(ctx);
        const int nchunks = (int)((n + s->chunk - 1) / s->chunk);
        while ((int)s->ev.size() < nchunks) {
            cudaEvent_t e;
            S2G_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->ev.push_back(e);
        }
        s->done.assign((size_t)nchunks, 0);
        // the device buffers may still be read by work of the previous call on ctx->stream
        S2G_CUDA(cudaEventRecord(s->start_ev, ctx->stream));
        for (int k = 0; k < s->nthreads; ++k) S2G_CUDA(cudaStreamWaitEvent(s->copy_stream[k], s->start_ev, 0));
        s->n = n; s->recorded = 0; s->error = 0; s->active = true;
        const int device = ctx->device;
        const size_t nim = (size_t)n_images;
        const int K = s->nthreads;
        int started = 0;
        try {
        for (int k = 0; k < K; ++k) {
        s->th[k] = std::thread([=]() {
            cudaSetDevice(device);
            cudaStream_t cs = s->copy_stream[k];
            unsigned long long bounce_turn = 0;
            bool bounce_used[s2g_stager::NBOUNCE] = {false, false, false};
            for (int c = k; c < nchunks; c += K) {
                const size_t o = (size_t)c * (size_t)s->chunk;
                const size_t cnt = std::min<size_t>((size_t)s->chunk, (size_t)n - o);
                cudaError_t e = cudaSuccess;
                auto cp = [&](void* d, const void* h, size_t per) {
                    char* dst = (char*)d + o * per * es;
                    const char* src = (const char*)h + o * per * es;
                    size_t left = cnt * per * es;
                    if (!s->bounce[k][0]) {   // no bounce buffers: the driver's pageable path
                        if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src, left, cudaMemcpyHostToDevice, cs);
                        return;
                    }
                    while (left > 0 && e == cudaSuccess) {
                        const int b = (int)(bounce_turn++ % s2g_stager::NBOUNCE);
                        const size_t piece = std::min(left, s2g_stager::BOUNCE_BYTES);
                        if (bounce_used[b]) e = cudaEventSynchronize(s->bounce_ev[k][b]);   // its previous DMA has read it
                        if (e != cudaSuccess) break;
                        memcpy(s->bounce[k][b], src, piece);
                        e = cudaMemcpyAsync(dst, s->bounce[k][b], piece, cudaMemcpyHostToDevice, cs);
                        if (e == cudaSuccess) e = cudaEventRecord(s->bounce_ev[k][b], cs);
                        bounce_used[b] = true;
                        dst += piece; src += piece; left -= piece;
                    }
                };
                cp(dpos, pos, 3); cp(dh, hsml, 1); cp(dm, m, 1); cp(dr, rho, 1); cp(dq, binq, nim); cp(dw, w, 1);
                if (e == cudaSuccess) e = cudaEventRecord(s->ev[(size_t)c], cs);
                bool stop = false;
                {
                    std::lock_guard<std::mutex> lk(s->mu);
                    if (e != cudaSuccess && s->error == 0) s->error = (int)e;
                    s->done[(size_t)c] = 1;
                    while (s->recorded < nchunks && s->done[(size_t)s->recorded]) ++s->recorded;
                    stop = s->error != 0;
                }
                s->cv.notify_all();
                if (stop) break;
            }
        });
        ++started;
        }
        threaded = true;
        } catch (...) {   // no (further) thread to be had
            if (started == 0)
                s->active = false;   // copy on the calling thread like a small input
            else {
                // some threads run: let them finish, then fall back to the calling thread for everything
                stager_join(ctx);
            }
        }
    }
    if (!threaded && n > 0) {
        if (ctx->stager) stager_join(ctx);
        S2G_CUDA(cudaMemcpyAsync(dpos, pos, 3 * (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
        S2G_CUDA(cudaMemcpyAsync(dh, hsml, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
        S2G_CUDA(cudaMemcpyAsync(dm, m, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
        S2G_CUDA(cudaMemcpyAsync(dr, rho, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
        S2G_CUDA(cudaMemcpyAsync(dq, binq, (size_t)n * es * (size_t)n_images, cudaMemcpyHostToDevice, ctx->stream));
        S2G_CUDA(cudaMemcpyAsync(dw, w, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    }
    memset(&P, 0, sizeof(P));
    P.pos = dpos; P.hsml = dh; P.m = dm; P.rho = dr; P.binq = dq; P.w = dw;
    P.n = n;
    P.in_dtype = in_dtype;
    return S2G_OK;
}

int s2g_stage_particles(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                        const void* binq, const void* w, int64_t n, int n_images, int in_dtype, s2g_particles& P)
{
    return stage_particles(ctx, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, P);
}

static s2g_particles dev_particles(const void* pos, const void* hsml, const void* m, const void* rho, const void* binq,
                                   const void* w, int64_t n, int in_dtype)
{
    s2g_particles P;
    memset(&P, 0, sizeof(P));
    P.pos = pos; P.hsml = hsml; P.m = m; P.rho = rho; P.binq = binq; P.w = w;
    P.n = n;
    P.in_dtype = in_dtype;
    return P;
}

static void set_center(s2g_particles& P, const double shift[3], int periodic, double boxsize, const double halfsize[3])
{
    P.fuse_center = 1;
    P.periodic = periodic;
    P.boxsize = boxsize;
    for (int d = 0; d < 3; ++d) {
        P.shift[d] = shift[d];
        P.halfsize[d] = halfsize[d];
    }
}

// ------------------------------------------------------------------------------------------------
// 2D / 3D deposit
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_deposit_2d_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                  const void* binq, const void* w, int64_t n, int32_t n_images, int32_t in_dtype,
                                  double len2pix, int64_t nx, int64_t ny, int32_t kernel, int32_t calc_mean,
                                  int32_t accumulate, double* image_dev)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, nx, kernel));
    S2G_CHECK(nx == ny, S2G_EINVAL, "%s: nx must equal ny (the reference always builds square maps, parameters.jl:105)",
              __func__);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", __func__);
    S2G_CHECK(image_dev != nullptr, S2G_EINVAL, "%s: image is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    if (!accumulate)
        S2G_CUDA(cudaMemsetAsync(image_dev, 0, sizeof(double) * (size_t)(nx * ny) * (size_t)(n_images + 1), ctx->stream));
    s2g_particles P = dev_particles(pos, hsml, m, rho, binq, w, n, in_dtype);
    s2g_geom G = make_geom(len2pix, nx, n_images, calc_mean);
    return s2g_launch_deposit_2d(ctx, P, G, kernel, image_dev);
}

extern "C" int s2g_deposit_2d(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                              const void* binq, const void* w, int64_t n, int32_t n_images, int32_t in_dtype,
                              double len2pix, int64_t nx, int64_t ny, int32_t kernel, int32_t calc_mean,
                              double* image_out, s2g_stats* stats)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, nx, kernel));
    S2G_CHECK(nx == ny, S2G_EINVAL, "%s: nx must equal ny", __func__);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", __func__);
    S2G_CHECK(image_out != nullptr, S2G_EINVAL, "%s: image_out is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t img_bytes = sizeof(double) * (size_t)(nx * ny) * (size_t)(n_images + 1);
    void* dimg;
    S2G_TRY(s2g_scratch(ctx, "image", img_bytes, &dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    s2g_particles P;
    S2G_TRY(stage_particles(ctx, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, P));
    StageGuard stage_guard{ctx};
    S2G_CUDA(cudaMemsetAsync(dimg, 0, img_bytes, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    s2g_geom G = make_geom(len2pix, nx, n_images, calc_mean);
    S2G_TRY(s2g_launch_deposit_2d(ctx, P, G, kernel, (double*)dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(image_out, dimg, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

// cic_mapping_2D with RM (cic_2D.jl:201-217): stokes == 0 -> RM is inert (faraday_rotate_pixel! does nothing), the
// ordinary deposit gives the reference result; stokes != 0 -> ordered compositing (s2g_stokes2d.cu)
extern "C" int s2g_deposit_2d_rm_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                     const void* binq, const void* w, const double* rm, int64_t n, int32_t n_images,
                                     int32_t in_dtype, double len2pix, int64_t nx, int64_t ny, int32_t kernel,
                                     int32_t calc_mean, int32_t stokes, double* image_dev)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, nx, kernel));
    S2G_CHECK(nx == ny, S2G_EINVAL, "%s: nx must equal ny", __func__);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", __func__);
    S2G_CHECK(image_dev != nullptr, S2G_EINVAL, "%s: image is NULL", __func__);
    S2G_CHECK(rm != nullptr || n == 0, S2G_EINVAL, "%s: rm is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    S2G_CUDA(cudaMemsetAsync(image_dev, 0, sizeof(double) * (size_t)(nx * ny) * (size_t)(n_images + 1), ctx->stream));
    s2g_particles P = dev_particles(pos, hsml, m, rho, binq, w, n, in_dtype);
    s2g_geom G = make_geom(len2pix, nx, n_images, calc_mean);
    if (!stokes) return s2g_launch_deposit_2d(ctx, P, G, kernel, image_dev);
    return s2g_launch_stokes_2d(ctx, P, G, kernel, rm, image_dev);
}

extern "C" int s2g_deposit_2d_rm(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                 const void* binq, const void* w, const double* rm, int64_t n, int32_t n_images,
                                 int32_t in_dtype, double len2pix, int64_t nx, int64_t ny, int32_t kernel,
                                 int32_t calc_mean, int32_t stokes, double* image_out, s2g_stats* stats)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, nx, kernel));
    S2G_CHECK(nx == ny, S2G_EINVAL, "%s: nx must equal ny", __func__);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", __func__);
    S2G_CHECK(image_out != nullptr, S2G_EINVAL, "%s: image_out is NULL", __func__);
    S2G_CHECK(rm != nullptr || n == 0, S2G_EINVAL, "%s: rm is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t img_bytes = sizeof(double) * (size_t)(nx * ny) * (size_t)(n_images + 1);
    void *dimg, *drm;
    S2G_TRY(s2g_scratch(ctx, "image", img_bytes, &dimg));
    S2G_TRY(s2g_scratch(ctx, "in_rm", sizeof(double) * (size_t)(n > 0 ? n : 1), &drm));
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    s2g_particles P;
    S2G_TRY(stage_particles(ctx, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, P));
    StageGuard stage_guard{ctx};
    if (n > 0) S2G_CUDA(cudaMemcpyAsync(drm, rm, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    S2G_CUDA(cudaMemsetAsync(dimg, 0, img_bytes, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    s2g_geom G = make_geom(len2pix, nx, n_images, calc_mean);
    if (!stokes)
        S2G_TRY(s2g_launch_deposit_2d(ctx, P, G, kernel, (double*)dimg));
    else
        S2G_TRY(s2g_launch_stokes_2d(ctx, P, G, kernel, (const double*)drm, (double*)dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(image_out, dimg, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

extern "C" int s2g_deposit_3d_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                  const void* binq, const void* w, int64_t n, int32_t in_dtype, double len2pix,
                                  int64_t npix, int32_t kernel, int32_t calc_mean, int32_t accumulate,
                                  double* image_dev)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, npix, kernel));
    S2G_CHECK(npix <= 2048, S2G_EINVAL, "%s: npix > 2048 not supported in 3D", __func__);
    S2G_CHECK(image_dev != nullptr, S2G_EINVAL, "%s: image is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    if (!accumulate)
        S2G_CUDA(cudaMemsetAsync(image_dev, 0, sizeof(double) * 2 * (size_t)(npix * npix * npix), ctx->stream));
    s2g_particles P = dev_particles(pos, hsml, m, rho, binq, w, n, in_dtype);
    s2g_geom G = make_geom(len2pix, npix, 1, calc_mean);
    return s2g_launch_deposit_3d(ctx, P, G, kernel, image_dev);
}

extern "C" int s2g_deposit_3d(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                              const void* binq, const void* w, int64_t n, int32_t in_dtype, double len2pix,
                              int64_t npix, int32_t kernel, int32_t calc_mean, double* image_out, s2g_stats* stats)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, npix, kernel));
    S2G_CHECK(npix <= 2048, S2G_EINVAL, "%s: npix > 2048 not supported in 3D", __func__);
    S2G_CHECK(image_out != nullptr, S2G_EINVAL, "%s: image_out is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t img_bytes = sizeof(double) * 2 * (size_t)(npix * npix * npix);
    void* dimg;
    S2G_TRY(s2g_scratch(ctx, "image", img_bytes, &dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    s2g_particles P;
    S2G_TRY(stage_particles(ctx, pos, hsml, m, rho, binq, w, n, 1, in_dtype, P));
    StageGuard stage_guard{ctx};
    S2G_CUDA(cudaMemsetAsync(dimg, 0, img_bytes, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    s2g_geom G = make_geom(len2pix, npix, 1, calc_mean);
    S2G_TRY(s2g_launch_deposit_3d(ctx, P, G, kernel, (double*)dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(image_out, dimg, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// footprints
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_footprints(s2g_ctx* ctx, const void* pos, const void* hsml, int64_t n, int32_t in_dtype,
                              double len2pix, int64_t npix, int32_t dims, int64_t* bounds_out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(dims == 2 || dims == 3, S2G_EINVAL, "%s: dims must be 2 or 3", __func__);
    S2G_CHECK(n >= 0 && (n == 0 || (pos && hsml && bounds_out)), S2G_EINVAL, "%s: bad arguments", __func__);
    S2G_CHECK(in_dtype == S2G_F32 || in_dtype == S2G_F64, S2G_EINVAL, "%s: bad in_dtype", __func__);
    S2G_CHECK(npix > 0 && npix <= 1000000, S2G_EINVAL, "%s: npix out of range", __func__);
    if (n == 0) return S2G_OK;
    const size_t es = esize(in_dtype);
    void *dpos, *dh, *dout;
    S2G_TRY(s2g_scratch(ctx, "in_pos", 3 * (size_t)n * es, &dpos));
    S2G_TRY(s2g_scratch(ctx, "in_hsml", (size_t)n * es, &dh));
    S2G_TRY(s2g_scratch(ctx, "bounds", (size_t)n * 2 * dims * sizeof(long long), &dout));
    S2G_CUDA(cudaMemcpyAsync(dpos, pos, 3 * (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(dh, hsml, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    s2g_particles P = dev_particles(dpos, dh, nullptr, nullptr, nullptr, nullptr, n, in_dtype);
    s2g_geom G = make_geom(len2pix, npix, 1, 1);
    S2G_TRY(s2g_launch_footprints(ctx, P, G, dims, (long long*)dout));
    S2G_CUDA(cudaMemcpyAsync(bounds_out, dout, (size_t)n * 2 * dims * sizeof(long long), cudaMemcpyDeviceToHost,
                             ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// reduce_image
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_reduce_image_2d_dev(s2g_ctx* ctx, const double* image_dev, int64_t nx, int64_t ny, int32_t n_images,
                                       int32_t reduce_image, double* out_dev)
{
    CTX_ENTER(ctx);
    S2G_CHECK(image_dev && out_dev && nx > 0 && nx == ny && n_images >= 1, S2G_EINVAL, "%s: bad arguments", __func__);
    return s2g_launch_reduce_2d(ctx, image_dev, nx, ny, n_images, reduce_image, out_dev);
}

extern "C" int s2g_reduce_image_2d(s2g_ctx* ctx, const double* image, int64_t nx, int64_t ny, int32_t n_images,
                                   int32_t reduce_image, double* out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(image && out && nx > 0 && nx == ny && n_images >= 1, S2G_EINVAL, "%s: bad arguments", __func__);
    const size_t npl = (size_t)(nx * ny);
    void *din, *dout;
    S2G_TRY(s2g_scratch(ctx, "image", sizeof(double) * npl * (size_t)(n_images + 1), &din));
    S2G_TRY(s2g_scratch(ctx, "reduced", sizeof(double) * npl * (size_t)n_images, &dout));
    S2G_CUDA(cudaMemcpyAsync(din, image, sizeof(double) * npl * (size_t)(n_images + 1), cudaMemcpyHostToDevice,
                             ctx->stream));
    S2G_TRY(s2g_launch_reduce_2d(ctx, (const double*)din, nx, ny, n_images, reduce_image, (double*)dout));
    S2G_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * npl * (size_t)n_images, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}

extern "C" int s2g_reduce_image_3d_dev(s2g_ctx* ctx, const double* image_dev, int64_t npix, int32_t reduce_image,
                                       double* out_dev)
{
    CTX_ENTER(ctx);
    S2G_CHECK(image_dev && out_dev && npix > 0, S2G_EINVAL, "%s: bad arguments", __func__);
    return s2g_launch_reduce_3d(ctx, image_dev, npix, reduce_image, out_dev);
}

extern "C" int s2g_reduce_image_3d(s2g_ctx* ctx, const double* image, int64_t npix, int32_t reduce_image, double* out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(image && out && npix > 0, S2G_EINVAL, "%s: bad arguments", __func__);
    const size_t nc = (size_t)(npix * npix * npix);
    void *din, *dout;
    S2G_TRY(s2g_scratch(ctx, "image", sizeof(double) * nc * 2, &din));
    S2G_TRY(s2g_scratch(ctx, "reduced", sizeof(double) * nc, &dout));
    S2G_CUDA(cudaMemcpyAsync(din, image, sizeof(double) * nc * 2, cudaMemcpyHostToDevice, ctx->stream));
    S2G_TRY(s2g_launch_reduce_3d(ctx, (const double*)din, npix, reduce_image, (double*)dout));
    S2G_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * nc, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// centre + filter
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_center_filter(s2g_ctx* ctx, const void* pos, int64_t n, int32_t in_dtype, const double shift[3],
                                 int32_t periodic, double boxsize, const double filter_center[3],
                                 const double filter_halfsize[3], void* pos_out, uint8_t* mask_out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(n >= 0 && (n == 0 || pos) && shift && filter_halfsize, S2G_EINVAL, "%s: bad arguments", __func__);
    S2G_CHECK(in_dtype == S2G_F32 || in_dtype == S2G_F64, S2G_EINVAL, "%s: bad in_dtype", __func__);
    S2G_CHECK(!filter_center || (filter_center[0] == 0.0 && filter_center[1] == 0.0 && filter_center[2] == 0.0),
              S2G_EINVAL, "%s: the recentred box is centred on 0 (filter_shift.jl:29)", __func__);
    if (n == 0) return S2G_OK;
    const size_t es = esize(in_dtype);
    void *dpos, *dout = nullptr, *dmask = nullptr;
    S2G_TRY(s2g_scratch(ctx, "in_pos", 3 * (size_t)n * es, &dpos));
    if (pos_out) S2G_TRY(s2g_scratch(ctx, "pos_out", 3 * (size_t)n * es, &dout));
    if (mask_out) S2G_TRY(s2g_scratch(ctx, "mask", (size_t)n, &dmask));
    S2G_CUDA(cudaMemcpyAsync(dpos, pos, 3 * (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    s2g_particles P = dev_particles(dpos, nullptr, nullptr, nullptr, nullptr, nullptr, n, in_dtype);
    set_center(P, shift, periodic, boxsize, filter_halfsize);
    S2G_TRY(s2g_launch_center_filter(ctx, P, dout, (uint8_t*)dmask));
    if (pos_out) S2G_CUDA(cudaMemcpyAsync(pos_out, dout, 3 * (size_t)n * es, cudaMemcpyDeviceToHost, ctx->stream));
    if (mask_out) S2G_CUDA(cudaMemcpyAsync(mask_out, dmask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// fused sphMapping body
// ------------------------------------------------------------------------------------------------
// projection of map_it (cic_interpolation.jl:331-345) fused into the position load: an axis permutation
// (rotate_to_xz_plane! / rotate_to_yz_plane!, rotate_particles.jl:35-73) or a 3x3 matrix (rotate_3D, :7-13)
static int set_projection(const char* fn, s2g_particles& P, const int32_t* perm, const double* rot)
{
    S2G_CHECK(!(perm && rot), S2G_EINVAL, "%s: give either an axis permutation or a rotation matrix, not both", fn);
    if (perm) {
        int seen = 0;
        for (int d = 0; d < 3; ++d) {
            S2G_CHECK(perm[d] >= 0 && perm[d] <= 2, S2G_EINVAL, "%s: perm[%d] = %d is not an axis", fn, d, perm[d]);
            seen |= 1 << perm[d];
            P.perm[d] = perm[d];
        }
        S2G_CHECK(seen == 7, S2G_EINVAL, "%s: perm is not a permutation of (0,1,2)", fn);
        P.proj = (perm[0] == 0 && perm[1] == 1 && perm[2] == 2) ? 0 : 1;
    } else if (rot) {
        for (int k = 0; k < 9; ++k) P.rot[k] = rot[k];
        P.proj = 2;
    }
    return S2G_OK;
}

static int sphmap_dev_impl(const char* fn, s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                           const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                           int32_t in_dtype, const int32_t* perm, const double* rot, const double shift[3],
                           int32_t periodic, double boxsize, const double halfsize[3], double len2pix, int64_t npix,
                           int32_t kernel, int32_t calc_mean, int32_t accumulate, double* image_dev)
{
    S2G_CHECK(ctx != nullptr, S2G_EINVAL, "%s: ctx is NULL", fn);
    S2G_CUDA(cudaSetDevice(ctx->device));
    S2G_CHECK(dims == 2 || dims == 3, S2G_EINVAL, "%s: dims must be 2 or 3", fn);
    S2G_TRY(check_common(fn, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, npix, kernel));
    S2G_CHECK(shift && halfsize && image_dev, S2G_EINVAL, "%s: NULL argument", fn);
    S2G_CHECK(dims == 2 || n_images == 1, S2G_EINVAL, "%s: 3D maps take a single quantity", fn);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", fn);
    S2G_CHECK(dims == 2 || npix <= 2048, S2G_EINVAL, "%s: npix > 2048 not supported in 3D", fn);
    S2G_TRY(stats_begin(ctx, n));
    const size_t ncell = dims == 2 ? (size_t)(npix * npix) : (size_t)(npix * npix * npix);
    const int planes = dims == 2 ? n_images + 1 : 2;
    if (!accumulate) S2G_CUDA(cudaMemsetAsync(image_dev, 0, sizeof(double) * ncell * planes, ctx->stream));
    s2g_particles P = dev_particles(pos, hsml, m, rho, binq, w, n, in_dtype);
    set_center(P, shift, periodic, boxsize, halfsize);
    S2G_TRY(set_projection(fn, P, perm, rot));
    // the reference does not forward calc_mean to cic_mapping_3D (cic_interpolation.jl:219-221): default false
    s2g_geom G = make_geom(len2pix, npix, dims == 2 ? n_images : 1, dims == 2 ? calc_mean : 0);
    if (dims == 2) return s2g_launch_deposit_2d(ctx, P, G, kernel, image_dev);
    return s2g_launch_deposit_3d(ctx, P, G, kernel, image_dev);
}

extern "C" int s2g_sphmap_dev(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                              const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                              int32_t in_dtype, const double shift[3], int32_t periodic, double boxsize,
                              const double halfsize[3], double len2pix, int64_t npix, int32_t kernel,
                              int32_t calc_mean, int32_t accumulate, double* image_dev)
{
    return sphmap_dev_impl(__func__, ctx, dims, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, nullptr, nullptr,
                           shift, periodic, boxsize, halfsize, len2pix, npix, kernel, calc_mean, accumulate, image_dev);
}

extern "C" int s2g_sphmap_projected_dev(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                                        const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                                        int32_t in_dtype, const int32_t* perm, const double* rot,
                                        const double shift[3], int32_t periodic, double boxsize,
                                        const double halfsize[3], double len2pix, int64_t npix, int32_t kernel,
                                        int32_t calc_mean, int32_t accumulate, double* image_dev)
{
    return sphmap_dev_impl(__func__, ctx, dims, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, perm, rot, shift,
                           periodic, boxsize, halfsize, len2pix, npix, kernel, calc_mean, accumulate, image_dev);
}

// First half of the host sphMapping call, shared with the device-group path (s2g_group.cu): argument checks, H2D
// staging of the six arrays, zero-filled flat image in the context's "image" scratch buffer, deposit with the fused
// centre + filter (+ projection), optional recentred positions back to the host.  Records ev[0] (start), ev[1]
// (inputs resident), ev[2] (deposit enqueued); leaves the stream un-synchronised.
int s2g_sphmap_stage_deposit(const char* fn, s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml,
                             const void* m, const void* rho, const void* binq, const void* w, int64_t n,
                             int32_t n_images, int32_t in_dtype, const int32_t* perm, const double* rot,
                             const double shift[3], int32_t periodic, double boxsize, const double halfsize[3],
                             double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean, void* pos_recentred_out,
                             double** image_dev_out)
{
    S2G_CHECK(ctx != nullptr, S2G_EINVAL, "%s: ctx is NULL", fn);
    S2G_CUDA(cudaSetDevice(ctx->device));
    S2G_CHECK(dims == 2 || dims == 3, S2G_EINVAL, "%s: dims must be 2 or 3", fn);
    S2G_TRY(check_common(fn, pos, hsml, m, rho, binq, w, n, in_dtype, len2pix, npix, kernel));
    S2G_CHECK(shift && halfsize, S2G_EINVAL, "%s: NULL argument", fn);
    S2G_CHECK(dims == 2 || n_images == 1, S2G_EINVAL, "%s: 3D maps take a single quantity", fn);
    S2G_CHECK(n_images >= 1 && n_images <= 64, S2G_EINVAL, "%s: n_images out of range", fn);
    S2G_CHECK(dims == 2 || npix <= 2048, S2G_EINVAL, "%s: npix > 2048 not supported in 3D", fn);
    S2G_TRY(stats_begin(ctx, n));
    const size_t ncell = dims == 2 ? (size_t)(npix * npix) : (size_t)(npix * npix * npix);
    const int planes = dims == 2 ? n_images + 1 : 2;
    void* dimg;
    S2G_TRY(s2g_scratch(ctx, "image", sizeof(double) * ncell * planes, &dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    s2g_particles P;
    S2G_TRY(stage_particles(ctx, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, P));
    StageGuard stage_guard{ctx};
    S2G_CUDA(cudaMemsetAsync(dimg, 0, sizeof(double) * ncell * planes, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    set_center(P, shift, periodic, boxsize, halfsize);
    S2G_TRY(set_projection(fn, P, perm, rot));
    S2G_CHECK(!(P.proj == 2 && pos_recentred_out && in_dtype == S2G_F32), S2G_EINVAL,
              "%s: rotated Float32 positions are Float64 in the reference; pos_recentred_out is not available", fn);
    // the reference does not forward calc_mean to cic_mapping_3D (cic_interpolation.jl:219-221): default false
    s2g_geom G = make_geom(len2pix, npix, dims == 2 ? n_images : 1, dims == 2 ? calc_mean : 0);
    if (dims == 2)
        S2G_TRY(s2g_launch_deposit_2d(ctx, P, G, kernel, (double*)dimg));
    else
        S2G_TRY(s2g_launch_deposit_3d(ctx, P, G, kernel, (double*)dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (pos_recentred_out && n > 0) {
        void* dpo;
        S2G_TRY(s2g_scratch(ctx, "pos_out", 3 * (size_t)n * esize(in_dtype), &dpo));
        S2G_TRY(s2g_launch_center_filter(ctx, P, dpo, nullptr));
        // (in place for a Julia caller: unstage_output first joins the threads that were reading the same array)
        S2G_TRY(unstage_output(ctx, pos_recentred_out, dpo, 3 * (size_t)n * esize(in_dtype)));
    }
    *image_dev_out = (double*)dimg;
    return S2G_OK;
}

static int sphmap_impl(const char* fn, s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                       const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images, int32_t in_dtype,
                       const int32_t* perm, const double* rot, const double shift[3], int32_t periodic, double boxsize,
                       const double halfsize[3], double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean,
                       int32_t reduce_image, int32_t return_both_maps, void* pos_recentred_out, double* out,
                       s2g_stats* stats)
{
    S2G_CHECK(out != nullptr, S2G_EINVAL, "%s: out is NULL", fn);
    double* dimg = nullptr;
    S2G_TRY(s2g_sphmap_stage_deposit(fn, ctx, dims, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, perm, rot, shift,
                                     periodic, boxsize, halfsize, len2pix, npix, kernel, calc_mean, pos_recentred_out,
                                     &dimg));
    const size_t ncell = dims == 2 ? (size_t)(npix * npix) : (size_t)(npix * npix * npix);
    const int planes = dims == 2 ? n_images + 1 : 2;
    const int out_planes = dims == 2 ? n_images : 1;
    const bool both = dims == 2 && return_both_maps;
    if (both) {
        S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
        S2G_TRY(unstage_output(ctx, out, dimg, sizeof(double) * ncell * planes));
    } else {
        void* dred = nullptr;
        S2G_TRY(s2g_scratch(ctx, "reduced", sizeof(double) * ncell * out_planes, &dred));
        if (dims == 2)
            S2G_TRY(s2g_launch_reduce_2d(ctx, (const double*)dimg, npix, npix, n_images, reduce_image, (double*)dred));
        else
            S2G_TRY(s2g_launch_reduce_3d(ctx, (const double*)dimg, npix, reduce_image, (double*)dred));
        S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
        S2G_TRY(unstage_output(ctx, out, dred, sizeof(double) * ncell * out_planes));
    }
    S2G_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_epilogue = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[3], ctx->ev[4]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[4]);
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

extern "C" int s2g_sphmap(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                          const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                          int32_t in_dtype, const double shift[3], int32_t periodic, double boxsize,
                          const double halfsize[3], double len2pix, int64_t npix, int32_t kernel, int32_t calc_mean,
                          int32_t reduce_image, int32_t return_both_maps, void* pos_recentred_out, double* out,
                          s2g_stats* stats)
{
    return sphmap_impl(__func__, ctx, dims, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, nullptr, nullptr, shift,
                       periodic, boxsize, halfsize, len2pix, npix, kernel, calc_mean, reduce_image, return_both_maps,
                       pos_recentred_out, out, stats);
}

extern "C" int s2g_sphmap_projected(s2g_ctx* ctx, int32_t dims, const void* pos, const void* hsml, const void* m,
                                    const void* rho, const void* binq, const void* w, int64_t n, int32_t n_images,
                                    int32_t in_dtype, const int32_t* perm, const double* rot, const double shift[3],
                                    int32_t periodic, double boxsize, const double halfsize[3], double len2pix,
                                    int64_t npix, int32_t kernel, int32_t calc_mean, int32_t reduce_image,
                                    int32_t return_both_maps, void* pos_recentred_out, double* out, s2g_stats* stats)
{
    return sphmap_impl(__func__, ctx, dims, pos, hsml, m, rho, binq, w, n, n_images, in_dtype, perm, rot, shift,
                       periodic, boxsize, halfsize, len2pix, npix, kernel, calc_mean, reduce_image, return_both_maps,
                       pos_recentred_out, out, stats);
}

// ------------------------------------------------------------------------------------------------
// HEALPix
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_healpix_deposit_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                       const void* binq, const void* w, int64_t n, int32_t in_dtype, int64_t nside,
                                       int32_t kernel, int32_t calc_mean, int32_t accumulate, double* map_dev,
                                       double* wmap_dev)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, 1.0, 1, kernel));
    S2G_CHECK(nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0, S2G_EINVAL,
              "%s: nside must be a power of two in [1, 8192]", __func__);
    S2G_CHECK(map_dev && wmap_dev, S2G_EINVAL, "%s: NULL map", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t npix = (size_t)(12 * nside * nside);
    if (!accumulate) {
        S2G_CUDA(cudaMemsetAsync(map_dev, 0, sizeof(double) * npix, ctx->stream));
        S2G_CUDA(cudaMemsetAsync(wmap_dev, 0, sizeof(double) * npix, ctx->stream));
    }
    s2g_particles P = dev_particles(pos, hsml, m, rho, binq, w, n, in_dtype);
    return s2g_launch_healpix(ctx, P, nside, kernel, calc_mean, nullptr, map_dev, wmap_dev);
}

extern "C" int s2g_healpix_deposit(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                   const void* binq, const void* w, int64_t n, int32_t in_dtype, int64_t nside,
                                   int32_t kernel, int32_t calc_mean, double* map_out, double* wmap_out,
                                   s2g_stats* stats)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, in_dtype, 1.0, 1, kernel));
    S2G_CHECK(nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0, S2G_EINVAL,
              "%s: nside must be a power of two in [1, 8192]", __func__);
    S2G_CHECK(map_out && wmap_out, S2G_EINVAL, "%s: NULL map", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t npix = (size_t)(12 * nside * nside);
    void* dmaps;
    S2G_TRY(s2g_scratch(ctx, "image", sizeof(double) * npix * 2, &dmaps));
    double* dmap = (double*)dmaps;
    double* dwmap = dmap + npix;
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    s2g_particles P;
    S2G_TRY(stage_particles(ctx, pos, hsml, m, rho, binq, w, n, 1, in_dtype, P));
    StageGuard stage_guard{ctx};
    S2G_CUDA(cudaMemsetAsync(dmaps, 0, sizeof(double) * npix * 2, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    S2G_TRY(s2g_launch_healpix(ctx, P, nside, kernel, calc_mean, nullptr, dmap, dwmap));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(map_out, dmap, sizeof(double) * npix, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(wmap_out, dwmap, sizeof(double) * npix, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

// fused healpix_map body: recentre on the observer + shell filter + far-to-near selection quirk + deposit
extern "C" int s2g_healpix_map_dev(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                                   const void* binq, const void* w, int64_t n, const double center[3],
                                   const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                                   int32_t accumulate, double* map_dev, double* wmap_dev, int64_t* n_selected)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_common(__func__, pos, hsml, m, rho, binq, w, n, S2G_F64, 1.0, 1, kernel));
    S2G_CHECK(nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0, S2G_EINVAL,
              "%s: nside must be a power of two in [1, 8192]", __func__);
    S2G_CHECK(map_dev && wmap_dev && center && radius_limits, S2G_EINVAL, "%s: NULL argument", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t npix = (size_t)(12 * nside * nside);
    if (!accumulate) {
        S2G_CUDA(cudaMemsetAsync(map_dev, 0, sizeof(double) * npix, ctx->stream));
        S2G_CUDA(cudaMemsetAsync(wmap_dev, 0, sizeof(double) * npix, ctx->stream));
    }
    s2g_particles P = dev_particles(pos, hsml, m, rho, binq, w, n, S2G_F64);
    P.fuse_center = 1;  // Pos .-= center (Float64), no periodic wrap, no box filter
    P.periodic = 0;
    for (int d = 0; d < 3; ++d) { P.shift[d] = center[d]; P.halfsize[d] = 0.0; }
    long long nsel = 0;
    S2G_TRY(s2g_launch_healpix_filtered(ctx, P, radius_limits[0], radius_limits[1], nside, kernel, calc_mean, map_dev,
                                        wmap_dev, &nsel));
    if (n_selected) *n_selected = nsel;
    return S2G_OK;
}

// First half of the host healpix_map call, shared with the device-group path: staging, zero-filled maps in the
// "image" scratch buffer (map at [0, npix), weight map at [npix, 2 npix)), recentre + shell filter + far-to-near
// selection + particle loop, optional recentred positions back.  Records ev[0..2] like s2g_sphmap_stage_deposit.
int s2g_healpix_stage_deposit(const char* fn, s2g_ctx* ctx, const void* pos, const void* hsml, const void* m,
                              const void* rho, const void* binq, const void* w, int64_t n, const double center[3],
                              const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                              void* pos_recentred_out, double** maps_dev_out, long long* n_selected)
{
    S2G_CHECK(ctx != nullptr, S2G_EINVAL, "%s: ctx is NULL", fn);
    S2G_CUDA(cudaSetDevice(ctx->device));
    S2G_TRY(check_common(fn, pos, hsml, m, rho, binq, w, n, S2G_F64, 1.0, 1, kernel));
    S2G_CHECK(nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0, S2G_EINVAL,
              "%s: nside must be a power of two in [1, 8192]", fn);
    S2G_CHECK(center && radius_limits, S2G_EINVAL, "%s: NULL argument", fn);
    S2G_TRY(stats_begin(ctx, n));
    const size_t npix = (size_t)(12 * nside * nside);
    void* dmaps;
    S2G_TRY(s2g_scratch(ctx, "image", sizeof(double) * npix * 2, &dmaps));
    double* dmap = (double*)dmaps;
    double* dwmap = dmap + npix;
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    s2g_particles P;
    S2G_TRY(stage_particles(ctx, pos, hsml, m, rho, binq, w, n, 1, S2G_F64, P));
    StageGuard stage_guard{ctx};
    S2G_CUDA(cudaMemsetAsync(dmaps, 0, sizeof(double) * npix * 2, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    P.fuse_center = 1;  // Pos .-= center (Float64), no periodic wrap, no box filter
    P.periodic = 0;
    for (int d = 0; d < 3; ++d) { P.shift[d] = center[d]; P.halfsize[d] = 0.0; }
    long long nsel = 0;
    S2G_TRY(s2g_launch_healpix_filtered(ctx, P, radius_limits[0], radius_limits[1], nside, kernel, calc_mean, dmap,
                                        dwmap, &nsel));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (pos_recentred_out && n > 0) {  // the reference recentres the caller's Pos in place (filter_particles.jl:20)
        void* dpo;
        S2G_TRY(s2g_scratch(ctx, "pos_out", 3 * (size_t)n * sizeof(double), &dpo));
        S2G_TRY(s2g_launch_center_filter(ctx, P, dpo, nullptr));
        S2G_TRY(unstage_output(ctx, pos_recentred_out, dpo, 3 * (size_t)n * sizeof(double)));
    }
    *maps_dev_out = dmap;
    if (n_selected) *n_selected = nsel;
    return S2G_OK;
}

extern "C" int s2g_healpix_map(s2g_ctx* ctx, const void* pos, const void* hsml, const void* m, const void* rho,
                               const void* binq, const void* w, int64_t n, const double center[3],
                               const double radius_limits[2], int64_t nside, int32_t kernel, int32_t calc_mean,
                               void* pos_recentred_out, double* map_out, double* wmap_out, s2g_stats* stats)
{
    S2G_CHECK(map_out && wmap_out, S2G_EINVAL, "%s: NULL argument", __func__);
    double* dmap = nullptr;
    long long nsel = 0;
    S2G_TRY(s2g_healpix_stage_deposit(__func__, ctx, pos, hsml, m, rho, binq, w, n, center, radius_limits, nside,
                                      kernel, calc_mean, pos_recentred_out, &dmap, &nsel));
    const size_t npix = (size_t)(12 * nside * nside);
    S2G_TRY(unstage_output(ctx, map_out, dmap, sizeof(double) * npix));
    S2G_TRY(unstage_output(ctx, wmap_out, dmap + npix, sizeof(double) * npix));
    S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    ctx->stats.n_in = nsel;  // particles selected by the shell filter
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

extern "C" int s2g_healpix_pixels(s2g_ctx* ctx, const double pos[3], double radius, int64_t nside, int64_t* out,
                                  int64_t cap, int64_t* count_out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(pos && out && count_out && cap > 0, S2G_EINVAL, "%s: bad arguments", __func__);
    S2G_CHECK(nside >= 1 && nside <= 8192 && (nside & (nside - 1)) == 0, S2G_EINVAL, "%s: bad nside", __func__);
    void *dout, *dcnt;
    S2G_TRY(s2g_scratch(ctx, "hp_pixels", sizeof(long long) * (size_t)cap, &dout));
    S2G_TRY(s2g_scratch(ctx, "hp_count", sizeof(long long), &dcnt));
    S2G_TRY(s2g_launch_healpix_pixels(ctx, pos, radius, nside, (long long*)dout, cap, (long long*)dcnt));
    long long cnt = 0;
    S2G_CUDA(cudaMemcpyAsync(&cnt, dcnt, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    *count_out = cnt;
    const long long ncopy = cnt < cap ? cnt : cap;
    if (ncopy > 0) {
        S2G_CUDA(cudaMemcpyAsync(out, dout, sizeof(long long) * (size_t)ncopy, cudaMemcpyDeviceToHost, ctx->stream));
        S2G_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return S2G_OK;
}

// ------------------------------------------------------------------------------------------------
// stencils
// ------------------------------------------------------------------------------------------------
static int check_stencil(const char* fn, int order, int dims, const void* pos, const void* q, int64_t n, int in_dtype,
                         int64_t npix)
{
    S2G_CHECK(order == 2 || order == 3, S2G_EINVAL, "%s: order must be 2 (CIC) or 3 (TSC)", fn);
    S2G_CHECK(dims == 2 || dims == 3, S2G_EINVAL, "%s: dims must be 2 or 3", fn);
    S2G_CHECK(n >= 0 && (n == 0 || (pos && q)), S2G_EINVAL, "%s: bad particle arrays", fn);
    S2G_CHECK(in_dtype == S2G_F32 || in_dtype == S2G_F64, S2G_EINVAL, "%s: bad in_dtype", fn);
    S2G_CHECK(npix > 0 && (dims == 2 ? npix <= 1000000 : npix <= 2048), S2G_EINVAL, "%s: npix out of range", fn);
    return S2G_OK;
}

extern "C" int s2g_stencil_deposit_dev(s2g_ctx* ctx, int32_t order, int32_t dims, const void* pos, const void* q,
                                       int64_t n, int32_t in_dtype, double len2pix, int64_t npix, int32_t periodic,
                                       int32_t accumulate, double* image_dev)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_stencil(__func__, order, dims, pos, q, n, in_dtype, npix));
    S2G_CHECK(image_dev != nullptr, S2G_EINVAL, "%s: image is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t ncell = dims == 2 ? (size_t)(npix * npix) : (size_t)(npix * npix * npix);
    if (!accumulate) S2G_CUDA(cudaMemsetAsync(image_dev, 0, sizeof(double) * ncell * 2, ctx->stream));
    return s2g_launch_stencil(ctx, order, dims, pos, q, n, in_dtype, len2pix, npix, periodic, image_dev);
}

extern "C" int s2g_stencil_deposit(s2g_ctx* ctx, int32_t order, int32_t dims, const void* pos, const void* q,
                                   int64_t n, int32_t in_dtype, double len2pix, int64_t npix, int32_t periodic,
                                   double* image_out, s2g_stats* stats)
{
    CTX_ENTER(ctx);
    S2G_TRY(check_stencil(__func__, order, dims, pos, q, n, in_dtype, npix));
    S2G_CHECK(image_out != nullptr, S2G_EINVAL, "%s: image_out is NULL", __func__);
    S2G_TRY(stats_begin(ctx, n));
    const size_t es = esize(in_dtype);
    const size_t ncell = dims == 2 ? (size_t)(npix * npix) : (size_t)(npix * npix * npix);
    void *dpos, *dq, *dimg;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    S2G_TRY(s2g_scratch(ctx, "in_pos", 3 * nn * es, &dpos));
    S2G_TRY(s2g_scratch(ctx, "in_q", nn * es, &dq));
    S2G_TRY(s2g_scratch(ctx, "image", sizeof(double) * ncell * 2, &dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (n > 0) {
        S2G_CUDA(cudaMemcpyAsync(dpos, pos, 3 * (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
        S2G_CUDA(cudaMemcpyAsync(dq, q, (size_t)n * es, cudaMemcpyHostToDevice, ctx->stream));
    }
    S2G_CUDA(cudaMemsetAsync(dimg, 0, sizeof(double) * ncell * 2, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    S2G_TRY(s2g_launch_stencil(ctx, order, dims, dpos, dq, n, in_dtype, len2pix, npix, periodic, (double*)dimg));
    S2G_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    S2G_CUDA(cudaMemcpyAsync(image_out, dimg, sizeof(double) * ncell * 2, cudaMemcpyDeviceToHost, ctx->stream));
    S2G_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    S2G_TRY(stats_collect(ctx));
    ctx->stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    ctx->stats.ms_compute = ev_ms(ctx->ev[1], ctx->ev[2]);
    ctx->stats.ms_d2h = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
    if (stats) *stats = ctx->stats;
    return S2G_OK;
}

extern "C" int s2g_accumulate_finite_dev(s2g_ctx* ctx, double* sum_dev, const double* local_dev, int64_t n)
{
    CTX_ENTER(ctx);
    S2G_CHECK(n >= 0 && (n == 0 || (sum_dev && local_dev)), S2G_EINVAL, "%s: bad arguments", __func__);
    return s2g_launch_accumulate_finite(ctx, sum_dev, local_dev, n);
}

extern "C" int s2g_divide_slice_dev(s2g_ctx* ctx, int32_t dims, double* q_slice_dev, const double* w_slice_dev,
                                    int64_t n, int64_t plane_stride, int32_t n_images, int32_t reduce_image)
{
    CTX_ENTER(ctx);
    S2G_CHECK(dims == 2 || dims == 3, S2G_EINVAL, "%s: dims must be 2 or 3", __func__);
    S2G_CHECK(n >= 0 && (n == 0 || (q_slice_dev && w_slice_dev)) && n_images >= 1 && plane_stride >= n, S2G_EINVAL,
              "%s: bad arguments", __func__);
    return s2g_launch_divide_slice(ctx, dims, q_slice_dev, w_slice_dev, n, plane_stride, n_images, reduce_image);
}

// ------------------------------------------------------------------------------------------------
// synthetic particles / microbenchmarks
// ------------------------------------------------------------------------------------------------
extern "C" int s2g_synth_particles_dev(s2g_ctx* ctx, uint64_t seed, int64_t first_id, int64_t n, int64_t n_total,
                                       double box, double n_ngb, double sigma_ln_rho, int32_t out_dtype, void* pos,
                                       void* hsml, void* m, void* rho, void* temp)
{
    CTX_ENTER(ctx);
    S2G_CHECK(n >= 0 && n_total > 0 && first_id >= 0, S2G_EINVAL, "%s: bad counts", __func__);
    S2G_CHECK(n == 0 || (pos && hsml && m && rho && temp), S2G_EINVAL, "%s: NULL output", __func__);
    S2G_CHECK(out_dtype == S2G_F32 || out_dtype == S2G_F64, S2G_EINVAL, "%s: bad out_dtype", __func__);
    return s2g_launch_synth(ctx, seed, first_id, n, n_total, box, n_ngb, sigma_ln_rho, out_dtype, pos, hsml, m, rho,
                            temp);
}

extern "C" int s2g_microbench(s2g_ctx* ctx, int32_t which, uint64_t bytes, int32_t iters, double* rate_out)
{
    CTX_ENTER(ctx);
    S2G_CHECK(rate_out != nullptr && iters > 0, S2G_EINVAL, "%s: bad arguments", __func__);
    return s2g_run_microbench(ctx, which, (size_t)bytes, iters, rate_out);
}
