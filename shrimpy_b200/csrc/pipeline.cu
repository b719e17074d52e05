// Host-buffer streaming pipeline: H2D slab -> deskew kernel -> D2H slab on three streams.
//
// Stands behind biahub.analysis.deskew.deskew_data as called at scripts/measure_psf.py:239-246
// (numpy in, numpy out).  The reference script cuts the stack along raw X "so that it fits in
// the GPU memory" (measure_psf.py:218-221); here the cut is along the tilt axis in multiples of
// average_n_slices rows, which is just as halo-free and makes every output slab
// one contiguous block of the result, so the D2H side is a single linear copy per slab.
//
// Pageable caller buffers (what an ordinary numpy array is): the driver's own staged copy of pageable memory
// measured 10.5 GB/s on the B200 boxes (70 ms for the 737 MB of one mantis channel against 13 ms from page-locked
// memory), so the pipeline stages such buffers itself -- a few host threads gather a slab's rows into a page-locked
// ring (input) or scatter a finished slab out of one (output) while the copies and kernels of the neighbouring slabs
// run -- and every transfer the GPU sees is asynchronous and at the PCIe rate.
#include "common.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

struct shrimpy_pipeline {
    int device = 0;
    size_t budget = 0;
    static constexpr int kBuf = 3;
    cudaStream_t s_h2d = nullptr, s_run = nullptr, s_d2h = nullptr;
    void *d_raw[kBuf] = {nullptr, nullptr, nullptr};
    float *d_out[kBuf] = {nullptr, nullptr, nullptr};
    size_t raw_cap = 0, out_cap = 0;
    // page-locked staging rings, allocated on the first call with a pageable input / output
    void *h_stage_in[kBuf] = {nullptr, nullptr, nullptr};
    float *h_stage_out[kBuf] = {nullptr, nullptr, nullptr};
    size_t stage_in_cap = 0, stage_out_cap = 0;
    cudaEvent_t ev_h2d[kBuf], ev_run[kBuf], ev_d2h[kBuf];
    bool events = false;
    int64_t launches = 0, h2d_bytes = 0, d2h_bytes = 0;
    int64_t staged_in_bytes = 0, staged_out_bytes = 0;
    std::mutex mutex;   // one call at a time per pipeline: the slots, events and streams above are shared state
};

using namespace shrimpy;

static void pipeline_free_buffers(shrimpy_pipeline *p) {
    for (int i = 0; i < shrimpy_pipeline::kBuf; ++i) {
        if (p->d_raw[i]) cudaFree(p->d_raw[i]);
        if (p->d_out[i]) cudaFree(p->d_out[i]);
        p->d_raw[i] = nullptr;
        p->d_out[i] = nullptr;
    }
    p->raw_cap = p->out_cap = 0;
}

static void pipeline_free_stage(shrimpy_pipeline *p, bool in, bool out) {
    for (int i = 0; i < shrimpy_pipeline::kBuf; ++i) {
        if (in && p->h_stage_in[i]) {
            cudaFreeHost(p->h_stage_in[i]);
            p->h_stage_in[i] = nullptr;
        }
        if (out && p->h_stage_out[i]) {
            cudaFreeHost(p->h_stage_out[i]);
            p->h_stage_out[i] = nullptr;
        }
    }
    if (in) p->stage_in_cap = 0;
    if (out) p->stage_out_cap = 0;
}

// The streams hold work that touches caller memory and this pipeline's slots: nothing may be in flight when a call
// returns, whatever it returns.
static void pipeline_drain(shrimpy_pipeline *p) {
    if (p->s_h2d) cudaStreamSynchronize(p->s_h2d);
    if (p->s_run) cudaStreamSynchronize(p->s_run);
    if (p->s_d2h) cudaStreamSynchronize(p->s_d2h);
}

static int host_threads() {
    static const int n = [] {
        const char *s = getenv("SHRIMPY_HOST_THREADS");
        if (s && *s) return std::max(1, atoi(s));
        const unsigned hc = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(8u, hc / 2));
    }();
    return n;
}

// A few persistent host threads that copy byte ranges (spawning threads per slab cost ~5 ms of a 30 ms call).
// One pool per process, created on first use; jobs are issued by one caller at a time (the pool has its own mutex).
class CopyPool {
  public:
    // two pools: the gather of the next input slab and the scatter of a finished output slab run side by side
    static CopyPool &get(int which) {
        static CopyPool gather(host_threads()), scatter(host_threads());
        return which == 0 ? gather : scatter;
    }
    // rows x width bytes, row r from src + r * src_pitch to dst + r * dst_pitch, cut over the threads by bytes
    void copy_rows(char *dst, size_t dst_pitch, const char *src, size_t src_pitch, size_t width, size_t rows) {
        const size_t total = width * rows;
        if (total < ((size_t)4 << 20) || workers_.empty()) {
            run_part(dst, dst_pitch, src, src_pitch, width, total, 0, 1);
            return;
        }
        std::lock_guard<std::mutex> one_job(issue_);
        {
            std::lock_guard<std::mutex> lock(m_);
            dst_ = dst; dst_pitch_ = dst_pitch; src_ = src; src_pitch_ = src_pitch; width_ = width; total_ = total;
            pending_ = (int)workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        run_part(dst, dst_pitch, src, src_pitch, width, total, 0, (int)workers_.size() + 1);
        std::unique_lock<std::mutex> lock(m_);
        done_.wait(lock, [&] { return pending_ == 0; });
    }

  private:
    explicit CopyPool(int threads) {
        for (int t = 1; t < threads; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    static void run_part(char *dst, size_t dst_pitch, const char *src, size_t src_pitch, size_t width, size_t total, int t,
                         int T) {
        // part t takes the byte range [t, t+1) * total / T of the row-major payload
        size_t a = total / T * t, b = (t == T - 1) ? total : total / T * (t + 1);
        while (a < b) {
            const size_t r = a / width, off = a - r * width;
            const size_t len = std::min(width - off, b - a);
            memcpy(dst + r * dst_pitch + off, src + r * src_pitch + off, len);
            a += len;
        }
    }
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lock(m_);
            cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            char *dst = dst_; const char *src = src_;
            const size_t dp = dst_pitch_, sp = src_pitch_, w = width_, total = total_;
            const int T = (int)workers_.size() + 1;
            lock.unlock();
            run_part(dst, dp, src, sp, w, total, t, T);
            lock.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, issue_;
    std::condition_variable cv_, done_;
    uint64_t generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    char *dst_ = nullptr; const char *src_ = nullptr;
    size_t dst_pitch_ = 0, src_pitch_ = 0, width_ = 0, total_ = 0;
};

static void parallel_copy_rows(int pool, char *dst, size_t dst_pitch, const char *src, size_t src_pitch, size_t width,
                               size_t rows) {
    CopyPool::get(pool).copy_rows(dst, dst_pitch, src, src_pitch, width, rows);
}

static bool is_pageable(const void *ptr) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// Page-locked host memory of exactly the size asked for (the host side keeps its own small cache of result blocks: a
// general-purpose caching allocator rounds 1.05 GB up to 2 GiB, and page-locking is what a cold call waits for).
extern "C" int shrimpy_host_alloc(size_t bytes, void **out) {
    if (!out || bytes == 0) return fail(SHRIMPY_EINVAL, "host_alloc: bad arguments");
    *out = nullptr;
    SHRIMPY_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return SHRIMPY_OK;
}

extern "C" int shrimpy_host_free(void *ptr) {
    if (!ptr) return SHRIMPY_OK;
    SHRIMPY_CUDA_TRY(cudaFreeHost(ptr));
    return SHRIMPY_OK;
}

extern "C" int shrimpy_pipeline_create(int device, size_t device_bytes_budget, shrimpy_pipeline **out) {
    if (!out) return fail(SHRIMPY_EINVAL, "pipeline_create: null out");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(SHRIMPY_ENOGPU, "no CUDA device visible");
    if (device < 0 || device >= count) return fail(SHRIMPY_EINVAL, "pipeline_create: device %d of %d", device, count);
    SHRIMPY_CUDA_TRY(cudaSetDevice(device));
    shrimpy_pipeline *p = new shrimpy_pipeline();
    p->device = device;
    p->budget = device_bytes_budget ? device_bytes_budget : (size_t)2 << 30;
    cudaError_t e = cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_run, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking);
    for (int i = 0; i < shrimpy_pipeline::kBuf && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&p->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_run[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_d2h[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        delete p;
        return fail(SHRIMPY_ECUDA, "pipeline_create: %s", cudaGetErrorString(e));
    }
    p->events = true;
    *out = p;
    return SHRIMPY_OK;
}

extern "C" void shrimpy_pipeline_destroy(shrimpy_pipeline *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(p->mutex);
        cudaSetDevice(p->device);
        pipeline_drain(p);
        pipeline_free_buffers(p);
        pipeline_free_stage(p, true, true);
        if (p->events)
            for (int i = 0; i < shrimpy_pipeline::kBuf; ++i) {
                cudaEventDestroy(p->ev_h2d[i]);
                cudaEventDestroy(p->ev_run[i]);
                cudaEventDestroy(p->ev_d2h[i]);
            }
        if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
        if (p->s_run) cudaStreamDestroy(p->s_run);
        if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
    }
    delete p;
}

extern "C" int shrimpy_pipeline_stats(const shrimpy_pipeline *p, int64_t *launches, int64_t *h2d_bytes,
                                      int64_t *d2h_bytes) {
    if (!p) return fail(SHRIMPY_EINVAL, "pipeline_stats: null pipeline");
    if (launches) *launches = p->launches;
    if (h2d_bytes) *h2d_bytes = p->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = p->d2h_bytes;
    return SHRIMPY_OK;
}

extern "C" int shrimpy_pipeline_staged_bytes(const shrimpy_pipeline *p, int64_t *in_bytes, int64_t *out_bytes) {
    if (!p) return fail(SHRIMPY_EINVAL, "pipeline_staged_bytes: null pipeline");
    if (in_bytes) *in_bytes = p->staged_in_bytes;
    if (out_bytes) *out_bytes = p->staged_out_bytes;
    return SHRIMPY_OK;
}

namespace {

// Scatters finished output slabs from the page-locked ring into a pageable result, in order, on its own thread, so
// that the main thread keeps enqueueing the next slabs meanwhile.
struct OutDrain {
    shrimpy_pipeline *p;
    std::vector<std::pair<float *, size_t>> jobs;   // destination and bytes of slab i
    std::mutex m;
    std::condition_variable cv;
    int enqueued = 0, finished = 0;
    bool stop = false;
    cudaError_t error = cudaSuccess;
    std::thread worker;

    explicit OutDrain(shrimpy_pipeline *pipe, int n_slabs) : p(pipe), jobs(n_slabs) {}

    void start() {
        worker = std::thread([this] {
            cudaSetDevice(p->device);
            for (int i = 0;; ++i) {
                {
                    std::unique_lock<std::mutex> lock(m);
                    cv.wait(lock, [&] { return enqueued > i || stop; });
                    if (enqueued <= i) return;
                }
                const int b = i % shrimpy_pipeline::kBuf;
                const cudaError_t e = cudaEventSynchronize(p->ev_d2h[b]);   // recorded for slab i before it was enqueued
                if (e == cudaSuccess)
                    parallel_copy_rows(1, reinterpret_cast<char *>(jobs[i].first), jobs[i].second,
                                       reinterpret_cast<const char *>(p->h_stage_out[b]), jobs[i].second,
                                       jobs[i].second, 1);
                {
                    std::lock_guard<std::mutex> lock(m);
                    if (e != cudaSuccess && error == cudaSuccess) error = e;
                    finished = i + 1;
                }
                cv.notify_all();
            }
        });
    }
    void push(int i, float *dst, size_t bytes) {
        {
            std::lock_guard<std::mutex> lock(m);
            jobs[i] = {dst, bytes};
            enqueued = i + 1;
        }
        cv.notify_all();
    }
    void wait_finished(int count) {   // slabs [0, count) have left the ring
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [&] { return finished >= count; });
    }
    cudaError_t finish() {
        if (!worker.joinable()) return cudaSuccess;
        {
            std::unique_lock<std::mutex> lock(m);
            cv.wait(lock, [&] { return finished >= enqueued; });
            stop = true;
        }
        cv.notify_all();
        worker.join();
        return error;
    }
    ~OutDrain() { finish(); }
};

}  // namespace

static int deskew_host_locked(shrimpy_pipeline *p, const void *h_raw, int raw_dtype, float *h_out, int Z, int Y,
                              int X, int Xp, int n_avg, double m00, double m02, double shift, float cval) {
    constexpr int kBuf = shrimpy_pipeline::kBuf;
    const size_t es = raw_dtype == SHRIMPY_U16 ? 2 : 4;
    const int Yn = (Y + n_avg - 1) / n_avg;
    // bytes per tilt block: n raw rows in, one output plane out
    const size_t raw_per_p = (size_t)Z * n_avg * X * es;
    const size_t out_per_p = (size_t)X * Xp * sizeof(float);
    // slab size: fits the budget with kBuf buffers in flight, and leaves >= ~12 slabs to overlap
    size_t ps_budget = p->budget / (kBuf * (raw_per_p + out_per_p));
    if (ps_budget == 0) ps_budget = 1;
    int ps = (int)std::min<size_t>(ps_budget, (size_t)std::max(1, (Yn + 11) / 12));
    // The D2H stream is the bottleneck stage and runs without gaps once it has started, so what the call pays on top
    // is the fill: the first slab's H2D + kernel, with nothing to overlap.  Slabs therefore ramp up 1, 2, 4, ... tilt
    // blocks before reaching the steady size.
    std::vector<int> starts, counts;
    for (int p0 = 0, ramp = 1; p0 < Yn; ramp *= 2) {
        const int pc = std::min(std::min(ramp, ps), Yn - p0);
        starts.push_back(p0);
        counts.push_back(pc);
        p0 += pc;
        if (ramp > ps) ramp = ps;
    }
    const int n_slabs = (int)starts.size();

    const size_t raw_need = (size_t)ps * raw_per_p, out_need = (size_t)ps * out_per_p;
    if (raw_need > p->raw_cap || out_need > p->out_cap) {
        pipeline_free_buffers(p);
        for (int i = 0; i < kBuf; ++i) {
            SHRIMPY_CUDA_TRY(cudaMalloc(&p->d_raw[i], raw_need));
            SHRIMPY_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&p->d_out[i]), out_need));
        }
        p->raw_cap = raw_need;
        p->out_cap = out_need;
    }
    static const bool no_stage = [] { const char *s = getenv("SHRIMPY_HOST_NO_STAGING"); return s && *s && *s != '0'; }();
    const bool stage_in = !no_stage && is_pageable(h_raw);
    const bool stage_out = !no_stage && is_pageable(h_out);
    if (stage_in && raw_need > p->stage_in_cap) {
        pipeline_free_stage(p, true, false);
        for (int i = 0; i < kBuf; ++i) SHRIMPY_CUDA_TRY(cudaHostAlloc(&p->h_stage_in[i], raw_need, cudaHostAllocDefault));
        p->stage_in_cap = raw_need;
    }
    if (stage_out && out_need > p->stage_out_cap) {
        pipeline_free_stage(p, false, true);
        for (int i = 0; i < kBuf; ++i)
            SHRIMPY_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&p->h_stage_out[i]), out_need, cudaHostAllocDefault));
        p->stage_out_cap = out_need;
    }
    OutDrain drain(p, n_slabs);
    if (stage_out) drain.start();

    const size_t src_pitch = (size_t)Y * X * es;  // one scan slice of the host stack
    for (int i = 0; i < n_slabs; ++i) {
        const int b = i % kBuf;
        const int p0 = starts[i], pc = counts[i];
        int32_t yr[2], zr[2];
        int rc = shrimpy_deskew_window_needs(Z, Y, n_avg, m00, m02, shift, p0, pc, 0, Xp, yr, zr);
        if (rc) return rc;
        const int yc = yr[1] - yr[0];
        // H2D: Z pieces of (yc rows * X) contiguous elements, pitch = one full slice
        const size_t width = (size_t)yc * X * es;
        const char *src = static_cast<const char *>(h_raw) + (size_t)yr[0] * X * es;
        if (stage_in) {
            // gather the slab into the ring on the host threads (the H2D of slab i - kBuf has left this slot) ...
            if (i >= kBuf) SHRIMPY_CUDA_TRY(cudaEventSynchronize(p->ev_h2d[b]));
            parallel_copy_rows(0, static_cast<char *>(p->h_stage_in[b]), width, src, src_pitch, width, (size_t)Z);
            p->staged_in_bytes += (int64_t)(width * Z);
        }
        if (i >= kBuf) SHRIMPY_CUDA_TRY(cudaStreamWaitEvent(p->s_h2d, p->ev_run[b], 0));  // slab i-kBuf consumed
        if (stage_in)   // ... and ship it as one linear page-locked copy
            SHRIMPY_CUDA_TRY(cudaMemcpyAsync(p->d_raw[b], p->h_stage_in[b], width * Z, cudaMemcpyHostToDevice, p->s_h2d));
        else
            SHRIMPY_CUDA_TRY(cudaMemcpy2DAsync(p->d_raw[b], width, src, src_pitch, width, (size_t)Z,
                                               cudaMemcpyHostToDevice, p->s_h2d));
        SHRIMPY_CUDA_TRY(cudaEventRecord(p->ev_h2d[b], p->s_h2d));
        p->h2d_bytes += (int64_t)(width * Z);

        SHRIMPY_CUDA_TRY(cudaStreamWaitEvent(p->s_run, p->ev_h2d[b], 0));
        if (i >= kBuf) SHRIMPY_CUDA_TRY(cudaStreamWaitEvent(p->s_run, p->ev_d2h[b], 0));  // out buffer drained
        shrimpy_window w = {p0, pc, 0, Xp, yr[0], yc, 0, Z};
        const int64_t before = shrimpy_launch_count();
        rc = shrimpy_deskew_window_device(p->d_raw[b], raw_dtype, p->d_out[b], Z, Y, X, Xp, n_avg, m00, m02, shift,
                                          cval, (int64_t)yc * X, X, (int64_t)X * Xp, Xp, &w, SHRIMPY_KERNEL_AUTO,
                                          p->s_run);
        if (rc) return rc;
        p->launches += shrimpy_launch_count() - before;
        SHRIMPY_CUDA_TRY(cudaEventRecord(p->ev_run[b], p->s_run));

        SHRIMPY_CUDA_TRY(cudaStreamWaitEvent(p->s_d2h, p->ev_run[b], 0));
        const size_t out_bytes = (size_t)pc * out_per_p;
        float *dst = h_out + (size_t)p0 * X * Xp;
        if (stage_out) {
            drain.wait_finished(i - kBuf + 1);      // slab i - kBuf has been scattered out of this ring slot
            SHRIMPY_CUDA_TRY(cudaMemcpyAsync(p->h_stage_out[b], p->d_out[b], out_bytes, cudaMemcpyDeviceToHost, p->s_d2h));
            SHRIMPY_CUDA_TRY(cudaEventRecord(p->ev_d2h[b], p->s_d2h));
            drain.push(i, dst, out_bytes);
            p->staged_out_bytes += (int64_t)out_bytes;
        } else {
            SHRIMPY_CUDA_TRY(cudaMemcpyAsync(dst, p->d_out[b], out_bytes, cudaMemcpyDeviceToHost, p->s_d2h));
            SHRIMPY_CUDA_TRY(cudaEventRecord(p->ev_d2h[b], p->s_d2h));
        }
        p->d2h_bytes += (int64_t)out_bytes;
    }
    SHRIMPY_CUDA_TRY(cudaStreamSynchronize(p->s_d2h));
    if (stage_out) SHRIMPY_CUDA_TRY(drain.finish());
    return SHRIMPY_OK;
}

extern "C" int shrimpy_deskew_host(shrimpy_pipeline *p, const void *h_raw, int raw_dtype, float *h_out, int Z, int Y,
                                   int X, int Xp, int n_avg, double m00, double m02, double shift, float cval) {
    if (!p) return fail(SHRIMPY_EINVAL, "deskew_host: null pipeline");
    if (Z <= 0 || Y <= 0 || X <= 0 || Xp < 0 || n_avg <= 0) return fail(SHRIMPY_EINVAL, "deskew_host: bad shape");
    if (raw_dtype != SHRIMPY_U16 && raw_dtype != SHRIMPY_F32) return fail(SHRIMPY_EINVAL, "deskew_host: bad dtype");
    std::lock_guard<std::mutex> lock(p->mutex);
    p->launches = p->h2d_bytes = p->d2h_bytes = p->staged_in_bytes = p->staged_out_bytes = 0;
    if (Xp == 0) return SHRIMPY_OK;
    if (!h_raw || !h_out) return fail(SHRIMPY_EINVAL, "deskew_host: null host pointer");
    SHRIMPY_CUDA_TRY(cudaSetDevice(p->device));
    const int rc = deskew_host_locked(p, h_raw, raw_dtype, h_out, Z, Y, X, Xp, n_avg, m00, m02, shift, cval);
    if (rc != SHRIMPY_OK) pipeline_drain(p);   // copies into / out of caller memory may still be in flight
    return rc;
}
