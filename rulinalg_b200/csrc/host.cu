// host.cu -- staging of pageable host operands (see context.cuh) and the small thread pool behind it.
//
// The reference hands over Vec<T> storage (src/matrix/mat_mul.rs:28-31, src/matrix/decomposition/lu.rs:166): pageable
// memory.  H2D: pool threads copy a piece (<= 16 MiB) into a pinned ring slot, the DMA engine takes it from there while
// the next piece is being copied.  D2H: the DMA engine fills a slot, and the piece is copied out to the user's
// memory once its event has fired (opportunistically on later calls, or in finish()).  Pinned operands never come here.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "context.cuh"

namespace rla {
namespace {

class Pool {
public:
    explicit Pool(int nthreads) {
        for (int i = 0; i < nthreads; ++i) std::thread([this] { loop(); }).detach();
        n_ = nthreads;
    }
    int size() const { return n_; }
    // f(part) for part in [0, parts); the caller takes parts too.  One run at a time.
    void run(int parts, const std::function<void(int)> &f) {
        if (parts <= 0) return;
        if (parts == 1 || n_ == 0) {
            for (int p = 0; p < parts; ++p) f(p);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mu_);
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &f;
            parts_ = parts;
            next_ = 0;
            pending_ = parts;
            ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void work() {
        for (;;) {
            int p;
            const std::function<void(int)> *f;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (job_ == nullptr || next_ >= parts_) return;
                p = next_++;
                f = job_;
            }
            (*f)(p);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
            }
            work();
        }
    }
    std::mutex mu_, run_mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    int parts_ = 0, next_ = 0, pending_ = 0, n_ = 0;
    uint64_t gen_ = 0;
};

Pool &pool() {
    // leaked on purpose: the workers are detached and may still be parked on the condition variable at exit
    static Pool *p = [] {
        int n = 0;
        if (const char *e = getenv("RLA_STAGE_THREADS")) n = atoi(e) - 1;
        else {
            const int hw = int(std::thread::hardware_concurrency());
            n = hw > 2 ? (hw - 1 > 15 ? 15 : hw - 1) : 1;
        }
        if (n < 0) n = 0;
        return new Pool(n);
    }();
    return *p;
}

}  // namespace

int staging_threads() { return pool().size() + 1; }

void parallel_copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows) {
    const size_t total = width * rows;
    if (total == 0) return;
    unsigned char *d = static_cast<unsigned char *>(dst);
    const unsigned char *s = static_cast<const unsigned char *>(src);
    // byte range [b0, b1) of the virtual contiguous rows x width space
    auto copy_range = [=](size_t b0, size_t b1) {
        size_t r = b0 / width, off = b0 - r * width;
        while (b0 < b1) {
            const size_t len = (width - off < b1 - b0) ? width - off : b1 - b0;
            memcpy(d + r * dpitch + off, s + r * spitch + off, len);
            b0 += len;
            ++r;
            off = 0;
        }
    };
    const size_t min_part = size_t(512) << 10;
    int parts = int((total + min_part - 1) / min_part);
    const int maxp = pool().size() + 1;
    if (parts > maxp) parts = maxp;
    if (parts <= 1) {
        copy_range(0, total);
        return;
    }
    const size_t chunk = ((total + parts - 1) / parts + 63) / 64 * 64;
    pool().run(parts, [&](int p) {
        const size_t b0 = size_t(p) * chunk;
        if (b0 >= total) return;
        copy_range(b0, b0 + chunk < total ? b0 + chunk : total);
    });
}

bool host_is_pinned(const void *p) {
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// The drainer: one thread per Stager that waits for each download's DMA and copies the slot out to the user's
// (pageable) rows, so the issuing thread never blocks on a kernel that has not run yet.
struct Stager::Drain {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv_work, cv_free;
    std::vector<Slot *> fifo;            // issued downloads, oldest first
    size_t head = 0;
    bool stop = false;
    int error = RLA_OK;
    void loop() {
        for (;;) {
            Slot *s;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || head < fifo.size(); });
                if (head >= fifo.size()) return;          // stop requested and nothing left
                s = fifo[head];
            }
            const cudaError_t e = cudaEventSynchronize(s->ev[s->dev]);
            if (e == cudaSuccess) parallel_copy2d(s->dst, s->dpitch, s->p, s->width, s->width, s->rows);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (e != cudaSuccess && error == RLA_OK) error = RLA_ERR_CUDA;
                s->busy = false;
                if (++head == fifo.size()) { fifo.clear(); head = 0; }
            }
            cv_free.notify_all();
        }
    }
};

int Stager::ensure_drain() {
    if (drain_) return RLA_OK;
    drain_ = new Drain();
    drain_->th = std::thread([d = drain_] { d->loop(); });
    return RLA_OK;
}

int Stager::ensure_slot(Slot &s) {
    if (s.p) return RLA_OK;
    void *p = nullptr;
    const cudaError_t e = cudaHostAlloc(&p, SLOT_BYTES, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        note_cuda_error(e);
        (void)cudaGetLastError();
        return RLA_ERR_NOMEM;
    }
    s.p = static_cast<unsigned char *>(p);
    return RLA_OK;
}

void Stager::release() {
    if (drain_) {
        (void)finish();
        {
            std::lock_guard<std::mutex> lk(drain_->mu);
            drain_->stop = true;
        }
        drain_->cv_work.notify_all();
        drain_->th.join();
        delete drain_;
        drain_ = nullptr;
    }
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    auto drop = [&](Slot &s) {
        if (s.busy) (void)cudaEventSynchronize(s.ev[s.dev]);
        s.busy = false;
        for (int d = 0; d < RLA_MAX_DEVICES; ++d)
            if (s.ev[d]) {
                if (have_cur && cudaSetDevice(d) == cudaSuccess) cudaEventDestroy(s.ev[d]);
                s.ev[d] = nullptr;
            }
        if (s.p) cudaFreeHost(s.p);
        s.p = nullptr;
    };
    for (int i = 0; i < NUP; ++i) drop(up_[i]);
    for (int i = 0; i < NDOWN; ++i) drop(down_[i]);
    if (have_cur) cudaSetDevice(cur);
    (void)cudaGetLastError();
}

int Stager::finish() {
    if (!drain_) return RLA_OK;
    std::unique_lock<std::mutex> lk(drain_->mu);
    drain_->cv_free.wait(lk, [&] { return drain_->head >= drain_->fifo.size(); });
    const int st = drain_->error;
    drain_->error = RLA_OK;
    return st;
}

int Stager::upload2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, bool pinned,
                     int dev, cudaStream_t st) {
    if (rows == 0 || width == 0) return RLA_OK;
    if (pinned || width * rows < DIRECT_BYTES) {          // small pageable copies: the driver's own bounce buffer is fine
        RLA_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, st));
        return RLA_OK;
    }
    unsigned char *d = static_cast<unsigned char *>(dst);
    const unsigned char *s = static_cast<const unsigned char *>(src);
    for (size_t c0 = 0; c0 < width; c0 += SLOT_BYTES) {                    // column segments (only for rows wider than a slot)
        const size_t w = width - c0 < SLOT_BYTES ? width - c0 : SLOT_BYTES;
        const size_t rpp = SLOT_BYTES / w;                                 // rows per piece (>= 1)
        for (size_t r0 = 0; r0 < rows; r0 += rpp) {
            const size_t nr = rows - r0 < rpp ? rows - r0 : rpp;
            Slot &sl = up_[next_up_];
            next_up_ = (next_up_ + 1) % NUP;
            RLA_TRY(ensure_slot(sl));
            if (sl.busy) RLA_CUDA(cudaEventSynchronize(sl.ev[sl.dev]));    // its DMA (copy-only stream) is long done
            sl.busy = false;
            if (!sl.ev[dev]) RLA_CUDA(cudaEventCreateWithFlags(&sl.ev[dev], cudaEventDisableTiming));
            parallel_copy2d(sl.p, w, s + r0 * spitch + c0, spitch, w, nr);
            RLA_CUDA(cudaMemcpy2DAsync(d + r0 * dpitch + c0, dpitch, sl.p, w, w, nr, cudaMemcpyHostToDevice, st));
            RLA_CUDA(cudaEventRecord(sl.ev[dev], st));
            sl.dev = dev;
            sl.busy = true;
            staged_bytes += w * nr;
        }
    }
    return RLA_OK;
}

int Stager::download2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows, bool pinned,
                       int dev, cudaStream_t st) {
    if (rows == 0 || width == 0) return RLA_OK;
    if (pinned || width * rows < DIRECT_BYTES) {
        // (a small pageable destination blocks this thread until the data has arrived -- the calls that move so
        // little are synchronous one-shot calls anyway)
        RLA_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToHost, st));
        return RLA_OK;
    }
    // larger pageable destinations go through the ring: cudaMemcpy2DAsync to pageable memory would block the issuing
    // thread until every kernel queued before it has run, and the pipeline behind it would drain
    RLA_TRY(ensure_drain());
    unsigned char *d = static_cast<unsigned char *>(dst);
    const unsigned char *s = static_cast<const unsigned char *>(src);
    for (size_t c0 = 0; c0 < width; c0 += SLOT_BYTES) {
        const size_t w = width - c0 < SLOT_BYTES ? width - c0 : SLOT_BYTES;
        const size_t rpp = SLOT_BYTES / w;
        for (size_t r0 = 0; r0 < rows; r0 += rpp) {
            const size_t nr = rows - r0 < rpp ? rows - r0 : rpp;
            Slot &sl = down_[next_down_];
            next_down_ = (next_down_ + 1) % NDOWN;
            RLA_TRY(ensure_slot(sl));
            {
                std::unique_lock<std::mutex> lk(drain_->mu);
                drain_->cv_free.wait(lk, [&] { return !sl.busy; });
            }
            if (!sl.ev[dev]) RLA_CUDA(cudaEventCreateWithFlags(&sl.ev[dev], cudaEventDisableTiming));
            RLA_CUDA(cudaMemcpy2DAsync(sl.p, w, s + r0 * spitch + c0, spitch, w, nr, cudaMemcpyDeviceToHost, st));
            RLA_CUDA(cudaEventRecord(sl.ev[dev], st));
            sl.dev = dev;
            sl.dst = d + r0 * dpitch + c0;
            sl.dpitch = dpitch;
            sl.width = w;
            sl.rows = nr;
            {
                std::lock_guard<std::mutex> lk(drain_->mu);
                sl.busy = true;
                drain_->fifo.push_back(&sl);
            }
            drain_->cv_work.notify_one();
            staged_bytes += w * nr;
        }
    }
    return RLA_OK;
}

}  // namespace rla
