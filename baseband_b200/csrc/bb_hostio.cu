// Host-side ingest helpers of the C ABI: a persistent pool of native threads
// that moves bytes from a file (its mmap'ed page cache, or pread) into the
// pinned staging buffers the H2D copies read from.  The reference reads with
// one `fh.readinto` per payload (baseband/base/payload.py:83-143), i.e. one
// core's copy rate; feeding a PCIe 5 x16 link (~55 GB/s) takes several cores
// and no per-slice interpreter overhead, hence native threads instead of a
// Python thread pool (profiles/r2_file_ingest.txt).
#include <errno.h>
#include <pthread.h>
#include <string.h>
#include <unistd.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "bb_runtime.cuh"

namespace bb {

struct CopyJob {
    uint8_t *dst = nullptr;
    const uint8_t *src = nullptr;       // memcpy source, or
    int fd = -1;                        // pread source (src == nullptr)
    long long offset = 0;
    long long nbytes = 0;
    long long done = 0;                 // bytes actually moved
};

class HostPool {
  public:
    // One pool per process, made on first use and never destroyed (its
    // threads are detached).  A forked child starts without threads: it
    // drops the parent's pool object (whose workers do not exist there, and
    // whose locks may be held) and makes its own on first use.
    static HostPool &get() {
        static std::once_flag once;
        std::call_once(once, [] {
            pthread_atfork(nullptr, nullptr, [] { instance() = nullptr; });
        });
        HostPool *&pool = instance();
        if (!pool) pool = new HostPool();
        return *pool;
    }

    // Run jobs[0..n) on the pool (job 0 on the calling thread); returns when
    // all are finished.  One batch at a time.
    void run(CopyJob *jobs, int n) {
        acquire();
        if (n > 1) {
            ensure_workers(n - 1);
            post(jobs, 1, n);
        }
        execute(jobs[0]);
        if (n > 1) finish();
        release();
    }

    // The same without the caller: the jobs (copied) run on the workers only
    // and `wait` collects them.  begin/wait pairs must not be nested.
    void begin(const CopyJob *jobs, int n) {
        acquire();
        for (int i = 0; i < n; ++i) held_[i] = jobs[i];
        nheld_ = n;
        ensure_workers(n < 1 ? 1 : n);
        post(held_, 0, n);
    }

    long long wait() {
        finish();
        long long total = 0;
        for (int i = 0; i < nheld_; ++i) {
            total += held_[i].done;
            if (held_[i].done < held_[i].nbytes) break;
        }
        nheld_ = 0;
        release();
        return total;
    }

  private:
    HostPool() = default;
    static HostPool *&instance() {
        static HostPool *pool = nullptr;
        return pool;
    }

    static void execute(CopyJob &job) {
        if (job.src) {
            memcpy(job.dst, job.src, (size_t)job.nbytes);
            job.done = job.nbytes;
            return;
        }
        long long got = 0;
        while (got < job.nbytes) {
            ssize_t k = pread(job.fd, job.dst + got,
                              (size_t)(job.nbytes - got), job.offset + got);
            if (k < 0 && errno == EINTR) continue;
            if (k <= 0) break;
            got += k;
        }
        job.done = got;
    }

    void acquire() {
        std::unique_lock<std::mutex> lock(mutex_);
        idle_.wait(lock, [this] { return !busy_; });
        busy_ = true;
    }

    void release() {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            busy_ = false;
        }
        idle_.notify_one();
    }

    void post(CopyJob *jobs, int first, int n) {
        {
            std::lock_guard<std::mutex> lock(mutex_);
            jobs_ = jobs;
            next_ = first;
            njobs_ = n;
            pending_ = n - first;
            ++generation_;
        }
        wake_.notify_all();
    }

    void finish() {
        std::unique_lock<std::mutex> lock(mutex_);
        finished_.wait(lock, [this] { return pending_ == 0; });
        jobs_ = nullptr;
    }

    void ensure_workers(int n) {
        while (nworkers_ < n) {
            std::thread([this] { loop(); }).detach();
            ++nworkers_;
        }
    }

    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            CopyJob *job = nullptr;
            {
                std::unique_lock<std::mutex> lock(mutex_);
                wake_.wait(lock, [&] {
                    return stop_ || (generation_ != seen && jobs_
                                     && next_ < njobs_);
                });
                if (stop_) return;
                job = &jobs_[next_++];
                if (next_ >= njobs_) seen = generation_;
            }
            execute(*job);
            {
                std::lock_guard<std::mutex> lock(mutex_);
                if (--pending_ == 0) finished_.notify_all();
            }
        }
    }

    std::mutex mutex_;
    std::condition_variable wake_, finished_, idle_;
    bool busy_ = false;
    CopyJob held_[64];
    int nheld_ = 0;
    int nworkers_ = 0;
    CopyJob *jobs_ = nullptr;
    int next_ = 0, njobs_ = 0, pending_ = 0;
    unsigned long long generation_ = 0;
    bool stop_ = false;
};

static int split_jobs(CopyJob *jobs, uint8_t *dst, const uint8_t *src, int fd,
                      long long offset, long long nbytes, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    long long step = (nbytes + nthreads - 1) / nthreads;
    if (step < (1ll << 20)) step = 1ll << 20;
    step = (step + 4095) / 4096 * 4096;
    int n = 0;
    for (long long lo = 0; lo < nbytes && n < 64; lo += step, ++n) {
        jobs[n].dst = dst + lo;
        jobs[n].src = src ? src + lo : nullptr;
        jobs[n].fd = fd;
        jobs[n].offset = offset + lo;
        jobs[n].nbytes = lo + step < nbytes ? step : nbytes - lo;
        jobs[n].done = 0;
    }
    return n;
}

static int run_split(uint8_t *dst, const uint8_t *src, int fd,
                     long long offset, long long nbytes, int nthreads,
                     long long *moved) {
    // slices on 4 KiB boundaries of the destination, at least 1 MiB each
    CopyJob jobs[64];
    const int n = split_jobs(jobs, dst, src, fd, offset, nbytes, nthreads);
    if (n) HostPool::get().run(jobs, n);
    long long total = 0;
    for (int i = 0; i < n; ++i) {
        total += jobs[i].done;
        if (jobs[i].done < jobs[i].nbytes) break;    // short read: stop here
    }
    *moved = total;
    return BB_OK;
}

}  // namespace bb

using namespace bb;

extern "C" int bb_host_copy(void *dst, const void *src, int64_t nbytes,
                            int32_t nthreads) {
    if (nbytes < 0 || (nbytes > 0 && (!dst || !src)))
        return set_error(BB_ERR_ARGUMENT, "bad host copy arguments");
    long long moved = 0;
    return run_split((uint8_t *)dst, (const uint8_t *)src, -1, 0, nbytes,
                     nthreads, &moved);
}

extern "C" int bb_host_pread(int32_t fd, void *dst, int64_t nbytes,
                             int64_t offset, int32_t nthreads,
                             int64_t *nread) {
    if (fd < 0 || nbytes < 0 || offset < 0 || (nbytes > 0 && !dst) || !nread)
        return set_error(BB_ERR_ARGUMENT, "bad host pread arguments");
    long long moved = 0;
    int rc = run_split((uint8_t *)dst, nullptr, fd, offset, nbytes, nthreads,
                       &moved);
    *nread = moved;
    return rc;
}

extern "C" int bb_host_copy_begin(void *dst, const void *src, int64_t nbytes,
                                  int32_t nthreads) {
    if (nbytes < 0 || (nbytes > 0 && (!dst || !src)))
        return set_error(BB_ERR_ARGUMENT, "bad host copy arguments");
    CopyJob jobs[64];
    const int n = split_jobs(jobs, (uint8_t *)dst, (const uint8_t *)src, -1,
                             0, nbytes, nthreads);
    HostPool::get().begin(jobs, n);
    return BB_OK;
}

extern "C" int bb_host_copy_wait(int64_t *nbytes) {
    const long long moved = HostPool::get().wait();
    if (nbytes) *nbytes = moved;
    return BB_OK;
}
