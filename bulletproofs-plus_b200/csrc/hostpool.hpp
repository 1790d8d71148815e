// Persistent host worker pool for the Fiat-Shamir replay (one per bpp_ctx): std::thread creation costs ~30 us per thread,
// which at three parallel regions per verification call was a third of the host time of a 1024-proof batch.
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace bpp {

class HostPool {
  public:
    explicit HostPool(int threads) { resize(threads); }
    ~HostPool() { stop(); }
    int size() const { return (int)workers_.size() + 1; }
    void resize(int threads) {
        stop();
        stopping_ = false;
        for (int t = 1; t < threads; t++) workers_.emplace_back([this] { loop(); });
    }
    // f(i) for i in [0, n), in grains of `grain`; the calling thread takes part
    void run(size_t n, size_t grain, const std::function<void(size_t)> &f) {
        if (n == 0) return;
        if (workers_.empty() || n <= grain) { for (size_t i = 0; i < n; i++) f(i); return; }
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &f; n_ = n; grain_ = grain; next_.store(0); pending_ = (int)workers_.size(); generation_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void work() {
        for (;;) {
            size_t i0 = next_.fetch_add(grain_);
            if (i0 >= n_) break;
            size_t i1 = i0 + grain_ < n_ ? i0 + grain_ : n_;
            for (size_t i = i0; i < i1; i++) (*fn_)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stopping_ || generation_ != seen; });
                if (stopping_) return;
                seen = generation_;
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stopping_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
        workers_.clear();
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_ = 0, grain_ = 1;
    std::atomic<size_t> next_{0};
    int pending_ = 0;
    uint64_t generation_ = 0;
    bool stopping_ = false;
};

} // namespace bpp
