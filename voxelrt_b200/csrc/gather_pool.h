// A pool of host threads that lives with a context and sleeps between calls (vrt_sync's staging gather).  Header-only, plain C++17: the CPU
// test tier exercises it without CUDA (tests/native/test_host.cpp).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace vrt {

// Host threads that gather dirty bricks from the caller's records into the pinned staging buffer (vrt_sync).  They live as long as the
// context and sleep between calls: a per-frame edit batch (thousands of 512-byte bricks, a few MB) is too short to pay for starting
// threads, yet one core's memcpy of it is the longest single piece of a frame's vrt_sync.
class GatherPool {
public:
    ~GatherPool() { stop(); }
    unsigned workers() const { return (unsigned)th_.size(); }
    void start(unsigned n) {
        quit_ = false;  // (no worker exists here)
        const unsigned g = gen_;
        for (unsigned i = 0; i < n; i++) th_.emplace_back([this, i, g] { loop(i, g); });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
        th_.clear();
    }
    // f(part, parts) on every worker (part 0 .. workers() - 1) and on the caller (part workers()); returns when all are done
    void run(const std::function<void(unsigned, unsigned)>& f) {
        const unsigned parts = workers() + 1;
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = &f, pending_ = workers(), gen_++;
        }
        cv_.notify_all();
        f(parts - 1, parts);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(unsigned idx, unsigned seen) {
        for (;;) {
            const std::function<void(unsigned, unsigned)>* f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return quit_ || gen_ != seen; });
                if (quit_) return;
                seen = gen_, f = job_;
            }
            (*f)(idx, workers() + 1);
            {
                std::lock_guard<std::mutex> l(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned, unsigned)>* job_ = nullptr;
    unsigned gen_ = 0, pending_ = 0;
    bool quit_ = false;
};

}  // namespace vrt
