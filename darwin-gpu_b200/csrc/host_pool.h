// host_pool.h -- helper threads of the engine's host-side passes (descriptor check + routing, copy-in, copy-out).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace gact {

// Persistent helper threads for the host-side passes over a large batch (spawning a thread per pass costs more than
// the pass saves).  run(parts, fn) calls fn(0 .. parts-1), part 0 on the calling thread; one call at a time.
class HostPool {
public:
    ~HostPool()
    {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void run(int parts, const std::function<void(int)> &fn)
    {
        if (parts <= 1) { fn(0); return; }
        // another thread (another engine of this process) is using the helpers: do the parts here, one after the other
        std::unique_lock<std::mutex> one(call_, std::try_to_lock);
        if (!one.owns_lock()) { for (int p = 0; p < parts; p++) fn(p); return; }
        while ((int)th_.size() < parts - 1) th_.emplace_back([this, id = (int)th_.size() + 1] { loop(id); });
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn; parts_ = parts; left_ = parts - 1; gen_++;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [this] { return left_ == 0; });
        fn_ = nullptr;
    }
private:
    void loop(int id)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *fn = nullptr;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
                if (id < parts_) fn = fn_;
            }
            if (fn) {
                (*fn)(id);
                std::lock_guard<std::mutex> l(m_);
                if (--left_ == 0) done_.notify_one();
            }
        }
    }
    std::mutex m_, call_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> th_;
    const std::function<void(int)> *fn_ = nullptr;
    int parts_ = 0, left_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

}  // namespace gact
