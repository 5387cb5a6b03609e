// frame-copy-pool.hpp -- a few helper threads that split one large memcpy (FrameBatcher::push copies a frame of
// natoms*12 bytes from the caller's pageable array into pinned staging memory: one host thread moves ~7 GB/s, i.e.
// ~6 k frames/s of the 100k-atom box, below what one GPU takes end to end).
//
// copy() is synchronous: it returns when every byte is in place (the caller's frame buffer is reused by the
// trajectory reader right after push()).  Helpers spin briefly after a job -- frames arrive back to back when the
// producer is fast -- and then sleep on a condition variable, so an idle pool costs nothing.
#ifndef WN_FRAME_COPY_POOL_HPP
#define WN_FRAME_COPY_POOL_HPP

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

class FrameCopyPool
{
public:
    // helpers = threads besides the caller (0: plain memcpy on the caller's thread)
    explicit FrameCopyPool(int helpers) : gen_(0), remaining_(0), quit_(false), dst_(nullptr), src_(nullptr), bytes_(0), parts_(1)
    {
        // fewer helpers than asked for if the system has no more threads to give: the ones that started are used
        try { for (int k = 0; k < helpers; ++k) threads_.emplace_back([this, k]() { run(k + 1); }); }
        catch (const std::system_error&) {}
    }
    ~FrameCopyPool()
    {
        { std::lock_guard<std::mutex> g(m_); quit_ = true; }
        cv_.notify_all();
        for (std::thread& t : threads_) t.join();
    }
    FrameCopyPool(const FrameCopyPool&) = delete;
    FrameCopyPool& operator=(const FrameCopyPool&) = delete;
    int helpers() const { return (int)threads_.size(); }

    void copy(void* dst, const void* src, size_t bytes)
    {
        const int parts = (int)threads_.size() + 1;
        if (parts == 1 || bytes < (size_t)parts * 65536) { std::memcpy(dst, src, bytes); return; }
        {
            std::lock_guard<std::mutex> g(m_);
            dst_ = static_cast<char*>(dst); src_ = static_cast<const char*>(src); bytes_ = bytes; parts_ = parts;
            remaining_.store(parts - 1, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        copyPart(0, static_cast<char*>(dst), static_cast<const char*>(src), bytes, parts);
        while (remaining_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }

private:
    static void copyPart(int k, char* dst, const char* src, size_t bytes, int parts)
    {
        // equal parts on 4 KB boundaries; the last part takes the remainder
        const size_t each = ((bytes / (size_t)parts) + 4095) & ~(size_t)4095;
        const size_t lo = std::min(bytes, each * (size_t)k), hi = k + 1 == parts ? bytes : std::min(bytes, each * (size_t)(k + 1));
        if (hi > lo) std::memcpy(dst + lo, src + lo, hi - lo);
    }
    void run(int k)
    {
        uint64_t seen = 0;
        for (;;) {
            // spin for a short while (back-to-back frames), then sleep
            const auto spin_until = std::chrono::steady_clock::now() + std::chrono::microseconds(60);
            bool have = false;
            while (std::chrono::steady_clock::now() < spin_until) {
                if (gen_.load(std::memory_order_acquire) != seen) { have = true; break; }
            }
            if (!have) {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&]() { return quit_ || gen_.load(std::memory_order_acquire) != seen; });
                if (quit_) return;
            }
            char* dst; const char* src; size_t bytes; int parts;
            {
                std::lock_guard<std::mutex> g(m_);      // the job fields are written under the same mutex
                if (quit_) return;
                seen = gen_.load(std::memory_order_acquire);
                dst = dst_; src = src_; bytes = bytes_; parts = parts_;
            }
            if (k < parts) copyPart(k, dst, src, bytes, parts);
            remaining_.fetch_sub(1, std::memory_order_release);
        }
    }

    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_;
    std::atomic<int> remaining_;
    bool quit_;
    char* dst_;
    const char* src_;
    size_t bytes_;
    int parts_;
    std::vector<std::thread> threads_;
};

#endif
