// Caching device allocator.  kvxopt's call pattern creates and destroys a factor object per linsolve call and, on the
// IPM path, once per iteration for the K = A S^-1 A' block (reference src/python/misc.py:1486): ~25 cudaMalloc +
// cudaFree pairs per handle cost more than the numeric work of the small configurations.  Freed blocks are kept per
// device in a size-ordered free list and handed out again when they are at most 2x the request; blocks above
// BIG_BLOCK bypass the cache (the multi-GB panels of the large configurations must go back to the driver).
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <unordered_map>

namespace b200s {

class DevPool {
  public:
    static DevPool& get() { static DevPool p; return p; }
    cudaError_t malloc(void** out, size_t bytes) {
        if (bytes == 0) bytes = 1;
        const size_t want = (bytes + 255) & ~(size_t)255;
        int dev = 0;
        cudaGetDevice(&dev);
        if (want <= BIG_BLOCK) {
            std::lock_guard<std::mutex> g(mu_);
            auto& fl = free_[dev];
            auto it = fl.lower_bound(want);
            if (it != fl.end() && it->first <= 2 * want) {
                *out = it->second;
                cached_ -= it->first;
                live_[it->second] = {it->first, dev};
                fl.erase(it);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(out, want);
        if (e == cudaErrorMemoryAllocation) {          // give the cache back to the driver and retry once
            cudaGetLastError();
            trim();
            e = cudaMalloc(out, want);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> g(mu_);
            live_[*out] = {want, dev};
        }
        return e;
    }
    // The caller guarantees that no kernel still uses the block (handles synchronise their stream before they die).
    void free(void* p) {
        if (!p) return;
        size_t sz = 0;
        int dev = 0;
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = live_.find(p);
            if (it == live_.end()) { cudaFree(p); return; }
            sz = it->second.first; dev = it->second.second;
            live_.erase(it);
            if (sz <= BIG_BLOCK && cached_ + sz <= CACHE_CAP) {
                free_[dev].emplace(sz, p);
                cached_ += sz;
                return;
            }
        }
        cudaFree(p);
    }
    void trim() {
        std::lock_guard<std::mutex> g(mu_);
        for (auto& d : free_) {
            for (auto& b : d.second) cudaFree(b.second);
            d.second.clear();
        }
        cached_ = 0;
    }

  private:
    static constexpr size_t BIG_BLOCK = (size_t)256 << 20, CACHE_CAP = (size_t)2 << 30;
    std::mutex mu_;
    std::map<int, std::multimap<size_t, void*>> free_;
    std::unordered_map<void*, std::pair<size_t, int>> live_;
    size_t cached_ = 0;
};

// Cached pinned host staging buffers: cudaHostAlloc / cudaFreeHost cost milliseconds each, which an object that lives for one
// small LP (a KKT solver per solvers.lp call) would pay every time.  Freed buffers are kept (grow-only, a few KB each).
class PinnedPool {
  public:
    static PinnedPool& get() { static PinnedPool p; return p; }
    cudaError_t malloc(void** out, size_t bytes) {
        const size_t want = ((bytes ? bytes : 1) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = free_.lower_bound(want);
            if (it != free_.end() && it->first <= 4 * want) { *out = it->second; live_[*out] = it->first; free_.erase(it); return cudaSuccess; }
        }
        cudaError_t e = cudaHostAlloc(out, want, cudaHostAllocPortable);
        if (e == cudaSuccess) { std::lock_guard<std::mutex> g(mu_); live_[*out] = want; }
        return e;
    }
    void free(void* p) {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu_);
        auto it = live_.find(p);
        if (it == live_.end()) { cudaFreeHost(p); return; }
        free_.emplace(it->second, p);
        live_.erase(it);
    }
  private:
    std::mutex mu_;
    std::multimap<size_t, void*> free_;
    std::unordered_map<void*, size_t> live_;
};
inline cudaError_t pinned_malloc(void** out, size_t bytes) { return PinnedPool::get().malloc(out, bytes); }
inline void pinned_free(void* p) { PinnedPool::get().free(p); }

inline cudaError_t pool_malloc(void** out, size_t bytes) { return DevPool::get().malloc(out, bytes); }
inline void pool_free(void* p) { DevPool::get().free(p); }

}  // namespace b200s
