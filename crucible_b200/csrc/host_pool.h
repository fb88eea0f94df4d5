// host_pool.h — process-wide cache of the big host blocks behind a scene's staging arrays.
//
// Why.  The reference rebuilds its world for every frame (Scene::render_image, src/scene/mod.rs:332-347), and so does a
// caller of this library that mirrors it: cr_scene_create -> cr_scene_add_* -> cr_scene_commit -> cr_render ->
// cr_scene_destroy per frame.  For the 10 M-triangle scene the staging arrays are 1.5 GB (vertices 720 MB, elements
// 640 MB, per-primitive metadata 160 MB); taken fresh from the OS every time they cost ~370 000 page faults plus the
// geometric regrowth of std::vector (measured: 1.65 s of a 3.2 s end-to-end frame, against 0.3 s for the commit).
// Blocks of HOST_POOL_MIN bytes and more therefore come from this cache and go back to it when a vector lets go; the
// second scene of a process touches memory that is already mapped.  cr_device_trim empties the cache; CRB_HOST_POOL_MB
// (default 6144) bounds what it may hold.  Smaller allocations use operator new as before.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <new>
#include <type_traits>
#include <utility>
#include <vector>

namespace crb {

static constexpr size_t HOST_POOL_MIN = (size_t)1 << 20;    // smaller requests bypass the cache
static constexpr size_t HOST_POOL_ALIGN = (size_t)2 << 20;  // block size granularity = alignment (transparent huge pages)

class HostBlockPool {
public:
    static HostBlockPool& instance() {
        static HostBlockPool* p = new HostBlockPool();  // never destroyed: vectors of static lifetime may outlive it otherwise
        return *p;
    }
    void* get(size_t bytes) {
        const size_t want = round_up(bytes);
        {
            std::lock_guard<std::mutex> lk(mu_);
            // best fit, but never a block more than twice the request (a 1 GB block must not serve a 3 MB vector)
            auto it = free_.lower_bound(want);
            if (it != free_.end() && it->first <= 2 * want) {
                void* p = it->second;
                cached_ -= it->first;
                live_[p] = it->first;
                free_.erase(it);
                ++hits_;
                return p;
            }
        }
        void* p = std::aligned_alloc(HOST_POOL_ALIGN, want);
        if (!p) throw std::bad_alloc();
        std::lock_guard<std::mutex> lk(mu_);
        live_[p] = want;
        ++misses_;
        return p;
    }
    void put(void* p) {
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto it = live_.find(p);
            if (it == live_.end()) {  // not ours (cannot happen: PoolAlloc routes by size on both sides)
                drop.push_back(p);
            } else {
                const size_t cap = it->second;
                live_.erase(it);
                free_.emplace(cap, p);
                cached_ += cap;
                while (cached_ > limit() && !free_.empty()) {  // over budget: the smallest blocks go first
                    auto sm = free_.begin();
                    cached_ -= sm->first;
                    drop.push_back(sm->second);
                    free_.erase(sm);
                }
            }
        }
        for (void* q : drop) std::free(q);
    }
    void trim() {
        std::multimap<size_t, void*> gone;
        {
            std::lock_guard<std::mutex> lk(mu_);
            gone.swap(free_);
            cached_ = 0;
        }
        for (auto& kv : gone) std::free(kv.second);
    }
    size_t cached_bytes() {
        std::lock_guard<std::mutex> lk(mu_);
        return cached_;
    }
    void counters(uint64_t& hits, uint64_t& misses) {
        std::lock_guard<std::mutex> lk(mu_);
        hits = hits_;
        misses = misses_;
    }

private:
    static size_t round_up(size_t b) { return (b + HOST_POOL_ALIGN - 1) / HOST_POOL_ALIGN * HOST_POOL_ALIGN; }
    static size_t limit() {
        static const size_t v = [] {
            const char* e = std::getenv("CRB_HOST_POOL_MB");
            const long long mb = e ? std::atoll(e) : 6144;
            return (size_t)(mb < 0 ? 0 : mb) << 20;
        }();
        return v;
    }
    std::mutex mu_;
    std::multimap<size_t, void*> free_;  // capacity -> block
    std::map<void*, size_t> live_;       // blocks handed out
    size_t cached_ = 0;
    uint64_t hits_ = 0, misses_ = 0;
};

// Types whose default construction a HostVec may skip entirely (every field is written before it is read); everything else
// is DEFAULT-initialised by resize() / the sizing constructor, i.e. trivial types (double, int32_t, uchar4, plain structs)
// are left uninitialised instead of being zero-filled: the arrays are hundreds of megabytes and are filled right after,
// often by several threads.  resize(n, value) and copies are unaffected.
template <typename T>
struct pool_no_init : std::false_type {};

template <typename T>
struct PoolAlloc {
    using value_type = T;
    template <typename U>
    void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
        if constexpr (!pool_no_init<U>::value) ::new (static_cast<void*>(p)) U;
    }
    template <typename U, typename A0, typename... A>
    void construct(U* p, A0&& a0, A&&... a) {
        ::new (static_cast<void*>(p)) U(std::forward<A0>(a0), std::forward<A>(a)...);
    }
    PoolAlloc() noexcept = default;
    template <typename U>
    PoolAlloc(const PoolAlloc<U>&) noexcept {}
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T);
        if (bytes >= HOST_POOL_MIN) return static_cast<T*>(HostBlockPool::instance().get(bytes));
        return static_cast<T*>(::operator new(bytes));
    }
    void deallocate(T* p, size_t n) noexcept {
        if (n * sizeof(T) >= HOST_POOL_MIN) HostBlockPool::instance().put(p);
        else ::operator delete(p);
    }
    template <typename U>
    bool operator==(const PoolAlloc<U>&) const noexcept { return true; }
    template <typename U>
    bool operator!=(const PoolAlloc<U>&) const noexcept { return false; }
};
// a std::vector whose big blocks come from the cache
template <typename T>
using HostVec = std::vector<T, PoolAlloc<T>>;

}  // namespace crb
