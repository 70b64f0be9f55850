/* STAND-IN for martinus/unordered_dense v3.0.1 (absent from /root/reference; pinned at
 * src/hashing/CMakeLists.txt.in:9).  TEST INFRASTRUCTURE: lets the reference's own
 * src/hashing/syncmer.cpp compile in place for oracle/_ref.  Only the members the reference calls exist.
 * PARITY UNPINNED: wyhash::hash(uint64_t) restates the published v3.0.1 algorithm
 * (mix(x, 0x9E3779B97F4A7C15); mix = lo64(a*b) ^ hi64(a*b)); `set` keeps the documented property that
 * iteration is in insertion order (dense vector of values + index). */
#pragma once
#include <cstdint>
#include <cstddef>
#include <unordered_map>
#include <vector>
#include <string>
namespace ankerl::unordered_dense
{
namespace detail::wyhash
{
[[nodiscard]] static inline uint64_t mix(uint64_t a, uint64_t b)
{
    __uint128_t r = a;
    r *= b;
    return static_cast<uint64_t>(r) ^ static_cast<uint64_t>(r >> 64U);
}
[[nodiscard]] static inline uint64_t hash(uint64_t x) { return mix(x, UINT64_C(0x9E3779B97F4A7C15)); }
} // namespace detail::wyhash

template <typename Key>
class set
{
    std::vector<Key> values_;
    std::unordered_map<Key, size_t> index_;
public:
    using value_type = Key;
    using iterator = typename std::vector<Key>::iterator;
    using const_iterator = typename std::vector<Key>::const_iterator;
    std::pair<iterator, bool> insert(Key const & k)
    {
        auto [it, fresh] = index_.try_emplace(k, values_.size());
        if (fresh)
            values_.push_back(k);
        return {values_.begin() + static_cast<std::ptrdiff_t>(it->second), fresh};
    }
    bool contains(Key const & k) const { return index_.count(k) != 0; }
    size_t size() const { return values_.size(); }
    bool empty() const { return values_.empty(); }
    void clear() { values_.clear(); index_.clear(); }
    iterator begin() { return values_.begin(); }
    iterator end() { return values_.end(); }
    const_iterator begin() const { return values_.begin(); }
    const_iterator end() const { return values_.end(); }
};
} // namespace ankerl::unordered_dense
