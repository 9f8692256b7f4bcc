#include "slot_allocator.h"

#include <algorithm>

namespace vrt {

void RangeArena::put_free(uint32_t base, uint32_t count) {
    free_[base] = count;
    by_size_.emplace(count, base);
}

void RangeArena::drop_free(std::map<uint32_t, uint32_t>::iterator it) {
    by_size_.erase(std::make_pair(it->second, it->first));
    free_.erase(it);
}

void RangeArena::reset(uint32_t capacity) {
    free_.clear();
    by_size_.clear();
    parked_.clear();
    capacity_ = capacity;
    allocated_ = 0;
    high_water_ = 0;
    if (capacity) put_free(0, capacity);
}

void RangeArena::grow(uint32_t new_capacity) {
    if (new_capacity <= capacity_) return;
    uint32_t old = capacity_;
    capacity_ = new_capacity;
    allocated_ += new_capacity - old;  // release() subtracts it again
    release(old, new_capacity - old);
}

uint32_t RangeArena::alloc(uint32_t count) {
    if (count == 0) return 0;
    // best fit = the smallest free range that holds `count` (least fragmentation), lowest
    // address among equals
    // (the size index makes this O(log n); an edit-heavy session leaves thousands of small free ranges behind)
    auto fit = by_size_.lower_bound(std::make_pair(count, 0u));
    if (fit == by_size_.end()) return kNone;
    uint32_t base = fit->second, size = fit->first;
    drop_free(free_.find(base));
    if (size > count) put_free(base + count, size - count);
    allocated_ += count;
    high_water_ = std::max(high_water_, base + count);
    return base;
}

bool RangeArena::extend(uint32_t base, uint32_t cur, uint32_t want) {
    if (want <= cur) return true;
    auto it = free_.find(base + cur);
    uint32_t need = want - cur;
    if (it == free_.end() || it->second < need) return false;
    uint32_t fbase = it->first, fsize = it->second;
    drop_free(it);
    if (fsize > need) put_free(fbase + need, fsize - need);
    allocated_ += need;
    high_water_ = std::max(high_water_, base + want);
    return true;
}

void RangeArena::release(uint32_t base, uint32_t count) {
    if (count == 0) return;
    allocated_ -= count;
    auto next = free_.lower_bound(base);
    // merge with the range that ends at `base`
    if (next != free_.begin()) {
        auto prev = std::prev(next);
        if (prev->first + prev->second == base) {
            base = prev->first;
            count += prev->second;
            drop_free(prev);
        }
    }
    // merge with the range that starts at the end
    if (next != free_.end() && base + count == next->first) {
        count += next->second;
        drop_free(next);
    }
    put_free(base, count);
}

void RangeArena::quarantine(uint32_t base, uint32_t count) {
    if (count) parked_.emplace_back(base, count);
}

void RangeArena::flush_quarantine() {
    for (auto& r : parked_) release(r.first, r.second);
    parked_.clear();
}

uint32_t RangeArena::largest_free() const {
    uint32_t m = 0;
    for (auto& r : free_) m = std::max(m, r.second);
    return m;
}

bool RangeArena::check_invariants() const {
    uint64_t free_total = 0;
    uint32_t prev_end = 0;
    bool first = true;
    for (auto& r : free_) {
        if (r.second == 0) return false;
        if (!first && r.first <= prev_end) return false;  // overlapping or not coalesced
        if ((uint64_t)r.first + r.second > capacity_) return false;
        prev_end = r.first + r.second;
        free_total += r.second;
        first = false;
    }
    if (by_size_.size() != free_.size()) return false;
    for (auto& r : by_size_) {
        auto it = free_.find(r.second);
        if (it == free_.end() || it->second != r.first) return false;
    }
    uint64_t parked = 0;
    for (auto& r : parked_) parked += r.second;
    return free_total + allocated_ == capacity_ && parked <= allocated_;
}

}  // namespace vrt
