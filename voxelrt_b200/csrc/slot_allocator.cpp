#include "slot_allocator.h"

#include <algorithm>

namespace vrt {

void RangeArena::reset(uint32_t capacity) {
    free_.clear();
    parked_.clear();
    capacity_ = capacity;
    allocated_ = 0;
    high_water_ = 0;
    if (capacity) free_[0] = capacity;
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
    auto best = free_.end();
    for (auto it = free_.begin(); it != free_.end(); ++it) {
        if (it->second >= count && (best == free_.end() || it->second < best->second)) best = it;
    }
    if (best == free_.end()) return kNone;
    uint32_t base = best->first, size = best->second;
    free_.erase(best);
    if (size > count) free_[base + count] = size - count;
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
    free_.erase(it);
    if (fsize > need) free_[fbase + need] = fsize - need;
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
            free_.erase(prev);
        }
    }
    // merge with the range that starts at the end
    if (next != free_.end() && base + count == next->first) {
        count += next->second;
        free_.erase(next);
    }
    free_[base] = count;
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
    uint64_t parked = 0;
    for (auto& r : parked_) parked += r.second;
    return free_total + allocated_ == capacity_ && parked <= allocated_;
}

}  // namespace vrt
