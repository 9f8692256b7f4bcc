#include "slot_allocator.h"

#include <algorithm>

namespace vrt {

// ---- the table ----------------------------------------------------------------------------------------------------------------
void RangeArena::Table::clear(uint32_t log2_size) {
    key.assign((size_t)1 << log2_size, kEmpty);
    val.assign((size_t)1 << log2_size, 0);
    mask = (1u << log2_size) - 1, shift = 32 - (int)log2_size, live = 0;
}

const uint32_t* RangeArena::Table::find(uint32_t k) const {
    for (uint32_t i = home(k);; i = (i + 1) & mask) {
        if (key[i] == k) return &val[i];
        if (key[i] == kEmpty) return nullptr;
    }
}

void RangeArena::Table::rehash(uint32_t log2_size) {
    std::vector<uint32_t> k0, v0;
    k0.swap(key), v0.swap(val);
    clear(log2_size);
    for (size_t i = 0; i < k0.size(); i++)
        if (k0[i] != kEmpty) put(k0[i], v0[i]);
}

void RangeArena::Table::put(uint32_t k, uint32_t v) {
    if ((live + 1) * 2 > mask + 1) rehash((uint32_t)(32 - shift) + 1);  // load <= 1/2
    for (uint32_t i = home(k);; i = (i + 1) & mask) {
        if (key[i] == k) {
            val[i] = v;
            return;
        }
        if (key[i] == kEmpty) {
            key[i] = k, val[i] = v, live++;
            return;
        }
    }
}

void RangeArena::Table::erase(uint32_t k) {
    uint32_t i = home(k);
    for (;; i = (i + 1) & mask) {
        if (key[i] == k) break;
        if (key[i] == kEmpty) return;
    }
    live--;
    // close the gap: every later entry of the probe run moves up if its home position allows it
    for (uint32_t j = (i + 1) & mask;; j = (j + 1) & mask) {
        if (key[j] == kEmpty) break;
        const uint32_t h = home(key[j]);
        if (((j - h) & mask) >= ((j - i) & mask)) {  // home is at or before the gap (cyclically): may move
            key[i] = key[j], val[i] = val[j];
            i = j;
        }
    }
    key[i] = kEmpty;
}

// ---- free ranges --------------------------------------------------------------------------------------------------------------
void RangeArena::put_free(uint32_t base, uint32_t count) {
    by_base_.put(base, count);
    by_end_.put(base + count, base);
    const int c = size_class(count);
    index_[c].push_back(base);
    nonempty_[c >> 6] |= 1ull << (c & 63);
    n_free_++, n_index_++;
    if (n_index_ > 4 * n_free_ + 4096) sweep_index();
}

void RangeArena::drop_free(uint32_t base, uint32_t count) {
    by_base_.erase(base);
    by_end_.erase(base + count);
    n_free_--;
}

void RangeArena::sweep_index() {
    for (auto& v : index_) v.clear();
    nonempty_[0] = nonempty_[1] = 0;
    n_index_ = 0;
    for (size_t i = 0; i < by_base_.key.size(); i++) {
        if (by_base_.key[i] == Table::kEmpty) continue;
        const int c = size_class(by_base_.val[i]);
        index_[c].push_back(by_base_.key[i]);
        nonempty_[c >> 6] |= 1ull << (c & 63);
        n_index_++;
    }
    // (table order is a hash order; keep the index deterministic and low addresses first out)
    for (auto& v : index_) std::sort(v.begin(), v.end(), std::greater<uint32_t>());
}

void RangeArena::reset(uint32_t capacity) {
    by_base_.clear(10);
    by_end_.clear(10);
    for (auto& v : index_) v.clear();
    nonempty_[0] = nonempty_[1] = 0;
    n_free_ = n_index_ = 0;
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
    for (int c = size_class(count); c < kClasses;) {
        // next class at or above c that has index entries
        uint64_t w = nonempty_[c >> 6] & (~0ull << (c & 63));
        if (!w) {
            if (c < 64 && nonempty_[1]) w = nonempty_[1], c = 64;
            else return kNone;
        }
        c = (c & ~63) + __builtin_ctzll(w);
        std::vector<uint32_t>& v = index_[c];
        // classes of exact sizes and classes wholly above `count` fit by construction; the power-of-two class `count` itself falls
        // into (count > 64 only) must be searched
        const bool must_check = c > 64 && c == size_class(count);
        for (size_t k = v.size(); k-- > 0;) {
            const uint32_t base = v[k];
            const uint32_t* sz = by_base_.find(base);
            if (!sz || size_class(*sz) != c) {  // stale: merged away, taken by address, or re-freed with another size
                v[k] = v.back(), v.pop_back(), n_index_--;
                continue;
            }
            if (must_check && *sz < count) continue;
            const uint32_t size = *sz;
            v[k] = v.back(), v.pop_back(), n_index_--;
            drop_free(base, size);
            if (size > count) put_free(base + count, size - count);
            if (index_[c].empty()) nonempty_[c >> 6] &= ~(1ull << (c & 63));
            allocated_ += count;
            high_water_ = std::max(high_water_, base + count);
            return base;
        }
        if (v.empty()) nonempty_[c >> 6] &= ~(1ull << (c & 63));
        c++;
    }
    return kNone;
}

bool RangeArena::extend(uint32_t base, uint32_t cur, uint32_t want) {
    if (want <= cur) return true;
    const uint32_t* sz = by_base_.find(base + cur);
    uint32_t need = want - cur;
    if (!sz || *sz < need) return false;
    uint32_t fbase = base + cur, fsize = *sz;
    drop_free(fbase, fsize);
    if (fsize > need) put_free(fbase + need, fsize - need);
    allocated_ += need;
    high_water_ = std::max(high_water_, base + want);
    return true;
}

void RangeArena::release(uint32_t base, uint32_t count) {
    if (count == 0) return;
    allocated_ -= count;
    // merge with the range that ends at `base`
    if (const uint32_t* pb = by_end_.find(base)) {
        const uint32_t pbase = *pb, psize = base - pbase;
        drop_free(pbase, psize);
        base = pbase, count += psize;
    }
    // merge with the range that starts at the end
    if (const uint32_t* ns = by_base_.find(base + count)) {
        const uint32_t nsize = *ns;
        drop_free(base + count, nsize);
        count += nsize;
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
    for (size_t i = 0; i < by_base_.key.size(); i++)
        if (by_base_.key[i] != Table::kEmpty) m = std::max(m, by_base_.val[i]);
    return m;
}

bool RangeArena::check_invariants() const {
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    for (size_t i = 0; i < by_base_.key.size(); i++)
        if (by_base_.key[i] != Table::kEmpty) ranges.emplace_back(by_base_.key[i], by_base_.val[i]);
    if (ranges.size() != n_free_ || by_base_.live != n_free_ || by_end_.live != n_free_) return false;
    std::sort(ranges.begin(), ranges.end());
    uint64_t free_total = 0;
    uint32_t prev_end = 0;
    bool first = true;
    for (auto& r : ranges) {
        if (r.second == 0) return false;
        if (!first && r.first <= prev_end) return false;  // overlapping or not coalesced
        if ((uint64_t)r.first + r.second > capacity_) return false;
        const uint32_t* b = by_end_.find(r.first + r.second);
        if (!b || *b != r.first) return false;
        // every free range is reachable through the index of its class
        const std::vector<uint32_t>& v = index_[size_class(r.second)];
        if (std::find(v.begin(), v.end(), r.first) == v.end()) return false;
        if (!(nonempty_[size_class(r.second) >> 6] >> (size_class(r.second) & 63) & 1)) return false;
        prev_end = r.first + r.second;
        free_total += r.second;
        first = false;
    }
    size_t n_index = 0;
    for (auto& v : index_) n_index += v.size();
    if (n_index != n_index_) return false;
    uint64_t parked = 0;
    for (auto& r : parked_) parked += r.second;
    return free_total + allocated_ == capacity_ && parked <= allocated_;
}

}  // namespace vrt
