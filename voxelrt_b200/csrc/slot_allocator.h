// Brick-slot arena for the sparse brickmap residency (host side).
//
// Plays the role of the reference's FreeList + BrickSlotAllocator
// (src/VoxelRT/BrickSlotAllocator.h:6-65, BrickSlotAllocator.cpp:5-94): every resident sector
// owns one contiguous range of brick slots, and brick i of the sector lives at
//   base + popcount(allocMask & ((1 << i) - 1))            (BrickSlotAllocator.h:37-41)
// so the device needs only {allocMask, base} per sector to find any brick.
//
// Differences by design (B200 residency): slots are 0-based, the arena grows
// (the device buffers are re-allocated and copied device-side, nothing is re-uploaded), and
// ranges released during one sync are quarantined until the sync's device-side moves have
// been issued, so a relocation can never overwrite a brick another relocation still reads.
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <utility>
#include <vector>

namespace vrt {

class RangeArena {
public:
    static constexpr uint32_t kNone = 0xFFFFFFFFu;

    explicit RangeArena(uint32_t capacity = 0) { reset(capacity); }

    void reset(uint32_t capacity);
    // Extends the arena to new_capacity slots (>= capacity()).
    void grow(uint32_t new_capacity);

    // Best-fit allocation of `count` contiguous slots; kNone when nothing fits.
    uint32_t alloc(uint32_t count);
    // Tries to extend [base, base+cur) to [base, base+want) in place.
    bool extend(uint32_t base, uint32_t cur, uint32_t want);
    // Returns a range to the arena immediately.
    void release(uint32_t base, uint32_t count);
    // Parks a range; it becomes allocatable again at flush_quarantine().
    void quarantine(uint32_t base, uint32_t count);
    void flush_quarantine();

    uint32_t capacity() const { return capacity_; }
    uint32_t allocated() const { return allocated_; }           // FreeList::NumAllocated
    size_t free_ranges() const { return free_.size(); }          // FreeList::FreeRanges.size()
    uint32_t largest_free() const;
    uint32_t high_water() const { return high_water_; }          // one past the highest slot ever handed out
    bool check_invariants() const;                               // ranges sorted, disjoint, coalesced

private:
    void put_free(uint32_t base, uint32_t count);
    void drop_free(std::map<uint32_t, uint32_t>::iterator it);

    std::map<uint32_t, uint32_t> free_;                  // base -> count
    std::set<std::pair<uint32_t, uint32_t>> by_size_;    // (count, base): best fit = lower_bound({count, 0}), lowest base among equals
    std::vector<std::pair<uint32_t, uint32_t>> parked_;
    uint32_t capacity_ = 0, allocated_ = 0, high_water_ = 0;
};

struct SectorSlots {
    uint64_t mask = 0;
    uint32_t base = 0;
};

inline uint32_t popcount64(uint64_t v) { return (uint32_t)__builtin_popcountll(v); }
inline uint32_t slot_of(const SectorSlots& s, uint32_t brick) {
    return s.base + popcount64(s.mask & ((1ull << brick) - 1));
}

}  // namespace vrt
