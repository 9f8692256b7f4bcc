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
#include <utility>
#include <vector>

namespace vrt {

// Host bookkeeping only; what makes it worth a page of code is the edit workload: a frame of scattered voxel edits re-sizes ~1,500
// sectors, i.e. 1,500 alloc + 1,500 release per vrt_sync, and with node-based trees (std::map by address + std::set by size, round 1)
// those cost 0.8 us a pair — 1.2 ms per frame, three frame kernels' worth.  Now: free ranges live in two open-addressing tables
// (by first slot, by one-past-last slot: O(1) coalescing with both neighbours), and are indexed by size class in plain vectors —
// one class per exact size up to 64 slots (a sector's range is 1..64 bricks) and one per power of two above.  Index entries are not
// removed when a range is merged or taken by address; they are checked against the tables when popped (and swept now and then).
class RangeArena {
public:
    static constexpr uint32_t kNone = 0xFFFFFFFFu;

    explicit RangeArena(uint32_t capacity = 0) { reset(capacity); }

    void reset(uint32_t capacity);
    // Extends the arena to new_capacity slots (>= capacity()).
    void grow(uint32_t new_capacity);

    // `count` contiguous slots from the smallest size class that holds a fitting free range (exact fit first for count <= 64);
    // kNone when nothing fits.
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
    size_t free_ranges() const { return n_free_; }               // FreeList::FreeRanges.size()
    uint32_t largest_free() const;
    uint32_t high_water() const { return high_water_; }          // one past the highest slot ever handed out
    bool check_invariants() const;                               // ranges disjoint, coalesced, both tables and the index agree

private:
    // uint32 -> uint32, linear probing, backward-shift deletion (no tombstones); keys are slot numbers (< 2^31)
    struct Table {
        static constexpr uint32_t kEmpty = 0xFFFFFFFFu;
        std::vector<uint32_t> key, val;
        uint32_t mask = 0, live = 0;
        int shift = 32;
        void clear(uint32_t log2_size);
        uint32_t home(uint32_t k) const { return (k * 2654435761u) >> shift; }
        const uint32_t* find(uint32_t k) const;
        void put(uint32_t k, uint32_t v);
        void erase(uint32_t k);
        void rehash(uint32_t log2_size);
    };
    static constexpr int kClasses = 96;
    static int size_class(uint32_t n) { return n <= 64 ? (int)n : 64 + (26 - __builtin_clz(n)); }  // 65..127 -> 65, 128..255 -> 66, ...

    void put_free(uint32_t base, uint32_t count);
    void drop_free(uint32_t base, uint32_t count);  // the tables only; the index entry goes stale
    void sweep_index();

    Table by_base_;  // first slot -> count
    Table by_end_;   // one past the last slot -> first slot
    std::vector<uint32_t> index_[kClasses];  // first slots of free ranges of the class (may hold stale entries)
    uint64_t nonempty_[2] = {0, 0};          // classes whose vector is not empty
    size_t n_free_ = 0, n_index_ = 0;
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
