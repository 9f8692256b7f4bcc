// Host-side voxel map with the reference's API surface (src/VoxelRT/VoxelMap.h:9-264), written for
// this repository: same type and member names, same indexing rules, same dirty-tracking contract,
// so code written against the reference's VoxelMap compiles against this one.  No glm, no SIMD
// dispatch helpers: the bulk-edit paths are scalar (the device does the heavy lifting).
//
//   Voxel / Material::GetEncoded      VoxelMap.h:9-51
//   LinearIndexer3D + the 3 indexers  VoxelMap.h:60-102   (x fastest, then z, then y)
//   Brick / Sector                    VoxelMap.h:110-169, VoxelMap.cpp:6-66
//   VoxelMap                          VoxelMap.h:180-264, VoxelMap.cpp:98-123
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

namespace vrt_host {

struct ivec3 {
    int32_t x = 0, y = 0, z = 0;
    ivec3() = default;
    ivec3(int32_t x_, int32_t y_, int32_t z_) : x(x_), y(y_), z(z_) {}
    ivec3 operator>>(int s) const { return {x >> s, y >> s, z >> s}; }
};
struct dvec3 {
    double x = 0, y = 0, z = 0;
};

struct Voxel {
    uint8_t Data = 0;
    bool IsEmpty() const { return Data == 0; }
    static Voxel CreateEmpty() { return {}; }
    static Voxel Create(uint32_t materialId) {
        assert(materialId < 256);
        return Voxel{(uint8_t)materialId};
    }
};

// IEEE binary32 -> binary16, round to nearest even (what glm::packHalf2x16 does, VoxelMap.h:36)
inline uint16_t FloatToHalf(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7FFFFFFFu;
    if (ax > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);
    if (ax >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);
    if (ax >= 0x38800000u) {
        uint32_t mant = ax & 0x7FFFFFu, h = (((ax >> 23) - 112) << 10) | (mant >> 13), rem = mant & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
        return (uint16_t)(sign | h);
    }
    if (ax < 0x33000000u) return (uint16_t)sign;
    uint32_t mant = (ax & 0x7FFFFFu) | 0x800000u, shift = 126 - (ax >> 23), h = mant >> shift, rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

struct Material {
    uint8_t Color[3] = {0, 0, 0};
    uint8_t MetalFuzziness = 255;
    float Emission = 0.0f;
    // RGB565 | f16 emission << 16 | fuzziness << 32 (VoxelMap.h:27-41)
    uint64_t GetEncoded() const {
        uint64_t p = 0;
        p |= (uint64_t)(Color[0] >> 3) << 11;
        p |= (uint64_t)(Color[1] >> 2) << 5;
        p |= (uint64_t)(Color[2] >> 3);
        p |= (uint64_t)FloatToHalf(Emission) << 16;
        p |= (uint64_t)MetalFuzziness << 32;
        return p;
    }
};

template <int ShiftXZ_, int ShiftY_, bool Signed_>
struct LinearIndexer3D {
    static constexpr int32_t ShiftXZ = ShiftXZ_, ShiftY = ShiftY_;
    static constexpr int32_t SizeXZ = 1 << ShiftXZ, SizeY = 1 << ShiftY;
    static constexpr int32_t MaskXZ = SizeXZ - 1, MaskY = SizeY - 1;
    static constexpr size_t MaxArea = (size_t)1 << (ShiftXZ * 2 + ShiftY);
    static bool CheckInBounds(ivec3 p) {
        if (Signed_) p = {p.x + SizeXZ / 2, p.y + SizeY / 2, p.z + SizeXZ / 2};
        return (uint32_t)(p.x | p.z) < (uint32_t)SizeXZ && (uint32_t)p.y < (uint32_t)SizeY;
    }
    static uint32_t GetIndex(ivec3 p) { return (uint32_t)((p.x & MaskXZ) | (p.z & MaskXZ) << ShiftXZ | (p.y & MaskY) << (ShiftXZ * 2)); }
    static ivec3 GetPos(uint32_t index) {
        if (Signed_) {
            int32_t x = (int32_t)(index << (32 - ShiftXZ)) >> (32 - ShiftXZ);
            int32_t z = (int32_t)(index << (32 - ShiftXZ * 2)) >> (32 - ShiftXZ);
            int32_t y = (int32_t)(index << (32 - ShiftXZ * 2 - ShiftY)) >> (32 - ShiftY);
            return {x, y, z};
        }
        return {(int32_t)(index & MaskXZ), (int32_t)(index >> (ShiftXZ * 2) & MaskY), (int32_t)(index >> ShiftXZ & MaskXZ)};
    }
};
using WorldSectorIndexer = LinearIndexer3D<12, 8, true>;
using MaskIndexer = LinearIndexer3D<2, 2, false>;
using BrickIndexer = LinearIndexer3D<3, 3, false>;

struct Brick {
    Voxel Data[BrickIndexer::MaxArea] = {};
    bool IsEmpty() const {
        for (const Voxel& v : Data)
            if (!v.IsEmpty()) return false;
        return true;
    }
};
static_assert(sizeof(Brick) == 512, "Brick must be 512 bytes (the ABI's brick payload)");

struct Sector {
    std::vector<Brick> Storage;
    uint8_t BrickSlots[64] = {};
    Brick* GetBrick(uint32_t index, bool create = false) {
        uint8_t& slot = BrickSlots[index];
        if (slot != 0) return &Storage[slot - 1u];
        if (!create) return nullptr;
        slot = (uint8_t)(Storage.size() + 1);
        Storage.emplace_back();
        return &Storage.back();
    }
    uint64_t GetAllocationMask() const {
        uint64_t m = 0;
        for (uint32_t i = 0; i < 64; i++) m |= (uint64_t)(BrickSlots[i] != 0) << i;
        return m;
    }
    void DeleteBricks(uint64_t mask) {
        if (mask == 0) return;
        Sector kept;
        for (uint64_t m = GetAllocationMask() & ~mask; m; m &= m - 1) {
            uint32_t i = (uint32_t)__builtin_ctzll(m);
            *kept.GetBrick(i, true) = *GetBrick(i);
        }
        *this = std::move(kept);
    }
    uint64_t DeleteEmptyBricks(uint64_t mask = ~0ull) {
        uint64_t empty = 0;
        for (uint64_t m = mask & GetAllocationMask(); m; m &= m - 1) {
            uint32_t i = (uint32_t)__builtin_ctzll(m);
            if (GetBrick(i)->IsEmpty()) empty |= 1ull << i;
        }
        DeleteBricks(empty);
        return empty;
    }
};

struct HitResult {
    double Distance = -1.0;
    float Normal[3] = {0, 0, 0};
    float UV[2] = {0, 0};
    ivec3 VoxelPos;
    bool IsMiss() const { return Distance <= 0.0; }
};

struct VoxelMap {
    std::unordered_map<uint32_t, Sector> Sectors;
    std::map<uint32_t, uint64_t> DirtyLocs;  // sector index -> 4x4x4 mask of dirty bricks
    Material Palette[256] = {};

    // pos in brick coordinates (VoxelMap.cpp:98-123); like the reference, an existing sector always
    // yields a brick (created on demand — quirk Q5)
    Brick* GetBrick(ivec3 pos, bool create = false, bool markAsDirty = false) {
        ivec3 sectorPos = pos >> MaskIndexer::ShiftXZ;
        if (!WorldSectorIndexer::CheckInBounds(sectorPos)) return nullptr;
        uint32_t sectorIdx = WorldSectorIndexer::GetIndex(sectorPos), brickIdx = MaskIndexer::GetIndex(pos);
        Sector* sector;
        auto it = Sectors.find(sectorIdx);
        if (it != Sectors.end()) sector = &it->second;
        else if (create) sector = &Sectors[sectorIdx];
        else return nullptr;
        if (markAsDirty) DirtyLocs[sectorIdx] |= 1ull << brickIdx;
        return sector->GetBrick(brickIdx, true);
    }
    Voxel Get(ivec3 pos) {
        Brick* b = GetBrick(pos >> BrickIndexer::ShiftXZ);
        return b ? b->Data[BrickIndexer::GetIndex(pos)] : Voxel::CreateEmpty();
    }
    void Set(ivec3 pos, Voxel v) {
        Brick* b = GetBrick(pos >> BrickIndexer::ShiftXZ, true, true);
        if (b) b->Data[BrickIndexer::GetIndex(pos)] = v;
    }
    void MarkAllDirty() {
        for (auto& [idx, sector] : Sectors) DirtyLocs[idx] = sector.GetAllocationMask();
    }
    // Scalar region edit (RegionDispatchSIMD, VoxelMap.h:222-263): fn(x, y, z, Voxel&) -> bool changed.
    // Marks changed/empty bricks dirty and garbage-collects empty bricks and sectors.
    template <typename F>
    void RegionDispatch(ivec3 regionMin, ivec3 regionMax, bool createEmpty, F fn) {
        std::unordered_map<uint32_t, uint64_t> emptyBricks;
        ivec3 bmin = regionMin >> 3, bmax = regionMax >> 3;
        for (int32_t by = bmin.y; by <= bmax.y; by++)
            for (int32_t bz = bmin.z; bz <= bmax.z; bz++)
                for (int32_t bx = bmin.x; bx <= bmax.x; bx++) {
                    ivec3 bp(bx, by, bz);
                    Brick* brick = GetBrick(bp, createEmpty);
                    if (!brick) continue;
                    bool changed = false;
                    for (uint32_t i = 0; i < 512; i++) {
                        int32_t x = bx * 8 + (int32_t)(i & 7), z = bz * 8 + (int32_t)(i >> 3 & 7), y = by * 8 + (int32_t)(i >> 6);
                        if (x < regionMin.x || x > regionMax.x || y < regionMin.y || y > regionMax.y || z < regionMin.z || z > regionMax.z) continue;
                        changed |= fn(x, y, z, brick->Data[i]);
                    }
                    bool empty = brick->IsEmpty();
                    if (changed || empty) {
                        uint32_t sectorIdx = WorldSectorIndexer::GetIndex(bp >> 2);
                        uint64_t bit = 1ull << MaskIndexer::GetIndex(bp);
                        DirtyLocs[sectorIdx] |= bit;
                        if (empty) emptyBricks[sectorIdx] |= bit;
                    }
                }
        for (auto [sectorIdx, emptyMask] : emptyBricks) {
            Sector& s = Sectors[sectorIdx];
            if ((s.GetAllocationMask() & ~emptyMask) != 0) s.DeleteBricks(emptyMask);
            else Sectors.erase(sectorIdx);
        }
    }
};

}  // namespace vrt_host
