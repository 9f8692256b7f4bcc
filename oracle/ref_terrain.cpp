// oracle/_ref/libref_terrain.so — the reference's own terrain generator (src/VoxelRT/TerrainGenerator.cpp, compiled from where it lies, with
// its vendored FastNoise2 as built by scenes/Makefile) driven through its public interface: RequestSector / Poll, i.e. the worker threads,
// GenerateSector and the copy of the non-empty bricks (TerrainGenerator.cpp:5-34,126-147).  TEST INFRASTRUCTURE: pins scenes/terrain.py
// (the INPUT generator of BASELINE configs 1, 2, 4, 5) sector by sector.  Nothing is copied; see oracle/Makefile.
#include "VoxelMap.cpp"
#include "TerrainGenerator.cpp"
#include <chrono>
#include <map>
#include <tuple>
// n sector positions in, per sector: allocation mask and 64 x 512 voxel ids (brick b at [b], zeros where the sector has no such brick)
extern "C" __attribute__((visibility("default"))) int ref_terrain_generate(uint32_t n, const int32_t* xyz, uint64_t* masks, uint8_t* bricks) {
    auto map = std::make_shared<VoxelMap>();
    TerrainGenerator gen(map);
    std::map<std::tuple<int, int, int>, uint32_t> index;
    for (uint32_t i = 0; i < n; i++) {
        index[{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}] = i;
        gen.RequestSector(glm::ivec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    }
    for (uint32_t done = 0; done < n;) {
        auto r = gen.Poll();
        if (!r.second) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
            continue;
        }
        auto it = index.find({r.first.x, r.first.y, r.first.z});
        if (it == index.end()) return -1;
        const uint32_t i = it->second;
        masks[i] = r.second->GetAllocationMask();
        for (uint32_t b = 0; b < 64; b++) {
            Brick* br = r.second->GetBrick(b, false);
            if (br) std::memcpy(bricks + ((size_t)i * 64 + b) * 512, br->Data, 512);
            else std::memset(bricks + ((size_t)i * 64 + b) * 512, 0, 512);
        }
        done++;
    }
    return 0;
}
