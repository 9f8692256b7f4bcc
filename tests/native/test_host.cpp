// Native (C++) tests of the host side: the VoxelMap mirror, the slot arena, and — with --gpu — the
// B200Renderer adapter end to end against the CPU oracle (edits included).  Test infrastructure.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../oracle/vrt_oracle.h"
#include "../../voxelrt_b200/csrc/gather_pool.h"
#include "../../voxelrt_b200/csrc/slot_allocator.h"
#include "../../voxelrt_b200/host/b200_renderer.h"

using namespace vrt_host;

static int g_fail = 0;
#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) {                                                     \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c);    \
            g_fail++;                                                   \
        }                                                               \
    } while (0)

static void test_indexers() {
    // x fastest, then z, then y (VoxelMap.h:94-97); signed world indexer round trip (:80-92)
    CHECK(BrickIndexer::GetIndex({1, 0, 0}) == 1 && BrickIndexer::GetIndex({0, 0, 1}) == 8 && BrickIndexer::GetIndex({0, 1, 0}) == 64);
    CHECK(MaskIndexer::GetIndex({3, 3, 3}) == 63 && MaskIndexer::GetIndex({5, 0, 0}) == 1);
    std::mt19937 rng(1);
    for (int i = 0; i < 10000; i++) {
        ivec3 p((int)(rng() % 4096) - 2048, (int)(rng() % 256) - 128, (int)(rng() % 4096) - 2048);
        ivec3 q = WorldSectorIndexer::GetPos(WorldSectorIndexer::GetIndex(p));
        CHECK(q.x == p.x && q.y == p.y && q.z == p.z);
        CHECK(WorldSectorIndexer::CheckInBounds(p));
    }
    CHECK(!WorldSectorIndexer::CheckInBounds({2048, 0, 0}) && !WorldSectorIndexer::CheckInBounds({0, -129, 0}));
}

static void test_voxel_map() {
    VoxelMap map;
    CHECK(map.Get({5, 6, 7}).IsEmpty());
    map.Set({5, 6, 7}, Voxel::Create(9));
    CHECK(map.Get({5, 6, 7}).Data == 9);
    CHECK(map.DirtyLocs.size() == 1 && map.DirtyLocs.begin()->second == 1ull);  // brick (0,0,0) of sector 0
    map.Set({40, 6, 7}, Voxel::Create(3));                                       // sector (1,0,0), brick (1,0,0)
    CHECK(map.DirtyLocs.size() == 2);
    CHECK(map.DirtyLocs[WorldSectorIndexer::GetIndex({1, 0, 0})] == 2ull);
    // quirk Q5: a read inside an existing sector allocates the brick (VoxelMap.cpp:122)
    uint64_t before = map.Sectors[0].GetAllocationMask();
    map.Get({20, 20, 20});
    CHECK(map.Sectors[0].GetAllocationMask() != before);
    // region clear garbage-collects empty bricks and sectors (VoxelMap.h:253-262)
    map.RegionDispatch({0, 0, 0}, {63, 31, 31}, false, [](int, int, int, Voxel& v) {
        bool ch = !v.IsEmpty();
        v = Voxel::CreateEmpty();
        return ch;
    });
    CHECK(map.Sectors.empty());
    Material m;
    m.Color[0] = 255, m.Color[1] = 128, m.Color[2] = 8, m.Emission = 0.8f, m.MetalFuzziness = 7;
    CHECK(m.GetEncoded() == orc_encode_material(255, 128, 8, 7, 0.8f));
}

static void test_arena() {
    vrt::RangeArena a(256);
    std::mt19937 rng(7);
    std::vector<std::pair<uint32_t, uint32_t>> live;
    for (int i = 0; i < 20000; i++) {
        if (live.empty() || rng() % 3) {
            uint32_t n = 1 + rng() % 64, base = a.alloc(n);
            if (base == vrt::RangeArena::kNone) {
                a.grow(a.capacity() * 2);
                base = a.alloc(n);
            }
            CHECK(base != vrt::RangeArena::kNone);
            for (auto& r : live) CHECK(base + n <= r.first || r.first + r.second <= base);  // disjoint
            live.push_back({base, n});
        } else {
            size_t k = rng() % live.size();
            if (rng() % 2) a.release(live[k].first, live[k].second);
            else {
                a.quarantine(live[k].first, live[k].second);
                a.flush_quarantine();
            }
            live.erase(live.begin() + (long)k);
        }
        if (i % 97 == 0) CHECK(a.check_invariants());
    }
    uint64_t sum = 0;
    for (auto& r : live) sum += r.second;
    CHECK(a.allocated() == sum && a.check_invariants());
    for (auto& r : live) a.release(r.first, r.second);
    CHECK(a.allocated() == 0 && a.free_ranges() == 1);  // fully coalesced
    // ranges beyond a sector's 64 slots (power-of-two size classes: a class member may be too small and must be passed over), growth
    // in place, and the steady churn of an edit session (sectors re-sized by +-1: the index must not pile up stale entries)
    live.clear();
    std::vector<uint8_t> used(a.capacity(), 0);
    auto mark = [&](uint32_t b, uint32_t n, uint8_t v) {
        if (used.size() < a.capacity()) used.resize(a.capacity(), 0);
        for (uint32_t i = b; i < b + n; i++) {
            CHECK(used[i] != v);
            used[i] = v;
        }
    };
    std::vector<std::pair<uint32_t, uint32_t>> parked;
    auto unpark = [&]() {
        a.flush_quarantine();
        for (auto& r : parked) mark(r.first, r.second, 0);
        parked.clear();
    };
    for (int i = 0; i < 60000; i++) {
        const unsigned op = rng() % 8;
        if (live.empty() || op < 3) {
            uint32_t n = (rng() % 16 == 0) ? 65 + rng() % 400 : 1 + rng() % 64, base = a.alloc(n);
            if (base == vrt::RangeArena::kNone) {
                a.grow(a.capacity() * 2);
                base = a.alloc(n);
            }
            CHECK(base != vrt::RangeArena::kNone && base + n <= a.capacity() && base + n <= a.high_water());
            mark(base, n, 1);
            live.push_back({base, n});
        } else if (op < 5) {
            auto& r = live[rng() % live.size()];
            uint32_t want = r.second + 1 + rng() % 3;
            bool room = r.first + want <= a.capacity();
            for (uint32_t j = r.first + r.second; room && j < r.first + want; j++) room = used[j] == 0;
            bool got = a.extend(r.first, r.second, want);
            CHECK(got == room);  // in place exactly when the slots behind the range are free
            if (got) mark(r.first + r.second, want - r.second, 1), r.second = want;
        } else {
            size_t k = rng() % live.size();
            mark(live[k].first, live[k].second, 2);  // parked: not yet allocatable
            parked.push_back(live[k]);
            a.quarantine(live[k].first, live[k].second);
            if (rng() % 4 == 0) unpark();
            live[k] = live.back(), live.pop_back();
        }
        if (i % 499 == 0) {
            unpark();
            CHECK(a.check_invariants());
        }
    }
    unpark();
    sum = 0;
    for (auto& r : live) sum += r.second;
    CHECK(a.allocated() == sum && a.check_invariants());
    for (auto& r : live) a.release(r.first, r.second);
    CHECK(a.allocated() == 0 && a.free_ranges() == 1 && a.largest_free() == a.capacity());
}

// vrt_sync's staging pool: every part of every job runs exactly once, on the workers AND the caller, job after job (the workers sleep in
// between), also after the pool was stopped and started again with another size, and for jobs smaller than the pool.
static void test_gather_pool() {
    vrt::GatherPool pool;
    CHECK(pool.workers() == 0);
    for (unsigned workers : {3u, 1u, 7u}) {
        pool.start(workers);
        CHECK(pool.workers() == workers);
        for (int job = 0; job < 200; job++) {
            const size_t n = (size_t)(job % 7 == 0 ? 2 : 1000 + 37 * job);
            std::vector<int> hits(n, 0);
            std::vector<unsigned> part_seen(workers + 1, 0);
            pool.run([&](unsigned part, unsigned parts) {
                CHECK(parts == workers + 1 && part < parts);
                part_seen[part]++;
                for (size_t i = n * part / parts; i < n * (part + 1) / parts; i++) hits[i] += 1 + job;
            });
            for (size_t i = 0; i < n; i++) CHECK(hits[i] == 1 + job);
            for (unsigned k = 0; k <= workers; k++) CHECK(part_seen[k] == 1);
        }
        pool.stop();
        CHECK(pool.workers() == 0);
    }
}

// ---- GPU: adapter vs oracle -----------------------------------------------------------------------
static void fill_scene(VoxelMap& map) {
    for (int z = 0; z < 160; z++)
        for (int x = 0; x < 160; x++) {
            int h = 20 + (int)(10.0 * std::sin(x * 0.07) + 8.0 * std::cos(z * 0.05) + 4.0 * std::sin((x + z) * 0.21));
            for (int y = 0; y <= h; y++) map.Set({x, y, z}, Voxel::Create(y + 2 >= h ? 245u + ((x ^ z) & 3) : 1u + ((x * 7 + y * 3 + z) % 200)));
        }
    for (int i = 1; i < 256; i++) {
        map.Palette[i].Color[0] = (uint8_t)(i * 37), map.Palette[i].Color[1] = (uint8_t)(i * 91), map.Palette[i].Color[2] = (uint8_t)(i * 53);
    }
    map.Palette[255].Emission = 10.0f;
}
static void oracle_sync_all(OrcMap* orc, VoxelMap& map) {
    for (auto& [idx, s] : map.Sectors) {
        ivec3 p = WorldSectorIndexer::GetPos(idx);
        std::vector<uint8_t> payload;
        uint64_t mask = s.GetAllocationMask();
        for (uint64_t m = mask; m; m &= m - 1) {
            const uint8_t* b = (const uint8_t*)s.GetBrick((uint32_t)__builtin_ctzll(m));
            payload.insert(payload.end(), b, b + 512);
        }
        VrtDirtySector r{p.x, p.y, p.z, 0, mask, mask, payload.data()};
        orc_map_sync(orc, 1, &r);
    }
    uint64_t enc[256];
    for (int i = 0; i < 256; i++) enc[i] = map.Palette[i].GetEncoded();
    orc_map_set_palette(orc, enc);
}
static bool frames_equal(B200Renderer& r, VoxelMap& map, Camera& cam, uvec2 size, uint32_t frameNo) {
    OrcMap* orc = orc_map_create(6, 4);  // rebuilt from scratch: the device's delta path must equal a full rebuild
    oracle_sync_all(orc, map);
    mat4 proj = cam.GetProjMatrix() * cam.GetViewMatrix(false);
    mat4 inv = GetInverseProjScreenMat(proj, size.x, size.y);
    VrtFrame f{};
    f.width = size.x, f.height = size.y;
    std::memcpy(f.inv_proj, inv.m, 64);
    std::memcpy(f.proj, proj.m, 64);
    const double p[3] = {cam.ViewPosition.x, cam.ViewPosition.y, cam.ViewPosition.z};
    for (int a = 0; a < 3; a++) f.world_origin[a] = (int32_t)std::floor(p[a]), f.origin_frac[a] = (float)(p[a] - std::floor(p[a]));
    f.frame_no = frameNo, f.bounces = 0, f.part_count = 1;
    std::vector<VrtTile> want((size_t)size.x * size.y / 16);
    orc_render(orc, &f, want.data(), nullptr, nullptr, 0, 0, size.y);
    orc_map_destroy(orc);
    return std::memcmp(want.data(), r.Tiles().data(), want.size() * sizeof(VrtTile)) == 0;
}
static void test_gpu() {
    auto map = std::make_shared<VoxelMap>();
    fill_scene(*map);
    B200Renderer r(map, 0);
    r.NumLightBounces = 0;
    Camera cam;
    cam.ViewPosition = {80.3, 60.7, 12.2};
    cam.Euler[0] = 0.1f, cam.Euler[1] = -0.5f;
    uvec2 size{322, 182};  // rounded down to 320x180
    cam.AspectRatio = 320.0f / 180.0f;
    r.RenderFrame(cam, size);
    CHECK(r.FrameSize().x == 320 && r.FrameSize().y == 180 && map->DirtyLocs.empty());
    CHECK(frames_equal(r, *map, cam, r.FrameSize(), r.FrameNo));
    // brush-like edits: carve a sphere, add a pillar -> only dirty bricks are uploaded, frame still equal
    map->RegionDispatch({60, 5, 40}, {100, 45, 80}, false, [](int x, int y, int z, Voxel& v) {
        int dx = x - 80, dy = y - 25, dz = z - 60;
        if (dx * dx + dy * dy + dz * dz > 18 * 18 || v.IsEmpty()) return false;
        v = Voxel::CreateEmpty();
        return true;
    });
    for (int y = 0; y < 90; y++) map->Set({70, y, 50}, Voxel::Create(255));
    CHECK(!map->DirtyLocs.empty());
    r.RenderFrame(cam, size);
    VrtStats st;
    vrt_get_stats(r.Handle(), &st);
    CHECK(st.bricks_uploaded > 0 && st.bricks_uploaded < st.resident_bricks / 4);  // delta, not a full re-upload
    CHECK(frames_equal(r, *map, cam, r.FrameSize(), r.FrameNo));
    // picking ray straight down onto the pillar top
    auto hits = r.RayCast({dvec3{70.5, 120.0, 50.5}}, {dvec3{0.0001, -1.0, 0.0001}});
    CHECK(!hits[0].IsMiss() && hits[0].VoxelPos.x == 70 && hits[0].VoxelPos.y == 89 && hits[0].VoxelPos.z == 50 && hits[0].Normal[1] == 1.0f);
    std::printf("gpu: %.1f Mrays/s (host-buffer frame), %llu bricks resident\n", r.RaysPerSecondOfLastFrame() / 1e6, (unsigned long long)st.resident_bricks);
}

// ---- GPU: RenderFrame through GBuffer::DenoiseAndPresent vs the two oracles chained ----------------------------
extern "C" {
struct PostOracle;
PostOracle* post_oracle_create(int w, int h);
void post_oracle_destroy(PostOracle*);
void post_oracle_set_camera(PostOracle*, const float proj[16], const float inv_proj[16], const double pos[3]);
void post_oracle_frame(PostOracle*, const uint32_t* tiles, int reset_history, int num_passes, int debug_channel, uint32_t* out_rgba8);
}
static void test_gpu_present() {
    auto map = std::make_shared<VoxelMap>();
    fill_scene(*map);
    B200Renderer r(map, 0);
    r.NumLightBounces = 0;
    r.DenoiseAndPresent = true;
    r.NumDenoiserPasses = 3;
    Camera cam;
    cam.Euler[0] = 0.1f, cam.Euler[1] = -0.5f;
    uvec2 size{320, 180};
    cam.AspectRatio = 320.0f / 180.0f;
    PostOracle* po = post_oracle_create(320, 180);
    OrcMap* orc = orc_map_create(6, 4);
    oracle_sync_all(orc, *map);
    for (int frame = 0; frame < 4; frame++) {
        cam.ViewPosition = {80.3 + 0.4 * frame, 60.7, 12.2 + 0.3 * frame};
        const bool worldChanged = !map->DirtyLocs.empty();  // frame 0: MarkAllDirty in the constructor
        r.RenderFrame(cam, size);
        mat4 proj = cam.GetProjMatrix() * cam.GetViewMatrix(false);
        mat4 inv = GetInverseProjScreenMat(proj, size.x, size.y);
        VrtFrame f{};
        f.width = size.x, f.height = size.y;
        std::memcpy(f.inv_proj, inv.m, 64);
        std::memcpy(f.proj, proj.m, 64);
        const double p[3] = {cam.ViewPosition.x, cam.ViewPosition.y, cam.ViewPosition.z};
        for (int a = 0; a < 3; a++) f.world_origin[a] = (int32_t)std::floor(p[a]), f.origin_frac[a] = (float)(p[a] - std::floor(p[a]));
        f.frame_no = r.FrameNo, f.bounces = 0, f.part_count = 1;
        std::vector<VrtTile> tiles((size_t)size.x * size.y / 16);
        orc_render(orc, &f, tiles.data(), nullptr, nullptr, 0, 0, size.y);
        std::vector<uint32_t> want((size_t)size.x * size.y);
        post_oracle_set_camera(po, proj.m, inv.m, p);
        post_oracle_frame(po, (const uint32_t*)tiles.data(), worldChanged ? 1 : 0, 3, 0, want.data());
        CHECK(r.Presented().size() == want.size());
        CHECK(std::memcmp(want.data(), r.Presented().data(), want.size() * 4) == 0);
    }
    orc_map_destroy(orc);
    post_oracle_destroy(po);
    std::printf("gpu-present: 4 frames through RenderFrame + DenoiseAndPresent equal the oracles\n");
}

int main(int argc, char** argv) {
    test_indexers();
    test_voxel_map();
    test_arena();
    test_gather_pool();
    if (argc > 1 && !std::strcmp(argv[1], "--gpu-present")) {
        try {
            test_gpu_present();
        } catch (const std::exception& e) {
            std::printf("FAIL exception: %s\n", e.what());
            g_fail++;
        }
    } else if (argc > 1 && !std::strcmp(argv[1], "--gpu")) {
        try {
            test_gpu();
        } catch (const std::exception& e) {
            std::printf("FAIL exception: %s\n", e.what());
            g_fail++;
        }
    } else {
        // without a GPU the adapter must refuse loudly, not fall back
        bool threw = false;
        try {
            B200Renderer r(std::make_shared<VoxelMap>(), 0);
        } catch (const std::runtime_error& e) {
            threw = std::strstr(e.what(), "no CPU fallback") != nullptr || std::strstr(e.what(), "CUDA") != nullptr;
        }
        if (!threw) std::printf("note: a CUDA device is present; run with --gpu for the device tests\n");
    }
    std::printf(g_fail ? "native tests: %d FAILED\n" : "native tests: ok\n", g_fail);
    return g_fail ? 1 : 0;
}
