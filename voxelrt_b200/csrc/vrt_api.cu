// C ABI of libvoxelrt_b200.so (include/voxelrt_b200.h): context, brickmap residency with
// sparse brick slots and dirty-brick delta upload, and the launches of the traversal kernels.
// Host side of the reference this replaces: FlatVoxelStorage / GpuVoxelStorage::SyncBuffers
// (src/VoxelRT/CpuRenderer.cpp:33-61, GpuRenderer.cpp:45-167) and the body of
// CpuRenderer::RenderFrame (CpuRenderer.cpp:415-464).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "gather_pool.h"
#include "slot_allocator.h"
#include "vrt_glsl.cuh"
#include "vrt_kernels.cuh"
#include "vrt_glsl_frame.cuh"

using namespace vrt;

namespace {

thread_local std::string g_create_error;

struct DeviceBuffer {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace

// VRT_SYNC_PROFILE=1 in the environment: where the HOST time of vrt_sync goes, by phase, printed when the context is destroyed
// (a diagnostic for the edit workload; costs seven clock reads per call when on, one branch when off)
struct SyncProfile {
    bool on = getenv("VRT_SYNC_PROFILE") != nullptr;
    double t[8] = {}, worst[8] = {};
    uint64_t calls = 0;
    std::chrono::steady_clock::time_point last;
    void start() {
        if (on) last = std::chrono::steady_clock::now(), calls++;
    }
    void mark(int phase) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        const double us = std::chrono::duration<double, std::micro>(now - last).count();
        t[phase] += us;
        if (us > worst[phase]) worst[phase] = us;
        last = now;
    }
    void report() const {
        if (!on || !calls) return;
        static const char* names[8] = {"validate (pass 1)", "arena growth + staging size", "commit (pass 2: arena, work lists)", "wait for staging halves",
                                       "gather into pinned staging", "H2D + kernel launches", "occupancy / boxes / quarantine", ""};
        // (the slowest call of every phase — the initial upload of the scene, with its one-time pinned allocation — is left out of the mean)
        fprintf(stderr, "[vrt_sync profile] %llu calls, host microseconds per call (mean without the slowest call | slowest call):\n", (unsigned long long)calls);
        for (int i = 0; i < 7; i++) fprintf(stderr, "  %-40s %8.1f | %10.1f\n", names[i], calls > 1 ? (t[i] - worst[i]) / (double)(calls - 1) : t[i], worst[i]);
    }
};
static SyncProfile g_sync_profile;

struct SyncUpload {  // vrt_sync: one brick to stage (payload in the caller's record, or the shared zero brick) and its slot
    const uint8_t* src;
    uint32_t slot;
};

struct VrtContext {
    int device = 0;
    std::vector<SyncUpload> sync_uploads;  // vrt_sync's work lists, kept between calls
    std::vector<uint2> sync_moves;
    std::vector<HeaderUpdate> sync_headers;
    std::vector<uint32_t> sync_seen;  // per sector: serial of the last vrt_sync call that named it (duplicate records)
    uint32_t sync_serial = 0;
    GatherPool gather_pool;  // started by the first vrt_sync that has enough bricks to share out
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H copies of finished bands overlap the next band's kernel
    cudaEvent_t ev_band[16] = {};
    cudaEvent_t ev_stage[2] = {};  // vrt_sync: a half of the staging buffers has been consumed
    bool stage_used[2] = {false, false};
    cudaEvent_t ev_sync = nullptr;    // end of the last vrt_sync / upload work on `stream`
    cudaEvent_t ev_render[16] = {};  // ends of the last 16 renders/traces on caller streams (callers may rotate several streams)
    uint32_t render_seq = 0;
    cudaStream_t gather_streams[VRT_GATHER_DEPTH] = {};  // a rank's copies to DIFFERENT presenting GPUs run concurrently (several copy engines)
    cudaEvent_t ev_gather_src[2 * VRT_GATHER_DEPTH] = {};  // vrt_render_gather: frame kernel done (ring), last gather copy done
    cudaEvent_t ev_gather_done[2 * VRT_GATHER_DEPTH] = {};
    uint32_t gather_seq = 0;
    bool gather_pending = false;
    bool render_pending = false;

    uint32_t sxz = 6, sy = 4, n_sectors = 0;
    uint32_t sxp = 0, syp = 0, n_hdr = 0;  // bordered header grid (one OUTSIDE sector on every side)
    // resident brickmap
    uint4* d_hdr = nullptr;        // entry 0 of the bordered grid, inside d_hdr_alloc
    uint4* d_hdr_alloc = nullptr;  // grid + a guard shell of OUTSIDE entries on both ends
    uint32_t hdr_guard = 0;
    uint32_t* d_occ = nullptr;     // 1 bit per header entry incl. the guards (k_build_occ)
    uint2* d_cells = nullptr;
    uint8_t* d_voxels = nullptr;
    uint2* d_palette = nullptr;
    uint32_t* d_albedo = nullptr;  // per palette entry: packed albedo of the primary-only frame kernel (k_palette_albedo)
    RangeArena arena;
    std::vector<SectorSlots> sectors;  // host mirror of d_hdr
    uint64_t resident_sectors = 0;
    bool have_palette = false;

    // staging (pinned host + device), grown on demand
    uint8_t* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    DeviceBuffer d_stage;
    // scratch for the host-pointer entry points
    DeviceBuffer d_rays_o, d_rays_d, d_hits, d_fb, d_aux, d_q_o, d_q_d, d_q_out;

    // shading inputs
    uint8_t* d_bn = nullptr;
    uint32_t* d_sky = nullptr;
    VrtSkyDesc sky{};

    uint32_t* d_sat = nullptr;  // summed-volume table of the box builder
    bool boxes_stale = true;    // some sector's emptiness changed since the boxes were built
    uint2* d_groups = nullptr;  // vrt_trace_glsl: SectorMasks of the GLSL renderer (one u64 per 4x4x4 sectors), built on demand
    uint2* d_lut = nullptr;     // vrt_trace_glsl: ray/cell interaction LUT
    bool groups_stale = true;
    cudaEvent_t ev_groups = nullptr;  // end of the last k_build_groups
    int macro_on = 1;  // 0 off, 1 on, 2 on + "metrics" launches count the macro loop's own trips (diagnostic)
    DevMetrics* d_metrics = nullptr;
    bool metrics_on = false;
    // persistent frame kernel: ring of ticket counters (one per launch in flight) and the host's view of each
    // frames with bounces: wavefront passes with trip budgets (k_wave_*, vrt_shade.cuh) against the one-thread-per-pixel kernel.
    // Bit-identical, but which is faster depends on the scene (open terrain, 2 bounces: +10..15 %; inside Sponza: -8..12 %), so
    // the default (2) measures: after every change of scene emptiness / bounce count / frame size the next four bounce frames
    // alternate between the forms between CUDA events, and the form with the faster frame is kept.  0 / 1 force a form.
    int gather_threads = 0;
    int wave_on = 2;
    int trace_refill = VRT_TRACE_REFILL;  // k_wave_trace: lanes in flight below which a warp refills (tuning knob, "trace_refill")
    int tile_order = 1;  // k_render: 0 top-to-bottom, 1 bottom-to-top (see FrameParams::work_add; "tile_order")
    int trace_ctas = VRT_TRACE_CTAS;      // k_wave_trace: resident CTAs per SM it is compiled for (8 / 10 / 12, "trace_ctas")
    int wave_choice = -1;         // -1 undecided, 0 per-pixel, 1 wavefront
    int wave_phase = 0;           // 0..3: the tuning frame to time next (even: per-pixel, odd: wavefront), 4: waiting for the events
    uint64_t wave_key = 0;        // (bounces, width, height, scene epoch) the decision was taken for
    uint64_t scene_epoch = 0;     // bumped when a sync changes some sector's emptiness
    cudaEvent_t ev_tune[8] = {};  // (begin, end) of the four tuning frames: per-pixel, wavefront, per-pixel, wavefront
    // wavefront frames: a ring of buffer sets, so that consecutive frames issued on different streams (multi-GPU: frames rotate over
    // streams to overlap one frame's tail with the next frame's start) do not wait for each other's queues
    struct WaveSet {
        DeviceBuffer rays, hits, path, n;
        cudaEvent_t ev = nullptr;          // last frame that used this set has finished
        cudaStream_t side = nullptr;       // k_wave_trace_generic runs beside k_wave_trace
        cudaEvent_t fork = nullptr, join = nullptr;
        bool used = false;
    };
    static constexpr uint32_t kWaveSets = 3;
    WaveSet wave[kWaveSets];
    uint32_t wave_seq = 0;
    int persist_on = 0;  // measured slower than the grid form on primary frames (tile order loses the L1 locality of 4 adjacent warp tiles per CTA)
    uint32_t* d_tickets = nullptr;
    uint32_t ticket_base[16] = {};
    uint32_t ticket_seq = 0;
    int sm_count = 0;

    std::vector<void*> exported, imported;
    VrtStats stats{};
    std::string err;
};

namespace {

int fail(VrtContext* c, int status, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return status;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(ctx, VRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Work issued by the caller on its own streams (up to 16 calls in flight) that residency changes must not overtake.
int wait_renders(VrtContext* ctx, cudaStream_t s) {
    for (uint32_t k = 0; k < 16u && k < ctx->render_seq; k++) CU(cudaStreamWaitEvent(s, ctx->ev_render[k], 0));
    return VRT_OK;
}
int sync_renders(VrtContext* ctx) {
    for (uint32_t k = 0; k < 16u && k < ctx->render_seq; k++) CU(cudaEventSynchronize(ctx->ev_render[k]));
    return VRT_OK;
}

int ensure(VrtContext* ctx, DeviceBuffer& b, size_t bytes) {
    if (b.bytes >= bytes) return VRT_OK;
    if (b.p) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFree(b.p));
        ctx->stats.device_bytes -= b.bytes;
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = std::max(bytes, (size_t)1 << 16);
    CU(cudaMalloc(&b.p, want));
    b.bytes = want;
    ctx->stats.device_bytes += want;
    return VRT_OK;
}

int ensure_host_stage(VrtContext* ctx, size_t bytes) {
    if (ctx->h_stage_bytes >= bytes) return VRT_OK;
    if (ctx->h_stage) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr;
    }
    size_t want = std::max(bytes + bytes / 2, (size_t)1 << 20);
    CU(cudaMallocHost((void**)&ctx->h_stage, want));
    ctx->h_stage_bytes = want;
    return VRT_OK;
}

// (Re)allocates the brick arena on the device to `capacity` slots, keeping resident bricks.
int resize_arena(VrtContext* ctx, uint32_t capacity) {
    uint32_t old_cap = ctx->arena.capacity();
    if (capacity <= old_cap && ctx->d_voxels) return VRT_OK;
    uint8_t* nv = nullptr;
    uint2* nc = nullptr;
    cudaError_t e = cudaMalloc((void**)&nv, (size_t)capacity * 512);
    if (e == cudaSuccess) e = cudaMalloc((void**)&nc, (size_t)capacity * 64);
    if (e != cudaSuccess) {
        if (nv) cudaFree(nv);
        cudaGetLastError();
        return fail(ctx, VRT_ERR_OOM, "Could not allocate brick slots");  // BrickSlotAllocator.cpp:15
    }
    if (ctx->d_voxels) {
        uint32_t used = ctx->arena.high_water();
        CU(cudaMemcpyAsync(nv, ctx->d_voxels, (size_t)used * 512, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(nc, ctx->d_cells, (size_t)used * 64, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->render_pending) { int st_ = sync_renders(ctx); if (st_) return st_; }
        CU(cudaFree(ctx->d_voxels));
        CU(cudaFree(ctx->d_cells));
        ctx->stats.device_bytes -= (size_t)old_cap * 576;
    }
    ctx->d_voxels = nv;
    ctx->d_cells = nc;
    ctx->stats.device_bytes += (size_t)capacity * 576;
    if (old_cap == 0) ctx->arena.reset(capacity);
    else ctx->arena.grow(capacity);
    return VRT_OK;
}

DevScene dev_scene(const VrtContext* ctx) {
    DevScene S;
    S.hdr = ctx->d_hdr;
    S.cells = ctx->d_cells;
    S.voxels = ctx->d_voxels;
    S.palette = ctx->d_palette;
    S.albedo = ctx->d_albedo;
    S.occ = ctx->d_occ;
    S.occ_bias = ctx->hdr_guard;
    S.sxz = ctx->sxz;
    S.sy = ctx->sy;
    S.lim_xz = 1u << (ctx->sxz + 5);
    S.lim_y = 1u << (ctx->sy + 5);
    S.sxp = ctx->sxp;
    S.sxzp = ctx->sxp * ctx->sxp;
    S.n_hdr = ctx->n_hdr;
    return S;
}

// Constants of the traversal frame for a world origin (see RayFrame).
RayFrame ray_frame(const VrtContext* ctx, const int32_t wo[3]) { return make_ray_frame(ctx->sxp, ctx->macro_on, wo); }

// Rebuilds the empty-sector boxes (k_box_*) on the context stream when they are stale.
int rebuild_boxes(VrtContext* ctx) {
    if (!ctx->boxes_stale) return VRT_OK;
    const uint32_t sxp = ctx->sxp, syp = ctx->syp, sxzp = sxp * sxp, n = ctx->n_hdr;
    if (!ctx->d_sat) {
        CU(cudaMalloc((void**)&ctx->d_sat, (size_t)n * sizeof(uint32_t)));
        ctx->stats.device_bytes += (size_t)n * sizeof(uint32_t);
    }
    cudaStream_t s = ctx->stream;
    k_box_occupancy<<<(n + 255) / 256, 256, 0, s>>>(ctx->d_hdr, ctx->d_sat, n);
    // along x: lines = (z, y); along z: lines = (x, y); along y: lines = (x, z)
    k_box_scan<<<(sxp * syp + 127) / 128, 128, 0, s>>>(ctx->d_sat, sxp * syp, sxp, 1u, sxp * syp, sxp, 0u);
    k_box_scan<<<(sxp * syp + 127) / 128, 128, 0, s>>>(ctx->d_sat, sxp * syp, sxp, sxp, sxp, 1u, sxzp);
    k_box_scan<<<(sxzp + 127) / 128, 128, 0, s>>>(ctx->d_sat, sxzp, syp, sxzp, sxzp, 1u, 0u);
    const uint32_t n_view = ctx->n_sectors;
    k_box_grow<<<(n_view + 127) / 128, 128, 0, s>>>(ctx->d_hdr, ctx->d_sat, sxp, sxzp, 1 << ctx->sxz, 1 << ctx->sy);
    CU(cudaGetLastError());
    ctx->stats.last_launches += 5;
    ctx->boxes_stale = false;
    return VRT_OK;
}

// Orders work on a caller stream after the context's residency work, and vice versa.
int begin_on_stream(VrtContext* ctx, cudaStream_t s) {
    if (s != ctx->stream) CU(cudaStreamWaitEvent(s, ctx->ev_sync, 0));
    return VRT_OK;
}
int end_on_stream(VrtContext* ctx, cudaStream_t s) {
    if (s != ctx->stream) {
        CU(cudaEventRecord(ctx->ev_render[ctx->render_seq++ & 15u], s));
        ctx->render_pending = true;
    }
    return VRT_OK;
}

int launch_trace(VrtContext* ctx, uint64_t n, const float* d_o, const float* d_d, const int32_t wo[3], uint32_t max_iters, VrtHit* d_out,
                 cudaStream_t s) {
    if (n == 0) return VRT_OK;
    if (max_iters == 0) max_iters = VRT_MAX_ITERS_DEFAULT;
    DevScene S = dev_scene(ctx);
    RayFrame W = ray_frame(ctx, wo);
    uint64_t blocks = (n + 127) / 128;
    if (blocks > 0x7FFFFFFFull) return fail(ctx, VRT_ERR_INVALID, "too many rays for one launch");
    if (ctx->metrics_on) {
        CU(cudaMemsetAsync(ctx->d_metrics, 0, sizeof(DevMetrics), s));
        k_trace<true><<<(unsigned)blocks, 128, 0, s>>>(S, W, d_o, d_d, max_iters, n, d_out, ctx->d_metrics);
    } else {
        k_trace<false><<<(unsigned)blocks, 128, 0, s>>>(S, W, d_o, d_d, max_iters, n, d_out, nullptr);
    }
    ctx->stats.last_launches = 1;
    CU(cudaGetLastError());
    return VRT_OK;
}

// Launches the frame kernel for macro-tile rows [row0, row1) (32-pixel rows; row1 = 0 means "to the end").
// Bands are only meaningful for an unpartitioned frame (part_count == 1).
// GenerateRayCellInteractionMaskLUT (GpuRenderer.cpp:193-210): cells of a 4x4x4 mask a ray of a given octant can reach from a cell
void interaction_lut(uint64_t table[512]) {
    for (int oct = 0; oct < 8; oct++) {
        const int sx = (oct & 1) ? 1 : -1, sy = (oct & 2) ? 1 : -1, sz = (oct & 4) ? 1 : -1;
        for (int origin = 0; origin < 64; origin++) {
            const int ox = origin & 3, oz = (origin >> 2) & 3, oy = origin >> 4;
            uint64_t m = 0;
            for (int jy = 0; jy < 4; jy++)
                for (int jz = 0; jz < 4; jz++)
                    for (int jx = 0; jx < 4; jx++) {
                        const int x = ox + jx * sx, y = oy + jy * sy, z = oz + jz * sz;
                        if (x >= 0 && x < 4 && y >= 0 && y < 4 && z >= 0 && z < 4) m |= 1ull << (x + 4 * z + 16 * y);
                    }
            table[origin + 64 * oct] = m;
        }
    }
}

// SectorMasks (the 128^3 level) and the interaction LUT of the GLSL renderer, built on demand on the context's stream (after the
// sync that changed the scene, which itself waited for the frames in flight); work on another stream `s` waits for the build
int glsl_scene(VrtContext* ctx, cudaStream_t s, GlslScene& G, uint64_t& launches) {
    if (ctx->sxz < 2 || ctx->sy < 2) return fail(ctx, VRT_ERR_UNSUPPORTED, "GLSL casts: the view must span at least 4 sectors per axis (128^3 level)");
    const uint32_t gxz = ctx->sxz - 2, n_groups = 1u << (2 * gxz + ctx->sy - 2);
    if (!ctx->d_lut) {
        uint64_t table[512];
        interaction_lut(table);
        CU(cudaMalloc((void**)&ctx->d_lut, sizeof(table)));
        CU(cudaMemcpy(ctx->d_lut, table, sizeof(table), cudaMemcpyHostToDevice));
    }
    if (!ctx->d_groups) CU(cudaMalloc((void**)&ctx->d_groups, (size_t)n_groups * 8));
    if (ctx->groups_stale) {
        k_build_groups<<<(n_groups + 127) / 128, 128, 0, ctx->stream>>>(dev_scene(ctx), ctx->d_groups, gxz, n_groups);
        CU(cudaGetLastError());
        if (!ctx->ev_groups) CU(cudaEventCreateWithFlags(&ctx->ev_groups, cudaEventDisableTiming));
        CU(cudaEventRecord(ctx->ev_groups, ctx->stream));
        ctx->groups_stale = false;
        launches++;
    }
    if (s != ctx->stream && ctx->ev_groups) CU(cudaStreamWaitEvent(s, ctx->ev_groups, 0));
    G = GlslScene{ctx->d_groups, ctx->d_lut, gxz};
    return VRT_OK;
}

int launch_render(VrtContext* ctx, const VrtFrame* f, void* d_out, VrtHit* d_aux, cudaStream_t s, uint32_t row0 = 0, uint32_t row1 = 0) {
    if (f->width == 0 || f->height == 0 || (f->width & 3u) || (f->height & 3u))
        return fail(ctx, VRT_ERR_INVALID, "frame size must be a non-zero multiple of 4 (CpuRenderer.cpp:419)");
    if (f->bounces > 7) return fail(ctx, VRT_ERR_INVALID, "bounces > 7");
    if ((f->flags & VRT_FRAME_COMPACT) && f->bounces != 0)
        return fail(ctx, VRT_ERR_INVALID, "VRT_FRAME_COMPACT drops the irradiance words, which are constant only for bounces == 0");
    if (f->bounces > 0 && ctx->d_bn == nullptr) return fail(ctx, VRT_ERR_STATE, "bounces > 0 needs vrt_set_blue_noise first");
    uint32_t part_count = f->part_count ? f->part_count : 1;
    if (f->part_index >= part_count) return fail(ctx, VRT_ERR_INVALID, "part_index >= part_count");

    FrameParams F;
    fill_frame_params(F, f, ctx->sxp, ctx->macro_on, ctx->d_bn, ctx->d_sky, &ctx->sky);
    F.out = d_out;
    F.aux = (f->flags & VRT_FRAME_AUX_HITS) ? d_aux : nullptr;
    F.metrics = ctx->d_metrics;

    if (!fill_frame_partition(F, f, row0, row1)) return fail(ctx, VRT_ERR_INVALID, "frame too large");
    if (F.n_work <= F.work_offset) return VRT_OK;
    F.work_add = ctx->tile_order ? F.n_work - 1u : F.work_offset;
    F.work_mul = ctx->tile_order ? -1 : 1;
    DevScene S = dev_scene(ctx);
    const unsigned wpb = VRT_RENDER_THREADS / 32;
    unsigned blocks = (F.n_work - F.work_offset + wpb - 1) / wpb;
    const bool rows = (F.flags & VRT_FRAME_PART_ROWS) != 0u, primary = F.bounces == 0;
    if (F.flags & VRT_FRAME_GLSL) {  // the GPU renderer's frame shader (vrt_glsl_frame.cuh): one kernel, one thread per pixel
        if (F.flags & (VRT_FRAME_COMPACT | VRT_FRAME_AUX_HITS)) return fail(ctx, VRT_ERR_INVALID, "VRT_FRAME_GLSL: no compact payload, no aux hits");
        GlslScene G;
        uint64_t launches = 0;
        int st = glsl_scene(ctx, s, G, launches);
        if (st) return st;
        const uint32_t cast_flags = (F.flags & VRT_FRAME_GLSL_ANISOTROPIC) ? VRT_GLSL_ANISOTROPIC : 0u;
        if (rows) k_render_glsl<true><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, G, F, cast_flags);
        else k_render_glsl<false><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, G, F, cast_flags);
        CU(cudaGetLastError());
        ctx->stats.last_launches += launches + 1;
        return VRT_OK;
    }
    if (ctx->metrics_on) CU(cudaMemsetAsync(ctx->d_metrics, 0, sizeof(DevMetrics), s));
    // which form traces a frame with bounces (see wave_on)
    bool use_wave = false;
    int tune_slot = -1;  // >= 0: this frame is one of the four timed ones (events tune_slot, tune_slot + 1)
    if (!primary && !ctx->metrics_on && !ctx->persist_on && ctx->wave_on) {
        if (ctx->wave_on == 1) use_wave = true;
        else if (row0 != 0 || row1 != 0) use_wave = ctx->wave_choice == 1;  // row-range launches of the band-pipelined host path never tune
        else {
            // (a rank of a screen split tunes on its own share of the frame; both forms give the same bytes, so ranks may differ)
            const uint64_t key = ((uint64_t)F.bounces << 56) ^ ((uint64_t)F.width << 40) ^ ((uint64_t)F.height << 24) ^ ((uint64_t)part_count << 60) ^
                                 ((uint64_t)(F.flags & VRT_FRAME_PART_ROWS) << 52) ^ (ctx->scene_epoch & 0xFFFFFFu);
            if (key != ctx->wave_key) ctx->wave_key = key, ctx->wave_choice = -1, ctx->wave_phase = 0;
            if (ctx->wave_choice < 0 && ctx->wave_phase == 4 && cudaEventQuery(ctx->ev_tune[7]) == cudaSuccess) {
                // two timed frames per form, alternating, the faster one of each counts: the very first frame of a form also pays
                // for its lazily loaded kernels and cold caches, which is not what the next thousand frames will see
                float t[4] = {};
                bool ok = true;
                for (int i = 0; i < 4; i++) ok = ok && cudaEventElapsedTime(&t[i], ctx->ev_tune[2 * i], ctx->ev_tune[2 * i + 1]) == cudaSuccess;
                if (ok) ctx->wave_choice = fminf(t[1], t[3]) < fminf(t[0], t[2]) ? 1 : 0;
                else ctx->wave_phase = 0;
                cudaGetLastError();
            }
            if (ctx->wave_choice >= 0) use_wave = ctx->wave_choice == 1;
            else if (ctx->wave_phase < 4) use_wave = (ctx->wave_phase & 1) != 0, tune_slot = 2 * ctx->wave_phase, ctx->wave_phase++;
            else use_wave = true;  // the timed frames are still in flight
        }
    }
    // (the timed region starts right before the first launch — after any one-time buffer allocation of the form)
    auto tune_begin = [&]() {
        if (tune_slot >= 0) cudaEventRecord(ctx->ev_tune[tune_slot], s);
    };
    struct TuneEnd {  // closes the timed region on every return path below
        VrtContext* c;
        int slot;
        cudaStream_t st;
        ~TuneEnd() {
            if (slot >= 0) cudaEventRecord(c->ev_tune[slot + 1], st);
        }
    } tune_end{ctx, tune_slot, s};
    if (!primary) ctx->stats.bounce_form = (use_wave ? 2u : 1u) | ((ctx->wave_on == 2 && !ctx->metrics_on && !ctx->persist_on && row0 == 0 && row1 == 0 && ctx->wave_choice < 0) ? 0x100u : 0u);
    if (use_wave) {
        // Wavefront: camera pass, then per bounce level a trace pass (persistent, lanes refilled from the level's queue) and a shade
        // pass (one thread per pixel in tile order: the packet coupling is two half-warp votes) — vrt_shade.cuh, vrt_kernels.cuh.
        const size_t cap = (size_t)(F.n_work - F.work_offset) * 32u;  // pixel slots of this launch; at most one ray per slot and level
        int st2;
        VrtContext::WaveSet& ws = ctx->wave[ctx->wave_seq++ % VrtContext::kWaveSets];
        const size_t path_bytes = cap * (sizeof(float4) + sizeof(float2)) + (cap / 16u) * sizeof(uint16_t) + 64u;
        const uint32_t n_counters = 3u * 10u;  // per level [0..9]: rays queued, refill cursor, generic rays queued
        if (ws.used && (ws.rays.bytes < cap * sizeof(RayRec) || ws.path.bytes < path_bytes)) CU(cudaEventSynchronize(ws.ev));  // growing: nobody may still use the set
        if ((st2 = ensure(ctx, ws.rays, cap * sizeof(RayRec)))) return st2;
        if ((st2 = ensure(ctx, ws.hits, cap * sizeof(HitRec)))) return st2;
        if (cap > 0x7FFFFFFFull) return fail(ctx, VRT_ERR_INVALID, "frame too large for the wavefront form");
        if ((st2 = ensure(ctx, ws.path, path_bytes))) return st2;
        if ((st2 = ensure(ctx, ws.n, n_counters * sizeof(uint32_t)))) return st2;
        if (ws.used) CU(cudaStreamWaitEvent(s, ws.ev, 0));
        tune_begin();
        uint32_t* cnt = static_cast<uint32_t*>(ws.n.p);
        CU(cudaMemsetAsync(cnt, 0, n_counters * sizeof(uint32_t), s));
        WaveBuffers B;
        B.rays = static_cast<RayRec*>(ws.rays.p);
        B.capacity = (uint32_t)cap;
        B.n_rays = cnt;
        B.head = cnt + 10;
        B.n_generic = cnt + 20;
        B.hits = static_cast<HitRec*>(ws.hits.p);
        B.path_a = static_cast<float4*>(ws.path.p);
        B.path_b = reinterpret_cast<float2*>(B.path_a + cap);
        B.pk_alive = reinterpret_cast<uint16_t*>(B.path_b + cap);
        if (rows) k_wave_primary<true><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F, B);
        else k_wave_primary<false><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F, B);
        ctx->stats.last_launches += 1;
        const int ctas = ctx->trace_ctas >= 12 ? 12 : (ctx->trace_ctas >= 10 ? 10 : 8);
        const unsigned grid = (unsigned)std::min<size_t>((cap + VRT_RENDER_THREADS - 1) / VRT_RENDER_THREADS, (size_t)ctx->sm_count * ctas);
        for (uint32_t level = 1; level <= F.bounces; level++) {
            TraceArgs A;
            A.rays = B.rays;
            A.n = B.n_rays + level;
            A.head = B.head + level;
            A.hits = B.hits;
            A.max_iters = F.max_iters;
            A.refill = (uint32_t)ctx->trace_refill;
            A.n_generic = B.n_generic + level;
            A.capacity = B.capacity;
            // the few generic rays of the level run beside the trace pass on a second stream (long serial chains, few warps)
            CU(cudaEventRecord(ws.fork, s));
            CU(cudaStreamWaitEvent(ws.side, ws.fork, 0));
            k_wave_trace_generic<<<(unsigned)ctx->sm_count * 4u, 32, 0, ws.side>>>(S, F.W, A);
            CU(cudaEventRecord(ws.join, ws.side));
            if (ctas == 12) k_wave_trace<12><<<grid, VRT_RENDER_THREADS, 0, s>>>(S, F.W, A);
            else if (ctas == 10) k_wave_trace<10><<<grid, VRT_RENDER_THREADS, 0, s>>>(S, F.W, A);
            else k_wave_trace<8><<<grid, VRT_RENDER_THREADS, 0, s>>>(S, F.W, A);
            CU(cudaStreamWaitEvent(s, ws.join, 0));
            if (rows) k_wave_shade<true><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F, B, level);
            else k_wave_shade<false><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F, B, level);
            ctx->stats.last_launches += 3;
        }
        CU(cudaEventRecord(ws.ev, s));
        ws.used = true;
        CU(cudaGetLastError());
        return VRT_OK;
    }
    tune_begin();
    if (false) {
    } else if (ctx->persist_on && !ctx->metrics_on && !rows) {
        // one resident grid; warps pull tiles from a ticket counter (see k_render_persist)
        const unsigned resident = (unsigned)ctx->sm_count * (unsigned)VRT_RENDER_CTAS(primary) * (unsigned)ctx->persist_on;
        const unsigned grid = std::min(blocks, resident);
        const uint32_t slot = ctx->ticket_seq++ & 15u;
        uint32_t* ticket = ctx->d_tickets + slot;
        const uint32_t base = ctx->ticket_base[slot];
        ctx->ticket_base[slot] += F.n_work - F.work_offset;
        if (primary) k_render_persist<true><<<grid, VRT_RENDER_THREADS, 0, s>>>(S, F, ticket, base);
        else k_render_persist<false><<<grid, VRT_RENDER_THREADS, 0, s>>>(S, F, ticket, base);
    } else {
#define VRT_LAUNCH(M, P, R) k_render<M, P, R><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F)
        const int variant = (ctx->metrics_on ? 4 : 0) | (primary ? 2 : 0) | (rows ? 1 : 0);
        // big views (header table >= 4 MB): bounce rays consult the one-bit sector table first (OCC kernels, see cast_loop_fast)
        const bool occ = (size_t)ctx->n_hdr * sizeof(uint4) >= ((size_t)4 << 20);
        switch (variant) {
            case 0:
                if (occ) k_render<false, false, false, true><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F);
                else VRT_LAUNCH(false, false, false);
                break;
            case 1:
                if (occ) k_render<false, false, true, true><<<blocks, VRT_RENDER_THREADS, 0, s>>>(S, F);
                else VRT_LAUNCH(false, false, true);
                break;
            case 2: VRT_LAUNCH(false, true, false); break;
            case 3: VRT_LAUNCH(false, true, true); break;
            case 4: VRT_LAUNCH(true, false, false); break;
            case 5: VRT_LAUNCH(true, false, true); break;
            case 6: VRT_LAUNCH(true, true, false); break;
            default: VRT_LAUNCH(true, true, true); break;
        }
#undef VRT_LAUNCH
    }
    ctx->stats.last_launches += 1;
    CU(cudaGetLastError());
    return VRT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
extern "C" int vrt_create(const VrtConfig* cfg, VrtContext** out) {
    VrtContext* ctx = nullptr;  // for the CU macro: errors go to g_create_error
    if (!cfg || !out) return fail(nullptr, VRT_ERR_INVALID, "null argument");
    if (cfg->struct_size != sizeof(VrtConfig)) return fail(nullptr, VRT_ERR_INVALID, "VrtConfig.struct_size mismatch");
    if (cfg->sectors_xz_log2 < 1 || cfg->sectors_xz_log2 > 10 || cfg->sectors_y_log2 < 1 || cfg->sectors_y_log2 > 10 ||
        2 * cfg->sectors_xz_log2 + cfg->sectors_y_log2 > 26)
        return fail(nullptr, VRT_ERR_INVALID, "view extent out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VRT_ERR_CUDA,
                    std::string("no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    int dev = cfg->device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= ndev) return fail(nullptr, VRT_ERR_INVALID, "device ordinal out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(nullptr, VRT_ERR_UNSUPPORTED, std::string("device ") + prop.name + " is not sm_100+; built for sm_100a only");

    VrtContext* c = new VrtContext();
    c->device = dev;
    c->sxz = cfg->sectors_xz_log2;
    c->sy = cfg->sectors_y_log2;
    c->n_sectors = 1u << (2 * c->sxz + c->sy);
    c->sectors.resize(c->n_sectors);
    DeviceGuard g(dev);
    ctx = c;
    auto bail = [&](int st) {
        g_create_error = c->err;
        vrt_destroy(c);
        return st;
    };
#define CUB(call)                                                                         \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return bail(VRT_ERR_CUDA);                                                    \
        }                                                                                 \
    } while (0)
    CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto& e : c->ev_band) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : c->ev_stage) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming));
    for (auto& e : c->ev_render) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& ws : c->wave) {
        CUB(cudaEventCreateWithFlags(&ws.ev, cudaEventDisableTiming));
        CUB(cudaStreamCreateWithFlags(&ws.side, cudaStreamNonBlocking));
        CUB(cudaEventCreateWithFlags(&ws.fork, cudaEventDisableTiming));
        CUB(cudaEventCreateWithFlags(&ws.join, cudaEventDisableTiming));
    }
    for (auto& e : c->ev_tune) CUB(cudaEventCreate(&e));
    for (auto& gs : c->gather_streams) CUB(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
    for (auto& e : c->ev_gather_src) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : c->ev_gather_done) CUB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->sxp = (1u << c->sxz) + 2u;
    c->syp = (1u << c->sy) + 2u;
    c->n_hdr = c->sxp * c->sxp * c->syp;
    // guard shell: two y-slabs of OUTSIDE entries before and after the bordered grid
    c->hdr_guard = 2u * c->sxp * c->sxp;
    CUB(cudaMalloc((void**)&c->d_hdr_alloc, ((size_t)c->n_hdr + 2u * c->hdr_guard) * sizeof(uint4)));
    c->d_hdr = c->d_hdr_alloc + c->hdr_guard;
    k_init_headers<<<(c->n_hdr + 2u * c->hdr_guard + 255) / 256, 256, 0, c->stream>>>(c->d_hdr, c->sxp, c->syp, c->hdr_guard);
    CUB(cudaMalloc((void**)&c->d_occ, (((size_t)c->n_hdr + 2u * c->hdr_guard + 31) / 32 + 1) * sizeof(uint32_t)));
    k_build_occ<<<(c->n_hdr + 2u * c->hdr_guard + 255) / 256, 256, 0, c->stream>>>(c->d_hdr_alloc, c->n_hdr + 2u * c->hdr_guard, c->d_occ);
    CUB(cudaGetLastError());
    CUB(cudaMalloc((void**)&c->d_palette, 256 * sizeof(uint2)));
    CUB(cudaMemsetAsync(c->d_palette, 0, 256 * sizeof(uint2), c->stream));
    CUB(cudaMalloc((void**)&c->d_albedo, 256 * sizeof(uint32_t)));
    CUB(cudaMemsetAsync(c->d_albedo, 0, 256 * sizeof(uint32_t), c->stream));
    CUB(cudaMalloc((void**)&c->d_tickets, 16 * sizeof(uint32_t)));
    CUB(cudaMemsetAsync(c->d_tickets, 0, 16 * sizeof(uint32_t), c->stream));
    CUB(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    CUB(cudaMalloc((void**)&c->d_metrics, sizeof(DevMetrics)));
    CUB(cudaMemsetAsync(c->d_metrics, 0, sizeof(DevMetrics), c->stream));
    c->stats.device_bytes = ((size_t)c->n_hdr + 2u * c->hdr_guard) * sizeof(uint4) + 256 * sizeof(uint2) + sizeof(DevMetrics);
    uint32_t cap = cfg->initial_brick_capacity ? cfg->initial_brick_capacity : (1u << 16);
    int st = resize_arena(c, cap);
    if (st != VRT_OK) return bail(st);
    CUB(cudaEventRecord(c->ev_sync, c->stream));
    CUB(cudaStreamSynchronize(c->stream));
#undef CUB
    *out = c;
    return VRT_OK;
}

extern "C" void vrt_destroy(VrtContext* ctx) {
    if (!ctx) return;
    g_sync_profile.report();
    ctx->gather_pool.stop();
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    for (void* p : ctx->imported) cudaIpcCloseMemHandle(p);
    for (void* p : ctx->exported) cudaFree(p);
    for (auto& ws : ctx->wave) {
        if (ws.ev) cudaEventDestroy(ws.ev);
        if (ws.fork) cudaEventDestroy(ws.fork);
        if (ws.join) cudaEventDestroy(ws.join);
        if (ws.side) cudaStreamDestroy(ws.side);
        for (DeviceBuffer* b : {&ws.rays, &ws.hits, &ws.path, &ws.n})
            if (b->p) cudaFree(b->p);
    }
    for (auto& e : ctx->ev_tune) if (e) cudaEventDestroy(e);
    if (ctx->ev_groups) cudaEventDestroy(ctx->ev_groups);
    DeviceBuffer* bufs[] = {&ctx->d_stage, &ctx->d_rays_o, &ctx->d_rays_d, &ctx->d_hits, &ctx->d_fb,
                            &ctx->d_aux,   &ctx->d_q_o,    &ctx->d_q_d,    &ctx->d_q_out};
    for (auto* b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    cudaFree(ctx->d_hdr_alloc);
    cudaFree(ctx->d_occ);
    cudaFree(ctx->d_cells);
    cudaFree(ctx->d_voxels);
    cudaFree(ctx->d_palette);
    cudaFree(ctx->d_albedo);
    cudaFree(ctx->d_tickets);
    for (auto& gs : ctx->gather_streams) if (gs) cudaStreamDestroy(gs);
    for (auto& e : ctx->ev_gather_src) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->ev_gather_done) if (e) cudaEventDestroy(e);
    cudaFree(ctx->d_bn);
    cudaFree(ctx->d_sky);
    cudaFree(ctx->d_metrics);
    cudaFree(ctx->d_sat);
    cudaFree(ctx->d_groups);
    cudaFree(ctx->d_lut);
    if (ctx->ev_sync) cudaEventDestroy(ctx->ev_sync);
    for (auto& e : ctx->ev_render) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& e : ctx->ev_band)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->ev_stage)
        if (e) cudaEventDestroy(e);
    cudaGetLastError();
    delete ctx;
}

extern "C" const char* vrt_last_error(const VrtContext* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int vrt_get_stats(const VrtContext* ctx, VrtStats* out) {
    if (!ctx || !out) return VRT_ERR_INVALID;
    *out = ctx->stats;
    out->resident_bricks = ctx->arena.allocated();
    out->brick_capacity = ctx->arena.capacity();
    out->free_ranges = ctx->arena.free_ranges();
    out->resident_sectors = ctx->resident_sectors;
    return VRT_OK;
}

extern "C" int vrt_set_option(VrtContext* ctx, const char* name, int64_t value) {
    if (!ctx || !name) return VRT_ERR_INVALID;
    if (!strcmp(name, "metrics")) ctx->metrics_on = value != 0;
    else if (!strcmp(name, "macro_steps")) ctx->macro_on = (int)value;
    else if (!strcmp(name, "persistent")) ctx->persist_on = (int)value;
    else if (!strcmp(name, "compact_bounces")) {
        // round 1's CTA-level re-dealing of bounce rays (measured slower) is gone: the wavefront trace pass refills lanes from a queue
        if (value != 0) return fail(ctx, VRT_ERR_UNSUPPORTED, "compact_bounces was removed; see the \"wavefront\" option");
    }
    else if (!strcmp(name, "wavefront")) ctx->wave_on = (int)value;
    else if (!strcmp(name, "gather_threads")) {  // host threads of vrt_sync's staging gather: 0 = min(8, cores), 1 = the calling thread only
        if (value < 0 || value > 64) return fail(ctx, VRT_ERR_INVALID, "gather_threads: 0..64");
        if ((int)value != ctx->gather_threads) ctx->gather_pool.stop();
        ctx->gather_threads = (int)value;
    }
    else if (!strcmp(name, "reserve_slots")) {
        // test hook: takes the first `value` brick slots out of the arena, so that a small scene lives at slot numbers a 10 GB scene
        // reaches (byte offsets beyond 2^32: every kernel must address bricks with 64-bit arithmetic).  Before the first vrt_sync.
        if (value < 0 || value > 0x7FFFFFFF || ctx->arena.allocated() != 0) return fail(ctx, VRT_ERR_STATE, "reserve_slots: only on an empty arena");
        DeviceGuard g(ctx->device);
        const uint64_t want = (uint64_t)value + ctx->arena.capacity();
        if (want > 0x7FFFFFFFull) return fail(ctx, VRT_ERR_OOM, "Could not allocate brick slots");
        int st = resize_arena(ctx, (uint32_t)want);
        if (st) return st;
        if (ctx->arena.alloc((uint32_t)value) == RangeArena::kNone) return fail(ctx, VRT_ERR_OOM, "Could not allocate brick slots");
    }
    else if (!strcmp(name, "trace_ctas")) ctx->trace_ctas = (int)value;
    else if (!strcmp(name, "tile_order")) ctx->tile_order = value != 0;
    else if (!strcmp(name, "trace_refill")) ctx->trace_refill = (int)std::min<int64_t>(32, std::max<int64_t>(1, value));
    else return fail(ctx, VRT_ERR_INVALID, std::string("unknown option ") + name);
    return VRT_OK;
}

extern "C" int vrt_get_metrics(VrtContext* ctx, VrtTraversalMetrics* out) {
    if (!ctx || !out) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    CU(cudaDeviceSynchronize());
    DevMetrics m;
    CU(cudaMemcpy(&m, ctx->d_metrics, sizeof(m), cudaMemcpyDeviceToHost));
    out->rays = m.rays;
    out->iters = m.iters;
    out->sector_fetches = m.sector_fetches;
    out->cell_fetches = m.cell_fetches;
    out->hits = m.hits;
    out->capped = m.capped;
    return VRT_OK;
}

// ------------------------------------------------------------------------------------------------
// residency
// ------------------------------------------------------------------------------------------------
extern "C" int vrt_set_palette(VrtContext* ctx, const uint64_t palette[256]) {
    if (!ctx || !palette) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    if (ctx->render_pending) { int st_ = wait_renders(ctx, ctx->stream); if (st_) return st_; }
    int st = ensure_host_stage(ctx, 256 * 8);
    if (st) return st;
    CU(cudaStreamSynchronize(ctx->stream));  // staging may still feed a previous copy
    memcpy(ctx->h_stage, palette, 256 * 8);
    CU(cudaMemcpyAsync(ctx->d_palette, ctx->h_stage, 256 * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_palette_albedo<<<1, 256, 0, ctx->stream>>>(ctx->d_palette, ctx->d_albedo);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_sync, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->have_palette = true;
    return VRT_OK;
}

extern "C" int vrt_sync(VrtContext* ctx, uint32_t n, const VrtDirtySector* recs) {
    if (!ctx || (n && !recs)) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    ctx->stats.bytes_uploaded = ctx->stats.bricks_uploaded = ctx->stats.bricks_relocated = 0;
    ctx->stats.last_launches = 0;
    if (n == 0) return VRT_OK;
    SyncProfile& prof = g_sync_profile;
    prof.start();

    using Upload = SyncUpload;
    // (work lists live with the context: a frame of edits lists ~25 k entries, and fresh vectors of that size come from mmap every call)
    std::vector<Upload>& uploads = ctx->sync_uploads;
    std::vector<uint2>& moves = ctx->sync_moves;
    std::vector<HeaderUpdate>& headers = ctx->sync_headers;
    uploads.clear(), moves.clear(), headers.clear();
    static const uint8_t kZeroBrick[VRT_BRICK_BYTES] = {};

    // The call is TRANSACTIONAL: pass 1 looks at every record without touching anything, and everything that can fail for
    // lack of memory (arena growth, staging buffers) happens between the passes — so a bad record or an allocation failure
    // returns with the host mirror, the slot arena and the device state exactly as they were.
    // Pass 1: payload pointers, duplicate sectors (two records for one sector would race in the single k_move / k_upload
    // launches below: the second record's moves read what the first one's moves write), and upper bounds for the sizes.
    uint64_t need_slots = 0, max_uploads = 0, max_moves = 0, max_headers = 0;
    {
        // duplicates: one stamp per sector of the view, holding the serial number of the last call that named the sector
        std::vector<uint32_t>& stamp = ctx->sync_seen;
        if (stamp.size() != ctx->sectors.size()) stamp.assign(ctx->sectors.size(), 0u);
        if (++ctx->sync_serial == 0u) {
            std::fill(stamp.begin(), stamp.end(), 0u);
            ctx->sync_serial = 1u;
        }
        const uint32_t serial = ctx->sync_serial;
        for (uint32_t r = 0; r < n; r++) {
            const VrtDirtySector& d = recs[r];
            // ViewSectorIndexer::CheckInBounds -> continue (CpuRenderer.cpp:40, GpuRenderer.cpp:54-55)
            if (((uint32_t)(d.sx | d.sz) >> ctx->sxz) != 0 || ((uint32_t)d.sy >> ctx->sy) != 0) continue;
            const uint32_t si = (uint32_t)d.sx | ((uint32_t)d.sz << ctx->sxz) | ((uint32_t)d.sy << (2 * ctx->sxz));
            if (stamp[si] == serial)
                return fail(ctx, VRT_ERR_INVALID, "the same sector appears in two records of one vrt_sync call (nothing was changed)");
            stamp[si] = serial;
            const SectorSlots& old = ctx->sectors[si];
            const uint64_t new_mask = (d.flags & VRT_SECTOR_REMOVED) ? 0ull : d.alloc_mask;  // GpuRenderer.cpp:59-67
            const uint64_t dirty = d.dirty_mask & new_mask;                                   // :62
            if (dirty && !d.bricks)
                return fail(ctx, VRT_ERR_INVALID, "record " + std::to_string(r) + ": dirty bricks without payload (nothing was changed)");
            const uint64_t fresh = new_mask & ~old.mask & ~dirty;  // allocated without content: zero-filled below
            max_uploads += popcount64(dirty) + popcount64(fresh);
            if (new_mask != old.mask) {
                need_slots += popcount64(new_mask);
                max_moves += popcount64(old.mask & new_mask & ~dirty);
                max_headers++;
            }
        }
    }
    prof.mark(0);
    // Every fresh range of this call fits behind the high-water mark (one coalesced free range): then no alloc() below can
    // fail, whatever the fragmentation.  Grown BEFORE the first mutation; a failure here leaves everything as it was.
    if (need_slots) {
        const uint64_t tail = (uint64_t)ctx->arena.capacity() - ctx->arena.high_water();
        if (tail < need_slots) {
            const uint64_t want = std::max<uint64_t>((uint64_t)ctx->arena.capacity() * 2, (uint64_t)ctx->arena.high_water() + need_slots);
            if (want > 0x7FFFFFFFull) return fail(ctx, VRT_ERR_OOM, "Could not allocate brick slots");  // BrickSlotAllocator.cpp:15
            int st = resize_arena(ctx, (uint32_t)want);
            if (st) return st;
        }
    }
    // staging for the largest chunk this call can need (same layout as below, from the upper bounds)
    const size_t kChunkBricks = (size_t)1 << 17;
    {
        const size_t mm = (max_moves * 8 + 15) & ~(size_t)15, mb = mm + max_headers * sizeof(HeaderUpdate);
        const size_t c0 = std::min<size_t>(max_uploads, kChunkBricks), half_max = ((c0 * 516 + 15) & ~(size_t)15) + mb + 16;
        if (max_uploads || mb) {
            int st = ensure_host_stage(ctx, 2 * half_max);
            if (st) return st;
            st = ensure(ctx, ctx->d_stage, 2 * half_max);
            if (st) return st;
        }
    }
    prof.mark(1);
    uploads.reserve(max_uploads);
    moves.reserve(max_moves);
    headers.reserve(max_headers);

    // Pass 2: commit to the host mirror and the arena, and list the device work.
    for (uint32_t r = 0; r < n; r++) {
        const VrtDirtySector& d = recs[r];
        if (((uint32_t)(d.sx | d.sz) >> ctx->sxz) != 0 || ((uint32_t)d.sy >> ctx->sy) != 0) continue;
        uint32_t si = (uint32_t)d.sx | ((uint32_t)d.sz << ctx->sxz) | ((uint32_t)d.sy << (2 * ctx->sxz));
        SectorSlots old = ctx->sectors[si];
        uint64_t new_mask = (d.flags & VRT_SECTOR_REMOVED) ? 0ull : d.alloc_mask;
        uint64_t dirty = d.dirty_mask & new_mask;

        SectorSlots cur = old;
        if (new_mask != old.mask) {
            uint32_t old_n = popcount64(old.mask), new_n = popcount64(new_mask);
            // every new brick sorts after every resident one -> resident slots keep their place
            bool appended = old.mask != 0 && (new_mask & old.mask) == old.mask &&
                            (uint32_t)__builtin_ctzll(new_mask & ~old.mask) > 63u - (uint32_t)__builtin_clzll(old.mask);
            if (new_n == 0) {
                ctx->arena.quarantine(old.base, old_n);
                cur.base = 0;
            } else if (appended && ctx->arena.extend(old.base, old_n, new_n)) {
                // new bricks all sort after the resident ones: the range grows in place, nothing moves
            } else {
                uint32_t base = ctx->arena.alloc(new_n);
                if (base == RangeArena::kNone)  // cannot happen: the tail behind the high-water mark holds need_slots
                    return fail(ctx, VRT_ERR_STATE, "internal: slot arena exhausted after pre-growth");
                cur.base = base;
                // resident bricks that survive and are not re-sent move device-side
                uint64_t keep = old.mask & new_mask & ~dirty;
                SectorSlots nw{new_mask, base};
                for (; keep; keep &= keep - 1) {
                    uint32_t b = (uint32_t)__builtin_ctzll(keep);
                    moves.push_back(make_uint2(slot_of(old, b), slot_of(nw, b)));
                }
                if (old_n) ctx->arena.quarantine(old.base, old_n);
            }
            cur.mask = new_mask;
            if ((old.mask == 0) != (new_mask == 0)) {
                ctx->resident_sectors += new_mask ? 1 : -1;
                ctx->boxes_stale = true;
                ctx->groups_stale = true;
                ctx->scene_epoch++;
            }
            ctx->sectors[si] = cur;
            headers.push_back(HeaderUpdate{hdr_index(ctx->sxp, ctx->sxp * ctx->sxp, d.sx, d.sy, d.sz), (uint32_t)new_mask,
                                           (uint32_t)(new_mask >> 32), cur.base});
        }
        const uint8_t* src = d.bricks;
        // payload order: ascending bit order over dirty_mask & alloc_mask
        for (uint64_t m = dirty; m; m &= m - 1) {
            uint32_t b = (uint32_t)__builtin_ctzll(m);
            uploads.push_back(Upload{src, slot_of(cur, b)});
            src += 512;
        }
        // Bricks that enter the allocation mask WITHOUT being dirty (VoxelMap::GetBrick creates the brick on a mere lookup,
        // VoxelMap.cpp:122 — quirk Q5) own a slot nothing else writes; slots are recycled, so it is filled with an empty
        // brick (what the reference's zero-initialised per-position storage holds for a brick that never had content).
        for (uint64_t m = new_mask & ~old.mask & ~dirty; m; m &= m - 1)
            uploads.push_back(Upload{kZeroBrick, slot_of(cur, (uint32_t)__builtin_ctzll(m))});
    }

    prof.mark(2);
    // Staging is chunked and double-buffered: a chunk of at most kChunkBricks bricks (64 MiB) is gathered into one half of the
    // pinned buffer (several host threads for big chunks) while the previous chunk's H2D copy and upload kernel run from the
    // other half — a multi-GB scene never needs a multi-GB pinned allocation, and the host gather overlaps the PCIe copy.
    // Chunk 0 also carries the relocation pairs and header records ("meta").
    const size_t nu = uploads.size(), nm = moves.size(), nh = headers.size();
    const size_t meta_moves = (nm * 8 + 15) & ~(size_t)15, meta_bytes = meta_moves + nh * sizeof(HeaderUpdate);
    const size_t chunk0 = std::min(nu, kChunkBricks);
    const size_t half = ((chunk0 * 516 + 15) & ~(size_t)15) + meta_bytes + 16;  // bricks, slots, meta
    const size_t n_chunks = nu ? (nu + kChunkBricks - 1) / kChunkBricks : (meta_bytes ? 1 : 0);
    size_t total = 0;
    if (n_chunks) {
        int st = VRT_OK;  // (both staging buffers were sized from the upper bounds before pass 2)
        // The previous frame may still read the arena from a caller stream: the kernels below wait for it on the device, but
        // the host gather and the H2D copies do not — they overlap the frame that is still being traced.
        // (the halves' offsets depend on this call's sizes, so both halves of the PREVIOUS call must have been consumed first;
        // their events follow that call's upload kernels, which never wait for the frame traced after them)
        for (int h2 = 0; h2 < 2; h2++)
            if (ctx->stage_used[h2]) CU(cudaEventSynchronize(ctx->ev_stage[h2]));
        bool waited = false;
        auto wait_once = [&]() -> int {
            if (waited || !ctx->render_pending) return VRT_OK;
            waited = true;
            return wait_renders(ctx, ctx->stream);
        };
        for (size_t c = 0; c < n_chunks; c++) {
            const size_t i0 = c * kChunkBricks, nb = std::min(kChunkBricks, nu - std::min(nu, i0));
            uint8_t* hs = ctx->h_stage + (c & 1) * half;
            uint8_t* ds = reinterpret_cast<uint8_t*>(ctx->d_stage.p) + (c & 1) * half;
            if (c >= 2) CU(cudaEventSynchronize(ctx->ev_stage[c & 1]));  // the chunk that used this half has been consumed
            const size_t off_slots = nb * 512, off_meta = (off_slots + nb * 4 + 15) & ~(size_t)15;
            uint32_t* slots = reinterpret_cast<uint32_t*>(hs + off_slots);
            prof.mark(3);
            auto gather = [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; i++) {
                    memcpy(hs + i * 512, uploads[i0 + i].src, 512);
                    slots[i] = uploads[i0 + i].slot;
                }
            };
            if (nb >= 8192 && ctx->gather_threads != 1) {  // (a frame's edit batch, ~3 k bricks, is gathered by the caller: with one process per GPU on a shared host, 8 pools waking for 100 us of copying each cost more than they save)
                if (ctx->gather_pool.workers() == 0) {
                    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
                    const unsigned want = ctx->gather_threads > 0 ? (unsigned)ctx->gather_threads : std::min(8u, hw);
                    if (want > 1) ctx->gather_pool.start(want - 1);
                }
                if (ctx->gather_pool.workers())
                    ctx->gather_pool.run([&](unsigned part, unsigned parts) { gather(nb * part / parts, nb * (part + 1) / parts); });
                else gather(0, nb);
            } else gather(0, nb);
            prof.mark(4);
            size_t bytes = off_meta;
            if (c == 0) {
                if (nm) memcpy(hs + off_meta, moves.data(), nm * 8);
                if (nh) memcpy(hs + off_meta + meta_moves, headers.data(), nh * sizeof(HeaderUpdate));
                bytes += meta_bytes;
            }
            CU(cudaMemcpyAsync(ds, hs, bytes, cudaMemcpyHostToDevice, ctx->stream));
            total += bytes;
            if ((st = wait_once())) return st;
            if (c == 0 && nm) {
                k_move_bricks<<<(unsigned)((nm + 7) / 8), 256, 0, ctx->stream>>>(reinterpret_cast<const uint2*>(ds + off_meta), (uint32_t)nm,
                                                                                 ctx->d_voxels, ctx->d_cells);
                ctx->stats.last_launches++;
            }
            if (nb) {
                k_upload_bricks<<<(unsigned)((nb + 7) / 8), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(ds),
                                                                                   reinterpret_cast<const uint32_t*>(ds + off_slots), (uint32_t)nb,
                                                                                   ctx->d_voxels, ctx->d_cells);
                ctx->stats.last_launches++;
            }
            if (c == 0 && nh) {
                k_write_headers<<<(unsigned)((nh + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<const HeaderUpdate*>(ds + off_meta + meta_moves),
                                                                                      (uint32_t)nh, ctx->d_hdr);
                ctx->stats.last_launches++;
            }
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev_stage[c & 1], ctx->stream));
            ctx->stage_used[c & 1] = true;
        }
    }
    prof.mark(5);
    if (ctx->render_pending) { int st_ = wait_renders(ctx, ctx->stream); if (st_) return st_; }  // (box rebuild / callers that sync nothing)
    if (nh) {  // some sector's allocation mask changed: refresh the one-bit view (78 k entries for the 2048x512x2048 view)
        const uint32_t n_all = ctx->n_hdr + 2u * ctx->hdr_guard;
        k_build_occ<<<(n_all + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_hdr_alloc, n_all, ctx->d_occ);
        ctx->stats.last_launches++;
    }
    ctx->arena.flush_quarantine();
    {
        int st = rebuild_boxes(ctx);
        if (st) return st;
    }
    CU(cudaEventRecord(ctx->ev_sync, ctx->stream));
    ctx->stats.bytes_uploaded = total;
    ctx->stats.bricks_uploaded = nu;
    ctx->stats.bricks_relocated = nm;
    prof.mark(6);
    return VRT_OK;
}

extern "C" int vrt_read_sector(VrtContext* ctx, int32_t sx, int32_t sy, int32_t sz, uint64_t* out_alloc_mask, uint32_t* out_base_slot,
                               uint8_t* out_bricks, uint64_t* out_cells) {
    if (!ctx) return VRT_ERR_INVALID;
    if (((uint32_t)(sx | sz) >> ctx->sxz) != 0 || ((uint32_t)sy >> ctx->sy) != 0) return fail(ctx, VRT_ERR_INVALID, "sector outside the view");
    DeviceGuard g(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    uint4 h;
    CU(cudaMemcpy(&h, ctx->d_hdr + hdr_index(ctx->sxp, ctx->sxp * ctx->sxp, sx, sy, sz), sizeof(h),
                  cudaMemcpyDeviceToHost));  // the DEVICE copy is what is inspected
    uint64_t mask = (uint64_t)h.x | ((uint64_t)h.y << 32);
    if (out_alloc_mask) *out_alloc_mask = mask;
    if (out_base_slot) *out_base_slot = h.z;
    if (out_bricks) memset(out_bricks, 0, 64 * 512);
    if (out_cells) memset(out_cells, 0, 64 * 8 * sizeof(uint64_t));
    uint32_t k = 0;
    for (uint64_t m = mask; m; m &= m - 1, k++) {
        uint32_t b = (uint32_t)__builtin_ctzll(m);
        if (out_bricks) CU(cudaMemcpy(out_bricks + b * 512, ctx->d_voxels + (size_t)(h.z + k) * 512, 512, cudaMemcpyDeviceToHost));
        if (out_cells) CU(cudaMemcpy(out_cells + b * 8, ctx->d_cells + (size_t)(h.z + k) * 8, 64, cudaMemcpyDeviceToHost));
    }
    return VRT_OK;
}

// ------------------------------------------------------------------------------------------------
// ray cast / hit query
// ------------------------------------------------------------------------------------------------
extern "C" int vrt_trace_device(VrtContext* ctx, uint64_t n, const float* d_origin3, const float* d_dir3, const int32_t wo[3],
                                uint32_t max_iters, VrtHit* d_out, void* stream) {
    if (!ctx || !wo || (n && (!d_origin3 || !d_dir3 || !d_out))) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    int st = begin_on_stream(ctx, s);
    if (st) return st;
    st = launch_trace(ctx, n, d_origin3, d_dir3, wo, max_iters, d_out, s);
    if (st) return st;
    return end_on_stream(ctx, s);
}

extern "C" int vrt_trace(VrtContext* ctx, uint64_t n, const float* origin3, const float* dir3, const int32_t wo[3], uint32_t max_iters,
                         VrtHit* out) {
    if (!ctx || !wo || (n && (!origin3 || !dir3 || !out))) return VRT_ERR_INVALID;
    if (n == 0) return VRT_OK;
    DeviceGuard g(ctx->device);
    int st;
    if ((st = ensure(ctx, ctx->d_rays_o, n * 12))) return st;
    if ((st = ensure(ctx, ctx->d_rays_d, n * 12))) return st;
    if ((st = ensure(ctx, ctx->d_hits, n * sizeof(VrtHit)))) return st;
    CU(cudaMemcpyAsync(ctx->d_rays_o.p, origin3, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_rays_d.p, dir3, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    st = launch_trace(ctx, n, (const float*)ctx->d_rays_o.p, (const float*)ctx->d_rays_d.p, wo, max_iters, (VrtHit*)ctx->d_hits.p, ctx->stream);
    if (st) return st;
    CU(cudaMemcpyAsync(out, ctx->d_hits.p, n * sizeof(VrtHit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

extern "C" int vrt_trace_glsl(VrtContext* ctx, uint64_t n, const float* origin3, const float* dir3, const int32_t wo[3], uint32_t flags,
                              VrtHit* out) {
    if (!ctx || !wo || (n && (!origin3 || !dir3 || !out))) return VRT_ERR_INVALID;
    if (flags & ~(VRT_GLSL_COARSE | VRT_GLSL_ANISOTROPIC)) return fail(ctx, VRT_ERR_INVALID, "vrt_trace_glsl: unknown flag");
    if (n == 0) return VRT_OK;
    if (n > 0x7FFFFFFFull * 128ull) return fail(ctx, VRT_ERR_INVALID, "too many rays for one launch");
    DeviceGuard g(ctx->device);
    int st;
    uint64_t launches = 0;
    GlslScene G;
    if ((st = glsl_scene(ctx, ctx->stream, G, launches))) return st;
    DevScene S = dev_scene(ctx);
    if ((st = ensure(ctx, ctx->d_rays_o, n * 12))) return st;
    if ((st = ensure(ctx, ctx->d_rays_d, n * 12))) return st;
    if ((st = ensure(ctx, ctx->d_hits, n * sizeof(VrtHit)))) return st;
    CU(cudaMemcpyAsync(ctx->d_rays_o.p, origin3, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_rays_d.p, dir3, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    k_trace_glsl<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(S, G, wo[0], wo[1], wo[2], (const float*)ctx->d_rays_o.p,
                                                                      (const float*)ctx->d_rays_d.p, flags, n, (VrtHit*)ctx->d_hits.p);
    launches++;
    CU(cudaGetLastError());
    ctx->stats.last_launches = launches;
    CU(cudaMemcpyAsync(out, ctx->d_hits.p, n * sizeof(VrtHit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

extern "C" int vrt_hit_query(VrtContext* ctx, uint64_t n, const double* origin3, const double* dir3, uint32_t max_iters, VrtHitD* out) {
    if (!ctx || (n && (!origin3 || !dir3 || !out))) return VRT_ERR_INVALID;
    if (n == 0) return VRT_OK;
    if (max_iters == 0) max_iters = 1024;  // VoxelMap.h:214
    DeviceGuard g(ctx->device);
    int st;
    if ((st = ensure(ctx, ctx->d_q_o, n * 24))) return st;
    if ((st = ensure(ctx, ctx->d_q_d, n * 24))) return st;
    if ((st = ensure(ctx, ctx->d_q_out, n * sizeof(VrtHitD)))) return st;
    CU(cudaMemcpyAsync(ctx->d_q_o.p, origin3, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_q_d.p, dir3, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    k_hit_query<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(dev_scene(ctx), (const double*)ctx->d_q_o.p, (const double*)ctx->d_q_d.p,
                                                                     max_iters, n, (VrtHitD*)ctx->d_q_out.p);
    ctx->stats.last_launches = 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, ctx->d_q_out.p, n * sizeof(VrtHitD), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

// ------------------------------------------------------------------------------------------------
// shading inputs
// ------------------------------------------------------------------------------------------------
extern "C" int vrt_set_blue_noise(VrtContext* ctx, const uint8_t* rg, size_t bytes) {
    if (!ctx || !rg) return VRT_ERR_INVALID;
    if (bytes != VRT_BLUE_NOISE_BYTES) return fail(ctx, VRT_ERR_INVALID, "blue noise must be 128x128x64 (R,G) bytes");
    DeviceGuard g(ctx->device);
    if (!ctx->d_bn) {
        CU(cudaMalloc((void**)&ctx->d_bn, VRT_BLUE_NOISE_BYTES));
        ctx->stats.device_bytes += VRT_BLUE_NOISE_BYTES;
    }
    if (ctx->render_pending) { int st_ = sync_renders(ctx); if (st_) return st_; }
    CU(cudaMemcpyAsync(ctx->d_bn, rg, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev_sync, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

extern "C" int vrt_set_sky(VrtContext* ctx, const VrtSkyDesc* desc, const uint32_t* texels) {
    if (!ctx || !desc || !texels) return VRT_ERR_INVALID;
    if (desc->face_size < 4 || (desc->face_size & (desc->face_size - 1)) || desc->mip_levels == 0 || desc->mip_levels > 16)
        return fail(ctx, VRT_ERR_INVALID, "bad sky descriptor");
    DeviceGuard g(ctx->device);
    if (ctx->render_pending) { int st_ = sync_renders(ctx); if (st_) return st_; }
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_sky) {
        CU(cudaFree(ctx->d_sky));
        ctx->stats.device_bytes -= ctx->sky.texel_count * 4;
        ctx->d_sky = nullptr;
    }
    CU(cudaMalloc((void**)&ctx->d_sky, desc->texel_count * 4));
    ctx->stats.device_bytes += desc->texel_count * 4;
    CU(cudaMemcpyAsync(ctx->d_sky, texels, desc->texel_count * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev_sync, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->sky = *desc;
    return VRT_OK;
}

// ------------------------------------------------------------------------------------------------
// frame
// ------------------------------------------------------------------------------------------------
extern "C" int vrt_render_device(VrtContext* ctx, const VrtFrame* frame, void* d_out, VrtHit* d_aux, void* stream) {
    if (!ctx || !frame || !d_out) return VRT_ERR_INVALID;
    if ((frame->flags & VRT_FRAME_AUX_HITS) && !d_aux) return fail(ctx, VRT_ERR_INVALID, "VRT_FRAME_AUX_HITS without aux buffer");
    DeviceGuard g(ctx->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    int st = begin_on_stream(ctx, s);
    if (st) return st;
    ctx->stats.last_launches = 0;
    st = launch_render(ctx, frame, d_out, d_aux, s);
    if (st) return st;
    return end_on_stream(ctx, s);
}

extern "C" int vrt_render(VrtContext* ctx, const VrtFrame* frame, void* out, VrtHit* aux_hits) {
    if (!ctx || !frame || !out) return VRT_ERR_INVALID;
    if ((frame->flags & VRT_FRAME_AUX_HITS) && !aux_hits) return fail(ctx, VRT_ERR_INVALID, "VRT_FRAME_AUX_HITS without aux buffer");
    DeviceGuard g(ctx->device);
    size_t npx = (size_t)frame->width * frame->height;
    const size_t px_bytes = (frame->flags & VRT_FRAME_COMPACT) ? 8u : 16u, tile_bytes = px_bytes * 16u;
    int st;
    if ((st = ensure(ctx, ctx->d_fb, npx * px_bytes))) return st;
    bool aux = (frame->flags & VRT_FRAME_AUX_HITS) != 0;
    if (aux && (st = ensure(ctx, ctx->d_aux, npx * sizeof(VrtHit)))) return st;
    uint32_t part_count = frame->part_count ? frame->part_count : 1;
    if (part_count > 1 && (frame->flags & VRT_FRAME_PART_ROWS) && !(frame->flags & VRT_FRAME_LINEAR_OUTPUT) && !aux) {
        // Band split: only this rank's 8-pixel bands are traced and only they cross PCIe (one strided D2H copy into the
        // same offsets of `out`; the other ranks' bands of `out` are left untouched).
        if (frame->part_index >= part_count) return fail(ctx, VRT_ERR_INVALID, "part_index >= part_count");
        ctx->stats.last_launches = 0;
        st = launch_render(ctx, frame, ctx->d_fb.p, nullptr, ctx->stream);
        if (st) return st;
        const size_t band = (size_t)(frame->width / 4) * (VRT_BAND_ROWS / 4u) * tile_bytes;
        const uint32_t rows_full = frame->height / VRT_BAND_ROWS, rows_all = (frame->height + VRT_BAND_ROWS - 1u) / VRT_BAND_ROWS, r = frame->part_index;
        const uint32_t mine_full = rows_full > r ? (rows_full - r + part_count - 1) / part_count : 0u;
        if (mine_full) CU(cudaMemcpy2DAsync((char*)out + r * band, part_count * band, (char*)ctx->d_fb.p + r * band, part_count * band, band, mine_full, cudaMemcpyDeviceToHost, ctx->stream));
        if (rows_all > rows_full && rows_full % part_count == r) {
            const size_t tail = (size_t)(frame->width / 4) * ((frame->height % VRT_BAND_ROWS) / 4u) * tile_bytes;
            CU(cudaMemcpyAsync((char*)out + rows_full * band, (char*)ctx->d_fb.p + rows_full * band, tail, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CU(cudaStreamSynchronize(ctx->stream));
        return VRT_OK;
    }
    if (part_count > 1) {  // pixels of other ranks stay zero in a partial frame
        CU(cudaMemsetAsync(ctx->d_fb.p, 0, npx * px_bytes, ctx->stream));
        if (aux) CU(cudaMemsetAsync(ctx->d_aux.p, 0, npx * sizeof(VrtHit), ctx->stream));
    }
    ctx->stats.last_launches = 0;
    const bool tiled = (frame->flags & VRT_FRAME_LINEAR_OUTPUT) == 0;
    const uint32_t macros_y = (frame->height + 31) / 32;
    if (part_count == 1 && tiled && !aux && macros_y >= 16) {
        // Band pipeline: the frame is rendered in up to 8 bands of macro-tile rows; a band's 16 B/px
        // tiles are contiguous in the reference's tile order, so its D2H copy runs on a second stream
        // while the next band is being traced.  The PCIe copy, not the kernel, bounds this call.
        const uint32_t n_bands = 8, rows_per = (macros_y + n_bands - 1) / n_bands;
        const size_t row_bytes = (size_t)32 * frame->width * px_bytes;  // one macro row = 8 tile rows
        for (uint32_t b = 0; b < n_bands; b++) {
            uint32_t r0 = b * rows_per, r1 = std::min(macros_y, r0 + rows_per);
            if (r0 >= r1) break;
            st = launch_render(ctx, frame, ctx->d_fb.p, nullptr, ctx->stream, r0, r1);
            if (st) return st;
            CU(cudaEventRecord(ctx->ev_band[b], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_band[b], 0));
            size_t off = (size_t)r0 * row_bytes, end = std::min((size_t)r1 * row_bytes, npx * px_bytes);
            CU(cudaMemcpyAsync((char*)out + off, (char*)ctx->d_fb.p + off, end - off, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        CU(cudaStreamSynchronize(ctx->copy_stream));
        CU(cudaStreamSynchronize(ctx->stream));
        return VRT_OK;
    }
    st = launch_render(ctx, frame, ctx->d_fb.p, (VrtHit*)ctx->d_aux.p, ctx->stream);
    if (st) return st;
    CU(cudaMemcpyAsync(out, ctx->d_fb.p, npx * px_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (aux) CU(cudaMemcpyAsync(aux_hits, ctx->d_aux.p, npx * sizeof(VrtHit), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU framebuffer sharing (CUDA IPC over NVLink peer mappings)
// ------------------------------------------------------------------------------------------------
extern "C" int vrt_fb_export(VrtContext* ctx, uint64_t bytes, uint8_t handle_out[64], void** d_ptr_out) {
    if (!ctx || !handle_out || !d_ptr_out || bytes == 0) return VRT_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(ctx->device);
    void* p = nullptr;
    CU(cudaMalloc(&p, bytes));
    CU(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(ctx, VRT_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    ctx->exported.push_back(p);
    ctx->stats.device_bytes += bytes;
    *d_ptr_out = p;
    return VRT_OK;
}

extern "C" int vrt_fb_import(VrtContext* ctx, const uint8_t handle[64], void** d_ptr_out) {
    if (!ctx || !handle || !d_ptr_out) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->imported.push_back(p);
    *d_ptr_out = p;
    return VRT_OK;
}

extern "C" int vrt_fb_release(VrtContext* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    auto it = std::find(ctx->imported.begin(), ctx->imported.end(), d_ptr);
    if (it != ctx->imported.end()) {
        CU(cudaIpcCloseMemHandle(d_ptr));
        ctx->imported.erase(it);
        return VRT_OK;
    }
    it = std::find(ctx->exported.begin(), ctx->exported.end(), d_ptr);
    if (it != ctx->exported.end()) {
        CU(cudaDeviceSynchronize());
        CU(cudaFree(d_ptr));
        ctx->exported.erase(it);
        return VRT_OK;
    }
    return fail(ctx, VRT_ERR_INVALID, "pointer not owned by this context");
}

// Pipelined tile gather (include/voxelrt_b200.h): frame kernel into the rank's own buffer on `stream`, then ONE strided
// device-to-device copy of the rank's 8-pixel bands into the owner's framebuffer on the copy stream.  Band b of a
// tile-layout framebuffer is the byte range [b * band, (b + 1) * band), band = (width / 4) * 2 tile rows * 256 B, and the rank
// owns bands part_index, part_index + part_count, ... — i.e. a 2-D copy with pitch part_count * band.  A last, shorter
// band (height % 8 != 0) goes in a second, 1-D copy.
extern "C" int vrt_render_gather(VrtContext* ctx, const VrtFrame* frame, void* d_local_fb, void* d_owner_fb, void* stream) {
    if (!ctx || !frame || !d_local_fb || !d_owner_fb) return VRT_ERR_INVALID;
    uint32_t part_count = frame->part_count ? frame->part_count : 1;
    if (!(frame->flags & VRT_FRAME_PART_ROWS) && part_count > 1) return fail(ctx, VRT_ERR_INVALID, "vrt_render_gather needs VRT_FRAME_PART_ROWS");
    if (frame->flags & (VRT_FRAME_LINEAR_OUTPUT | VRT_FRAME_AUX_HITS)) return fail(ctx, VRT_ERR_INVALID, "vrt_render_gather: tile layout, no aux records");
    DeviceGuard g(ctx->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    int st = begin_on_stream(ctx, s);
    if (st) return st;
    ctx->stats.last_launches = 0;
    // a local buffer may be reused every VRT_GATHER_DEPTH-th call: this frame's kernel waits for the copy issued that many calls ago
    if (ctx->gather_seq >= VRT_GATHER_DEPTH) CU(cudaStreamWaitEvent(s, ctx->ev_gather_done[(ctx->gather_seq - VRT_GATHER_DEPTH) % (2u * VRT_GATHER_DEPTH)], 0));
    st = launch_render(ctx, frame, d_local_fb, nullptr, s);
    if (st) return st;
    const uint32_t seq = ctx->gather_seq++;
    cudaStream_t gs = ctx->gather_streams[seq % VRT_GATHER_DEPTH];
    if (d_owner_fb != d_local_fb) {
        cudaEvent_t ev = ctx->ev_gather_src[seq % (2u * VRT_GATHER_DEPTH)];
        CU(cudaEventRecord(ev, s));
        CU(cudaStreamWaitEvent(gs, ev, 0));
        const size_t tile_bytes = (frame->flags & VRT_FRAME_COMPACT) ? sizeof(VrtTileAD) : sizeof(VrtTile);
        const size_t band = (size_t)(frame->width / 4) * (VRT_BAND_ROWS / 4u) * tile_bytes;
        const uint32_t rows_full = frame->height / VRT_BAND_ROWS, rows_all = (frame->height + VRT_BAND_ROWS - 1u) / VRT_BAND_ROWS;
        const uint32_t r = frame->part_index;
        const uint32_t mine_full = rows_full > r ? (rows_full - r + part_count - 1) / part_count : 0u;
        uint8_t* dst = static_cast<uint8_t*>(d_owner_fb);
        const uint8_t* src = static_cast<const uint8_t*>(d_local_fb);
        static const int gather_1d = getenv("VRT_GATHER_1D") ? atoi(getenv("VRT_GATHER_1D")) : 0;
        if (mine_full && gather_1d) {
            for (uint32_t k = 0; k < mine_full; k++) {
                size_t off = (size_t)(r + k * part_count) * band;
                CU(cudaMemcpyAsync(dst + off, src + off, band, cudaMemcpyDeviceToDevice, gs));
            }
        } else if (mine_full) CU(cudaMemcpy2DAsync(dst + r * band, part_count * band, src + r * band, part_count * band, band, mine_full, cudaMemcpyDeviceToDevice, gs));
        if (rows_all > rows_full && rows_full % part_count == r) {
            const size_t tail = (size_t)(frame->width / 4) * ((frame->height % VRT_BAND_ROWS) / 4u) * tile_bytes;
            CU(cudaMemcpyAsync(dst + rows_full * band, src + rows_full * band, tail, cudaMemcpyDeviceToDevice, gs));
        }
        ctx->gather_pending = true;
    }
    CU(cudaEventRecord(ctx->ev_gather_done[seq % (2u * VRT_GATHER_DEPTH)], gs));  // (copy-stream order also covers the calls without a copy)
    return end_on_stream(ctx, s);
}

extern "C" int vrt_gather_wait(VrtContext* ctx, void* stream) {
    if (!ctx) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    if (ctx->gather_pending)  // the last copy on each of the gather streams
        for (uint32_t k = 1; k <= VRT_GATHER_DEPTH && k <= ctx->gather_seq; k++)
            CU(cudaStreamWaitEvent(s, ctx->ev_gather_done[(ctx->gather_seq - k) % (2u * VRT_GATHER_DEPTH)], 0));
    return VRT_OK;
}

// Internal diagnostic hook (not part of include/voxelrt_b200.h): reads and clears the macro-loop event
// counters filled by "metrics" launches while macro_steps = 2.  Used by tools_macro_stats.py only.
extern "C" __attribute__((visibility("default"))) int vrt_debug_macro_diag(VrtContext* ctx, uint64_t out[8]) {
    if (!ctx || !out) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, g_macro_diag, 8 * sizeof(uint64_t)));
    uint64_t zero[8] = {};
    CU(cudaMemcpyToSymbol(g_macro_diag, zero, sizeof(zero)));
    return VRT_OK;
}

// Internal self-check hook (not part of include/voxelrt_b200.h): compares rcp_rn_normal with rcp.rn (both signs) and sqrt_rn_normal
// with sqrt.rn for EVERY binary32 bit pattern in [lo_bits, hi_bits]; *mismatches receives the count.  out_sample (host, 2*n floats, may be null)
// receives rcp_rn_normal of the first n patterns and of their negations, for a comparison against the host's own 1.0f/x.
namespace {
__global__ void k_debug_rcp_check(uint32_t lo_bits, uint64_t count, unsigned long long* mismatches, float* sample, uint32_t n_sample) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned bad = 0;
    for (; i < count; i += stride) {
        float x = __uint_as_float(lo_bits + (uint32_t)i);
        float a = vrt::rcp_rn_normal(x), b = __frcp_rn(x);
        float c = vrt::rcp_rn_normal(-x), d = __frcp_rn(-x);
        float e = vrt::sqrt_rn_normal(x), g = __fsqrt_rn(x);
        bad += (__float_as_uint(a) != __float_as_uint(b)) + (__float_as_uint(c) != __float_as_uint(d)) + (__float_as_uint(e) != __float_as_uint(g));
        if (sample && i < n_sample) {
            sample[i] = a;
            sample[n_sample + i] = c;
        }
    }
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}
}  // namespace
extern "C" __attribute__((visibility("default"))) int vrt_debug_rcp_check(VrtContext* ctx, uint32_t lo_bits, uint32_t hi_bits, uint64_t* mismatches,
                                                                          float* out_sample, uint32_t n_sample) {
    if (!ctx || !mismatches || hi_bits < lo_bits) return VRT_ERR_INVALID;
    DeviceGuard g(ctx->device);
    unsigned long long* d_bad = nullptr;
    float* d_sample = nullptr;
    CU(cudaMalloc((void**)&d_bad, sizeof(unsigned long long)));
    CU(cudaMemset(d_bad, 0, sizeof(unsigned long long)));
    if (out_sample && n_sample) CU(cudaMalloc((void**)&d_sample, 2ull * n_sample * sizeof(float)));
    uint64_t count = (uint64_t)hi_bits - lo_bits + 1;
    k_debug_rcp_check<<<148 * 16, 256, 0, ctx->stream>>>(lo_bits, count, d_bad, d_sample, d_sample ? n_sample : 0u);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned long long bad = 0;
    CU(cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost));
    if (d_sample) CU(cudaMemcpy(out_sample, d_sample, 2ull * n_sample * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    cudaFree(d_sample);
    *mismatches = bad;
    return VRT_OK;
}
