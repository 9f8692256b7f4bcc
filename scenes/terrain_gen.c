/* Bulk terrain generation for the big-view bench scene (BASELINE configs[3]: a 4096-wide terrain with >= 10 GB of bricks).
 * INPUT generation only (not the product, not the oracle).  Per sector this is TerrainGenerator::GenerateSector of the reference
 * (VoxelRT/TerrainGenerator.cpp:5-34): FastNoise2's GenUniformGrid3D over the 32^3 voxels at frequency 0.004 / seed 12345 with the
 * reference's node tree (the density comes from the reference's own vendored FastNoise2: libFastNoise.so is resolved at load time through
 * function pointers handed in by scenes/terrain.py), voxel id = noise < 0 ? 245 + (trunc(noise * 1234.5678f) & 3) : 0, non-empty bricks
 * only — here for many sectors at once with OpenMP, because the Python per-sector path costs more than the noise itself.
 * `y_shift` moves the terrain up by that many voxels (the noise is sampled at y - 96 - y_shift; the reference uses y - 96): everything
 * below the surface is solid, so the shift sets how many GB of bricks the scene holds. */
#include <omp.h>
#include <stdint.h>
#include <string.h>

typedef void (*gen_grid_fn)(void* node, float* out, int x, int y, int z, int sx, int sy, int sz, float freq, int seed, float* minmax);

/* sectors (x, y, z) in [0,nx) x [0,ny) x [0,nz), index i = x + nx * (z + nz * y).  masks[i] = allocation mask of the non-empty bricks;
 * the k = popcount(mask) bricks of sector i (ascending brick index bx | bz<<2 | by<<4, voxel index x | z<<3 | y<<6) are written to
 * bricks + (i * 64) * 512, packed at the front of the sector's 64-brick block.  Returns the number of non-empty bricks. */
uint64_t terrain_gen_sectors(gen_grid_fn gen, void* node, int nx, int ny, int nz, int y_shift, uint64_t* masks, uint8_t* bricks, int threads) {
    const int64_t n = (int64_t)nx * ny * nz;
    uint64_t total = 0;
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads) reduction(+ : total)
    {
        float noise[32 * 32 * 32];
        uint8_t ids[32 * 32 * 32]; /* [y][z][x] */
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n; i++) {
            const int sx = (int)(i % nx), sz = (int)((i / nx) % nz), sy = (int)(i / ((int64_t)nx * nz));
            gen(node, noise, sx * 32, sy * 32 - 96 - y_shift, sz * 32, 32, 32, 32, 0.004f, 12345, 0); /* :10, output index x + 32 y + 1024 z */
            for (int z = 0; z < 32; z++)
                for (int y = 0; y < 32; y++)
                    for (int x = 0; x < 32; x++) {
                        const float v = noise[x + 32 * y + 1024 * z];
                        ids[(y * 32 + z) * 32 + x] = v < 0.0f ? (uint8_t)(245 + ((int)(v * 1234.5678f) & 3)) : 0; /* :21-23 */
                    }
            uint64_t mask = 0;
            uint8_t* dst = bricks + (size_t)i * 64 * 512;
            for (int b = 0; b < 64; b++) {
                const int bx = b & 3, bz = (b >> 2) & 3, by = b >> 4;
                uint8_t brick[512];
                int any = 0;
                for (int y = 0; y < 8; y++)
                    for (int z = 0; z < 8; z++) {
                        const uint8_t* row = &ids[((by * 8 + y) * 32 + bz * 8 + z) * 32 + bx * 8];
                        memcpy(&brick[(y * 8 + z) * 8], row, 8);
                        uint64_t w;
                        memcpy(&w, row, 8);
                        any |= w != 0;
                    }
                if (any) {
                    memcpy(dst, brick, 512);
                    dst += 512;
                    mask |= 1ull << b;
                }
            }
            masks[i] = mask;
            total += (uint64_t)__builtin_popcountll(mask);
        }
    }
    return total;
}
