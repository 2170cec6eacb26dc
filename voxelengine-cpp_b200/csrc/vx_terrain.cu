// libvxgpu: terrain fill.  WorldGenerator::generate (world/generator/
// WorldGenerator.cpp:419-461) without structure placements: the chunk is
// produced in HBM from its 16x16 heightmap and per-column biome.
//
// The reference walks every column top-down and writes one voxel at a time
// (generate_pole, :87-116; stride 1 KiB between writes).  Here a column first
// turns its layers into at most 2 * VX_GEN_MAX_LAYERS (y_hi, y_lo, id) spans in
// shared memory — the same unsigned arithmetic, nothing written yet — and the
// CTA then streams the whole 256 KiB chunk once, in memory order, with 16-byte
// stores, each voxel taking the last span that covers it (later writes of the
// reference overwrite earlier ones).  Plants (generatePlants, :362-401) use one
// util::PseudoRandom stream per chunk whose consumption depends on what was
// placed, so one thread walks the 256 columns in the reference's order; a
// plant only ever lands on the (empty) voxel above its own column, so it is a
// 257th span of that column.
#include <cstring>
#include <string>
#include <vector>

#include "vx_ctx.hpp"

using namespace vx;

namespace {

constexpr int MAX_SPANS = 2 * VX_GEN_MAX_LAYERS;
constexpr uint32_t BLOCK_STRUCT_AIR = 2;  // core_defs.hpp

struct GenTables {
    const vx_gen_biome* biomes;
    const vx_gen_layer* layers;
    const vx_gen_plant* plants;
    const float* weight_sums;  // BiomeElementList::weightsSum per biome
    int n_biomes, n_plants;
    uint32_t sea_level;
};

struct Span {
    int16_t hi, lo;  // voxels lo..hi inclusive
    uint16_t id;
};

// generate_pole (:87-116) with the writes replaced by spans
__device__ __forceinline__ int pole_spans(const vx_gen_layer* layers, int count, uint32_t last_layers_height,
                                          int top, int bottom, uint32_t sea_level, Span* out, int n) {
    uint32_t y = (uint32_t)top;
    uint32_t extension = 0;
    for (int k = 0; k < count; k++) {
        const vx_gen_layer L = layers[k];
        if (y < sea_level && !L.below_sea_level) {
            extension = (uint32_t)max(0, (int)L.height);
            continue;
        }
        int h = L.height;
        if (h == -1) h = (int)(y - last_layers_height - (uint32_t)bottom + 1u);  // the resizeable layer
        else h += (int)extension;
        h = (int)min((uint32_t)h, y + 1u);
        if (h > 0) {
            // voxels y, y-1, ..., y-h+1; rows at or above CHUNK_H are out of bounds in the reference
            out[n++] = Span {(int16_t)min(y, (uint32_t)CH - 1u), (int16_t)(y - (uint32_t)h + 1u), L.block};
            if (y >= (uint32_t)CH && y - (uint32_t)h + 1u >= (uint32_t)CH) n--;
            y -= (uint32_t)h;
        }
        extension = 0;
    }
    return n;
}

// util::PseudoRandom (maths/util.hpp:19-87)
struct Prng {
    uint16_t seed;
    __device__ int rand() {
        seed = (uint16_t)(seed + 0x7ed5 + (seed << 6));
        seed = (uint16_t)(seed ^ 0xc23c ^ (seed >> 9));
        seed = (uint16_t)(seed + 0x1656 + (seed << 3));
        seed = (uint16_t)((seed + 0xa264) ^ (seed << 4));
        seed = (uint16_t)(seed + 0xfd70 - (seed << 3));
        seed = (uint16_t)(seed ^ 0xba49 ^ (seed >> 8));
        return (int)seed;
    }
    __device__ void set_seed(int a, int b) {  // :72-77
        seed = (uint16_t)(((uint16_t)(a * 23729) | (uint16_t)(b % 16786)) ^ (uint16_t)(b * a));
        rand();
    }
    __device__ float rand_float() {  // :61-63; the first rand() is the high half (g++ evaluation order)
        const uint32_t hi = (uint32_t)rand(), lo = (uint32_t)rand();
        return __fdiv_rn((float)((hi << 16) | lo), 4294967296.0f);
    }
};

__device__ __forceinline__ uint32_t span_id(const Span* s, int n, int y) {
    uint32_t id = 0;
    for (int k = 0; k < n; k++)
        if (y <= s[k].hi && y >= s[k].lo) id = s[k].id;
    return id;
}

// grid (n), 256 threads
__global__ void __launch_bounds__(256) k_terrain_fill(DevWorld W, const int32_t* pools, GenTables G,
                                                      const float* heights, const uint8_t* biomes) {
    const int c = blockIdx.x, tid = threadIdx.x;
    const int p = pools[c];
    const ChunkMeta& m = W.meta[p];
    __shared__ Span spans[256][MAX_SPANS + 1];
    __shared__ uint8_t nspan[256];
    __shared__ uint8_t col_biome[256];
    __shared__ int16_t col_height[256];
    __shared__ int s_top;
    // the plants walk below is one thread: everything it touches sits in shared memory
    constexpr int SB = 16, SP = 64;  // biomes / plant entries staged (larger tables stay in HBM)
    __shared__ vx_gen_biome s_biomes[SB];
    __shared__ vx_gen_plant s_plants[SP];
    __shared__ float s_sums[SB];
    __shared__ uint32_t s_pflags[SP];  // DevDef::flags of the staged plants
    const bool staged = G.n_biomes <= SB && G.n_plants <= SP;
    if (staged) {
        if (tid < G.n_biomes) {
            s_biomes[tid] = G.biomes[tid];
            s_sums[tid] = G.weight_sums[tid];
        }
        if (tid < G.n_plants) {
            s_plants[tid] = G.plants[tid];
            const uint32_t pb = G.plants[tid].block;
            s_pflags[tid] = pb < (uint32_t)W.n_defs ? W.defs[pb].flags : 0u;
        }
    }
    if (tid == 0) s_top = 0;
    __syncthreads();
    // ---- generateLand (:403-417): spans of column (x, z) = (tid & 15, tid >> 4)
    {
        const int b = min((int)biomes[c * 256 + tid], G.n_biomes - 1);
        col_biome[tid] = (uint8_t)b;
        const vx_gen_biome B = staged ? s_biomes[b] : G.biomes[b];
        int height = (int)(heights[c * 256 + tid] * (float)CH);
        height = max(0, height);
        int n = pole_spans(G.layers + B.sea_begin, B.sea_count, B.sea_last_layers_height, (int)G.sea_level,
                           height, G.sea_level, spans[tid], 0);
        n = pole_spans(G.layers + B.ground_begin, B.ground_count, B.ground_last_layers_height, height, 0,
                       G.sea_level, spans[tid], n);
        nspan[tid] = (uint8_t)n;
        col_height[tid] = (int16_t)min(height, CH - 1);  // generatePlants clamps (:379)
        int top = 0;
        for (int k = 0; k < n; k++) top = max(top, (int)spans[tid][k].hi);
        atomicMax(&s_top, top + 1);  // + 1: a plant may stand on the top voxel
    }
    __syncthreads();
    // ---- generatePlants (:362-401): one sequential random stream per chunk
    if (tid == 0) {
        Prng rnd;
        rnd.set_seed(m.cx, m.cz);
        for (int col = 0; col < 256; col++) {  // z outer, x inner
            const int height = col_height[col];
            if (!(height + 1 > (int)G.sea_level && height + 1 < CH)) continue;
            const int b = col_biome[col];
            const vx_gen_biome B = staged ? s_biomes[b] : G.biomes[b];
            const vx_gen_plant* plants = staged ? s_plants : G.plants;
            float r = rnd.rand_float();
            // BiomeElementList::choose (GeneratorDef.hpp:92-105)
            int pick = -1;
            if (!(B.plants_count == 0 || r > B.plants_chance || B.plants_chance < 1e-6f)) {
                r = __fdiv_rn(r, B.plants_chance);
                r *= staged ? s_sums[b] : G.weight_sums[b];
                pick = B.plants_begin + B.plants_count - 1;
                for (int k = 0; k < B.plants_count; k++) {
                    r -= plants[B.plants_begin + k].weight;
                    if (r <= 0.0f) {
                        pick = B.plants_begin + k;
                        break;
                    }
                }
            }
            const uint32_t plant = pick >= 0 ? plants[pick].block : 0u;
            if (!plant) continue;
            const int n = nspan[col];
            if (span_id(spans[col], n, height + 1) != 0) continue;
            const uint32_t ground = span_id(spans[col], n, height);
            const uint32_t gf = ground < (uint32_t)W.n_defs ? __ldg(&W.defs[ground].flags) : 0u;
            if (!(gf & DF_SOLID)) continue;
            uint32_t state = 0;
            const uint32_t pf = staged ? s_pflags[pick] : (plant < (uint32_t)W.n_defs ? W.defs[plant].flags : 0u);
            const uint32_t prof = (pf >> DF_ROT_SHIFT) & 3u;
            if (prof != VX_ROT_NONE) state = (uint32_t)rnd.rand() % (prof == VX_ROT_PIPE ? 6u : 4u);  // :392-395
            // a plant is the last span of its column; the state rides in `lo`'s place via a marker
            spans[col][MAX_SPANS] = Span {(int16_t)(height + 1), (int16_t)state, (uint16_t)plant};
            nspan[col] = (uint8_t)(n | 0x80);
        }
    }
    __syncthreads();
    // ---- stream the chunk: 4 voxels per thread per step, memory order
    const int top = s_top;  // rows above hold air only
    uint4* out = reinterpret_cast<uint4*>(W.vox_of(p));
    for (int g = tid; g < CVOL / 4; g += 256) {
        const int y = g >> 6, z = (g >> 2) & 15, x0 = (g & 3) * 4;
        uint32_t v[4] = {0, 0, 0, 0};
        if (y <= top) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int col = z * 16 + x0 + j;
                const int n = nspan[col];
                uint32_t id = span_id(spans[col], n & 0x7F, y);
                if (id == BLOCK_STRUCT_AIR) id = 0;  // :452-456
                if ((n & 0x80) && spans[col][MAX_SPANS].hi == y) {
                    const Span P = spans[col][MAX_SPANS];
                    id = (P.id == BLOCK_STRUCT_AIR ? 0u : (uint32_t)P.id) | ((uint32_t)(uint16_t)P.lo << 16);
                }
                v[j] = id;
            }
        }
        out[g] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

}  // namespace

extern "C" {

int vx_terrain_fill(vx_ctx* ctx, const vx_gen_def* def, const vx_gen_chunk* chunks, int32_t n) {
    if (!ctx) return -1;
    if (ctx->pool_cap == 0) {
        ctx->fail("vx_terrain_fill: configure the window first");
        return -1;
    }
    if (n <= 0) return 0;
    if (!def || !chunks || def->n_biomes <= 0 || def->n_biomes > 256) {
        ctx->fail("vx_terrain_fill: a generator definition with 1..256 biomes is required");
        return -1;
    }
    std::vector<float> sums(def->n_biomes, 0.0f);
    for (int b = 0; b < def->n_biomes; b++) {
        const vx_gen_biome& B = def->biomes[b];
        if (B.ground_count < 0 || B.sea_count < 0 || B.ground_count > VX_GEN_MAX_LAYERS ||
            B.sea_count > VX_GEN_MAX_LAYERS || B.ground_begin < 0 || B.sea_begin < 0 ||
            B.ground_begin + B.ground_count > def->n_layers || B.sea_begin + B.sea_count > def->n_layers ||
            B.plants_count < 0 || B.plants_begin < 0 || B.plants_begin + B.plants_count > def->n_plants) {
            ctx->fail("vx_terrain_fill: biome " + std::to_string(b) + " has layer / plant ranges outside the tables "
                      "(at most " + std::to_string(VX_GEN_MAX_LAYERS) + " layers per list)");
            return -1;
        }
        // BiomeElementList's constructor: weightsSum += entry.weight, in order (GeneratorDef.hpp:82-87)
        for (int k = 0; k < B.plants_count; k++) sums[b] += def->plants[B.plants_begin + k].weight;
    }
    for (int i = 0; i < n; i++)
        if (!chunks[i].heights || !chunks[i].biomes) {
            ctx->fail("vx_terrain_fill: chunk without heights / biomes");
            return -1;
        }
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    std::vector<vx_chunk_key> keys(n);
    std::vector<uint8_t> none(n, 0);
    for (int i = 0; i < n; i++) keys[i] = vx_chunk_key {chunks[i].cx, chunks[i].cz};
    std::vector<int> pools;
    if (vx_put_slots(ctx, keys.data(), none.data(), n, pools)) return -1;
    // ---- one upload: tables, heights, biome indices
    auto pad = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_biomes = 0;
    const size_t o_layers = o_biomes + pad(sizeof(vx_gen_biome) * def->n_biomes);
    const size_t o_plants = o_layers + pad(sizeof(vx_gen_layer) * (size_t)std::max(1, def->n_layers));
    const size_t o_sums = o_plants + pad(sizeof(vx_gen_plant) * (size_t)std::max(1, def->n_plants));
    const size_t o_heights = o_sums + pad(sizeof(float) * def->n_biomes);
    const size_t o_cbiomes = o_heights + pad(sizeof(float) * 256 * (size_t)n);
    const size_t total = o_cbiomes + pad(256 * (size_t)n);
    uint8_t* host_buf = vx_host_stage(ctx, total);
    if (!host_buf) return -1;
    std::memcpy(host_buf + o_biomes, def->biomes, sizeof(vx_gen_biome) * def->n_biomes);
    if (def->n_layers > 0) std::memcpy(host_buf + o_layers, def->layers, sizeof(vx_gen_layer) * def->n_layers);
    if (def->n_plants > 0) std::memcpy(host_buf + o_plants, def->plants, sizeof(vx_gen_plant) * def->n_plants);
    std::memcpy(host_buf + o_sums, sums.data(), sizeof(float) * def->n_biomes);
    for (int i = 0; i < n; i++) {
        std::memcpy(host_buf + o_heights + sizeof(float) * 256 * (size_t)i, chunks[i].heights, sizeof(float) * 256);
        std::memcpy(host_buf + o_cbiomes + 256 * (size_t)i, chunks[i].biomes, 256);
    }
    if (total > ctx->codec_stage_bytes) {
        cudaFree(ctx->d_codec_stage);
        ctx->d_codec_stage = nullptr;
        ctx->codec_stage_bytes = 0;
        VX_CUDA(cudaMalloc((void**)&ctx->d_codec_stage, total + total / 4));
        ctx->codec_stage_bytes = total + total / 4;
    }
    uint8_t* d = ctx->d_codec_stage;
    VX_CUDA(cudaMemcpyAsync(d, host_buf, total, cudaMemcpyHostToDevice, st));
    GenTables G {reinterpret_cast<const vx_gen_biome*>(d + o_biomes),
                 reinterpret_cast<const vx_gen_layer*>(d + o_layers),
                 reinterpret_cast<const vx_gen_plant*>(d + o_plants),
                 reinterpret_cast<const float*>(d + o_sums), def->n_biomes, def->n_plants, def->sea_level};
    if (vx_upload_list(ctx, pools, 0)) return -1;
    VX_LAUNCH(ctx, KID_TERRAIN, k_terrain_fill<<<n, 256, 0, st>>>(ctx->W, ctx->d_list, G,
                                                                   reinterpret_cast<const float*>(d + o_heights),
                                                                   d + o_cbiomes));
    VX_CUDA(cudaGetLastError());
    if (vx_zero_light(ctx, pools)) return -1;
    if (vx_derive_chunks(ctx, pools, false)) return -1;
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
