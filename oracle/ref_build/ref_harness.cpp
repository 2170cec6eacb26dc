// oracle/_ref/libvxref.so — flat C API (oracle/vxoracle.h) over the
// reference's OWN translation units, compiled unmodified from
// /root/reference/src by oracle/ref_build/Makefile.
//
// TEST INFRASTRUCTURE ONLY.  This file is the link seam, nothing more:
//   * it builds a Content / ContentIndices / Block table from a vx_content,
//   * it supplies the handful of symbols whose real definitions need OpenGL,
//     libpng, Lua or the asset system (Mesh ctor/dtor, ContentGfxCache members,
//     ContentPackRuntime dtor) — none of them is reached from
//     Lighting::* or BlocksRenderer::build/createMesh,
//   * it forwards every vxo_* call to the reference class method named in
//     oracle/vxoracle.h.
// All lighting and meshing arithmetic executed through this library is the
// reference's (lighting/*.cpp, voxels/Chunks.cpp, voxels/blocks_agent.cpp,
// voxels/VoxelsVolume.cpp, graphics/render/BlocksRenderer.cpp).

#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "coders/rle.hpp"
#include "content/Content.hpp"
#include "content/ContentPack.hpp"
#include "frontend/ContentGfxCache.hpp"
#include "graphics/core/Mesh.hpp"
#include "graphics/render/BlocksRenderer.hpp"
#include "items/ItemDef.hpp"
#include "lighting/Lighting.hpp"
#include "objects/EntityDef.hpp"
#include "objects/rigging.hpp"
#include "settings.hpp"
#include "util/ThreadPool.hpp"
#include "voxels/Block.hpp"
#include "voxels/Chunk.hpp"
#include "voxels/Chunks.hpp"
#include "voxels/blocks_agent.hpp"
#include "world/LevelEvents.hpp"
#include "world/generator/GeneratorDef.hpp"
#include "world/generator/VoxelFragment.hpp"
#include "world/generator/WorldGenerator.hpp"
#include <map>
#include "world/generator/GeneratorDef.hpp"
#include "world/generator/VoxelFragment.hpp"

#include "../vxoracle.h"

// ---- symbols the light/mesh path never calls but the linker wants ----------

ContentPackRuntime::~ContentPackRuntime() = default;

int Mesh::meshesCount = 0;
int Mesh::drawCalls = 0;
Mesh::Mesh(const float*, size_t, const int*, size_t, const VertexAttribute*)
    : vao(0), vbo(0), ibo(0), vertices(0), indices(0), vertexSize(0) {
}
Mesh::~Mesh() = default;

// ContentGfxCache: the real constructor needs Assets + Atlas (libpng, GL).
// The harness fills the same two private tables from the vx_content instead.
namespace {
    const vx_content* g_pending_content = nullptr;
}

ContentGfxCache::ContentGfxCache(
    const Content& content, const Assets& assets, const GraphicsSettings& settings
)
    : content(content), assets(assets), settings(settings) {
    refresh();
}
ContentGfxCache::~ContentGfxCache() = default;

void ContentGfxCache::refresh() {
    const vx_content* src = g_pending_content;
    sideregions = std::make_unique<UVRegion[]>(src->n_defs * 6);
    for (int i = 0; i < src->n_defs * 6; i++) {
        const float* r = src->uv_regions + i * 4;
        sideregions[i] = UVRegion(r[0], r[1], r[2], r[3]);
    }
    models.clear();
    for (int id = 0; id < src->n_defs; id++) {
        const vx_block_def& d = src->defs[id];
        if (d.model != VX_MODEL_CUSTOM) {
            continue;
        }
        model::Model m;
        for (int k = 0; k < d.model_mesh_count; k++) {
            const vx_model_mesh& sm = src->meshes[d.model_mesh_begin + k];
            model::Mesh& mesh = m.addMesh("");
            for (int v = 0; v < sm.vertex_count; v++) {
                const vx_model_vertex& sv = src->vertices[sm.vertex_begin + v];
                mesh.vertices.push_back(model::Vertex {
                    glm::vec3(sv.coord[0], sv.coord[1], sv.coord[2]),
                    glm::vec2(sv.uv[0], sv.uv[1]),
                    glm::vec3(sv.normal[0], sv.normal[1], sv.normal[2])});
            }
        }
        models[static_cast<blockid_t>(id)] = std::move(m);
    }
}

const model::Model& ContentGfxCache::getModel(blockid_t id) const {
    const auto& found = models.find(id);
    if (found == models.end()) {
        throw std::runtime_error("model not found");
    }
    return found->second;
}

// ---- world ------------------------------------------------------------------

#include "ref_world.hpp"

static std::unique_ptr<Content> make_content(const vx_content* src) {
    UptrsMap<std::string, Block> defs;
    std::vector<Block*> order;
    auto groups = std::make_unique<DrawGroups>();
    for (int id = 0; id < src->n_defs; id++) {
        const vx_block_def& s = src->defs[id];
        std::string name = "vx:block" + std::to_string(id);
        auto block = std::make_unique<Block>(name);
        Block& def = *block;
        for (int c = 0; c < 4; c++) def.emission[c] = s.emission[c];
        def.drawGroup = s.draw_group;
        def.model = static_cast<BlockModel>(s.model);
        def.culling = static_cast<CullingMode>(s.culling);
        def.lightPassing = s.light_passing;
        def.skyLightPassing = s.sky_light_passing;
        def.shadeless = s.shadeless;
        def.ambientOcclusion = s.ambient_occlusion;
        def.translucent = s.translucent;
        def.rotatable = s.rotation_profile != VX_ROT_NONE;
        if (s.rotation_profile == VX_ROT_PIPE) {
            def.rotations = BlockRotProfile::PIPE;
        } else if (s.rotation_profile == VX_ROT_PANE) {
            def.rotations = BlockRotProfile::PANE;
        }
        def.hitboxes.clear();
        if (s.has_hitbox) {
            def.hitboxes.push_back(AABB(
                glm::vec3(s.hitbox_a[0], s.hitbox_a[1], s.hitbox_a[2]),
                glm::vec3(s.hitbox_b[0], s.hitbox_b[1], s.hitbox_b[2])));
        }
        // runtime info, as content/ContentBuilder.cpp:29-37 derives it
        def.rt.id = static_cast<blockid_t>(id);
        def.rt.emissive = s.emission[0] | s.emission[1] | s.emission[2] | s.emission[3];
        def.rt.solid = s.solid;
        order.push_back(&def);
        groups->insert(def.drawGroup);
        defs[name] = std::move(block);
    }
    auto indices = std::make_unique<ContentIndices>(
        ContentUnitIndices<Block>(order),
        ContentUnitIndices<ItemDef>({}),
        ContentUnitIndices<EntityDef>({}));
    ResourceIndicesSet resources {};
    return std::make_unique<Content>(
        std::move(indices), std::move(groups),
        ContentUnitDefs<Block>(std::move(defs)),
        ContentUnitDefs<ItemDef>({}),
        ContentUnitDefs<EntityDef>({}),
        ContentUnitDefs<GeneratorDef>({}),
        UptrsMap<std::string, ContentPackRuntime> {},
        UptrsMap<std::string, BlockMaterial> {},
        UptrsMap<std::string, rigging::SkeletonConfig> {},
        std::move(resources), dv::value(nullptr));
}

static BlocksRenderer& renderer_for(vxo_world* w, uint64_t capacity) {
    if (!w->renderer || w->rendererCapacity != capacity) {
        w->renderer = std::make_unique<BlocksRenderer>(
            capacity, *w->content, *w->cache, w->settings);
        w->rendererCapacity = capacity;
    }
    return *w->renderer;
}

// ---- terrain fill through the reference's own WorldGenerator -----------------
// GeneratorScript is the reference's abstract scripting seam (world/generator/
// GeneratorDef.hpp:125-187); the engine implements it in Lua, this harness
// natively: heightmaps and biome-parameter maps come from the tables the test
// supplies (bpd = 1, so a 17x17 request is the chunk's 16x16 map + one padding
// row / column that WorldGenerator crops away), no structures.  One biome
// parameter whose value is the biome's index makes choose_biome (WorldGenerator.
// cpp:118-143) pick exactly the biome the test names.
namespace {

struct TableScript : GeneratorScript {
    std::map<std::pair<int, int>, const vx_gen_chunk*> chunks;

    const vx_gen_chunk* find(const glm::ivec2& offset) const {
        auto it = chunks.find({floordiv<CHUNK_W>(offset.x), floordiv<CHUNK_D>(offset.y)});
        return it == chunks.end() ? nullptr : it->second;
    }
    template <class F>
    static std::shared_ptr<Heightmap> make(const glm::ivec2& size, F value) {
        std::vector<float> buf((size_t)size.x * size.y);
        for (int z = 0; z < size.y; z++)
            for (int x = 0; x < size.x; x++)
                buf[(size_t)z * size.x + x] = value(std::min(x, CHUNK_W - 1), std::min(z, CHUNK_D - 1));
        return std::make_shared<Heightmap>(size.x, size.y, std::move(buf));
    }
    void initialize(uint64_t) override {}
    std::shared_ptr<Heightmap> generateHeightmap(const glm::ivec2& offset, const glm::ivec2& size, uint,
                                                 const std::vector<std::shared_ptr<Heightmap>>&) override {
        const vx_gen_chunk* c = find(offset);
        return make(size, [&](int x, int z) { return c ? c->heights[z * CHUNK_W + x] : 0.0f; });
    }
    std::vector<std::shared_ptr<Heightmap>> generateParameterMaps(const glm::ivec2& offset, const glm::ivec2& size,
                                                                  uint) override {
        const vx_gen_chunk* c = find(offset);
        return {make(size, [&](int x, int z) { return c ? (float)c->biomes[z * CHUNK_W + x] : 0.0f; })};
    }
    std::vector<Placement> placeStructuresWide(const glm::ivec2&, const glm::ivec2&, uint) override { return {}; }
    std::vector<Placement> placeStructures(const glm::ivec2&, const glm::ivec2&, const std::shared_ptr<Heightmap>&,
                                           uint) override { return {}; }
};

BlocksLayers make_layers(const vx_gen_layer* layers, int count, uint32_t last) {
    BlocksLayers out;
    for (int k = 0; k < count; k++) {
        BlocksLayer layer {"", layers[k].height, layers[k].below_sea_level != 0, {}};
        layer.rt.id = layers[k].block;  // what GeneratorDef::prepare resolves from the name
        out.layers.push_back(layer);
    }
    out.lastLayersHeight = last;
    return out;
}

}  // namespace


extern "C" {

const char* vxo_kind(void) {
    return "reference";
}

vxo_world* vxo_create(const vx_content* content, const vx_settings* settings,
                      int32_t ox, int32_t oz, int32_t w, int32_t d) {
    auto world = new vxo_world();
    world->content = make_content(content);
    world->settings.graphics.backlight.set(settings->backlight != 0);
    world->settings.graphics.denseRender.set(settings->dense_render != 0);
    g_pending_content = content;
    static const int dummy_assets = 0;
    world->cache = std::make_unique<ContentGfxCache>(
        *world->content, reinterpret_cast<const Assets&>(dummy_assets),
        world->settings.graphics);
    g_pending_content = nullptr;
    world->chunks = std::make_unique<Chunks>(
        w, d, 0, 0, &world->events, *world->content->getIndices());
    // the engine always follows the ctor with setCenter (LevelController.cpp:58-60)
    world->chunks->setCenter((ox + w / 2) * CHUNK_W, (oz + d / 2) * CHUNK_D);
    if (world->chunks->getOffsetX() != ox || world->chunks->getOffsetY() != oz) {
        delete world;
        return nullptr;
    }
    world->lighting = std::make_unique<Lighting>(*world->content, *world->chunks);
    return world;
}

void vxo_destroy(vxo_world* world) {
    delete world;
}

int vxo_put_chunk(vxo_world* world, int32_t cx, int32_t cz,
                  const vx_voxel* voxels, const vx_light* light) {
    auto chunk = std::make_shared<Chunk>(cx, cz);
    static_assert(sizeof(voxel) == sizeof(vx_voxel));
    std::memcpy(chunk->voxels, voxels, sizeof(vx_voxel) * CHUNK_VOL);
    if (light) {
        chunk->lightmap.set(light);
    }
    chunk->updateHeights();
    chunk->flags.loaded = true;
    return world->chunks->putChunk(chunk) ? 0 : -1;
}

int vxo_put_chunk_encoded(vxo_world* world, int32_t cx, int32_t cz, const uint8_t* voxels,
                          const uint8_t* lights) {
    auto chunk = std::make_shared<Chunk>(cx, cz);
    chunk->decode(voxels);  // voxels/Chunk.cpp:88-97
    if (lights) {
        auto decoded = Lightmap::decode(lights);  // lighting/Lightmap.cpp:26-34
        chunk->lightmap.set(decoded.get());
    }
    chunk->updateHeights();
    chunk->flags.loaded = true;
    return world->chunks->putChunk(chunk) ? 0 : -1;
}

int vxo_encode_chunk(vxo_world* world, int32_t cx, int32_t cz, uint8_t* voxels_out,
                     uint8_t* lights_out) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    if (voxels_out) {
        auto data = chunk->encode();  // voxels/Chunk.cpp:78-86
        std::memcpy(voxels_out, data.get(), CHUNK_DATA_LEN);
    }
    if (lights_out) {
        auto data = chunk->lightmap.encode();  // lighting/Lightmap.cpp:18-24
        std::memcpy(lights_out, data.get(), LIGHTMAP_DATA_LEN);
    }
    return 0;
}

int vxo_prebuild_sky(vxo_world* world, int32_t cx, int32_t cz) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    Lighting::prebuildSkyLight(*chunk, *world->content->getIndices());
    return 0;
}

int vxo_light_chunk(vxo_world* world, int32_t cx, int32_t cz, int32_t expand) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    // logic/ChunksController.cpp:116-123
    if (!(expand & VX_LIGHT_NO_SKY)) world->lighting->buildSkyLight(cx, cz);
    world->lighting->onChunkLoaded(cx, cz, (expand & VX_LIGHT_EXPAND) != 0);
    chunk->flags.lighted = true;
    return 0;
}

int vxo_light_clear(vxo_world* world) {
    world->lighting->clear();
    return 0;
}

int vxo_set_block(vxo_world* world, int32_t x, int32_t y, int32_t z,
                  uint16_t id, uint16_t state) {
    // logic/BlocksController.cpp:63-104
    blocks_agent::set(*world->chunks, x, y, z, id, int2blockstate(state));
    world->lighting->onBlockSet(x, y, z, id);
    return 0;
}

int vxo_get_light(vxo_world* world, int32_t cx, int32_t cz, vx_light* out) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    std::memcpy(out, chunk->lightmap.getLights(), sizeof(vx_light) * CHUNK_VOL);
    return 0;
}

int vxo_get_voxels(vxo_world* world, int32_t cx, int32_t cz, vx_voxel* out) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    std::memcpy(out, chunk->voxels, sizeof(vx_voxel) * CHUNK_VOL);
    return 0;
}

int vxo_get_info(vxo_world* world, int32_t cx, int32_t cz, vx_chunk_info* out) {
    std::memset(out, 0, sizeof(*out));
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return 0;
    out->present = 1;
    out->bottom = chunk->bottom;
    out->top = chunk->top;
    out->highest_point = chunk->lightmap.highestPoint;
    out->modified = chunk->flags.modified;
    out->lighted = chunk->flags.lighted;
    return 0;
}

uint64_t vxo_extrle_encode(const uint8_t* src, uint64_t n, uint8_t* dst, int32_t wide) {
    return wide ? extrle::encode16(src, n, dst) : extrle::encode(src, n, dst);  // coders/rle.cpp
}

uint64_t vxo_extrle_decode(const uint8_t* src, uint64_t n, uint8_t* dst, int32_t wide) {
    return wide ? extrle::decode16(src, n, dst) : extrle::decode(src, n, dst);
}

int vxo_terrain_fill(vxo_world* world, const vx_gen_def* src, const vx_gen_chunk* chunks, int32_t n) {
    if (n <= 0) return 0;
    GeneratorDef def("vx:tables");
    auto script = std::make_unique<TableScript>();
    int x0 = chunks[0].cx, x1 = x0, z0 = chunks[0].cz, z1 = z0;
    for (int i = 0; i < n; i++) {
        script->chunks[{chunks[i].cx, chunks[i].cz}] = &chunks[i];
        x0 = std::min(x0, chunks[i].cx); x1 = std::max(x1, chunks[i].cx);
        z0 = std::min(z0, chunks[i].cz); z1 = std::max(z1, chunks[i].cz);
    }
    def.script = std::move(script);
    def.seaLevel = src->sea_level;
    def.biomeParameters = 1;
    def.biomesBPD = 1;
    def.heightsBPD = 1;
    def.biomesInterpolation = InterpolationType::NEAREST;
    def.heightsInterpolation = InterpolationType::NEAREST;
    def.wideStructsChunksRadius = 1;  // 0 would alias the prototype-creation level (WorldGenerator.cpp:48-56)
    for (int b = 0; b < src->n_biomes; b++) {
        const vx_gen_biome& B = src->biomes[b];
        Biome biome;
        biome.name = "biome" + std::to_string(b);
        biome.parameters = {BiomeParameter {(float)b, 1.0f}};
        std::vector<WeightedEntry> plants;
        for (int k = 0; k < B.plants_count; k++) {
            WeightedEntry e {"", src->plants[B.plants_begin + k].weight, {}};
            e.rt.id = src->plants[B.plants_begin + k].block;
            plants.push_back(e);
        }
        biome.plants = BiomeElementList(std::move(plants), B.plants_chance);
        biome.structures = BiomeElementList({}, 0.0f);
        biome.groundLayers = make_layers(src->layers + B.ground_begin, B.ground_count, B.ground_last_layers_height);
        biome.seaLayers = make_layers(src->layers + B.sea_begin, B.sea_count, B.sea_last_layers_height);
        def.biomes.push_back(std::move(biome));
    }
    try {
        WorldGenerator generator(def, *world->content, 0);
        // ChunksController::update -> generator->update(centerX, centerY, loadDistance)
        generator.update((x0 + x1) / 2, (z0 + z1) / 2, std::max(x1 - x0, z1 - z0) / 2 + 2);
        std::vector<vx_voxel> voxels(CHUNK_VOL);
        for (int i = 0; i < n; i++) {
            static_assert(sizeof(voxel) == sizeof(vx_voxel));
            generator.generate(reinterpret_cast<voxel*>(voxels.data()), chunks[i].cx, chunks[i].cz);
            if (vxo_put_chunk(world, chunks[i].cx, chunks[i].cz, voxels.data(), nullptr)) return -1;
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "vxo_terrain_fill: %s\n", e.what());
        return -1;
    }
    return 0;
}

int vxo_mark_modified(vxo_world* world, int32_t cx, int32_t cz) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    chunk->flags.modified = true;
    return 0;
}

int vxo_clear_modified(vxo_world* world) {
    for (const auto& chunk : world->chunks->getChunks()) {
        if (chunk) chunk->flags.modified = false;
    }
    return 0;
}

static void fill_info(const MeshResult& r, int cx, int cz, vx_mesh_info* info) {
    std::memset(info, 0, sizeof(*info));
    info->cx = cx;
    info->cz = cz;
    info->cancelled = r.cancelled;
    if (r.cancelled) return;
    info->n_vertex_floats = r.data.mesh.vertices.size();
    info->n_indices = r.data.mesh.indices.size();
    info->n_entries = r.data.sortingMesh.entries.size();
    uint32_t total = 0;
    for (const auto& e : r.data.sortingMesh.entries) total += e.vertexData.size();
    info->n_entry_floats = total;
}

int vxo_mesh_chunk(vxo_world* world, int32_t cx, int32_t cz, uint64_t capacity,
                   vx_mesh_info* info) {
    Chunk* chunk = world->chunks->getChunk(cx, cz);
    if (!chunk) return -1;
    BlocksRenderer& renderer = renderer_for(world, capacity);
    // graphics/render/ChunksRenderer.cpp:42-51 (RendererWorker::operator())
    renderer.build(chunk, world->chunks.get());
    world->last = MeshResult {};
    if (!renderer.isCancelled()) {
        world->last.cancelled = false;
        world->last.data = renderer.createMesh();
    }
    fill_info(world->last, cx, cz, info);
    return 0;
}

int vxo_mesh_read(vxo_world* world, float* vertices, int32_t* indices,
                  vx_sort_entry* entries, float* entry_floats) {
    const MeshResult& r = world->last;
    if (r.cancelled) return -1;
    if (vertices) {
        std::memcpy(vertices, r.data.mesh.vertices.data(),
                    r.data.mesh.vertices.size() * sizeof(float));
    }
    if (indices) {
        std::memcpy(indices, r.data.mesh.indices.data(),
                    r.data.mesh.indices.size() * sizeof(int));
    }
    uint32_t offset = 0;
    size_t i = 0;
    for (const auto& e : r.data.sortingMesh.entries) {
        if (entries) {
            entries[i].position[0] = e.position.x;
            entries[i].position[1] = e.position.y;
            entries[i].position[2] = e.position.z;
            entries[i].first_float = offset;
            entries[i].n_floats = e.vertexData.size();
        }
        if (entry_floats) {
            std::memcpy(entry_floats + offset, e.vertexData.data(),
                        e.vertexData.size() * sizeof(float));
        }
        offset += e.vertexData.size();
        i++;
    }
    return 0;
}

// ---- CPU baseline helpers ----------------------------------------------------

namespace {
    struct BatchResult {
        bool cancelled;
        uint64_t indices;
    };

    // Same shape as RendererWorker (graphics/render/ChunksRenderer.cpp:21-52):
    // one BlocksRenderer per pool worker, build + createMesh per job.
    class BatchWorker : public util::Worker<Chunk*, BatchResult> {
        const Chunks& chunks;
        BlocksRenderer renderer;
    public:
        BatchWorker(vxo_world* w, uint64_t capacity)
            : chunks(*w->chunks),
              renderer(capacity, *w->content, *w->cache, w->settings) {
        }
        BatchResult operator()(Chunk* const& chunk) override {
            renderer.build(chunk, &chunks);
            if (renderer.isCancelled()) {
                return BatchResult {true, 0};
            }
            auto mesh = renderer.createMesh();
            return BatchResult {false, mesh.mesh.indices.size()};
        }
    };
}

double vxo_mesh_batch(vxo_world* world, const vx_chunk_key* keys, int32_t n,
                      uint64_t capacity, int32_t threads, uint64_t* faces_out) {
    std::vector<Chunk*> jobs;
    for (int i = 0; i < n; i++) {
        if (Chunk* c = world->chunks->getChunk(keys[i].cx, keys[i].cz)) {
            jobs.push_back(c);
        }
    }
    uint64_t indices = 0;
    size_t done = 0;
    // the reference's own pool (util/ThreadPool.hpp), as ChunksRenderer uses it
    util::ThreadPool<Chunk*, BatchResult> pool(
        "vxref-mesh-pool",
        [&]() { return std::make_shared<BatchWorker>(world, capacity); },
        [&](BatchResult& r) {
            indices += r.indices;
            done++;
        },
        threads);
    pool.setStopOnFail(false);
    auto t0 = std::chrono::steady_clock::now();
    for (Chunk* c : jobs) pool.enqueueJob(c);
    while (done < jobs.size()) {
        pool.update();
        std::this_thread::yield();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (faces_out) *faces_out = indices / 6;
    return std::chrono::duration<double>(t1 - t0).count();
}

double vxo_light_batch(vxo_world* world, const vx_chunk_key* keys, int32_t n,
                       int32_t expand) {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) {
        vxo_light_chunk(world, keys[i].cx, keys[i].cz, expand);
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
