// See GpuChunkBuild.hpp.  Reference lines cited are under /root/reference/src.
#include "GpuChunkBuild.hpp"

#include <algorithm>
#include <cstring>

#include "graphics/render/commons.hpp"
#include "lighting/Lightmap.hpp"
#include "maths/aabb.hpp"
#include "voxels/Block.hpp"

namespace vxhost {

static_assert(sizeof(voxel) == sizeof(vx_voxel), "voxel[] is vx_voxel[] byte for byte");
static_assert(sizeof(light_t) == sizeof(vx_light), "Lightmap::map is vx_light[] byte for byte");
static_assert(CHUNK_VOL == VX_CHUNK_VOL, "chunk geometry");
static_assert(CHUNK_VERTEX_SIZE == VX_VERTEX_SIZE, "vertex layout");

vx_content FlatContent::view() const {
    return vx_content {defs.data(), (int32_t)defs.size(), uv.data(), meshes.data(), (int32_t)meshes.size(),
                       vertices.data(), (int32_t)vertices.size()};
}

FlatContent flattenContent(const Content& content, const ContentGfxCache& cache) {
    FlatContent out;
    const auto& blocks = content.getIndices()->blocks;
    const size_t n = blocks.count();
    out.defs.resize(n);
    out.uv.resize(n * 6 * 4);
    for (size_t id = 0; id < n; id++) {
        const Block& b = blocks.require(id);
        vx_block_def& d = out.defs[id];
        std::memset(&d, 0, sizeof(d));
        std::memcpy(d.emission, b.emission, 4);
        d.draw_group = b.drawGroup;
        d.model = static_cast<uint8_t>(b.model);      // BlockModel order == vx_block_model (Block.hpp:86-97)
        d.culling = static_cast<uint8_t>(b.culling);  // CullingMode order == vx_culling_mode
        d.rotation_profile = !b.rotatable ? VX_ROT_NONE
                             : b.rotations.name == BlockRotProfile::PIPE_NAME ? VX_ROT_PIPE : VX_ROT_PANE;
        d.light_passing = b.lightPassing;
        d.sky_light_passing = b.skyLightPassing;
        d.shadeless = b.shadeless;
        d.ambient_occlusion = b.ambientOcclusion;
        d.translucent = b.translucent;
        d.solid = b.rt.solid;  // content/ContentBuilder.cpp:31-37
        d.has_hitbox = !b.hitboxes.empty();
        if (!b.hitboxes.empty()) {  // BlocksRenderer::blockAABB unites them (BlocksRenderer.cpp:233-237)
            AABB box = b.hitboxes[0];
            for (const AABB& h : b.hitboxes) {
                box.a = glm::min(box.a, h.a);
                box.b = glm::max(box.b, h.b);
            }
            d.hitbox_a[0] = box.a.x; d.hitbox_a[1] = box.a.y; d.hitbox_a[2] = box.a.z;
            d.hitbox_b[0] = box.b.x; d.hitbox_b[1] = box.b.y; d.hitbox_b[2] = box.b.z;
        }
        d.model_mesh_begin = (int32_t)out.meshes.size();
        if (b.model == BlockModel::custom) {
            for (const auto& mesh : cache.getModel(static_cast<blockid_t>(id)).meshes) {
                out.meshes.push_back(vx_model_mesh {(int32_t)out.vertices.size(), (int32_t)mesh.vertices.size()});
                for (const auto& v : mesh.vertices)
                    out.vertices.push_back(vx_model_vertex {{v.coord.x, v.coord.y, v.coord.z},
                                                            {v.uv.x, v.uv.y},
                                                            {v.normal.x, v.normal.y, v.normal.z}});
            }
        }
        d.model_mesh_count = (int32_t)out.meshes.size() - d.model_mesh_begin;
        for (int side = 0; side < 6; side++) {
            const UVRegion& r = cache.getRegion(static_cast<blockid_t>(id), side);
            float* o = &out.uv[(id * 6 + side) * 4];
            o[0] = r.u1; o[1] = r.v1; o[2] = r.u2; o[3] = r.v2;
        }
    }
    return out;
}

// ---------------------------------------------------------------------------
GpuContext::GpuContext(const Content& content, const ContentGfxCache& cache, const EngineSettings& settings,
                       int device) {
    FlatContent flat = flattenContent(content, cache);
    vx_content c = flat.view();
    // settings.hpp:64-75 (the EngineSettings object is non-const in the engine; get() is not const-qualified)
    auto& gs = const_cast<EngineSettings&>(settings).graphics;
    vx_settings st {gs.backlight.get() ? 1 : 0, gs.denseRender.get() ? 1 : 0};
    ctx = vx_ctx_create(&c, &st, device);
    if (!ctx) throw GpuError(std::string("vx_ctx_create: ") + vx_last_error(nullptr));
}

GpuContext::~GpuContext() {
    if (ctx) vx_ctx_destroy(ctx);
}

void GpuContext::check(int rc, const char* what) const {
    if (rc != 0) throw GpuError(std::string(what) + ": " + vx_last_error(ctx));
}

void GpuContext::bind(const Chunks& chunks) {
    const int nox = chunks.getOffsetX(), noz = chunks.getOffsetY();
    const int nw = chunks.getWidth(), nd = chunks.getHeight();
    if (nox == ox && noz == oz && nw == w && nd == d) return;
    check(vx_window_configure(ctx, nox, noz, nw, nd), "vx_window_configure");
    ox = nox; oz = noz; w = nw; d = nd;
}

void GpuContext::upload(const std::vector<const Chunk*>& batch) {
    std::vector<vx_chunk_in> in;
    in.reserve(batch.size());
    for (const Chunk* c : batch)
        in.push_back(vx_chunk_in {c->x, c->z, reinterpret_cast<const vx_voxel*>(c->voxels),
                                  c->flags.loadedLights ? reinterpret_cast<const vx_light*>(c->lightmap.map) : nullptr});
    if (!in.empty()) check(vx_chunks_put(ctx, in.data(), (int32_t)in.size()), "vx_chunks_put");
}

void GpuContext::upload(const Chunk& chunk) {
    upload(std::vector<const Chunk*> {&chunk});
}

void GpuContext::remove(int cx, int cz) {
    vx_chunk_key k {cx, cz};
    check(vx_chunks_remove(ctx, &k, 1), "vx_chunks_remove");
}

// ---------------------------------------------------------------------------
GpuLighting::GpuLighting(GpuContext& gpu, Chunks& chunks) : gpu(gpu), chunks(chunks) {
}

// Lightmaps and the flags the solver touches, device -> the host's Chunk objects.
void GpuLighting::pull(const std::vector<glm::ivec2>& keys) {
    for (const glm::ivec2& k : keys) {
        Chunk* chunk = chunks.getChunk(k.x, k.y);
        if (chunk == nullptr) continue;
        vx_chunk_info info;
        gpu.check(vx_chunk_get_info(gpu.handle(), k.x, k.y, &info), "vx_chunk_get_info");
        if (!info.present) continue;
        gpu.check(vx_light_get(gpu.handle(), k.x, k.y, reinterpret_cast<vx_light*>(chunk->lightmap.map)),
                  "vx_light_get");
        chunk->lightmap.highestPoint = info.highest_point;
        // LightSolver sets flags.modified and never clears it (LightSolver.cpp:29,74,111)
        if (info.modified) chunk->flags.modified = true;
    }
}

void GpuLighting::pullAround(int cx, int cz) {
    std::vector<glm::ivec2> keys;
    for (int dz = -1; dz <= 1; dz++)
        for (int dx = -1; dx <= 1; dx++) keys.emplace_back(cx + dx, cz + dz);
    pull(keys);
}

void GpuLighting::clear() {
    gpu.check(vx_light_clear(gpu.handle()), "vx_light_clear");
    for (const auto& chunk : chunks.getChunks()) {  // Lighting.cpp:28-39
        if (chunk == nullptr) continue;
        chunk->lightmap.clear();
    }
}

void GpuLighting::prebuildSkyLight(const std::vector<Chunk*>& batch) {
    std::vector<vx_chunk_key> keys;
    for (Chunk* c : batch) keys.push_back(vx_chunk_key {c->x, c->z});
    if (keys.empty()) return;
    gpu.check(vx_light_prebuild(gpu.handle(), keys.data(), (int32_t)keys.size()), "vx_light_prebuild");
    std::vector<glm::ivec2> k2;
    for (Chunk* c : batch) k2.emplace_back(c->x, c->z);
    pull(k2);
}

void GpuLighting::prebuildSkyLight(Chunk& chunk) {
    prebuildSkyLight(std::vector<Chunk*> {&chunk});
}

void GpuLighting::buildSkyLight(int cx, int cz) {
    skyPending.emplace_back(cx, cz);
}

void GpuLighting::onChunkLoaded(int cx, int cz, bool expand) {
    const glm::ivec2 key(cx, cz);
    auto it = std::find(skyPending.begin(), skyPending.end(), key);
    const bool sky = it != skyPending.end();
    if (sky) skyPending.erase(it);
    lightChunks({key}, expand, sky);
}

void GpuLighting::lightChunks(const std::vector<glm::ivec2>& keys, bool expand, bool sky) {
    if (keys.empty()) return;
    std::vector<vx_chunk_key> k;
    for (const auto& key : keys) k.push_back(vx_chunk_key {key.x, key.y});
    const int32_t flags = (expand ? VX_LIGHT_EXPAND : 0) | (sky ? 0 : VX_LIGHT_NO_SKY);
    gpu.check(vx_light_build(gpu.handle(), k.data(), (int32_t)k.size(), flags), "vx_light_build");
    // light reaches the eight neighbours at most (15 voxels): pull each touched chunk once
    std::vector<glm::ivec2> touched;
    for (const auto& key : keys)
        for (int dz = -1; dz <= 1; dz++)
            for (int dx = -1; dx <= 1; dx++) touched.emplace_back(key.x + dx, key.y + dz);
    std::sort(touched.begin(), touched.end(), [](const glm::ivec2& a, const glm::ivec2& b) {
        return a.y != b.y ? a.y < b.y : a.x < b.x;
    });
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    pull(touched);
}

void GpuLighting::onBlockSet(int x, int y, int z, blockid_t id) {
    // the host already wrote the voxel (Chunks::set -> blocks_agent::set); the device repeats the
    // write on its mirror and relights (blocks_agent.cpp:10-77 + Lighting.cpp:163-215)
    const voxel* vox = chunks.get(x, y, z);
    if (vox == nullptr) return;
    vx_edit e {x, y, z, id, blockstate2int(vox->state)};
    gpu.check(vx_light_edit(gpu.handle(), &e, 1), "vx_light_edit");
    pullAround(floordiv<CHUNK_W>(x), floordiv<CHUNK_D>(z));
}

void GpuLighting::setBlocks(const std::vector<vx_edit>& edits) {
    if (edits.empty()) return;
    gpu.check(vx_light_edit(gpu.handle(), edits.data(), (int32_t)edits.size()), "vx_light_edit");
    std::vector<glm::ivec2> touched;
    for (const vx_edit& e : edits) {
        const int cx = floordiv<CHUNK_W>(e.x), cz = floordiv<CHUNK_D>(e.z);
        for (int dz = -1; dz <= 1; dz++)
            for (int dx = -1; dx <= 1; dx++) touched.emplace_back(cx + dx, cz + dz);
    }
    std::sort(touched.begin(), touched.end(), [](const glm::ivec2& a, const glm::ivec2& b) {
        return a.y != b.y ? a.y < b.y : a.x < b.x;
    });
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    // voxels, Chunk::bottom/top and the lightmaps come back from the device mirror
    for (const glm::ivec2& k : touched) {
        Chunk* chunk = chunks.getChunk(k.x, k.y);
        if (chunk == nullptr) continue;
        vx_chunk_info info;
        gpu.check(vx_chunk_get_info(gpu.handle(), k.x, k.y, &info), "vx_chunk_get_info");
        if (!info.present) continue;
        gpu.check(vx_chunk_get_voxels(gpu.handle(), k.x, k.y, reinterpret_cast<vx_voxel*>(chunk->voxels)),
                  "vx_chunk_get_voxels");
        chunk->bottom = info.bottom;
        chunk->top = info.top;
    }
    pull(touched);
}

// ---------------------------------------------------------------------------
GpuRendererBatcher::GpuRendererBatcher(GpuContext& gpu, const Chunks& chunks, const EngineSettings& settings,
                                       Consumer consumer)
    : gpu(gpu), chunks(chunks), settings(settings), consumer(std::move(consumer)) {
}

bool GpuRendererBatcher::request(const std::shared_ptr<Chunk>& chunk) {
    chunk->flags.modified = false;  // ChunksRenderer.cpp:98
    const glm::ivec2 key(chunk->x, chunk->z);
    if (inwork.find(key) != inwork.end()) return false;  // :107-109
    inwork[key] = true;
    queue.push_back(chunk);
    return true;
}

ChunkMeshData GpuRendererBatcher::fetch(int index) const {
    vx_mesh_info info;
    gpu.check(vx_mesh_get_info(gpu.handle(), index, &info), "vx_mesh_get_info");
    util::Buffer<float> vertices(info.n_vertex_floats);
    util::Buffer<int> indices(info.n_indices);
    std::vector<vx_sort_entry> entries(info.n_entries);
    std::vector<float> entryFloats(info.n_entry_floats);
    gpu.check(vx_mesh_read(gpu.handle(), index, vertices.data(), indices.data(), entries.data(), entryFloats.data()),
              "vx_mesh_read");
    ChunkMeshData out {
        MeshData(std::move(vertices), std::move(indices),
                 util::Buffer<VertexAttribute>(CHUNK_VATTRS, sizeof(CHUNK_VATTRS) / sizeof(VertexAttribute))),
        SortingMeshData {}};  // BlocksRenderer::createMesh (BlocksRenderer.cpp:654-664)
    out.sortingMesh.entries.reserve(entries.size());
    for (const vx_sort_entry& e : entries)
        out.sortingMesh.entries.push_back(SortingMeshEntry {
            glm::vec3(e.position[0], e.position[1], e.position[2]),
            util::Buffer<float>(entryFloats.data() + e.first_float, e.n_floats), 0});
    return out;
}

ChunkMeshData GpuRendererBatcher::renderNow(const std::shared_ptr<Chunk>& chunk) {
    chunk->flags.modified = false;
    vx_chunk_key k {chunk->x, chunk->z};
    auto& gs = const_cast<EngineSettings&>(settings).graphics;
    // the main-thread renderer is created with chunkMaxVertices (ChunksRenderer.cpp:85-88)
    gpu.check(vx_mesh_build(gpu.handle(), &k, 1, (uint64_t)gs.chunkMaxVertices.get()), "vx_mesh_build");
    return fetch(0);
}

void GpuRendererBatcher::update() {
    if (queue.empty()) return;
    std::vector<std::shared_ptr<Chunk>> batch;
    batch.swap(queue);
    std::vector<vx_chunk_key> keys;
    for (const auto& c : batch) keys.push_back(vx_chunk_key {c->x, c->z});
    auto& gs = const_cast<EngineSettings&>(settings).graphics;
    // pool workers use the dense capacity under denseRender (ChunksRenderer.cpp:33-39)
    const uint64_t capacity = gs.denseRender.get() ? gs.chunkMaxVerticesDense.get() : gs.chunkMaxVertices.get();
    gpu.check(vx_mesh_build(gpu.handle(), keys.data(), (int32_t)keys.size(), capacity), "vx_mesh_build");
    for (size_t i = 0; i < batch.size(); i++) {
        const glm::ivec2 key(batch[i]->x, batch[i]->z);
        vx_mesh_info info;
        gpu.check(vx_mesh_get_info(gpu.handle(), (int32_t)i, &info), "vx_mesh_get_info");
        RendererResult result {key, info.cancelled != 0, ChunkMeshData {}};
        if (!result.cancelled) result.meshData = fetch((int)i);
        // the pool's result consumer (ChunksRenderer.cpp:73-82): the mesh is created there unless the
        // job was cancelled; the key leaves `inwork` either way
        if (consumer) consumer(result);
        inwork.erase(key);
    }
}

void GpuRendererBatcher::unload(const Chunk* chunk) {
    const glm::ivec2 key(chunk->x, chunk->z);
    for (auto it = queue.begin(); it != queue.end();) {
        if ((*it)->x == key.x && (*it)->z == key.y) {
            it = queue.erase(it);
            inwork.erase(key);
        } else {
            ++it;
        }
    }
}

void GpuRendererBatcher::clear() {
    inwork.clear();  // ChunksRenderer.cpp:122-126
    queue.clear();
}

}  // namespace vxhost
