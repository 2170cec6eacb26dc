// C++ host side of libvxgpu inside VoxelCore: the classes a maintainer drops in next to
// Lighting (lighting/Lighting.hpp:11-28) and ChunksRenderer (graphics/render/ChunksRenderer.hpp:44-82).
// Compiled against the reference's own headers (Chunks, Chunk, Lightmap, Content, ContentGfxCache,
// ChunkMeshData, RendererResult) and include/vxgpu.h; nothing here computes light or geometry —
// every call forwards to the C ABI and copies results into the reference's objects.
//
//   GpuContext          one vx_ctx per Level: Block -> vx_block_def flattening (content/
//                       ContentBuilder.cpp:29-37 runtime fields, frontend/ContentGfxCache.hpp:35-37 UV
//                       regions and custom models), the Chunks window, chunk uploads
//   GpuLighting         same public methods as Lighting; results land in Chunk::lightmap and
//                       Chunk::flags exactly where the reference's solver writes them
//   GpuRendererBatcher  the request / inwork / update / unload / clear flow of ChunksRenderer
//                       (ChunksRenderer.cpp:95-139) with the thread pool replaced by one batched
//                       vx_mesh_build per update(); results are RendererResult objects handed to the
//                       same consumer the reference gives its pool (:73-82)
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <glm/glm.hpp>
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/gtx/hash.hpp>

#include "content/Content.hpp"
#include "frontend/ContentGfxCache.hpp"
#include "graphics/core/Mesh.hpp"
#include "graphics/render/ChunksRenderer.hpp"  // RendererResult, ChunkMeshData
#include "settings.hpp"
#include "voxels/Chunk.hpp"
#include "voxels/Chunks.hpp"

#include "vxgpu.h"
#include "maths/voxmaths.hpp"

namespace vxhost {

struct GpuError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

/// Content + ContentGfxCache as the device wants them (vx_content); owns the arrays.
struct FlatContent {
    std::vector<vx_block_def> defs;
    std::vector<float> uv;  // n_defs * 6 * 4
    std::vector<vx_model_mesh> meshes;
    std::vector<vx_model_vertex> vertices;
    vx_content view() const;
};
/// The cache's regions are taken as they are after ContentGfxCache::refresh, i.e. with the
/// `_opaque` texture swap of culling == OPTIONAL blocks already applied (ContentGfxCache.cpp:27-31).
FlatContent flattenContent(const Content& content, const ContentGfxCache& cache);

class GpuContext {
    vx_ctx* ctx = nullptr;
    int ox = 0, oz = 0, w = 0, d = 0;
public:
    GpuContext(const Content& content, const ContentGfxCache& cache, const EngineSettings& settings,
               int device = 0);
    ~GpuContext();
    GpuContext(const GpuContext&) = delete;
    GpuContext& operator=(const GpuContext&) = delete;

    vx_ctx* handle() const { return ctx; }
    void check(int rc, const char* what) const;  // throws GpuError with vx_last_error

    /// Chunks::Chunks / setCenter / resize: the device window follows the AreaMap2D area.
    void bind(const Chunks& chunks);
    /// Chunks::putChunk: voxels, and the lightmap when it came from the lights cache
    /// (flags.loadedLights).  Call again after the host rewrote a chunk wholesale.
    void upload(const Chunk& chunk);
    void upload(const std::vector<const Chunk*>& chunks);
    void remove(int cx, int cz);
};

class GpuLighting {
    GpuContext& gpu;
    Chunks& chunks;
    std::vector<glm::ivec2> skyPending;  // buildSkyLight calls waiting for their onChunkLoaded

    void pull(const std::vector<glm::ivec2>& keys);
    void pullAround(int cx, int cz);
public:
    GpuLighting(GpuContext& gpu, Chunks& chunks);

    void clear();                                        // Lighting::clear
    void prebuildSkyLight(Chunk& chunk);                 // Lighting::prebuildSkyLight (static there)
    void buildSkyLight(int cx, int cz);                  // folded into the onChunkLoaded that follows it
    void onChunkLoaded(int cx, int cz, bool expand);     // (logic/ChunksController.cpp:118-122 call pairing)
    void onBlockSet(int x, int y, int z, blockid_t id);  // after Chunks::set, as BlocksController does

    /// Batch forms: what ChunksController::buildLights amounts to over a frame, in one call each.
    void prebuildSkyLight(const std::vector<Chunk*>& batch);
    void lightChunks(const std::vector<glm::ivec2>& keys, bool expand, bool sky = true);
    void setBlocks(const std::vector<vx_edit>& edits);   // blocks_agent::set + onBlockSet, host voxels updated too
};

class GpuRendererBatcher {
public:
    using Consumer = std::function<void(RendererResult&)>;
private:
    GpuContext& gpu;
    const Chunks& chunks;
    const EngineSettings& settings;
    Consumer consumer;
    std::unordered_map<glm::ivec2, bool> inwork;     // ChunksRenderer::inwork
    std::vector<std::shared_ptr<Chunk>> queue;       // the pool's job queue

    ChunkMeshData fetch(int index) const;
public:
    GpuRendererBatcher(GpuContext& gpu, const Chunks& chunks, const EngineSettings& settings,
                       Consumer consumer);

    /// ChunksRenderer::render(chunk, important = false): clears flags.modified, refuses a chunk that
    /// is already in work, queues it otherwise.  Returns whether it was queued.
    bool request(const std::shared_ptr<Chunk>& chunk);
    /// ChunksRenderer::render(chunk, important = true): built at once with chunkMaxVertices.
    ChunkMeshData renderNow(const std::shared_ptr<Chunk>& chunk);
    /// ChunksRenderer::update(): builds everything queued (one vx_mesh_build) and hands each result
    /// to the consumer on the calling thread; cancelled results only leave `inwork`.
    void update();
    /// ChunksRenderer::unload: the chunk left the world; a queued request for it is dropped.
    void unload(const Chunk* chunk);
    /// ChunksRenderer::clear: inwork.clear() + threadPool.clearQueue().
    void clear();

    size_t inWorkCount() const { return inwork.size(); }
    size_t queued() const { return queue.size(); }
};

}  // namespace vxhost
