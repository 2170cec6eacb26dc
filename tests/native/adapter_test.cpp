// Self-test of the C++ host adapter (voxelengine-cpp_b200/host/GpuChunkBuild.{hpp,cpp}) against the
// reference's own classes.  TEST INFRASTRUCTURE: built by oracle/ref_build/Makefile into
// oracle/_ref/libvxadapter_test.so together with the reference translation units, and driven by
// tests/test_host_adapter.py (-m gpu), which supplies the block table and a world.
//
// Two identical worlds are made of the reference's objects (Content, Chunks, Chunk).  World A is
// lit and meshed by the reference (Lighting, BlocksRenderer); world B by GpuLighting /
// GpuRendererBatcher, which must leave B's Chunk objects and ChunkMeshData byte-identical to A's:
//   1  Block -> vx_block_def flattening round trip (content/ContentBuilder.cpp:29-37)
//   2  prebuildSkyLight + buildSkyLight + onChunkLoaded (single-call and batch forms)
//   3  the request / inwork / update / unload / clear contract of ChunksRenderer.cpp:95-139 and the
//      meshes themselves (RendererWorker path and the `important` path)
//   4  Chunks::set + onBlockSet edits, then Lighting::clear
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "GpuChunkBuild.hpp"
#include "ref_world.hpp"
#include "voxels/blocks_agent.hpp"

#include "../../oracle/vxoracle.h"

using namespace vxhost;

namespace {

struct Fail {
    char* buf;
    int cap;
    int operator()(int code, const char* fmt, ...) const {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, cap, fmt, ap);
        va_end(ap);
        return code;
    }
};

bool same_chunk(const Chunk& a, const Chunk& b, std::string& why) {
    if (std::memcmp(a.lightmap.map, b.lightmap.map, sizeof(a.lightmap.map)) != 0) {
        int n = 0, first = -1;
        for (int i = 0; i < CHUNK_VOL; i++)
            if (a.lightmap.map[i] != b.lightmap.map[i]) {
                if (first < 0) first = i;
                n++;
            }
        why = "lightmap of (" + std::to_string(a.x) + "," + std::to_string(a.z) + "): " + std::to_string(n) +
              " voxels differ, first " + std::to_string(first);
        return false;
    }
    if (a.lightmap.highestPoint != b.lightmap.highestPoint) {
        why = "highestPoint of (" + std::to_string(a.x) + "," + std::to_string(a.z) + ")";
        return false;
    }
    if (a.flags.modified != b.flags.modified) {
        why = "flags.modified of (" + std::to_string(a.x) + "," + std::to_string(a.z) + "): reference " +
              std::to_string(a.flags.modified) + " adapter " + std::to_string(b.flags.modified);
        return false;
    }
    if (std::memcmp(a.voxels, b.voxels, sizeof(a.voxels)) != 0 || a.bottom != b.bottom || a.top != b.top) {
        why = "voxels / bottom / top of (" + std::to_string(a.x) + "," + std::to_string(a.z) + ")";
        return false;
    }
    return true;
}

bool same_world(vxo_world* A, vxo_world* B, std::string& why) {
    const auto& ca = A->chunks->getChunks();
    const auto& cb = B->chunks->getChunks();
    for (size_t i = 0; i < ca.size(); i++) {
        if (!ca[i] || !cb[i]) continue;
        if (!same_chunk(*ca[i], *cb[i], why)) return false;
    }
    return true;
}

bool same_mesh(const ChunkMeshData& a, const ChunkMeshData& b, std::string& why) {
    if (a.mesh.vertices.size() != b.mesh.vertices.size() || a.mesh.indices.size() != b.mesh.indices.size()) {
        why = "mesh sizes: reference " + std::to_string(a.mesh.vertices.size()) + " floats / " +
              std::to_string(a.mesh.indices.size()) + " indices, adapter " + std::to_string(b.mesh.vertices.size()) +
              " / " + std::to_string(b.mesh.indices.size());
        return false;
    }
    if (std::memcmp(a.mesh.vertices.data(), b.mesh.vertices.data(), a.mesh.vertices.size() * sizeof(float)) != 0) {
        why = "vertex words";
        return false;
    }
    if (std::memcmp(a.mesh.indices.data(), b.mesh.indices.data(), a.mesh.indices.size() * sizeof(int)) != 0) {
        why = "indices";
        return false;
    }
    if (a.mesh.attrs.size() != b.mesh.attrs.size() ||
        std::memcmp(a.mesh.attrs.data(), b.mesh.attrs.data(), a.mesh.attrs.size() * sizeof(VertexAttribute)) != 0) {
        why = "vertex attributes";
        return false;
    }
    const auto& ea = a.sortingMesh.entries;
    const auto& eb = b.sortingMesh.entries;
    if (ea.size() != eb.size()) {
        why = "sorting entries: " + std::to_string(ea.size()) + " vs " + std::to_string(eb.size());
        return false;
    }
    for (size_t i = 0; i < ea.size(); i++) {
        if (std::memcmp(&ea[i].position, &eb[i].position, sizeof(glm::vec3)) != 0 ||
            ea[i].vertexData.size() != eb[i].vertexData.size() ||
            std::memcmp(ea[i].vertexData.data(), eb[i].vertexData.data(), ea[i].vertexData.size() * sizeof(float)) != 0) {
            why = "sorting entry " + std::to_string(i);
            return false;
        }
    }
    return true;
}

}  // namespace

extern "C" int vxh_adapter_selftest(const vx_content* content, const vx_settings* settings, int32_t ox, int32_t oz,
                                    int32_t w, int32_t d, const vx_voxel* voxels /* w*d chunks, z-major */,
                                    const vx_edit* edits, int32_t n_edits, int32_t device, char* msg, int32_t msg_cap) {
    Fail fail {msg, msg_cap};
    msg[0] = 0;
    std::string why;
    vxo_world* A = vxo_create(content, settings, ox, oz, w, d);
    vxo_world* B = vxo_create(content, settings, ox, oz, w, d);
    if (!A || !B) return fail(1, "vxo_create failed");
    struct Guard {
        vxo_world *a, *b;
        ~Guard() {
            vxo_destroy(a);
            vxo_destroy(b);
        }
    } guard {A, B};
    for (int cz = 0; cz < d; cz++)
        for (int cx = 0; cx < w; cx++) {
            const vx_voxel* v = voxels + ((size_t)cz * w + cx) * VX_CHUNK_VOL;
            if (vxo_put_chunk(A, ox + cx, oz + cz, v, nullptr) || vxo_put_chunk(B, ox + cx, oz + cz, v, nullptr))
                return fail(2, "vxo_put_chunk failed");
        }
    try {
        // ---- 1. flattening: reference Content/ContentGfxCache (built from `content`) -> vx_content again
        FlatContent flat = flattenContent(*B->content, *B->cache);
        if ((int)flat.defs.size() != content->n_defs) return fail(10, "flatten: %zu defs for %d", flat.defs.size(), content->n_defs);
        for (int i = 0; i < content->n_defs; i++) {
            const vx_block_def &p = content->defs[i], &q = flat.defs[i];
            const bool same = !std::memcmp(p.emission, q.emission, 4) && p.draw_group == q.draw_group && p.model == q.model &&
                              p.culling == q.culling && p.rotation_profile == q.rotation_profile &&
                              p.light_passing == q.light_passing && p.sky_light_passing == q.sky_light_passing &&
                              p.shadeless == q.shadeless && p.ambient_occlusion == q.ambient_occlusion &&
                              p.translucent == q.translucent && p.solid == q.solid && p.has_hitbox == q.has_hitbox &&
                              (!p.has_hitbox || (!std::memcmp(p.hitbox_a, q.hitbox_a, 12) && !std::memcmp(p.hitbox_b, q.hitbox_b, 12))) &&
                              (p.model != VX_MODEL_CUSTOM || p.model_mesh_count == q.model_mesh_count);
            if (!same) return fail(11, "flatten: block %d differs", i);
        }
        if (std::memcmp(flat.uv.data(), content->uv_regions, flat.uv.size() * sizeof(float)) != 0) return fail(12, "flatten: uv regions differ");
        if ((int)flat.vertices.size() != content->n_vertices ||
            (content->n_vertices && std::memcmp(flat.vertices.data(), content->vertices, flat.vertices.size() * sizeof(vx_model_vertex)) != 0))
            return fail(13, "flatten: custom model vertices differ");

        GpuContext gpu(*B->content, *B->cache, B->settings, device);
        gpu.bind(*B->chunks);
        std::vector<const Chunk*> all;
        std::vector<Chunk*> allB;
        for (const auto& c : B->chunks->getChunks())
            if (c) {
                all.push_back(c.get());
                allB.push_back(c.get());
            }
        gpu.upload(all);
        GpuLighting lighting(gpu, *B->chunks);

        // ---- 2. lighting
        for (const auto& c : A->chunks->getChunks())
            if (c) Lighting::prebuildSkyLight(*c, *A->content->getIndices());
        lighting.prebuildSkyLight(allB);
        if (!same_world(A, B, why)) return fail(20, "after prebuildSkyLight: %s", why.c_str());
        std::vector<glm::ivec2> inner;
        for (int cz = oz + 1; cz < oz + d - 1; cz++)
            for (int cx = ox + 1; cx < ox + w - 1; cx++) inner.emplace_back(cx, cz);
        // the first half chunk by chunk through the Lighting-shaped methods, the rest as one batch
        const size_t half = inner.size() / 2;
        for (size_t i = 0; i < inner.size(); i++) {
            A->lighting->buildSkyLight(inner[i].x, inner[i].y);  // logic/ChunksController.cpp:118-122
            A->lighting->onChunkLoaded(inner[i].x, inner[i].y, true);
            A->chunks->getChunk(inner[i].x, inner[i].y)->flags.lighted = true;
        }
        for (size_t i = 0; i < half; i++) {
            lighting.buildSkyLight(inner[i].x, inner[i].y);
            lighting.onChunkLoaded(inner[i].x, inner[i].y, true);
        }
        lighting.lightChunks(std::vector<glm::ivec2>(inner.begin() + half, inner.end()), true);
        for (const auto& k : inner) B->chunks->getChunk(k.x, k.y)->flags.lighted = true;
        if (!same_world(A, B, why)) return fail(21, "after buildSkyLight + onChunkLoaded: %s", why.c_str());

        // ---- 3. the ChunksRenderer flow
        std::vector<RendererResult> delivered;
        GpuRendererBatcher batcher(gpu, *B->chunks, B->settings, [&](RendererResult& r) { delivered.push_back(std::move(r)); });
        std::vector<std::shared_ptr<Chunk>> innerB;
        for (const auto& c : B->chunks->getChunks())
            if (c && c->x > ox && c->x < ox + w - 1 && c->z > oz && c->z < oz + d - 1) innerB.push_back(c);
        for (const auto& c : innerB) {
            c->flags.modified = true;
            if (!batcher.request(c)) return fail(30, "first request refused");
            if (c->flags.modified) return fail(31, "request did not clear flags.modified");
            if (batcher.request(c)) return fail(32, "a chunk in work was queued twice");
        }
        if (batcher.inWorkCount() != innerB.size()) return fail(33, "inwork holds %zu of %zu", batcher.inWorkCount(), innerB.size());
        batcher.unload(innerB.back().get());  // the chunk left the world before its turn came
        if (batcher.inWorkCount() != innerB.size() - 1 || batcher.queued() != innerB.size() - 1) return fail(34, "unload kept the request");
        if (!batcher.request(innerB.back())) return fail(35, "request after unload refused");
        batcher.update();
        if (batcher.inWorkCount() != 0 || batcher.queued() != 0) return fail(36, "update left %zu in work", batcher.inWorkCount());
        if (delivered.size() != innerB.size()) return fail(37, "%zu results for %zu requests", delivered.size(), innerB.size());
        auto& gs = A->settings.graphics;
        const size_t capacity = gs.denseRender.get() ? gs.chunkMaxVerticesDense.get() : gs.chunkMaxVertices.get();
        BlocksRenderer worker(capacity, *A->content, *A->cache, A->settings);  // RendererWorker (ChunksRenderer.cpp:22-40)
        for (const RendererResult& r : delivered) {
            const Chunk* ca = A->chunks->getChunk(r.key.x, r.key.y);
            worker.build(ca, A->chunks.get());
            if (worker.isCancelled() != r.cancelled) return fail(38, "cancelled flag of (%d,%d)", r.key.x, r.key.y);
            if (r.cancelled) continue;
            ChunkMeshData want = worker.createMesh();
            if (!same_mesh(want, r.meshData, why)) return fail(39, "mesh of (%d,%d): %s", r.key.x, r.key.y, why.c_str());
        }
        {
            BlocksRenderer important(gs.chunkMaxVertices.get(), *A->content, *A->cache, A->settings);  // :85-88
            const auto& cb = innerB.front();
            important.build(A->chunks->getChunk(cb->x, cb->z), A->chunks.get());
            ChunkMeshData want = important.createMesh();
            ChunkMeshData got = batcher.renderNow(cb);
            if (!same_mesh(want, got, why)) return fail(40, "important mesh of (%d,%d): %s", cb->x, cb->z, why.c_str());
        }
        batcher.request(innerB.front());
        batcher.clear();
        if (batcher.inWorkCount() != 0 || batcher.queued() != 0) return fail(41, "clear left work behind");
        delivered.clear();
        batcher.update();
        if (!delivered.empty()) return fail(42, "a cleared request was built");

        // ---- 4. edits: Chunks::set then onBlockSet, as BlocksController does (logic/BlocksController.cpp:63-104)
        for (const auto& c : A->chunks->getChunks())
            if (c) c->flags.modified = false;
        for (const auto& c : B->chunks->getChunks())
            if (c) c->flags.modified = false;
        gpu.check(vx_chunks_clear_modified(gpu.handle()), "vx_chunks_clear_modified");
        for (int i = 0; i < n_edits; i++) {
            const vx_edit& e = edits[i];
            blocks_agent::set(*A->chunks, e.x, e.y, e.z, e.id, int2blockstate(e.state));
            A->lighting->onBlockSet(e.x, e.y, e.z, e.id);
            if (i % 2 == 0) {
                blocks_agent::set(*B->chunks, e.x, e.y, e.z, e.id, int2blockstate(e.state));
                lighting.onBlockSet(e.x, e.y, e.z, e.id);
            } else {
                lighting.setBlocks(std::vector<vx_edit> {e});  // the device sets the block, the host copy follows
            }
            if (!same_world(A, B, why)) return fail(50, "after edit %d: %s", i, why.c_str());
        }
        A->lighting->clear();
        lighting.clear();
        for (const auto& c : A->chunks->getChunks())
            if (c) c->flags.modified = false;
        for (const auto& c : B->chunks->getChunks())
            if (c) c->flags.modified = false;
        if (!same_world(A, B, why)) return fail(51, "after clear: %s", why.c_str());
    } catch (const std::exception& e) {
        return fail(90, "exception: %s", e.what());
    }
    snprintf(msg, msg_cap, "ok");
    return 0;
}
