// libvxgpu: scheduling.  The decisions of
// ChunksController::loadVisible / buildLights (logic/ChunksController.cpp:63-128)
// and ChunksRenderer::retrieveChunk / getOrRender (graphics/render/
// ChunksRenderer.cpp:128-183), taken for the whole window at once from the chunk
// flags that live on the device, instead of <= 128 items / 4 ms per frame.
#include <algorithm>
#include <vector>

#include "vx_ctx.hpp"

using namespace vx;

namespace {

// lists[0]: chunks to light with buildSkyLight + onChunkLoaded(expand = true);
// lists[1]: chunks whose lights came with them (flags.loadedLights): onChunkLoaded(false) only.
__global__ void k_schedule_light(DevWorld W, int padding, int32_t* lists, int32_t* counts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W.w * W.d) return;
    const int lx = t % W.w, lz = t / W.w;
    if (lx < padding || lz < padding || lx >= W.w - padding || lz >= W.d - padding) return;  // :72-73
    const int p = W.slot[t];
    if (p < 0) return;
    const ChunkMeta m = W.meta[p];
    if (m.lighted) return;  // chunk->flags.loaded && !chunk->flags.lighted (:77)
    for (int oz = -1; oz <= 1; oz++)  // buildLights: MIN_SURROUNDING == 9 (:109-116)
        for (int ox = -1; ox <= 1; ox++)
            if (W.pool_of(m.cx + ox, m.cz + oz) < 0) return;
    const int which = m.loaded_lights ? 1 : 0;
    lists[(size_t)which * W.pool_cap + atomicAdd(&counts[which], 1)] = p;
}

// retrieveChunk skips chunks that are not lighted; getOrRender renders a chunk
// that has no mesh yet or is modified (ChunksRenderer.cpp:128-139,150-158)
__global__ void k_schedule_mesh(DevWorld W, int32_t* list, int32_t* count) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W.w * W.d) return;
    const int p = W.slot[t];
    if (p < 0) return;
    const ChunkMeta m = W.meta[p];
    if (!m.lighted) return;
    if (m.meshed && !m.modified) return;
    list[atomicAdd(count, 1)] = p;
}

// ChunksRenderer::render clears flags.modified (:96) and the mesh joins `meshes`
__global__ void k_after_mesh(DevWorld W, const int32_t* pools, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || pools[t] < 0) return;
    W.meta[pools[t]].meshed = 1;
    W.meta[pools[t]].modified = 0;
}

}  // namespace

extern "C" int vx_window_update(vx_ctx* ctx, int32_t padding, uint64_t capacity, vx_update_stats* out) {
    if (!ctx) return -1;
    if (out) out->n_lit = out->n_meshed = 0;
    if (ctx->pool_cap == 0) return 0;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    cudaStream_t st = ctx->stream;
    const int slots = ctx->w * ctx->d;
    // scratch: the three list slots after the five of vx_upload_list are not used here;
    // a private buffer keeps this independent of the list cache
    if (!ctx->d_sched) {
        VX_CUDA(cudaMalloc((void**)&ctx->d_sched, ((size_t)ctx->pool_cap * 3 + 8) * sizeof(int32_t)));
        ctx->sched_cap = ctx->pool_cap;
    }
    int32_t* d_lists = ctx->d_sched + 8;
    int32_t* d_counts = ctx->d_sched;
    VX_CUDA(cudaMemsetAsync(d_counts, 0, 8 * sizeof(int32_t), st));
    VX_LAUNCH(ctx, KID_SCHED, k_schedule_light<<<(slots + 255) / 256, 256, 0, st>>>(ctx->W, padding, d_lists, d_counts));
    int32_t counts[3] = {0, 0, 0};
    VX_CUDA(cudaMemcpyAsync(counts, d_counts, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    auto fetch_keys = [&](int which, int n, std::vector<vx_chunk_key>& keys) -> int {
        std::vector<int32_t> pools(n);
        if (n)
            VX_CUDA(cudaMemcpy(pools.data(), d_lists + (size_t)which * ctx->pool_cap, n * sizeof(int32_t),
                               cudaMemcpyDeviceToHost));
        keys.clear();
        for (int p : pools) keys.push_back(vx_chunk_key {ctx->h_meta[p].cx, ctx->h_meta[p].cz});
        std::sort(keys.begin(), keys.end(), [](const vx_chunk_key& a, const vx_chunk_key& b) {
            return a.cz != b.cz ? a.cz < b.cz : a.cx < b.cx;
        });
        return 0;
    };
    std::vector<vx_chunk_key> keys;
    if (fetch_keys(0, counts[0], keys)) return -1;
    if (!keys.empty() && vx_light_build(ctx, keys.data(), (int32_t)keys.size(), VX_LIGHT_EXPAND)) return -1;
    if (fetch_keys(1, counts[1], keys)) return -1;
    if (!keys.empty() && vx_light_build(ctx, keys.data(), (int32_t)keys.size(), VX_LIGHT_NO_SKY)) return -1;
    if (out) out->n_lit = counts[0] + counts[1];
    // ---- meshing
    VX_LAUNCH(ctx, KID_SCHED, k_schedule_mesh<<<(slots + 255) / 256, 256, 0, st>>>(
                                  ctx->W, d_lists + 2 * (size_t)ctx->pool_cap, d_counts + 2));
    VX_CUDA(cudaMemcpyAsync(counts + 2, d_counts + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    if (fetch_keys(2, counts[2], keys)) return -1;
    if (vx_mesh_build(ctx, keys.data(), (int32_t)keys.size(), capacity)) return -1;  // also resets the result set
    if (!keys.empty()) {
        VX_LAUNCH(ctx, KID_SCHED, k_after_mesh<<<((int)keys.size() + 255) / 256, 256, 0, st>>>(
                                      ctx->W, ctx->d_mesh_pools, (int)keys.size()));
        VX_CUDA(cudaGetLastError());
        VX_CUDA(cudaStreamSynchronize(st));
    }
    if (out) out->n_meshed = counts[2];
    return 0;
}
