// libvxgpu: context, chunk window and transfers (C ABI of include/vxgpu.h).
// No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>

#include "vx_ctx.hpp"

using namespace vx;

static thread_local std::string g_create_error;

// one pinned staging buffer per context, grown on demand (compressed blobs, generator tables);
// contents are not kept across a growth
uint8_t* vx_host_stage(vx_ctx* ctx, size_t need) {
    if (need <= ctx->h_rle_bytes) return ctx->h_rle;
    if (ctx->h_rle) cudaFreeHost(ctx->h_rle);
    ctx->h_rle = nullptr;
    ctx->h_rle_bytes = 0;
    const size_t bytes = need + need / 4 + 256;
    if (cudaHostAlloc((void**)&ctx->h_rle, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        ctx->fail("pinned staging allocation failed");
        return nullptr;
    }
    ctx->h_rle_bytes = bytes;
    return ctx->h_rle;
}

const char* const kKernelNames[KID_N] = {
    "k_derive", "k_faces_refresh", "k_laycnt", "k_put_meta", "k_prebuild", "k_clear",
    "k_light_begin", "k_ystar", "k_seed", "k_worklist0", "k_solve", "k_light_end", "k_light_reset",
    "k_clear_modified", "k_mesh_lm", "k_mesh_classify", "k_mesh_scan", "k_mesh_offsets", "k_mesh_emit", "k_mesh_final", "k_edit", "k_decode", "k_encode", "k_schedule", "k_rle_decode", "k_rle_encode", "k_terrain_fill", "k_mesh_emit_generic"};

void vx_prof_begin(vx_ctx* ctx, int kid) {
    ctx->launches++;
    if (!ctx->prof_on) return;
    ProfPending p;
    p.kid = kid;
    for (cudaEvent_t* e : {&p.a, &p.b}) {
        if (!ctx->prof_free.empty()) {
            *e = ctx->prof_free.back();
            ctx->prof_free.pop_back();
        } else {
            cudaEventCreate(e);
        }
    }
    cudaEventRecord(p.a, ctx->stream);
    ctx->prof_pending.push_back(p);
}

void vx_prof_end(vx_ctx* ctx) {
    if (!ctx->prof_on) return;
    cudaEventRecord(ctx->prof_pending.back().b, ctx->stream);
}

static void prof_resolve(vx_ctx* ctx) {
    if (ctx->prof_pending.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (const ProfPending& p : ctx->prof_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            ctx->prof_ms[p.kid] += ms;
            ctx->prof_launches[p.kid]++;
        }
        ctx->prof_free.push_back(p.a);
        ctx->prof_free.push_back(p.b);
    }
    ctx->prof_pending.clear();
}

namespace {
__global__ void k_put_meta(vx::DevWorld W, const int32_t* pools, const vx::ChunkMeta* metas, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) W.meta[pools[t]] = metas[t];
}
}  // namespace

namespace {

template <class T>
cudaError_t dev_alloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

void free_pool_arrays(vx_ctx* ctx) {
    DevWorld& W = ctx->W;
    cudaFree(W.vox);
    cudaFree(W.light);
    cudaFree(W.lm);
    cudaFree(W.lp);
    cudaFree(W.emit);
    cudaFree(W.mblk);
    cudaFree(W.msc);
    cudaFree(W.moth);
    cudaFree(W.skycol);
    cudaFree(W.ystar0);
    cudaFree(W.ystar);
    cudaFree(W.laycnt);
    cudaFree(W.faceL);
    cudaFree(W.faceA);
    cudaFree(W.layerA);
    cudaFree(ctx->d_wl);
    cudaFree(ctx->d_wl_dirty);
    ctx->d_wl = ctx->d_wl_dirty = nullptr;
    cudaFree(W.sflag);
    cudaFree(W.nbr);
    cudaFree(W.a0);
    cudaFree(W.meta);
    cudaFree(ctx->d_slot);
    cudaFree(ctx->d_list);
    W.vox = nullptr;
    W.light = nullptr;
    W.lm = nullptr;
    W.lp = W.emit = nullptr;
    W.mblk = W.msc = W.moth = nullptr;
    W.skycol = W.ystar0 = W.ystar = W.laycnt = nullptr;
    W.faceL = nullptr;
    W.faceA = nullptr;
    W.layerA = nullptr;
    W.sflag = nullptr;
    W.nbr = nullptr;
    W.a0 = nullptr;
    W.meta = nullptr;
    ctx->d_slot = nullptr;
    ctx->d_list = nullptr;
}

}  // namespace

int vx_sync_slots(vx_ctx* ctx) {
    VX_CUDA(cudaMemcpyAsync(ctx->d_slot, ctx->h_slot.data(), ctx->h_slot.size() * sizeof(int),
                            cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int vx_upload_list(vx_ctx* ctx, const std::vector<int>& pools, int which) {
    if (pools.empty()) return 0;
    if (pools.size() > (size_t)ctx->pool_cap) {  // a list segment holds pool_cap entries (duplicate keys can exceed it)
        ctx->fail("more keys than window slots in one call");
        return -1;
    }
    // steady-state callers pass the same list again and again: skip the copy
    // (a pageable-memory H2D copy also stalls the host until the stream drains)
    if (ctx->list_cache[which] == pools) return 0;
    VX_CUDA(cudaMemcpyAsync(ctx->d_list + (size_t)which * ctx->pool_cap, pools.data(),
                            pools.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    ctx->list_cache[which] = pools;
    return 0;
}

// Chunks::putChunk bookkeeping for a batch: window slots -> pool records, fresh
// ChunkMeta on the device.  pools[i] is the pool index of keys[i].
int vx_put_slots(vx_ctx* ctx, const vx_chunk_key* keys, const uint8_t* loaded_lights, int n,
                 std::vector<int>& pools) {
    cudaStream_t st = ctx->stream;
    std::vector<ChunkMeta> metas;
    pools.clear();
    pools.reserve(n);
    metas.reserve(n);
    // a failed put puts nothing: every key is checked before any state changes
    if (n > ctx->pool_cap) {
        ctx->fail("vx_chunks_put: more chunks than window slots in one call");
        return -1;
    }
    for (int i = 0; i < n; i++)
        if (ctx->slot_index(keys[i].cx, keys[i].cz) < 0) {
            ctx->fail("vx_chunks_put: chunk outside the window");  // AreaMap2D::set -> false
            return -1;
        }
    for (int i = 0; i < n; i++) {
        int s = ctx->slot_index(keys[i].cx, keys[i].cz);
        int p = ctx->h_slot[s];
        if (p < 0) {
            p = ctx->free_pool.back();
            ctx->free_pool.pop_back();
            ctx->h_slot[s] = p;
            ctx->slot_epoch++;
        }
        ChunkMeta m {};
        m.cx = keys[i].cx;
        m.cz = keys[i].cz;
        m.bottom = 0;
        m.top = CH;  // voxels/Chunk.cpp:16-19
        m.present = 1;
        m.loaded_lights = loaded_lights ? loaded_lights[i] : 0;
        if (ctx->virt_on && !(loaded_lights && loaded_lights[i])) {
            // no lightmap came with the chunk: all zero, and not written out until something reads it
            m.lstate = LS_ZERO;
            m.lvirt = 1;
        }
        ctx->h_meta[p] = m;
        metas.push_back(m);
        pools.push_back(p);
    }
    if (vx_sync_slots(ctx)) return -1;
    // metas: one upload + a scatter kernel instead of one tiny copy per chunk
    if ((size_t)n > ctx->put_meta_cap) {
        cudaFree(ctx->d_put_meta);
        ctx->d_put_meta = nullptr;
        ctx->put_meta_cap = 0;
        VX_CUDA(cudaMalloc((void**)&ctx->d_put_meta, (size_t)n * sizeof(ChunkMeta)));
        ctx->put_meta_cap = n;
    }
    VX_CUDA(cudaMemcpyAsync(ctx->d_put_meta, metas.data(), (size_t)n * sizeof(ChunkMeta),
                            cudaMemcpyHostToDevice, st));
    if (vx_upload_list(ctx, pools, 0)) return -1;
    VX_LAUNCH(ctx, KID_PUTMETA,
              k_put_meta<<<(n + 255) / 256, 256, 0, st>>>(ctx->W, ctx->d_list, ctx->d_put_meta, n));
    VX_CUDA(cudaStreamSynchronize(st));  // `metas` is stack-owned
    return 0;
}

extern "C" {

const char* vx_last_error(const vx_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

vx_ctx* vx_ctx_create(const vx_content* content, const vx_settings* settings, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("vxgpu needs a CUDA device (no CPU fallback): ") +
                         cudaGetErrorString(e);
        return nullptr;
    }
    if (device < 0 || device >= count) {
        g_create_error = "vxgpu: bad device index";
        return nullptr;
    }
    if (!content || content->n_defs <= 0 || content->n_defs >= (int)VX_BLOCK_VOID) {
        g_create_error = "vxgpu: bad content table";
        return nullptr;
    }
    cudaSetDevice(device);
    auto ctx = new vx_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (settings) ctx->settings = *settings;
    ctx->virt_on = !getenv("VXGPU_NO_VIRTUAL") && !getenv("VXGPU_NO_PRISTINE");
    auto bail = [&](const char* what, cudaError_t err) -> vx_ctx* {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        vx_ctx_destroy(ctx);
        return nullptr;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail("cudaStreamCreate", e);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreate(&ctx->ev_l0);
    cudaEventCreate(&ctx->ev_l1);
    cudaEventCreate(&ctx->tm0);
    cudaEventCreate(&ctx->tm1);

    const int n = content->n_defs;
    ctx->defs.assign(content->defs, content->defs + n);
    std::set<uint8_t> groups;  // Content::drawGroups is a std::set<ubyte> (content/Content.hpp:13)
    for (auto& d : ctx->defs) groups.insert(d.draw_group);
    ctx->groups.assign(groups.begin(), groups.end());
    std::vector<DevDef> dd(n);
    std::vector<DevGeom> dg(n);
    for (int i = 0; i < n; i++) {
        const vx_block_def& s = ctx->defs[i];
        uint32_t rank = (uint32_t)(std::lower_bound(ctx->groups.begin(), ctx->groups.end(),
                                                    s.draw_group) - ctx->groups.begin());
        uint32_t f = s.draw_group;
        f |= (uint32_t)(s.model & 7) << DF_MODEL_SHIFT;
        f |= (uint32_t)(s.culling & 3) << DF_CULL_SHIFT;
        f |= (uint32_t)(s.rotation_profile & 3) << DF_ROT_SHIFT;
        if (s.light_passing) f |= DF_LP;
        if (s.sky_light_passing) f |= DF_SLP;
        if (s.shadeless) f |= DF_SHADELESS;
        if (s.ambient_occlusion) f |= DF_AO;
        if (s.translucent) f |= DF_TRANSLUCENT;
        if (s.solid) f |= DF_SOLID;
        if (s.has_hitbox) f |= DF_HITBOX;
        // rt.emissive = any of emission[0..3] (content/ContentBuilder.cpp:30)
        if (s.emission[0] | s.emission[1] | s.emission[2] | s.emission[3]) f |= DF_EMISSIVE;
        f |= rank << DF_RANK_SHIFT;
        dd[i].flags = f;
        dd[i].emission = (s.emission[0] & 15u) | ((s.emission[1] & 15u) << 4) |
                         ((s.emission[2] & 15u) << 8) | ((s.emission[3] & 15u) << 12);
        for (int k = 0; k < 3; k++) {
            dg[i].ha[k] = s.hitbox_a[k];
            dg[i].hb[k] = s.hitbox_b[k];
        }
        dg[i].mesh_begin = s.model_mesh_begin;
        dg[i].mesh_count = s.model == VX_MODEL_CUSTOM ? s.model_mesh_count : 0;
    }
    if ((e = dev_alloc(&ctx->d_defs, n)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = dev_alloc(&ctx->d_geom, n)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = dev_alloc(&ctx->d_uv, (size_t)n * 6)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = dev_alloc(&ctx->d_meshes, content->n_meshes)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = dev_alloc(&ctx->d_verts, content->n_vertices)) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemcpy(ctx->d_defs, dd.data(), n * sizeof(DevDef), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_geom, dg.data(), n * sizeof(DevGeom), cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->d_uv, content->uv_regions, (size_t)n * 6 * sizeof(float4), cudaMemcpyHostToDevice);
    if (content->n_meshes)
        cudaMemcpy(ctx->d_meshes, content->meshes, content->n_meshes * sizeof(vx_model_mesh),
                   cudaMemcpyHostToDevice);
    if (content->n_vertices)
        cudaMemcpy(ctx->d_verts, content->vertices, content->n_vertices * sizeof(vx_model_vertex),
                   cudaMemcpyHostToDevice);
    if ((e = dev_alloc(&ctx->d_counters, 64)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMallocHost((void**)&ctx->h_counters, 64 * sizeof(int32_t))) != cudaSuccess)
        return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_light_verdict, 16 * sizeof(int32_t))) != cudaSuccess)
        return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_mesh_totals, 8 * sizeof(uint64_t))) != cudaSuccess)
        return bail("cudaMallocHost", e);
    if ((e = dev_alloc(&ctx->d_mesh_totals, 8)) != cudaSuccess) return bail("cudaMalloc", e);
    DevWorld& W = ctx->W;
    W.defs = ctx->d_defs;
    W.geom = ctx->d_geom;
    W.uv = ctx->d_uv;
    W.meshes = ctx->d_meshes;
    W.verts = ctx->d_verts;
    W.n_defs = n;
    W.n_ranks = (int)ctx->groups.size();
    W.backlight = ctx->settings.backlight;
    W.dense_render = ctx->settings.dense_render;
    if ((e = cudaGetLastError()) != cudaSuccess) return bail("content upload", e);
    if (vx_mesh_init(ctx) || vx_light_init(ctx)) {
        g_create_error = ctx->err;
        vx_ctx_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

void vx_ctx_destroy(vx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    vx_light_finish(ctx);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_pool_arrays(ctx);
    cudaFree(ctx->d_defs);
    cudaFree(ctx->d_geom);
    cudaFree(ctx->d_uv);
    cudaFree(ctx->d_meshes);
    cudaFree(ctx->d_verts);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_mesh_out);
    cudaFree(ctx->d_emit_hdr);
    cudaFree(ctx->d_mesh_pools);
    cudaFree(ctx->d_units);
    cudaFree(ctx->d_units_t);
    cudaFree(ctx->d_tile_cnt);
    cudaFree(ctx->d_vertices);
    cudaFree(ctx->d_indices);
    cudaFree(ctx->d_tfloats);
    cudaFree(ctx->d_entries);
    cudaFree(ctx->d_edit_scratch);
    cudaFree(ctx->d_edit_buf);
    if (ctx->h_edit_buf) cudaFreeHost(ctx->h_edit_buf);
    cudaFree(ctx->d_codec_stage);
    cudaFree(ctx->d_rle);
    if (ctx->h_rle) cudaFreeHost(ctx->h_rle);
    cudaFree(ctx->d_sched);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_mesh_totals) cudaFreeHost(ctx->h_mesh_totals);
    cudaFree(ctx->d_mesh_totals);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    prof_resolve(ctx);
    for (cudaEvent_t e : ctx->prof_free) cudaEventDestroy(e);
    cudaFree(ctx->d_put_meta);
    if (ctx->tm0) cudaEventDestroy(ctx->tm0);
    if (ctx->tm1) cudaEventDestroy(ctx->tm1);
    if (ctx->ev_l0) cudaEventDestroy(ctx->ev_l0);
    if (ctx->ev_l1) cudaEventDestroy(ctx->ev_l1);
    if (ctx->h_light_verdict) cudaFreeHost(ctx->h_light_verdict);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void* vx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}

void vx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int vx_window_configure(vx_ctx* ctx, int32_t ox, int32_t oz, int32_t w, int32_t d) {
    if (!ctx) return -1;
    if (w <= 0 || d <= 0) {
        ctx->fail("vx_window_configure: empty window");
        return -1;
    }
    cudaSetDevice(ctx->device);
    if (w == ctx->w && d == ctx->d && ctx->pool_cap > 0) {
        // AreaMap2D::translate (util/AreaMap2D.hpp:23-49): chunks that stay in
        // the window keep their (pool-resident) state, the others are dropped.
        std::vector<int> ns((size_t)w * d, -1);
        for (int p = 0; p < ctx->pool_cap; p++) {
            ChunkMeta& m = ctx->h_meta[p];
            if (!m.present) continue;
            int lx = m.cx - ox, lz = m.cz - oz;
            if (lx < 0 || lz < 0 || lx >= w || lz >= d) {
                m.present = 0;
                ctx->free_pool.push_back(p);
                int32_t zero = 0;
                VX_CUDA(cudaMemcpyAsync(&ctx->W.meta[p].present, &zero, sizeof(zero),
                                        cudaMemcpyHostToDevice, ctx->stream));
                continue;
            }
            ns[(size_t)lz * w + lx] = p;
        }
        ctx->h_slot.swap(ns);
        ctx->slot_epoch++;
        ctx->ox = ox;
        ctx->oz = oz;
        ctx->W.ox = ox;
        ctx->W.oz = oz;
        if (vx_sync_slots(ctx)) return -1;
        VX_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    free_pool_arrays(ctx);
    // until every array below exists the window is empty: a failed allocation must not leave a
    // half-configured window that the next call with the same size would take for a valid one
    ctx->w = ctx->d = 0;
    ctx->pool_cap = 0;
    ctx->W.pool_cap = 0;
    const size_t cap = (size_t)w * d;
    for (auto& c : ctx->list_cache) c.clear();
    cudaFree(ctx->d_sched);
    ctx->d_sched = nullptr;
    ctx->h_slot.assign(cap, -1);
    ctx->slot_epoch++;
    ctx->h_meta.assign(cap, ChunkMeta {});
    ctx->free_pool.clear();
    for (int p = (int)cap - 1; p >= 0; p--) ctx->free_pool.push_back(p);
    DevWorld& W = ctx->W;
    VX_CUDA(dev_alloc(&W.vox, cap * CVOL));
    VX_CUDA(dev_alloc(&W.light, cap * CVOL));
    VX_CUDA(dev_alloc(&W.lm, cap * CVOL));
    VX_CUDA(dev_alloc(&W.lp, cap * PLANE_WORDS));
    VX_CUDA(dev_alloc(&W.emit, cap * PLANE_WORDS));
    VX_CUDA(dev_alloc(&W.mblk, cap * PLANE_WORDS));
    VX_CUDA(dev_alloc(&W.msc, cap * PLANE_WORDS));
    VX_CUDA(dev_alloc(&W.moth, cap * PLANE_WORDS));
    VX_CUDA(dev_alloc(&W.skycol, cap * 256));
    VX_CUDA(dev_alloc(&W.ystar0, cap * 256));
    VX_CUDA(dev_alloc(&W.ystar, cap * 256));
    VX_CUDA(dev_alloc(&W.laycnt, cap * 256));
    VX_CUDA(dev_alloc(&W.faceL, cap * 4 * CH * 16));
    VX_CUDA(dev_alloc(&W.faceA, cap * 4 * CH));
    VX_CUDA(dev_alloc(&W.layerA, cap * 2 * NSLAB * 16));
    VX_CUDA(dev_alloc(&ctx->d_wl, cap * 3 * NSLAB));
    VX_CUDA(dev_alloc(&ctx->d_wl_dirty, cap * 3 * NSLAB));
    VX_CUDA(cudaMemsetAsync(ctx->d_wl_dirty, 0, cap * 3 * NSLAB * sizeof(int32_t), ctx->stream));
    VX_CUDA(cudaMemsetAsync(ctx->d_wl, 0, cap * 3 * NSLAB * sizeof(int32_t), ctx->stream));  // ring slots: 0 = empty
    VX_CUDA(dev_alloc(&W.sflag, cap * NSLAB));
    VX_CUDA(dev_alloc(&W.nbr, cap));
    VX_CUDA(dev_alloc(&W.a0, cap * (CVOL / 2)));
    VX_CUDA(dev_alloc(&W.meta, cap));
    VX_CUDA(dev_alloc(&ctx->d_slot, cap));
    VX_CUDA(dev_alloc(&ctx->d_list, cap * 5));
    VX_CUDA(cudaMemsetAsync(W.meta, 0, cap * sizeof(ChunkMeta), ctx->stream));
    ctx->ox = ox;
    ctx->oz = oz;
    ctx->w = w;
    ctx->d = d;
    ctx->pool_cap = (int)cap;
    W.pool_cap = (int)cap;
    W.ox = ox;
    W.oz = oz;
    W.w = w;
    W.d = d;
    W.slot = ctx->d_slot;
    if (vx_sync_slots(ctx)) return -1;
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vx_chunks_put(vx_ctx* ctx, const vx_chunk_in* chunks, int32_t n) {
    if (!ctx) return -1;
    if (ctx->pool_cap == 0) {
        ctx->fail("vx_chunks_put: configure the window first");
        return -1;
    }
    if (n <= 0) return 0;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    std::vector<vx_chunk_key> keys(n);
    for (int i = 0; i < n; i++) keys[i] = vx_chunk_key {chunks[i].cx, chunks[i].cz};
    std::vector<int> pools, zero_pools, given_pools;
    std::vector<uint8_t> has_light(n);
    for (int i = 0; i < n; i++) has_light[i] = chunks[i].light != nullptr;
    if (vx_put_slots(ctx, keys.data(), has_light.data(), n, pools)) return -1;
    for (int i = 0; i < n; i++) {
        const vx_chunk_in& c = chunks[i];
        if (c.light) {
            VX_CUDA(cudaMemcpyAsync(ctx->W.light + (size_t)pools[i] * CVOL, c.light,
                                    sizeof(vx_light) * CVOL, cudaMemcpyHostToDevice, st));
            given_pools.push_back(pools[i]);
        } else {
            zero_pools.push_back(pools[i]);
        }
    }
    // voxel uploads: consecutive chunks whose host buffers and pool slots are both
    // contiguous go out as one copy (a world kept in one allocation uploads in a
    // handful of transfers instead of one per chunk)
    for (int i = 0; i < n;) {
        int j = i + 1;
        while (j < n && pools[j] == pools[j - 1] + 1 &&
               chunks[j].voxels == chunks[j - 1].voxels + CVOL)
            j++;
        VX_CUDA(cudaMemcpyAsync(ctx->W.vox + (size_t)pools[i] * CVOL, chunks[i].voxels,
                                sizeof(vx_voxel) * CVOL * (size_t)(j - i), cudaMemcpyHostToDevice, st));
        i = j;
    }
    if (!ctx->virt_on && !zero_pools.empty() && vx_zero_light(ctx, zero_pools)) return -1;
    if (!given_pools.empty() && vx_derive_chunks(ctx, given_pools, true)) return -1;
    if (!zero_pools.empty() && vx_derive_chunks(ctx, zero_pools, false)) return -1;
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int vx_chunks_remove(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < n; i++) {
        int s = ctx->slot_index(keys[i].cx, keys[i].cz);
        if (s < 0 || ctx->h_slot[s] < 0) continue;
        int p = ctx->h_slot[s];
        ctx->h_slot[s] = -1;
        ctx->slot_epoch++;
        ctx->h_meta[p].present = 0;
        ctx->free_pool.push_back(p);
        int32_t zero = 0;
        VX_CUDA(cudaMemcpyAsync(&ctx->W.meta[p].present, &zero, sizeof(zero),
                                cudaMemcpyHostToDevice, ctx->stream));
    }
    if (vx_sync_slots(ctx)) return -1;
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vx_chunk_get_info(vx_ctx* ctx, int32_t cx, int32_t cz, vx_chunk_info* out) {
    if (!ctx || !out) return -1;
    std::memset(out, 0, sizeof(*out));
    int p = ctx->pool_of(cx, cz);
    if (p < 0) return 0;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    ChunkMeta m;
    VX_CUDA(cudaMemcpyAsync(&m, &ctx->W.meta[p], sizeof(m), cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    out->present = 1;
    out->bottom = m.bottom;
    out->top = m.top;
    out->highest_point = m.highest;
    out->modified = m.modified;
    out->lighted = m.lighted;
    return 0;
}

int vx_chunk_get_voxels(vx_ctx* ctx, int32_t cx, int32_t cz, vx_voxel* out) {
    if (!ctx) return -1;
    int p = ctx->pool_of(cx, cz);
    if (p < 0) {
        ctx->fail("vx_chunk_get_voxels: no such chunk");
        return -1;
    }
    cudaSetDevice(ctx->device);
    VX_CUDA(cudaMemcpyAsync(out, ctx->W.vox + (size_t)p * CVOL, sizeof(vx_voxel) * CVOL,
                            cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vx_light_get(vx_ctx* ctx, int32_t cx, int32_t cz, vx_light* out) {
    if (!ctx) return -1;
    int p = ctx->pool_of(cx, cz);
    if (p < 0) {
        ctx->fail("vx_light_get: no such chunk");
        return -1;
    }
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    if (vx_materialize(ctx, std::vector<int> {p})) return -1;
    VX_CUDA(cudaMemcpyAsync(out, ctx->W.light + (size_t)p * CVOL, sizeof(vx_light) * CVOL,
                            cudaMemcpyDeviceToHost, ctx->stream));
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vx_light_get_batch(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, vx_light* out) {
    if (!ctx || !out) return -1;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    {
        std::vector<int> pools(n);
        for (int i = 0; i < n; i++) pools[i] = ctx->pool_of(keys[i].cx, keys[i].cz);
        if (vx_materialize(ctx, pools)) return -1;
    }
    for (int i = 0; i < n;) {
        const int p = ctx->pool_of(keys[i].cx, keys[i].cz);
        vx_light* dst = out + (size_t)i * CVOL;
        if (p < 0) {
            std::memset(dst, 0, sizeof(vx_light) * CVOL);
            i++;
            continue;
        }
        int j = i + 1;  // runs of consecutive pool slots leave as one copy
        while (j < n && ctx->pool_of(keys[j].cx, keys[j].cz) == p + (j - i)) j++;
        VX_CUDA(cudaMemcpyAsync(dst, ctx->W.light + (size_t)p * CVOL,
                                sizeof(vx_light) * CVOL * (size_t)(j - i), cudaMemcpyDeviceToHost,
                                ctx->stream));
        i = j;
    }
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vx_timer_begin(vx_ctx* ctx) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    VX_CUDA(cudaEventRecord(ctx->tm0, ctx->stream));
    return 0;
}

int vx_timer_end(vx_ctx* ctx, float* ms) {
    if (!ctx || !ms) return -1;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    VX_CUDA(cudaEventRecord(ctx->tm1, ctx->stream));
    VX_CUDA(cudaEventSynchronize(ctx->tm1));
    VX_CUDA(cudaEventElapsedTime(ms, ctx->tm0, ctx->tm1));
    return 0;
}

int vx_profile_enable(vx_ctx* ctx, int32_t on) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    prof_resolve(ctx);
    if (on) {
        for (int k = 0; k < KID_N; k++) {
            ctx->prof_ms[k] = 0;
            ctx->prof_launches[k] = 0;
        }
    }
    ctx->prof_on = on != 0;
    return 0;
}

int vx_profile_get(vx_ctx* ctx, int32_t i, char* name, int32_t name_cap, float* ms, int64_t* launches) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    prof_resolve(ctx);
    if (i >= 0 && i < KID_N) {
        if (name && name_cap > 0) {
            std::strncpy(name, kKernelNames[i], name_cap - 1);
            name[name_cap - 1] = 0;
        }
        if (ms) *ms = ctx->prof_ms[i];
        if (launches) *launches = ctx->prof_launches[i];
    }
    return KID_N;
}

int vx_last_timing(vx_ctx* ctx, float* light_ms, float* mesh_ms, int64_t* launches) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    if (light_ms) *light_ms = ctx->last_light_ms;
    if (mesh_ms) *mesh_ms = ctx->last_mesh_ms;
    if (launches) *launches = ctx->launches;
    return 0;
}

}  // extern "C"
