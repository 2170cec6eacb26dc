// Host-side context of libvxgpu (internal; the public surface is include/vxgpu.h).
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "vx_common.cuh"

#define VX_CUDA(call)                                                          \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            ctx->fail(std::string(#call) + ": " + cudaGetErrorString(e_));     \
            return -1;                                                         \
        }                                                                      \
    } while (0)

// kernel classes of the per-kernel profile (vx_profile_get)
enum KernelId {
    KID_DERIVE, KID_FACES, KID_LAYCNT, KID_PUTMETA, KID_PREBUILD, KID_CLEAR, KID_BEGIN, KID_YSTAR,
    KID_SEED, KID_WL0, KID_SOLVE, KID_END, KID_RESET, KID_CLEARMOD, KID_LM, KID_CLASSIFY, KID_SCAN, KID_OFFSETS, KID_EMIT, KID_FINAL, KID_EDIT, KID_DECODE, KID_ENCODE, KID_SCHED, KID_RLE_DECODE, KID_RLE_ENCODE, KID_TERRAIN, KID_EMIT_GENERIC,
    KID_N
};
extern const char* const kKernelNames[KID_N];

struct ProfPending {
    int kid;
    cudaEvent_t a, b;
};

struct MeshChunkOut {  // device-written summary of one meshed chunk
    uint32_t tot_floats, tot_idx;      // un-truncated opaque totals
    uint32_t kept_floats, kept_idx;    // what survives the capacity cut-off
    uint32_t t_floats, t_entries;      // translucent totals
    uint32_t v_off, i_off;             // arena offsets (floats / ints)
    uint32_t tf_off, te_off;           // translucent arena offsets
    uint32_t aabb_min[3], aabb_max[3]; // ordered-int encoded floats
    uint32_t n_entries_final;          // after the thin-AABB merge rule
    int32_t cancelled;
    int32_t pool;
    int32_t pad;
};

struct vx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // content
    std::vector<vx_block_def> defs;
    std::vector<uint8_t> groups;  // sorted unique draw groups (Content::drawGroups)
    vx::DevDef* d_defs = nullptr;
    vx::DevGeom* d_geom = nullptr;
    float4* d_uv = nullptr;
    vx_model_mesh* d_meshes = nullptr;
    vx_model_vertex* d_verts = nullptr;
    vx_settings settings {1, 1};

    // window + pool
    int ox = 0, oz = 0, w = 0, d = 0;
    int pool_cap = 0;
    std::vector<int> h_slot;           // [w*d] -> pool or -1
    std::vector<int> free_pool;        // free pool indices
    std::vector<vx::ChunkMeta> h_meta; // host shadow (cx, cz, present only)
    int32_t* d_slot = nullptr;
    vx::DevWorld W {};                 // device pointers, refreshed by sync_world()

    // scratch
    int32_t* d_list = nullptr;   // chunk pool-index lists (5 * pool_cap)
    std::vector<int> list_cache[5];  // what each device list currently holds
    vx::ChunkMeta* d_put_meta = nullptr;
    size_t put_meta_cap = 0;
    int32_t* d_wl = nullptr;        // solver work lists [3][pool_cap * NSLAB]
    int32_t* d_wl_dirty = nullptr;  // matching queued flags
    int n_sm = 148;
    int32_t* d_counters = nullptr;
    int32_t* h_counters = nullptr;  // pinned
    int coop_grid_async = 0;        // CTAs of the solver's cooperative launch (vx_light.cu)
    bool virt_on = true;            // Lighting::clear / prebuildSkyLight of untouched lightmaps are lazy (ChunkMeta::lvirt)
    void* h_stage = nullptr;        // pinned staging
    size_t h_stage_bytes = 0;

    // mesh outputs (resident until the next vx_mesh_build)
    std::vector<MeshChunkOut> mesh_out;
    std::vector<vx_chunk_key> mesh_keys;
    MeshChunkOut* d_mesh_out = nullptr;
    size_t mesh_out_cap = 0;
    uint4* d_emit_hdr = nullptr;  // 4 x uint4 per chunk of the batch (EmitHdr, vx_mesh.cu)
    size_t emit_hdr_cap = 0;
    int32_t* d_mesh_pools = nullptr;
    size_t mesh_pools_cap = 0;
    uint32_t* d_tile_cnt = nullptr;  // [n][16 tiles][ranks][4]
    size_t tile_cnt_cap = 0;
    uint32_t* d_units = nullptr;      // unit list, 4 words per unit
    size_t units_words = 0;
    size_t units_cap = 0;             // in units
    uint32_t* d_units_t = nullptr;    // translucent cube faces: a list of their own
    size_t units_t_words = 0;
    size_t units_t_cap = 0;
    bool split_transl = false;        // the next build lists translucent cube faces apart (vx_mesh_build's policy)
    unsigned long long slow_units_hint = 0;  // whole-voxel units of the last mesh build + a margin (0: not known yet): k_mesh_emit_generic's grid
    std::vector<int> mesh_pools_cache;
    // vx_mesh_build's host-side preparation, reused while the keys and the slot map stay the same
    uint64_t slot_epoch = 1;            // bumped whenever a window slot is assigned, dropped or moved
    uint64_t mesh_prep_epoch = 0;
    std::vector<vx_chunk_key> mesh_prep_keys;
    std::vector<int> mesh_prep_pools, mesh_prep_lm_pools;
    unsigned long long* d_mesh_totals = nullptr;  // {vertex floats, indices, entry floats, entries}
    uint64_t* h_mesh_totals = nullptr;            // pinned
    float* d_vertices = nullptr;
    size_t vertices_cap = 0;
    int32_t* d_indices = nullptr;
    size_t indices_cap = 0;
    float* d_tfloats = nullptr;
    size_t tfloats_cap = 0;
    vx_sort_entry* d_entries = nullptr;
    size_t entries_cap = 0;
    uint64_t mesh_capacity = 0;
    uint64_t arena_sizes[4] = {0, 0, 0, 0};  // vertices, indices, entries, entry floats

    // scheduler scratch (vx_sched.cu): 8 counters + 3 pool lists
    int32_t* d_sched = nullptr;
    int sched_cap = 0;

    // wire-format staging (vx_codec.cu)
    uint8_t* d_codec_stage = nullptr;
    size_t codec_stage_bytes = 0;
    // compressed blobs + stream descriptors (vx_rle.cu): device buffer and its pinned mirror
    uint8_t* d_rle = nullptr;
    size_t d_rle_bytes = 0;
    uint8_t* h_rle = nullptr;
    size_t h_rle_bytes = 0;

    // edits
    uint32_t* d_edit_scratch = nullptr;
    int32_t* h_edit_buf = nullptr;    // pinned: what one vx_light_edit call uploads (counters, flags, predecessors,
    int32_t* d_edit_buf = nullptr;    // touched chunks, the edits), and its copy on the device
    size_t edit_buf_cap = 0;          // words
    std::vector<uint32_t> edit_seen;  // per pool slot: the last vx_light_edit call (edit_epoch) that touched it
    uint32_t edit_epoch = 0;
    int last_edit_global = 0;         // edits of the last batch that ran on the global lightmaps (left their box)
    int edit_resident = 0;            // k_edit CTAs that can be resident at once (cooperative launch)
    float last_edit_ms = 0;
    int last_edit_waves = 0;
    int last_edit_launches = 0;

    // timing
    bool prof_on = false;
    std::vector<ProfPending> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    float prof_ms[KID_N] = {};
    int64_t prof_launches[KID_N] = {};
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // vx_light_build returns with its kernels queued: what the solver's queue said (and the build's device time)
    // is looked at by vx_light_finish — the next call that synchronises anyway, or that needs the answer
    cudaEvent_t ev_l0 = nullptr, ev_l1 = nullptr;
    int32_t* h_light_verdict = nullptr;  // pinned, 16 words
    bool light_pending = false;
    int light_pending_stride = 0;
    float last_light_ms = 0, last_mesh_ms = 0;
    int64_t launches = 0;

    void fail(const std::string& m) { err = m; }
    int slot_index(int cx, int cz) const {
        int lx = cx - ox, lz = cz - oz;
        if (lx < 0 || lz < 0 || lx >= w || lz >= d) return -1;
        return lz * w + lx;
    }
    int pool_of(int cx, int cz) const {
        int s = slot_index(cx, cz);
        return s < 0 ? -1 : h_slot[s];
    }
};

// implemented in vx_ctx.cu
void vx_prof_begin(vx_ctx* ctx, int kid);
void vx_prof_end(vx_ctx* ctx);
// every kernel launch goes through this: counts it and, when the profile is
// enabled, brackets it with events on the ctx stream
#define VX_LAUNCH(ctx, kid, ...)   \
    do {                           \
        vx_prof_begin(ctx, kid);   \
        __VA_ARGS__;               \
        vx_prof_end(ctx);          \
    } while (0)
int vx_sync_slots(vx_ctx* ctx);
int vx_upload_list(vx_ctx* ctx, const std::vector<int>& pools, int which);
int vx_put_slots(vx_ctx* ctx, const vx_chunk_key* keys, const uint8_t* loaded_lights, int n,
                 std::vector<int>& pools);
// implemented in vx_mesh.cu
int vx_mesh_init(vx_ctx* ctx);
// implemented in vx_light.cu
int vx_light_init(vx_ctx* ctx);
int vx_light_finish(vx_ctx* ctx);  // waits for a queued light build and reports its verdict (0: fine / nothing pending)
int vx_derive_chunks(vx_ctx* ctx, const std::vector<int>& pools, bool light_given);
int vx_zero_light(vx_ctx* ctx, const std::vector<int>& pools);
// writes out the lightmaps among `pools` (entries < 0 are skipped) that only exist as a state (ChunkMeta::lvirt);
// every path that reads a lightmap outside vx_light_build calls this first
int vx_materialize(vx_ctx* ctx, const std::vector<int>& pools);
uint8_t* vx_host_stage(vx_ctx* ctx, size_t need);  // pinned, ctx->h_rle; nullptr + error on failure
// vx_codec.cu: wire-format staging area <-> window slots
int vx_put_staged(vx_ctx* ctx, const vx_chunk_key* keys, const uint8_t* has_light, int32_t n);
int vx_encode_staged(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, bool voxels, bool lights);
