/*
 * vxgpu.h — C ABI of the B200-native chunk light + mesh path.
 *
 * Drop-in boundary for VoxelCore's (MihailRis/VoxelEngine-Cpp 0.28) chunk
 * build hot path: Lighting/LightSolver flood fill followed by
 * BlocksRenderer::build.  The reference has no FFI for this path;
 * the seam is a set of C++ class methods.  Every entry point below names the
 * reference method(s) it replaces (paths relative to the reference's src/).
 *
 * Conventions: plain pointers and sizes only; `int` status (0 = ok, <0 = error,
 * text via vx_last_error); the caller owns host chunk memory, the library owns
 * device mirrors, staging and output arenas; calls on one vx_ctx are
 * serialised by the caller (one ctx per host thread / CUDA stream).
 * There is NO CPU fallback: every entry point fails loudly when no CUDA device
 * is usable.
 */
#ifndef VXGPU_H
#define VXGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- data layouts (identical to the reference's, so buffers memcpy) ------ */

#define VX_CHUNK_W 16   /* constants.hpp:32 */
#define VX_CHUNK_H 256  /* constants.hpp:33 */
#define VX_CHUNK_D 16   /* constants.hpp:34 */
#define VX_CHUNK_VOL (VX_CHUNK_W * VX_CHUNK_H * VX_CHUNK_D) /* constants.hpp:45 */
#define VX_BLOCK_VOID 0xFFFFu /* constants.hpp:48 */
#define VX_VERTEX_SIZE 6      /* graphics/render/commons.hpp:14 (floats) */

/* voxel = {u16 id; u16 state} (voxels/voxel.hpp:12-40); state =
 * rotation:3 | segment:3 | reserved:2 | userbits:8.  On little-endian hosts a
 * reference `voxel[65536]` IS a vx_voxel[65536].  Index (y*16+z)*16+x
 * (constants.hpp:55-57). */
typedef uint32_t vx_voxel;
#define VX_VOXEL(id, state) ((vx_voxel)((uint32_t)(id) | ((uint32_t)(state) << 16)))
#define VX_VOXEL_ID(v) ((uint16_t)((v) & 0xFFFFu))
#define VX_VOXEL_STATE(v) ((uint16_t)((v) >> 16))
#define VX_STATE_ROTATION(s) ((s) & 7u)
#define VX_STATE_SEGMENT(s) (((s) >> 3) & 7u)

/* light_t = R | G<<4 | B<<8 | S<<12 (lighting/Lightmap.hpp:82-88). */
typedef uint16_t vx_light;

enum vx_block_model { /* voxels/Block.hpp:86-97 */
    VX_MODEL_NONE = 0,
    VX_MODEL_BLOCK = 1,
    VX_MODEL_XSPRITE = 2,
    VX_MODEL_AABB = 3,
    VX_MODEL_CUSTOM = 4
};

enum vx_culling_mode { /* voxels/Block.hpp:107-111 */
    VX_CULLING_DEFAULT = 0,
    VX_CULLING_OPTIONAL = 1,
    VX_CULLING_DISABLED = 2
};

enum vx_rotation_profile { /* voxels/Block.cpp:57-92 (BlockRotProfile) */
    VX_ROT_NONE = 0, /* not rotatable */
    VX_ROT_PIPE = 1,
    VX_ROT_PANE = 2
};

/* The Block fields the path reads (voxels/Block.hpp:133-263), flattened. */
typedef struct vx_block_def {
    uint8_t emission[4];       /* Block::emission R,G,B,S */
    uint8_t draw_group;        /* Block::drawGroup */
    uint8_t model;             /* vx_block_model */
    uint8_t culling;           /* vx_culling_mode */
    uint8_t rotation_profile;  /* vx_rotation_profile; != NONE <=> Block::rotatable */
    uint8_t light_passing;     /* Block::lightPassing */
    uint8_t sky_light_passing; /* Block::skyLightPassing */
    uint8_t shadeless;         /* Block::shadeless */
    uint8_t ambient_occlusion; /* Block::ambientOcclusion */
    uint8_t translucent;       /* Block::translucent */
    uint8_t solid;             /* Block::rt.solid (content/ContentBuilder.cpp:31-37) */
    uint8_t has_hitbox;        /* !Block::hitboxes.empty() */
    uint8_t reserved0;
    float hitbox_a[3];         /* union of Block::hitboxes, min corner
                                  (graphics/render/BlocksRenderer.cpp:233-237) */
    float hitbox_b[3];         /* union of Block::hitboxes, max corner */
    int32_t model_mesh_begin;  /* first vx_model_mesh of the custom model */
    int32_t model_mesh_count;  /* model::Model::meshes.size() */
} vx_block_def;

/* model::Vertex (graphics/commons/Model.hpp:10-14) */
typedef struct vx_model_vertex {
    float coord[3];
    float uv[2];
    float normal[3];
} vx_model_vertex;

/* model::Mesh (graphics/commons/Model.hpp:16-19): a triangle list */
typedef struct vx_model_mesh {
    int32_t vertex_begin;
    int32_t vertex_count; /* multiple of 3 */
} vx_model_mesh;

/* Content + ContentGfxCache as read by the path. */
typedef struct vx_content {
    const vx_block_def* defs; /* indexed by block id (Block::rt.id) */
    int32_t n_defs;
    const float* uv_regions;  /* n_defs*6*4: UVRegion{u1,v1,u2,v2} per
                                 id*6+side, sides -x,+x,-y,+y,-z,+z
                                 (frontend/ContentGfxCache.hpp:35-37) */
    const vx_model_mesh* meshes;
    int32_t n_meshes;
    const vx_model_vertex* vertices;
    int32_t n_vertices;
} vx_content;

/* EngineSettings keys that change results (settings.hpp:64-75). */
typedef struct vx_settings {
    int32_t backlight;    /* graphics.backlight, default 1 */
    int32_t dense_render; /* graphics.dense-render, default 1 */
} vx_settings;

typedef struct vx_chunk_key {
    int32_t cx, cz;
} vx_chunk_key;

/* One chunk handed to the library (voxels/Chunk.hpp:25-40). */
typedef struct vx_chunk_in {
    int32_t cx, cz;
    const vx_voxel* voxels; /* VX_CHUNK_VOL */
    const vx_light* light;  /* VX_CHUNK_VOL or NULL (= all zero) */
} vx_chunk_in;

typedef struct vx_chunk_info {
    int32_t present;
    int32_t bottom, top;    /* Chunk::bottom/top (voxels/Chunk.cpp:21-34) */
    int32_t highest_point;  /* Lightmap::highestPoint */
    int32_t modified;       /* Chunk::flags.modified as set by the light path */
    int32_t lighted;        /* Chunk::flags.lighted */
} vx_chunk_info;

/* One block edit = blocks_agent::set + Lighting::onBlockSet
 * (voxels/blocks_agent.cpp:10-77, lighting/Lighting.cpp:163-215). */
typedef struct vx_edit {
    int32_t x, y, z;
    uint16_t id;
    uint16_t state;
} vx_edit;

/* SortingMeshEntry (graphics/render/commons.hpp:18-27), flattened. */
typedef struct vx_sort_entry {
    float position[3];
    uint32_t first_float; /* offset into the chunk's translucent float array */
    uint32_t n_floats;
} vx_sort_entry;

/* Sizes of one built ChunkMeshData (graphics/render/commons.hpp:33-36). */
typedef struct vx_mesh_info {
    int32_t cx, cz;
    int32_t cancelled;        /* BlocksRenderer::isCancelled() */
    uint32_t n_vertex_floats; /* MeshData::vertices.size() */
    uint32_t n_indices;       /* MeshData::indices.size() */
    uint32_t n_entries;       /* SortingMeshData::entries.size() */
    uint32_t n_entry_floats;  /* sum of entries[i].vertexData.size() */
} vx_mesh_info;

typedef struct vx_ctx vx_ctx;

/* ---- lifecycle ----------------------------------------------------------- */

/* Replaces the constructors Lighting(content, chunks) (lighting/Lighting.cpp:17)
 * and BlocksRenderer(capacity, content, cache, settings)
 * (graphics/render/BlocksRenderer.cpp:14-33): uploads the block LUT, UV
 * regions and custom models once.  NULL on failure; vx_last_error(NULL). */
vx_ctx* vx_ctx_create(const vx_content* content, const vx_settings* settings, int device);
void vx_ctx_destroy(vx_ctx* ctx);
const char* vx_last_error(const vx_ctx* ctx);

/* Pinned host memory for chunk uploads / mesh downloads (optional; any host
 * pointer is accepted by the calls below, pinned ones copy faster). */
void* vx_host_alloc(size_t bytes);
void vx_host_free(void* p);

/* ---- chunk window (Chunks / AreaMap2D) ----------------------------------- */

/* Chunks::Chunks + setCenter/resize (voxels/Chunks.cpp:24-39,315-321): the
 * window covers chunk coords [ox,ox+w) x [oz,oz+d).  Chunks that stay inside
 * keep their device state; others are dropped (AreaMap2D.hpp:23-49). */
int vx_window_configure(vx_ctx* ctx, int32_t ox, int32_t oz, int32_t w, int32_t d);

/* Chunks::putChunk (voxels/Chunks.cpp:323-331) for a batch; also runs
 * Chunk::updateHeights (voxels/Chunk.cpp:21-34) on the device. */
int vx_chunks_put(vx_ctx* ctx, const vx_chunk_in* chunks, int32_t n);
int vx_chunks_remove(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n);
int vx_chunk_get_info(vx_ctx* ctx, int32_t cx, int32_t cz, vx_chunk_info* out);
int vx_chunk_get_voxels(vx_ctx* ctx, int32_t cx, int32_t cz, vx_voxel* out);
/* Clears Chunk::flags.modified for every chunk in the window. */
int vx_chunks_clear_modified(vx_ctx* ctx);

/* ---- chunk wire format (world files / lights cache) ----------------------- */

#define VX_CHUNK_DATA_LEN (VX_CHUNK_VOL * 4)    /* constants.hpp: CHUNK_DATA_LEN */
#define VX_LIGHTMAP_DATA_LEN (VX_CHUNK_VOL / 2) /* lighting/Lightmap.hpp: LIGHTMAP_DATA_LEN */

/* A chunk as the region files hold it: Chunk::encode layout (voxels/Chunk.cpp:
 * 70-86) = u16 ids[VOL] then u16 states[VOL], little endian; lights (optional)
 * = Lightmap::encode layout (lighting/Lightmap.cpp:18-24): the sky nibbles of
 * two voxels per byte. */
typedef struct vx_chunk_encoded {
    int32_t cx, cz;
    const uint8_t* voxels; /* VX_CHUNK_DATA_LEN */
    const uint8_t* lights; /* VX_LIGHTMAP_DATA_LEN or NULL */
} vx_chunk_encoded;

/* Chunks::putChunk of chunks still in wire format: Chunk::decode
 * (voxels/Chunk.cpp:88-97) and Lightmap::decode (lighting/Lightmap.cpp:26-34)
 * run on the device, so chunks stream from disk to HBM without a CPU reformat
 * (GlobalChunks::create, voxels/GlobalChunks.cpp:91-132). */
int vx_chunks_put_encoded(vx_ctx* ctx, const vx_chunk_encoded* chunks, int32_t n);
/* Chunk::encode / Lightmap::encode for a batch: voxels_out receives n blocks of
 * VX_CHUNK_DATA_LEN bytes, lights_out n blocks of VX_LIGHTMAP_DATA_LEN (either
 * may be NULL); missing chunks are zero-filled. */
int vx_chunks_encode(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, uint8_t* voxels_out,
                     uint8_t* lights_out);

/* A chunk as a region file holds it (world/files/WorldRegions.cpp:60-72): the
 * voxels layer = extrle::encode16 of the Chunk::encode bytes, the lights layer
 * (optional) = extrle::encode of the Lightmap::encode bytes (coders/rle.cpp:
 * 108-205, coders/compression.cpp:58-78). */
typedef struct vx_chunk_compressed {
    int32_t cx, cz;
    const uint8_t* voxels;
    uint32_t voxels_len;
    const uint8_t* lights; /* NULL: no cached lights */
    uint32_t lights_len;
} vx_chunk_compressed;

/* WorldRegions::getVoxels / getLights (world/files/WorldRegions.cpp:215-242) +
 * Chunks::putChunk with everything after the file read on the device:
 * compression::decompress (coders/compression.cpp:80-110), Chunk::decode and
 * Lightmap::decode.  A blob that does not decompress to exactly
 * VX_CHUNK_DATA_LEN / VX_LIGHTMAP_DATA_LEN bytes fails the call ("expected
 * decompressed size ... got ...", compression.cpp:93-98) and puts nothing. */
int vx_chunks_put_compressed(vx_ctx* ctx, const vx_chunk_compressed* chunks, int32_t n);

#define VX_LAYER_VOXELS 0 /* REGION_LAYER_VOXELS: EXTRLE16 */
#define VX_LAYER_LIGHTS 1 /* REGION_LAYER_LIGHTS: EXTRLE8 */
/* WorldRegions::put of one layer for a batch (world/files/WorldRegions.cpp:
 * 105-120): Chunk::encode / Lightmap::encode + compression::compress on the
 * device.  Blob i is out[offsets[i] .. offsets[i + 1]); offsets has n + 1
 * entries and is filled even when out_capacity is too small (the call then
 * fails and offsets[n] is the size needed). */
int vx_chunks_compress(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, int32_t layer, uint8_t* out,
                       uint64_t out_capacity, uint64_t* offsets);

/* ---- terrain fill (world generator: land + plants) ------------------------ */

/* BlocksLayer (world/generator/GeneratorDef.hpp:24-39) */
typedef struct vx_gen_layer {
    uint16_t block;          /* rt.id */
    int16_t height;          /* -1: the resizeable layer */
    uint8_t below_sea_level; /* 0: skipped under the sea level, its height extends the next layer */
    uint8_t pad[3];
} vx_gen_layer;

/* WeightedEntry of Biome::plants (GeneratorDef.hpp:57-68) */
typedef struct vx_gen_plant {
    uint16_t block; /* rt.id */
    uint16_t pad;
    float weight;
} vx_gen_plant;

/* The land / plants fields of Biome (GeneratorDef.hpp:108-123); begin/count index
 * vx_gen_def::layers and ::plants.  Plants are in BiomeElementList::entries order
 * (the loader sorts them by weight, descending). */
typedef struct vx_gen_biome {
    int32_t ground_begin, ground_count;      /* groundLayers.layers, top first */
    int32_t sea_begin, sea_count;            /* seaLayers.layers */
    uint32_t ground_last_layers_height;      /* BlocksLayers::lastLayersHeight */
    uint32_t sea_last_layers_height;
    int32_t plants_begin, plants_count;
    float plants_chance;                     /* BiomeElementList::chance */
} vx_gen_biome;

typedef struct vx_gen_def {
    const vx_gen_biome* biomes;
    int32_t n_biomes;
    const vx_gen_layer* layers;
    int32_t n_layers;
    const vx_gen_plant* plants;
    int32_t n_plants;
    uint32_t sea_level; /* GeneratorDef::seaLevel */
} vx_gen_def;

/* One chunk prototype at ChunkPrototypeLevel::HEIGHTMAP (WorldGenerator.hpp:24-38):
 * the cropped 16x16 heightmap (values in [0, 1), index z * 16 + x) and the biome
 * chosen for every column. */
typedef struct vx_gen_chunk {
    int32_t cx, cz;
    const float* heights;  /* 256 */
    const uint8_t* biomes; /* 256 indices into vx_gen_def::biomes */
} vx_gen_chunk;

#define VX_GEN_MAX_LAYERS 8 /* layers per BlocksLayers this library accepts */

/* WorldGenerator::generate (world/generator/WorldGenerator.cpp:419-461) for a
 * batch, without structure placements: generate_pole for the sea and ground
 * layers of every column (:87-116), generatePlants (:362-401, one
 * util::PseudoRandom stream per chunk seeded with (cx, cz)), struct_air -> air;
 * then Chunks::putChunk + Chunk::updateHeights with a zero lightmap, as
 * ChunksController::createChunk does (logic/ChunksController.cpp:130-152).  The
 * voxels are produced in HBM; only the 1 KiB + 256 B per chunk above cross PCIe. */
int vx_terrain_fill(vx_ctx* ctx, const vx_gen_def* def, const vx_gen_chunk* chunks, int32_t n);

/* ---- lighting ------------------------------------------------------------ */

/* Lighting::prebuildSkyLight (lighting/Lighting.cpp:41-63) for a batch. */
int vx_light_prebuild(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n);
/* `expand` argument of vx_light_build: bit 0 = the `expand` flag of
 * Lighting::onChunkLoaded; bit 1 = skip Lighting::buildSkyLight, which is what
 * ChunksController::buildLights does for chunks whose lights came from the
 * cache (flags.loadedLights, logic/ChunksController.cpp:116-123). */
#define VX_LIGHT_EXPAND 1
#define VX_LIGHT_NO_SKY 2
/* For each key: Lighting::buildSkyLight(cx,cz) + Lighting::onChunkLoaded(cx,cz,
 * expand) (lighting/Lighting.cpp:65-161) followed by flags.lighted = true
 * (logic/ChunksController.cpp:106-128).  The batch result equals the
 * reference applied key by key in any order.
 * Stream-ordered like every call here: it returns with its kernels queued; a
 * solver that stalled (an internal watchdog, not a data condition) is reported
 * by the next call that waits for the device (vx_mesh_build, the getters,
 * vx_last_timing, the next vx_light_build). */
int vx_light_build(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, int32_t expand);
/* Lighting::clear (lighting/Lighting.cpp:28-39). */
int vx_light_clear(vx_ctx* ctx);
/* blocks_agent::set + Lighting::onBlockSet per edit, in submission order. */
int vx_light_edit(vx_ctx* ctx, const vx_edit* edits, int32_t n);
/* Copies Lightmap::map of a chunk to the host. */
int vx_light_get(vx_ctx* ctx, int32_t cx, int32_t cz, vx_light* out);

/* ---- meshing ------------------------------------------------------------- */

/* BlocksRenderer::build + createMesh (graphics/render/BlocksRenderer.cpp:
 * 610-664) for a batch; `capacity` is the BlocksRenderer ctor argument.
 * Results stay resident on the device until the next vx_mesh_build. */
int vx_mesh_build(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, uint64_t capacity);
int vx_mesh_get_info(vx_ctx* ctx, int32_t i, vx_mesh_info* out);
/* Any destination may be NULL to skip it. */
int vx_mesh_read(vx_ctx* ctx, int32_t i, float* vertices, int32_t* indices,
                 vx_sort_entry* entries, float* entry_floats);

/* Batched forms for callers that move whole worlds: `out` receives n
 * consecutive lightmaps (VX_CHUNK_VOL each); missing chunks are zero-filled. */
int vx_light_get_batch(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, vx_light* out);

/* The results of the last vx_mesh_build live in four device arenas (vertex
 * floats, indices, sort entries, entry floats); a chunk's data starts at the
 * offsets of its vx_mesh_slice (in elements).  vx_mesh_read_arena copies whole
 * arenas in four transfers instead of four per chunk. */
typedef struct vx_mesh_slice {
    uint64_t vertex_off, index_off, entry_off, entry_float_off;
} vx_mesh_slice;
int vx_mesh_get_slice(vx_ctx* ctx, int32_t i, vx_mesh_slice* out);
int vx_mesh_arena_sizes(vx_ctx* ctx, uint64_t sizes[4]);
int vx_mesh_read_arena(vx_ctx* ctx, float* vertices, int32_t* indices,
                       vx_sort_entry* entries, float* entry_floats);

/* ---- scheduling ----------------------------------------- */

typedef struct vx_update_stats {
    int32_t n_lit;    /* chunks lit by this call */
    int32_t n_meshed; /* chunks (re)meshed by this call; read them with vx_mesh_get_info / vx_mesh_read */
} vx_update_stats;

/* One frame's worth of the reference's scheduling, for the whole window at once
 * instead of <= 128 items / 4 ms per frame:
 *  - ChunksController::loadVisible + buildLights (logic/ChunksController.cpp:
 *    63-128): every present chunk inside the `padding` margin that is not
 *    lighted and has all 8 neighbours present (MIN_SURROUNDING = 9) is lit —
 *    with buildSkyLight + onChunkLoaded(expand = true), or, when its lights
 *    came with it (vx_chunks_put* with a lightmap = flags.loadedLights), with
 *    onChunkLoaded(expand = false) only;
 *  - ChunksRenderer::retrieveChunk + getOrRender (graphics/render/
 *    ChunksRenderer.cpp:128-183): every lighted chunk that has no mesh yet, or
 *    whose flags.modified is set, is meshed and its modified flag cleared.
 * The decisions are taken on the device from the chunk flags kept there. */
int vx_window_update(vx_ctx* ctx, int32_t padding, uint64_t capacity, vx_update_stats* out);

/* ---- measurement hooks (bench.py) ---------------------------------------- */

/* CUDA events on the ctx stream bracketing any sequence of calls. */
int vx_timer_begin(vx_ctx* ctx);
int vx_timer_end(vx_ctx* ctx, float* ms);
/* Per-kernel device time: while enabled, every kernel launch is bracketed by
 * events on the ctx stream.  vx_profile_get(i) returns the number of kernel
 * classes; name/ms/launches describe class i since the last enable. */
int vx_profile_enable(vx_ctx* ctx, int32_t on);
int vx_profile_get(vx_ctx* ctx, int32_t i, char* name, int32_t name_cap, float* ms,
                   int64_t* launches);

/* Device milliseconds spent in the last vx_light_build / vx_mesh_build call
 * (CUDA events on the ctx stream) and the number of kernel launches made. */
int vx_last_timing(vx_ctx* ctx, float* light_ms, float* mesh_ms, int64_t* launches);
/* The last vx_light_edit call: device milliseconds, the number of waves its edits were grouped
 * into (an edit runs after every earlier edit within its reach) and of k_edit launches. */
int vx_last_edit_stats(vx_ctx* ctx, float* ms, int32_t* waves, int32_t* launches);
/* How many edits of the last vx_light_edit call ran on the lightmaps in global memory because they
 * reached outside the shared-memory box an edit is first tried in (-1: no context). */
int32_t vx_last_edit_global(vx_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VXGPU_H */
