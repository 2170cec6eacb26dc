/*
 * vxoracle.h — C API shared by the two CPU checkers of the chunk light + mesh
 * path.  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load these libraries.
 * The product (voxelengine-cpp_b200/) never links or calls them.
 *
 *   oracle/_ref/libvxref.so   the reference's own translation units compiled
 *                             from /root/reference (oracle/ref_build/Makefile)
 *   oracle/libvxoracle.so     a from-scratch CPU restatement (oracle/vxoracle.cpp)
 *
 * Both export exactly these symbols, so tests can swap one for the other.
 */
#ifndef VXORACLE_H
#define VXORACLE_H

#include "../include/vxgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vxo_world vxo_world;

/* "reference" or "port" */
const char* vxo_kind(void);

/* Window of chunk coords [ox,ox+w) x [oz,oz+d) (Chunks + AreaMap2D). */
vxo_world* vxo_create(const vx_content* content, const vx_settings* settings,
                      int32_t ox, int32_t oz, int32_t w, int32_t d);
void vxo_destroy(vxo_world* world);

/* Chunks::putChunk + Chunk::updateHeights; light may be NULL (zeros). */
int vxo_put_chunk(vxo_world* world, int32_t cx, int32_t cz,
                  const vx_voxel* voxels, const vx_light* light);
/* Chunk::decode + Lightmap::decode (lights may be NULL), then as vxo_put_chunk */
int vxo_put_chunk_encoded(vxo_world* world, int32_t cx, int32_t cz,
                          const uint8_t* voxels, const uint8_t* lights);
/* Chunk::encode / Lightmap::encode; either output may be NULL */
int vxo_encode_chunk(vxo_world* world, int32_t cx, int32_t cz, uint8_t* voxels_out,
                     uint8_t* lights_out);
/* extrle::encode16 / encode (wide = 1 / 0), coders/rle.cpp:108-131,160-205: the
 * region-file compression of the voxels / lights layers (world/files/
 * WorldRegions.cpp:60-72).  dst must hold 2 * n bytes; returns the encoded size. */
uint64_t vxo_extrle_encode(const uint8_t* src, uint64_t n, uint8_t* dst, int32_t wide);
/* extrle::decode16 / decode, coders/rle.cpp:93-106,133-158; returns the decoded
 * size in bytes.  The caller sizes dst (the reference does not bound its writes). */
uint64_t vxo_extrle_decode(const uint8_t* src, uint64_t n, uint8_t* dst, int32_t wide);
/* WorldGenerator::generate without structure placements (world/generator/
 * WorldGenerator.cpp:87-116,362-461) for every chunk of the batch, then as
 * vxo_put_chunk with a zero lightmap.  See vx_terrain_fill. */
int vxo_terrain_fill(vxo_world* world, const vx_gen_def* def, const vx_gen_chunk* chunks, int32_t n);
/* Lighting::prebuildSkyLight */
int vxo_prebuild_sky(vxo_world* world, int32_t cx, int32_t cz);
/* Lighting::buildSkyLight (unless VX_LIGHT_NO_SKY) + Lighting::onChunkLoaded
 * (expand = VX_LIGHT_EXPAND bit) + flags.lighted = true */
int vxo_light_chunk(vxo_world* world, int32_t cx, int32_t cz, int32_t expand);
/* Lighting::clear */
int vxo_light_clear(vxo_world* world);
/* blocks_agent::set + Lighting::onBlockSet */
int vxo_set_block(vxo_world* world, int32_t x, int32_t y, int32_t z,
                  uint16_t id, uint16_t state);
int vxo_get_light(vxo_world* world, int32_t cx, int32_t cz, vx_light* out);
int vxo_get_voxels(vxo_world* world, int32_t cx, int32_t cz, vx_voxel* out);
int vxo_get_info(vxo_world* world, int32_t cx, int32_t cz, vx_chunk_info* out);
int vxo_clear_modified(vxo_world* world);
/* sets Chunk::flags.modified of one chunk (test plumbing) */
int vxo_mark_modified(vxo_world* world, int32_t cx, int32_t cz);

/* BlocksRenderer(capacity,...).build + createMesh; result kept in the world
 * until the next call. */
int vxo_mesh_chunk(vxo_world* world, int32_t cx, int32_t cz, uint64_t capacity,
                   vx_mesh_info* info);
int vxo_mesh_read(vxo_world* world, float* vertices, int32_t* indices,
                  vx_sort_entry* entries, float* entry_floats);

/* CPU baseline helper: meshes `n` chunks with `threads` workers (one renderer
 * per worker, as ChunksRenderer's pool does) and returns wall seconds; the
 * meshes are discarded, `faces_out` (nullable) receives total index count/6. */
double vxo_mesh_batch(vxo_world* world, const vx_chunk_key* keys, int32_t n,
                      uint64_t capacity, int32_t threads, uint64_t* faces_out);
/* CPU baseline helper: lights `n` chunks in order on one thread; seconds. */
double vxo_light_batch(vxo_world* world, const vx_chunk_key* keys, int32_t n,
                       int32_t expand);

#ifdef __cplusplus
}
#endif
#endif
