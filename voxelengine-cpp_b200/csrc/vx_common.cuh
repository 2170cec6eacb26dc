// Shared device/host definitions of libvxgpu (sm_100a only).
//
// Device-resident world layout (one vx_ctx):
//   * a window of w*d chunk slots (Chunks / AreaMap2D, voxels/Chunks.hpp) that
//     maps to a pool of chunk records; every per-chunk array is indexed by the
//     pool index p and is contiguous per chunk, so a chunk is a handful of
//     large, 16-byte aligned, fully coalescable streams:
//       vox    u32[65536]   id | state<<16          (reference `voxel` layout)
//       light  u16[65536]   R | G<<4 | B<<8 | S<<12 (reference `light_t` layout)
//       lm     u16[65536]   the light the mesher samples (backlight, isOpenForLight)
//       lp     u32[2048]    1 bit/voxel: Block::lightPassing       (derived)
//       emit   u32[2048]    1 bit/voxel: Block::rt.emissive        (derived)
//       mblk / msc / moth  u32[2048] each, 1 bit/voxel, for the mesher (derived):
//                           blocks a simple cube's face / is a drawable simple cube /
//                           is any other drawable candidate (vx_mesh.cu)
//       skycol i16[256]     per column: highest !skyLightPassing y, -1 if none
//       ystar  i16[256]     per column: Lighting::buildSkyLight seed row, -1 if none
//       laycnt i16[256]     per layer: number of non-air voxels
//       faceL  u16[4][256][16]  copy of the four vertical border planes of
//                               `light` (x=0, x=15, z=0, z=15), always in sync
//       faceA  u64[4][256]      per border row: 4 "activated" bits per voxel
//       layerA u64[16][16]      same for the top/bottom layer of every y-slab
//   * bit planes use voxel index i = (y*16+z)*16+x: word i>>5 = y*8 + (z>>1),
//     bit i&31 = (z&1)*16 + x.  One y layer = 8 words.
//   * the light solver works on y-slabs of SLAB_H layers (NSLAB per chunk).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vxgpu.h"

namespace vx {

constexpr int CW = VX_CHUNK_W, CH = VX_CHUNK_H, CD = VX_CHUNK_D, CVOL = VX_CHUNK_VOL;
constexpr int PLANE_WORDS = CVOL / 32;  // 2048
constexpr int LAYER_WORDS = 8;
constexpr int SLAB_H = 32;
constexpr int NSLAB = CH / SLAB_H;  // 8
constexpr int SLAB_WORDS = SLAB_H * LAYER_WORDS;  // 256

// faces of a chunk, in the order used by faceL/faceA
enum Face { FACE_XM = 0, FACE_XP = 1, FACE_ZM = 2, FACE_ZP = 3 };

// packed Block flags (DevDef::flags)
constexpr uint32_t DF_GROUP_MASK = 0xFFu;         // Block::drawGroup
constexpr int DF_MODEL_SHIFT = 8;                 // 3 bits, vx_block_model
constexpr int DF_CULL_SHIFT = 11;                 // 2 bits, vx_culling_mode
constexpr int DF_ROT_SHIFT = 13;                  // 2 bits, vx_rotation_profile
constexpr uint32_t DF_LP = 1u << 15;              // lightPassing
constexpr uint32_t DF_SLP = 1u << 16;             // skyLightPassing
constexpr uint32_t DF_SHADELESS = 1u << 17;
constexpr uint32_t DF_AO = 1u << 18;
constexpr uint32_t DF_TRANSLUCENT = 1u << 19;
constexpr uint32_t DF_SOLID = 1u << 20;           // rt.solid
constexpr uint32_t DF_HITBOX = 1u << 21;          // !hitboxes.empty()
constexpr uint32_t DF_EMISSIVE = 1u << 22;        // rt.emissive
constexpr int DF_RANK_SHIFT = 24;                 // rank of drawGroup in the sorted set

struct DevDef {
    uint32_t flags;
    uint32_t emission;  // R | G<<4 | B<<8 | S<<12 (each 0..15)
};

struct DevGeom {  // rarely-read per-block geometry (aabb / custom models)
    float ha[3], hb[3];
    int32_t mesh_begin, mesh_count;
};

struct ChunkMeta {
    int32_t cx, cz;
    int32_t bottom, top;   // Chunk::bottom/top
    int32_t highest;       // Lightmap::highestPoint
    int32_t modified;      // Chunk::flags.modified
    int32_t lighted;       // Chunk::flags.lighted
    int32_t present;
    int32_t lit_now;       // member of the current vx_light_build batch
    int32_t any_act;       // some voxel of the chunk was activated in this build
    int32_t pad0, pad1;    // in_set / faces that hold a seed, during a light build
    int32_t loaded_lights; // Chunk::flags.loadedLights: the lightmap came with the chunk
    int32_t meshed;        // a mesh of this chunk has been built (ChunksRenderer::meshes)
    int32_t lstate;        // what is known about `light`: LS_UNKNOWN / LS_ZERO / LS_PRISTINE
    int32_t face_dirty;    // faces (1 << FACE_*) whose faceL copy an edit batch has yet to refresh from `light`
    int32_t lvirt;         // `light` and `faceL` are not materialised: they hold what lstate says (ZERO / PRISTINE)
    int32_t sky_min, sky_max;  // min / max of skycol[]
    // filled by k_light_begin for the chunks of a build (read by every slab's k_seed block):
    int32_t nbp[4];            // the four side neighbours present in the window (pool or -1), FACE_* order
    int32_t nb_pristine;       // this chunk and every present side neighbour hold a pristine lightmap
    int32_t nb_sky_hi;         // max of sky_max over this chunk and those neighbours
};

// ChunkMeta::lstate.  PRISTINE = exactly what Lighting::prebuildSkyLight leaves on a
// zero lightmap: S = 15 above skycol, everything else 0.  Lets the seeding
// kernels decide whole slabs from the column tables without reading the lightmap.
constexpr int LS_UNKNOWN = 0, LS_ZERO = 1, LS_PRISTINE = 2;

// Everything a kernel needs to find chunk data; passed by value.
struct DevWorld {
    int32_t ox, oz, w, d;
    const int32_t* slot;  // [w*d] pool index or -1
    vx_voxel* vox;
    vx_light* light;
    vx_light* lm;     // mesher light of the last vx_mesh_build (k_mesh_lm)
    uint32_t* lp;
    uint32_t* emit;
    uint32_t* mblk;   // mesher planes (k_derive / k_edit keep them): see mesher_bits()
    uint32_t* msc;
    uint32_t* moth;
    int16_t* skycol;
    int16_t* ystar0;  // per column: where buildSkyLight starts seeding on a pristine lightmap (static, like skycol)
    int16_t* ystar;
    int16_t* laycnt;  // [256] non-air voxels per layer (Chunk::updateHeights)
    vx_light* faceL;
    unsigned long long* faceA;
    unsigned long long* layerA;
    uint8_t* sflag;  // [pool_cap][NSLAB] per-slab seed flags of the current build
    int4* nbr;       // [pool_cap] the four side neighbours that take part in the current build (pool or -1)
    uint8_t* a0;     // [pool_cap][32768] seed nibbles (2 voxels per byte) of the current build
    ChunkMeta* meta;
    const DevDef* defs;
    const DevGeom* geom;
    const float4* uv;
    const vx_model_mesh* meshes;
    const vx_model_vertex* verts;
    int32_t pool_cap;
    int32_t n_defs;
    int32_t n_ranks;
    int32_t backlight, dense_render;

    __device__ int pool_of(int cx, int cz) const {
        int lx = cx - ox, lz = cz - oz;
        if (lx < 0 || lz < 0 || lx >= w || lz >= d) return -1;
        return __ldg(slot + (size_t)lz * w + lx);
    }
    __device__ vx_voxel* vox_of(int p) const { return vox + (size_t)p * CVOL; }
    __device__ vx_light* light_of(int p) const { return light + (size_t)p * CVOL; }
    __device__ vx_light* lm_of(int p) const { return lm + (size_t)p * CVOL; }
    __device__ uint32_t* lp_of(int p) const { return lp + (size_t)p * PLANE_WORDS; }
    __device__ uint32_t* emit_of(int p) const { return emit + (size_t)p * PLANE_WORDS; }
    __device__ uint32_t* mblk_of(int p) const { return mblk + (size_t)p * PLANE_WORDS; }
    __device__ uint32_t* msc_of(int p) const { return msc + (size_t)p * PLANE_WORDS; }
    __device__ uint32_t* moth_of(int p) const { return moth + (size_t)p * PLANE_WORDS; }
    __device__ int16_t* skycol_of(int p) const { return skycol + (size_t)p * 256; }
    __device__ int16_t* ystar0_of(int p) const { return ystar0 + (size_t)p * 256; }
    __device__ int16_t* ystar_of(int p) const { return ystar + (size_t)p * 256; }
    __device__ int16_t* laycnt_of(int p) const { return laycnt + (size_t)p * 256; }
    __device__ vx_light* faceL_of(int p, int f) const {
        return faceL + ((size_t)p * 4 + f) * (CH * 16);
    }
    __device__ unsigned long long* faceA_of(int p, int f) const {
        return faceA + ((size_t)p * 4 + f) * CH;
    }
    // li = 2*slab + (0: bottom layer of the slab, 1: top layer); 16 z-rows each
    __device__ unsigned long long* layerA_of(int p, int li) const {
        return layerA + ((size_t)p * (2 * NSLAB) + li) * 16;
    }
    __device__ uint8_t* sflag_of(int p) const { return sflag + (size_t)p * NSLAB; }
    __device__ uint8_t* a0_of(int p) const { return a0 + (size_t)p * (CVOL / 2); }
};

__device__ __forceinline__ int vidx(int x, int y, int z) { return (y * CD + z) * CW + x; }

// ---------------------------------------------------------------------------
// Bulk asynchronous copy (TMA, non-tensor form) global -> shared with mbarrier completion.
// A chunk's arrays are contiguous per y range, so a solver slab (32 layers of light = 16 KiB)
// or a run of layers moves with ONE instruction issued by one thread; the other threads keep
// initialising shared memory and everybody waits on the barrier's phase.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
// one thread: expect `bytes`, then start the copy (dst, src and bytes multiples of 16)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// "simple cube": the bulk of any world — cube model, draw group 0, default culling, not
// rotatable, not translucent, lit with ambient occlusion.  For those BlocksRenderer::isOpen
// (BlocksRenderer.hpp:121-139) reduces to "the neighbour is neither void nor a solid block of
// draw group 0", which the mesher evaluates for a whole 16-voxel row at once from bit planes.
__host__ __device__ inline bool is_simple_cube(uint32_t f) {
    return ((f >> DF_MODEL_SHIFT) & 7u) == VX_MODEL_BLOCK && (f & DF_GROUP_MASK) == 0 &&
           ((f >> DF_CULL_SHIFT) & 3u) == VX_CULLING_DEFAULT &&
           ((f >> DF_ROT_SHIFT) & 3u) == VX_ROT_NONE && !(f & DF_TRANSLUCENT);
}
// does block (id, flags) close the face of a simple cube next to it?
__host__ __device__ inline bool blocks_simple(uint32_t nid, uint32_t nf) {
    return nid == VX_BLOCK_VOID || ((nf & DF_SOLID) && (nf & DF_GROUP_MASK) == 0 && nid != 0);
}
// The three mesher bits of a voxel (id | state << 16): bit 0 -> mblk, 1 -> msc, 2 -> moth.
// msc / moth follow what BlocksRenderer::render looks at (:447-456): not air, a known block,
// segment 0.
__host__ __device__ inline uint32_t mesher_bits(uint32_t vox, uint32_t f, uint32_t n_defs) {
    const uint32_t id = vox & 0xFFFFu, state = vox >> 16;
    const bool known = id != 0 && id < n_defs;
    const bool cand = known && !VX_STATE_SEGMENT(state);
    const bool simple = is_simple_cube(f);
    return (blocks_simple(id, f) ? 1u : 0u) | (cand && simple ? 2u : 0u) | (cand && !simple ? 4u : 0u);
}

}  // namespace vx
