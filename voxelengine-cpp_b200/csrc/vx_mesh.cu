// libvxgpu: meshing.  Replaces BlocksRenderer::build + createMesh
// (graphics/render/BlocksRenderer.cpp:40-664) and the padded-volume gather of
// Chunks::getVoxels (voxels/Chunks.cpp:336-418) for a batch of chunks.
//
// The reference walks draw groups in ascending order and, inside a group, the
// voxels in index order, appending to one vertex/index cursor.  Here every
// chunk is cut into 16 tiles of 16 layers and the work is split into a
// latency-bound classification pass and a streaming emission pass; offsets come
// from scans, so the output order is the reference's without any atomics on it:
//   k_mesh_classify  stage the tile's block ids (+1-voxel halo) in shared memory,
//                    classify every voxel (draw group, open-face mask via
//                    isOpen, translucent), reduce per (tile, draw group): opaque
//                    floats / indices and translucent floats / entries; list the
//                    tile's emission units (a cube face, or a whole voxel of
//                    another model) with tile-relative offsets
//   k_mesh_scan      per chunk: exclusive offsets in reference order
//   k_mesh_offsets   over the batch: arena offsets and totals
//   k_mesh_emit      one thread per unit: vertices from the light cells around
//                    the face, written at the final offset; the capacity cut-off
//                    (BlocksRenderer.cpp:84,124,162,292) becomes "keep a
//                    primitive iff its end offset <= capacity"
//   k_mesh_final     thin-AABB merge rule of renderTranslucent (:588-606)
// All vertex arithmetic is IEEE single without FMA contraction (-fmad=false),
// in the reference's evaluation order.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vx_ctx.hpp"

using namespace vx;

namespace {

constexpr int TILE_H = 16;
constexpr int NTILE = CH / TILE_H;  // 16
constexpr int TW = 18;
constexpr int TL = TW * TW;
constexpr int TSZ = (TILE_H + 2) * TL;
constexpr int VSZ = VX_VERTEX_SIZE;

// per-voxel classification (s_info)
constexpr uint32_t VI_RANK = 0xFFu;
constexpr int VI_MASK_SHIFT = 8;  // 6 bits: faces to emit
constexpr uint32_t VI_TRANSL = 1u << 14;
constexpr uint32_t VI_DRAW = 1u << 15;

struct V3 {
    float x, y, z;
};
struct V4 {
    float r, g, b, a;
};
struct I3 {
    int x, y, z;
};

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3 {x, y, z}; }
__device__ __forceinline__ V3 v3i(I3 i) { return V3 {(float)i.x, (float)i.y, (float)i.z}; }
__device__ __forceinline__ I3 trunc3(V3 v) { return I3 {(int)v.x, (int)v.y, (int)v.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3 {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3 {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3 {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3 {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3 {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ I3 operator+(I3 a, I3 b) { return I3 {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ I3 operator-(I3 a, I3 b) { return I3 {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ I3 operator-(I3 a) { return I3 {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V4 operator+(V4 a, V4 b) {
    return V4 {a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a};
}
__device__ __forceinline__ V4 operator*(V4 a, V4 b) {
    return V4 {a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a};
}
__device__ __forceinline__ V4 operator*(V4 a, float s) { return V4 {a.r * s, a.g * s, a.b * s, a.a * s}; }
__device__ __forceinline__ V4 splat4(float s) { return V4 {s, s, s, s}; }

// glm::dot / normalize / cross, scalar definitions (glm 0.9.9 without intrinsics)
__device__ __forceinline__ float dot3(V3 a, V3 b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return px + py + pz;
}
__device__ __forceinline__ V3 normalize3(V3 v) {
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(dot3(v, v)));
    return v * inv;
}
__device__ __forceinline__ V3 cross3(V3 x, V3 y) {
    return V3 {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}

// voxels/Block.cpp:57-92 (BlockRotProfile PIPE / PANE); axes as 2-bit-per-
// component codes would be smaller, a constant table is simpler.
struct Orient {
    int8_t ax[3][3];
    int8_t fix[3];
};
__constant__ Orient c_orient[3][8];

struct Vtx {
    float x, y, z, u, v;
    uint32_t l;
};

// BlocksRenderer::vertex light packing (BlocksRenderer.cpp:50-60).  x86
// converts a negative float through a 64-bit integer and truncates.
__device__ __forceinline__ uint32_t pack_light(V4 l) {
    uint32_t p = ((uint32_t)(long long)(l.r * 255) & 0xffu) << 24;
    p |= ((uint32_t)(long long)(l.g * 255) & 0xffu) << 16;
    p |= ((uint32_t)(long long)(l.b * 255) & 0xffu) << 8;
    p |= ((uint32_t)(long long)(l.a * 255) & 0xffu);
    return p;
}
__device__ __forceinline__ Vtx mkv(V3 c, float u, float v, V4 light) {
    return Vtx {c.x, c.y, c.z, u, v, pack_light(light)};
}

// ordered-uint encoding of floats for atomicMin/Max
__device__ __forceinline__ uint32_t f2o(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float o2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ---------------------------------------------------------------------------
// Tile: block ids and mesher light of a 16-layer slice with 1-voxel halo
// ---------------------------------------------------------------------------
struct Tile {
    const DevWorld* W;
    const uint16_t* id;
    const uint16_t* lm;
    int ty0;
    const int* pools;  // 3x3 neighbourhood, [dz+1][dx+1] (staged mode: shared memory; an array member
                       // would be indexed dynamically and drag the whole struct into local memory)
    const uint32_t* dflags;  // shared-memory copy of DevDef::flags for ids < 256
    int cx, cz, own;  // global mode (id == nullptr): neighbour chunks are looked up on demand
    // global mode, cells within one voxel of (x, z): the chunk across the x border this voxel
    // touches (if any), across the z border, and across both; hdr_pools = the 3x3 table in HBM
    int px, pz, pxz;
    const int* hdr_pools;
};

__device__ __forceinline__ uint32_t flags_of(const DevWorld& W, const uint32_t* dflags, uint32_t id) {
    if (id < 256u) return dflags[id];  // 0 beyond n_defs
    return id < (uint32_t)W.n_defs ? __ldg(&W.defs[id].flags) : 0u;
}

// RGB + 1 saturating at 15, S unchanged (Chunks.cpp:395-410)
__device__ __forceinline__ uint32_t backlight(uint32_t l) {
    uint32_t rgb = l & 0x0FFFu;
    uint32_t is15 = rgb & (rgb >> 1) & (rgb >> 2) & (rgb >> 3) & 0x0111u;
    return (l & 0xF000u) | (rgb + (0x0111u ^ is15));
}

// mesher light of one voxel: zero unless the block lets light through
// (isOpenForLight, BlocksRenderer.cpp:385-400), RGB+1 with backlight
__device__ __forceinline__ uint32_t mesher_light(const DevWorld& W, uint32_t f, uint32_t id, uint32_t l) {
    if (!((f & DF_LP) || id == 0)) return 0;
    if (W.backlight && (f & DF_LP)) l = backlight(l);
    return l;
}

// What VoxelsVolume holds for chunk-local (x, y, z) (x, z in [-2, 17]):
// returns the id, *lm = the light pickLight would see (BlocksRenderer.cpp:385-411)
__device__ __forceinline__ uint32_t sample_global(const Tile& T, int x, int y, int z, uint32_t* lm) {
    const DevWorld& W = *T.W;
    const uint32_t* dflags = T.dflags;
    *lm = 0;
    if (x < -2 || x > 17 || z < -2 || z > 17 || y < 0 || y >= CH) return VX_BLOCK_VOID;
    int p;
    if (T.id)
        p = T.pools[((z >> 4) + 1) * 3 + (x >> 4) + 1];
    else
        p = ((x | z) >> 4) == 0 ? T.own : W.pool_of(T.cx + (x >> 4), T.cz + (z >> 4));
    if (p < 0) return VX_BLOCK_VOID;
    const int i = vidx(x & 15, y, z & 15);
    // both loads are issued before anything depends on either
    const uint32_t v = W.vox_of(p)[i];
    const uint32_t l = W.light_of(p)[i];
    const uint32_t id = v & 0xFFFFu;
    *lm = mesher_light(W, flags_of(W, dflags, id), id, l);
    return id;
}

__device__ __forceinline__ int tidx(int ly, int z, int x) { return (ly + 1) * TL + (z + 1) * TW + (x + 1); }

// Stages ids (+ mesher light, + the low state byte of the own voxels) of layers
// ty0-1 .. ty0+16.  Own-chunk rows are read as uint4 (4 voxels) / uint2 (4
// lights); both loads are issued before anything depends on them.
template <bool WITH_LIGHT>
__device__ void load_tile(const DevWorld& W, const int* pools, const uint32_t* dflags, int ty0,
                          uint16_t* s_id, uint16_t* s_lm, uint8_t* s_st) {
    const int pc = pools[4];
    const vx_voxel* cvox = W.vox_of(pc);
    const vx_light* clight = W.light_of(pc);
    for (int u = threadIdx.x; u < (TILE_H + 2) * 16 * 4; u += blockDim.x) {
        const int q = u & 3, z = (u >> 2) & 15, l = u >> 6;
        const int y = ty0 + l - 1;
        const int dst = tidx(l - 1, z, q * 4);
        uint32_t ids[4] = {VX_BLOCK_VOID, VX_BLOCK_VOID, VX_BLOCK_VOID, VX_BLOCK_VOID};
        uint32_t lms[4] = {0, 0, 0, 0};
        if (y >= 0 && y < CH) {
            const int i = vidx(q * 4, y, z);
            const uint4 v = *reinterpret_cast<const uint4*>(cvox + i);
            uint2 lt = make_uint2(0, 0);
            if (WITH_LIGHT) lt = *reinterpret_cast<const uint2*>(clight + i);
            const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
            const uint32_t ll[4] = {lt.x & 0xFFFFu, lt.x >> 16, lt.y & 0xFFFFu, lt.y >> 16};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                ids[j] = vv[j] & 0xFFFFu;
                if (WITH_LIGHT) lms[j] = mesher_light(W, flags_of(W, dflags, ids[j]), ids[j], ll[j]);
            }
            if (l >= 1 && l <= TILE_H) {
                const int o = (l - 1) * 256 + z * 16 + q * 4;
                *reinterpret_cast<uint32_t*>(s_st + o) =
                    ((vv[0] >> 16) & 0xFFu) | (((vv[1] >> 16) & 0xFFu) << 8) |
                    (((vv[2] >> 16) & 0xFFu) << 16) | (((vv[3] >> 16) & 0xFFu) << 24);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            s_id[dst + j] = (uint16_t)ids[j];
            if (WITH_LIGHT) s_lm[dst + j] = (uint16_t)lms[j];
        }
    }
    // the ring of 68 halo cells per layer
    for (int h = threadIdx.x; h < (TILE_H + 2) * 68; h += blockDim.x) {
        const int l = h / 68, e = h % 68;
        int x, z;
        if (e < 18) { z = -1; x = e - 1; }
        else if (e < 36) { z = 16; x = e - 19; }
        else { z = (e - 36) >> 1; x = (e & 1) ? 16 : -1; }
        const int y = ty0 + l - 1;
        uint32_t lm = 0, id = VX_BLOCK_VOID;
        if (y >= 0 && y < CH) {
            const int p = pools[((z >> 4) + 1) * 3 + (x >> 4) + 1];
            if (p >= 0) {
                const int i = vidx(x & 15, y, z & 15);
                const uint32_t v = W.vox_of(p)[i];
                uint32_t lt = 0;
                if (WITH_LIGHT) lt = W.light_of(p)[i];
                id = v & 0xFFFFu;
                if (WITH_LIGHT) lm = mesher_light(W, flags_of(W, dflags, id), id, lt);
            }
        }
        s_id[tidx(l - 1, z, x)] = (uint16_t)id;
        if (WITH_LIGHT) s_lm[tidx(l - 1, z, x)] = (uint16_t)lm;
    }
}

__device__ __forceinline__ bool in_tile(const Tile& T, int x, int y, int z) {
    if (!T.id) return false;  // global mode (k_mesh_emit): every cell is read from HBM / L2
    return x >= -1 && x <= 16 && z >= -1 && z <= 16 && y >= T.ty0 - 1 && y <= T.ty0 + TILE_H;
}
__device__ __forceinline__ uint32_t pick_id(const Tile& T, int x, int y, int z) {
    if (in_tile(T, x, y, z)) return T.id[tidx(y - T.ty0, z, x)];
    uint32_t lm;
    return sample_global(T, x, y, z, &lm);
}
__device__ __forceinline__ uint32_t pick_lm(const Tile& T, int x, int y, int z) {
    if (in_tile(T, x, y, z)) return T.lm[tidx(y - T.ty0, z, x)];
    if (!T.id) {
        // global mode: the mesher-light plane k_mesh_lm prepared (one 2-byte load)
        const DevWorld& W = *T.W;
        if (x < -2 || x > 17 || z < -2 || z > 17 || y < 0 || y >= CH) return 0;
        const int p = ((x | z) >> 4) == 0 ? T.own : __ldg(T.hdr_pools + ((z >> 4) + 1) * 3 + (x >> 4) + 1);
        if (p < 0) return 0;
        return W.lm_of(p)[vidx(x & 15, y, z & 15)];
    }
    uint32_t lm;
    sample_global(T, x, y, z, &lm);
    return lm;
}
// global mode, |x - voxel.x| <= 1 and |z - voxel.z| <= 1: the chunk is one of four known
// pools, the load is always issued (clamped address) and masked afterwards, so the
// nine cells of a face are nine independent loads
__device__ __forceinline__ uint32_t near_lm(const Tile& T, int x, int y, int z) {
    const bool ox = (unsigned)x > 15u, oz = (unsigned)z > 15u;
    const int p = ox ? (oz ? T.pxz : T.px) : (oz ? T.pz : T.own);
    const bool ok = p >= 0 && (unsigned)y < (unsigned)CH;
    const uint32_t l = T.W->lm_of(max(p, 0))[vidx(x & 15, min(max(y, 0), CH - 1), z & 15)];
    return ok ? l : 0u;
}
// pickLight, BlocksRenderer.cpp:402-411
__device__ __forceinline__ V4 pick_light(const Tile& T, int x, int y, int z) {
    const uint32_t l = pick_lm(T, x, y, z);
    return V4 {__fdiv_rn((float)(l & 0xFu), 15.0f), __fdiv_rn((float)((l >> 4) & 0xFu), 15.0f),
               __fdiv_rn((float)((l >> 8) & 0xFu), 15.0f), __fdiv_rn((float)(l >> 12), 15.0f)};
}
// pickSoftLight, BlocksRenderer.cpp:413-433
__device__ __forceinline__ V4 pick_soft(const Tile& T, I3 c, I3 right, I3 up) {
    I3 a = c - right, b = c - right - up, d = c - up;
    return (pick_light(T, c.x, c.y, c.z) + pick_light(T, a.x, a.y, a.z) +
            pick_light(T, b.x, b.y, b.z) + pick_light(T, d.x, d.y, d.z)) *
           0.25f;
}

// BlocksRenderer::isOpen, BlocksRenderer.hpp:121-139
__device__ __forceinline__ bool is_open(const DevWorld& W, const uint32_t* dflags, uint32_t nid,
                                        uint32_t def_flags, uint32_t def_id) {
    if (nid == VX_BLOCK_VOID) return false;
    const uint32_t bf = flags_of(W, dflags, nid);
    const uint32_t bg = bf & DF_GROUP_MASK;
    if ((bg != (def_flags & DF_GROUP_MASK) && bg) || !(bf & DF_SOLID)) return true;
    const uint32_t cull = (def_flags >> DF_CULL_SHIFT) & 3u;
    if ((cull == VX_CULLING_DISABLED || (cull == VX_CULLING_OPTIONAL && W.dense_render)) &&
        nid == def_id)
        return true;
    return nid == 0;
}

__device__ __forceinline__ void axes_of(uint32_t def_flags, uint32_t state, I3& X, I3& Y, I3& Z,
                                        I3* fix) {
    const uint32_t prof = (def_flags >> DF_ROT_SHIFT) & 3u;
    X = I3 {1, 0, 0};
    Y = I3 {0, 1, 0};
    Z = I3 {0, 0, 1};
    if (fix) *fix = I3 {0, 0, 0};
    if (prof != VX_ROT_NONE) {
        const Orient& o = c_orient[prof][VX_STATE_ROTATION(state)];
        X = I3 {o.ax[0][0], o.ax[0][1], o.ax[0][2]};
        Y = I3 {o.ax[1][0], o.ax[1][1], o.ax[1][2]};
        Z = I3 {o.ax[2][0], o.ax[2][1], o.ax[2][2]};
        if (fix) *fix = I3 {o.fix[0], o.fix[1], o.fix[2]};
    }
}

// Classification of one voxel of the chunk being meshed (BlocksRenderer::render
// / renderTranslucent filters, :452-463, and the face culling of blockCube).
__device__ uint32_t classify(const DevWorld& W, const Tile& T, uint32_t vox, int x, int y, int z) {
    const uint32_t id = vox & 0xFFFFu, state = vox >> 16;
    if (id == 0 || id >= (uint32_t)W.n_defs || VX_STATE_SEGMENT(state)) return 0;
    const uint32_t f = flags_of(W, T.dflags, id);
    const uint32_t model = (f >> DF_MODEL_SHIFT) & 7u;
    uint32_t mask = 0;
    bool draw = false;
    if (model == VX_MODEL_BLOCK) {
        I3 X, Y, Z;
        axes_of(f, state, X, Y, Z, nullptr);
        const I3 c {x, y, z};
        const I3 probes[6] = {c + Z, c - Z, c + Y, c - Y, c + X, c - X};
#pragma unroll
        for (int k = 0; k < 6; k++)
            if (is_open(W, T.dflags, pick_id(T, probes[k].x, probes[k].y, probes[k].z), f, id)) mask |= 1u << k;
        draw = mask != 0;
    } else if (model == VX_MODEL_XSPRITE) {
        mask = 0xFu;
        draw = true;
    } else if (model == VX_MODEL_AABB) {
        if (f & DF_HITBOX) {
            mask = 0x3Fu;
            draw = true;
        }
    } else if (model == VX_MODEL_CUSTOM) {
        draw = W.geom[id].mesh_count > 0;
    }
    if (!draw) return 0;
    return (f >> DF_RANK_SHIFT) | (mask << VI_MASK_SHIFT) | ((f & DF_TRANSLUCENT) ? VI_TRANSL : 0u) |
           VI_DRAW;
}

// (vertex floats, indices) a drawable voxel appends
__device__ __forceinline__ void voxel_counts(const DevWorld& W, uint32_t info, uint32_t id,
                                             uint32_t& floats, uint32_t& idx) {
    const uint32_t mask = (info >> VI_MASK_SHIFT) & 0x3Fu;
    if (mask) {
        const uint32_t n = __popc(mask);
        floats = n * 4 * VSZ;
        idx = n * 6;
        return;
    }
    floats = idx = 0;
    const DevGeom g = W.geom[id];
    for (int m = 0; m < g.mesh_count; m++) {
        const uint32_t vc = W.meshes[g.mesh_begin + m].vertex_count;
        floats += vc * VSZ;
        idx += vc;
    }
}

// ---------------------------------------------------------------------------
// output cursor of one voxel
// ---------------------------------------------------------------------------
struct Writer {
    uint32_t* vtx;   // chunk's opaque vertex words / translucent de-indexed words
    int32_t* idx;    // chunk's index array (opaque only)
    uint32_t fpos;   // float cursor, relative to the chunk
    uint32_t ipos;   // index cursor (opaque) / vertex count of this voxel (translucent)
    uint32_t cap;    // capacity in floats (opaque)
    bool transl;
    bool dead;       // a primitive did not fit: BlocksRenderer::overflow
    float wx, wz;    // translucent world offset (cx*16 + 0.5, cz*16 + 0.5)
    float mn[3], mx[3];
    bool has_pts;

    __device__ __forceinline__ void put(uint32_t at, const Vtx& v) {
        uint2* d = reinterpret_cast<uint2*>(vtx + at);
        d[0] = make_uint2(__float_as_uint(v.x), __float_as_uint(v.y));
        d[1] = make_uint2(__float_as_uint(v.z), __float_as_uint(v.u));
        d[2] = make_uint2(__float_as_uint(v.v), v.l);
    }
    __device__ __forceinline__ void put_world(uint32_t at, const Vtx& v) {
        // SortingMeshEntry vertices are moved to world space (:560-572)
        Vtx t = v;
        t.x = v.x + wx;
        t.y = v.y + 0.5f;
        t.z = v.z + wz;
        put(at, t);
        if (!has_pts) {
            has_pts = true;
            mn[0] = mx[0] = v.x;
            mn[1] = mx[1] = v.y;
            mn[2] = mx[2] = v.z;
        } else {
            mn[0] = fminf(mn[0], v.x);
            mx[0] = fmaxf(mx[0], v.x);
            mn[1] = fminf(mn[1], v.y);
            mx[1] = fmaxf(mx[1], v.y);
            mn[2] = fminf(mn[2], v.z);
            mx[2] = fmaxf(mx[2], v.z);
        }
    }
    // BlocksRenderer.cpp:84,124,162,292: vertexOffset + floats > capacity
    __device__ __forceinline__ bool full(uint32_t floats) {
        if (transl) return false;
        if (dead || fpos + floats > cap) {
            dead = true;
            return true;
        }
        return false;
    }
    // 4 vertices + index pattern (a,b,c,d,e,f) relative to the first vertex
    __device__ __forceinline__ void quad(const Vtx (&v)[4], int a, int b, int c, int d, int e, int f) {
        if (transl) {
            put_world(fpos, v[a]);
            put_world(fpos + VSZ, v[b]);
            put_world(fpos + 2 * VSZ, v[c]);
            put_world(fpos + 3 * VSZ, v[d]);
            put_world(fpos + 4 * VSZ, v[e]);
            put_world(fpos + 5 * VSZ, v[f]);
            fpos += 6 * VSZ;
        } else {
            put(fpos, v[0]);
            put(fpos + VSZ, v[1]);
            put(fpos + 2 * VSZ, v[2]);
            put(fpos + 3 * VSZ, v[3]);
            const int base = (int)(fpos / VSZ);
            int32_t* o = idx + ipos;
            o[0] = base + a;
            o[1] = base + b;
            o[2] = base + c;
            o[3] = base + d;
            o[4] = base + e;
            o[5] = base + f;
            fpos += 4 * VSZ;
            ipos += 6;
        }
    }
    // one vertex of a triangle list with sequential indices (blockCustomModel)
    __device__ __forceinline__ void tri_vertex(const Vtx& v) {
        if (transl) {
            put_world(fpos, v);
        } else {
            put(fpos, v);
            idx[ipos] = (int)(fpos / VSZ);
            ipos++;
        }
        fpos += VSZ;
    }
};

struct UV {
    float u1, v1, u2, v2;
};
__device__ __forceinline__ UV uv_of(const DevWorld& W, uint32_t id, int side) {
    const float4 r = __ldg(&W.uv[id * 6 + side]);
    return UV {r.x, r.y, r.z, r.w};
}

__device__ __forceinline__ V3 sun() { return V3 {0.2275f, 0.9388f, -0.1005f}; }  // :12

// face with precalculated lights, BlocksRenderer.cpp:74-97
__device__ void face_lights(Writer& wr, V3 coord, float fw, float fh, float fd, V3 ax, V3 ay, V3 az,
                            const UV& r, const V4 (&lights)[4], V4 tint) {
    if (wr.full(VSZ * 4)) return;
    const V3 X = ax * fw, Y = ay * fh, Z = az * fd;
    const float s = 0.5f;
    Vtx v[4];
    v[0] = mkv(coord + (-X - Y + Z) * s, r.u1, r.v1, lights[0] * tint);
    v[1] = mkv(coord + (X - Y + Z) * s, r.u2, r.v1, lights[1] * tint);
    v[2] = mkv(coord + (X + Y + Z) * s, r.u2, r.v2, lights[2] * tint);
    v[3] = mkv(coord + (-X + Y + Z) * s, r.u1, r.v2, lights[3] * tint);
    wr.quad(v, 0, 1, 3, 1, 2, 3);
}

// vertexAO, BlocksRenderer.cpp:99-114
__device__ __forceinline__ Vtx vertex_ao(const Tile& T, V3 coord, float u, float v, V4 tint, V3 ax,
                                         V3 ay, V3 az) {
    const V3 pos = coord + az * 0.5f + (ax + ay) * 0.5f;
    const I3 ip {(int)roundf(pos.x), (int)roundf(pos.y), (int)roundf(pos.z)};
    const V4 light = pick_soft(T, ip, trunc3(ax), trunc3(ay));
    return mkv(coord, u, v, light * tint);
}

// faceAO, BlocksRenderer.cpp:116-151
__device__ void face_ao(Writer& wr, const Tile& T, V3 coord, V3 X, V3 Y, V3 Z, const UV& r,
                        bool lights) {
    if (wr.full(VSZ * 4)) return;
    const float s = 0.5f;
    Vtx v[4];
    if (lights) {
        float d = dot3(normalize3(Z), sun());
        d = 0.7f + d * 0.3f;
        const V3 ax = normalize3(X), ay = normalize3(Y), az = normalize3(Z);
        const V4 tint = splat4(d);
        v[0] = vertex_ao(T, coord + (-X - Y + Z) * s, r.u1, r.v1, tint, ax, ay, az);
        v[1] = vertex_ao(T, coord + (X - Y + Z) * s, r.u2, r.v1, tint, ax, ay, az);
        v[2] = vertex_ao(T, coord + (X + Y + Z) * s, r.u2, r.v2, tint, ax, ay, az);
        v[3] = vertex_ao(T, coord + (-X + Y + Z) * s, r.u1, r.v2, tint, ax, ay, az);
    } else {
        const V4 tint = splat4(1.0f);
        v[0] = mkv(coord + (-X - Y + Z) * s, r.u1, r.v1, tint);
        v[1] = mkv(coord + (X - Y + Z) * s, r.u2, r.v1, tint);
        v[2] = mkv(coord + (X + Y + Z) * s, r.u2, r.v2, tint);
        v[3] = mkv(coord + (-X + Y + Z) * s, r.u1, r.v2, tint);
    }
    wr.quad(v, 0, 1, 2, 0, 2, 3);
}

// flat-lit face, BlocksRenderer.cpp:153-178
__device__ void face_flat(Writer& wr, V3 coord, V3 X, V3 Y, V3 Z, const UV& r, V4 tint,
                          bool lights) {
    if (wr.full(VSZ * 4)) return;
    const float s = 0.5f;
    if (lights) {
        float d = dot3(normalize3(Z), sun());
        d = 0.7f + d * 0.3f;
        tint = tint * d;
    }
    Vtx v[4];
    v[0] = mkv(coord + (-X - Y + Z) * s, r.u1, r.v1, tint);
    v[1] = mkv(coord + (X - Y + Z) * s, r.u2, r.v1, tint);
    v[2] = mkv(coord + (X + Y + Z) * s, r.u2, r.v2, tint);
    v[3] = mkv(coord + (-X + Y + Z) * s, r.u1, r.v2, tint);
    wr.quad(v, 0, 1, 2, 0, 2, 3);
}

// blockCube, BlocksRenderer.cpp:324-383 (the open-face mask was computed by
// classify() with the same probes)
__device__ void block_cube(Writer& wr, const DevWorld& W, const Tile& T, I3 c, uint32_t id,
                           uint32_t f, uint32_t state, uint32_t mask) {
    I3 X, Y, Z;
    axes_of(f, state, X, Y, Z, nullptr);
    const bool lights = !(f & DF_SHADELESS), ao = (f & DF_AO) != 0;
    const V3 fc = v3i(c);
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
        if (!((mask >> k) & 1u)) continue;
        I3 probe, fx, fy, fz;
        int tex;
        switch (k) {
            case 0: probe = c + Z; fx = X; fy = Y; fz = Z; tex = 5; break;
            case 1: probe = c - Z; fx = -X; fy = Y; fz = -Z; tex = 4; break;
            case 2: probe = c + Y; fx = X; fy = -Z; fz = Y; tex = 3; break;
            case 3: probe = c - Y; fx = X; fy = Z; fz = -Y; tex = 2; break;
            case 4: probe = c + X; fx = -Z; fy = Y; fz = X; tex = 1; break;
            default: probe = c - X; fx = Z; fy = Y; fz = -X; tex = 0; break;
        }
        const UV r = uv_of(W, id, tex);
        if (ao) {
            face_ao(wr, T, fc, v3i(fx), v3i(fy), v3i(fz), r, lights);
        } else {
            face_flat(wr, fc, v3i(fx), v3i(fy), v3i(fz), r, pick_light(T, probe.x, probe.y, probe.z),
                      lights);
        }
    }
}

// util::PseudoRandom, maths/util.hpp:29-48,81-87
struct PseudoRandom {
    uint16_t seed;
    __device__ int rand() {
        seed = (uint16_t)(seed + 0x7ed5 + (seed << 6));
        seed = (uint16_t)(seed ^ 0xc23c ^ (seed >> 9));
        seed = (uint16_t)(seed + 0x1656 + (seed << 3));
        seed = (uint16_t)((seed + 0xa264) ^ (seed << 4));
        seed = (uint16_t)(seed + 0xfd70 - (seed << 3));
        seed = (uint16_t)(seed ^ 0xba49 ^ (seed >> 8));
        return (int)seed;
    }
    __device__ static unsigned long long step(unsigned long long x, unsigned long long m,
                                              unsigned shift) {
        unsigned long long t = ((x >> shift) ^ x) & m;
        return (x ^ t) ^ (t << shift);
    }
    __device__ void set_seed(long long number) {
        unsigned long long n = (unsigned long long)number;
        n = step(n, 0x2222222222222222ull, 1);
        n = step(n, 0x0c0c0c0c0c0c0c0cull, 2);
        n = step(n, 0x00f000f000f000f0ull, 4);
        seed = (uint16_t)n;
        rand();
    }
};

// blockXSprite, BlocksRenderer.cpp:180-217
__device__ void block_xsprite(Writer& wr, const DevWorld& W, const Tile& T, int x, int y, int z,
                              uint32_t id) {
    const I3 px {1, 0, 0}, nx {-1, 0, 0}, up {0, 1, 0};
    const V4 la = pick_soft(T, I3 {x, y + 1, z}, px, up);
    const V4 lb = pick_soft(T, I3 {x + 1, y + 1, z}, px, up);
    const V4 lc = pick_soft(T, I3 {x, y + 1, z}, nx, up);
    const V4 ld = pick_soft(T, I3 {x - 1, y + 1, z}, nx, up);
    const V4 l1[4] = {la, lb, lb, la};
    const V4 l2[4] = {lc, ld, ld, lc};
    PseudoRandom rnd;
    rnd.set_seed((long long)((x * 52321) ^ (z * 389) ^ y));
    // rand32() = (rand() << 16) | rand(); only the low half survives the
    // `short` conversion and it is the second value drawn
    rnd.rand();
    const short r16 = (short)rnd.rand();
    const float xs = ((float)(signed char)r16 / 512) * 1.0f;
    const float zs = ((float)(signed char)(r16 >> 8) / 512) * 1.0f;
    const float fw = __fdiv_rn(1.0f, 1.41f);
    const V4 tint = splat4(0.8f);
    const V3 c = v3(x + xs, (float)y, z + zs);
    const V3 Y1 = v3(0, 1, 0), Z0 = v3(0, 0, 0);
    const UV t1 = uv_of(W, id, 0 /* FACE_MX */), t2 = uv_of(W, id, 4 /* FACE_MZ */);
    face_lights(wr, c, fw, 1.0f, 0, v3(-1, 0, 1), Y1, Z0, t1, l2, tint);
    face_lights(wr, c, fw, 1.0f, 0, v3(1, 0, 1), Y1, Z0, t1, l1, tint);
    face_lights(wr, c, fw, 1.0f, 0, v3(-1, 0, -1), Y1, Z0, t2, l2, tint);
    face_lights(wr, c, fw, 1.0f, 0, v3(1, 0, -1), Y1, Z0, t2, l1, tint);
}

// blockAABB, BlocksRenderer.cpp:222-273
__device__ void block_aabb(Writer& wr, const DevWorld& W, const Tile& T, I3 ic, uint32_t id,
                           uint32_t f, uint32_t state) {
    const DevGeom g = W.geom[id];
    V3 a = v3(g.ha[0], g.ha[1], g.ha[2]);
    V3 b = v3(g.hb[0], g.hb[1], g.hb[2]);
    const V3 size = v3(fabsf(b.x - a.x), fabsf(b.y - a.y), fabsf(b.z - a.z));
    I3 Xi, Yi, Zi, fix;
    axes_of(f, state, Xi, Yi, Zi, &fix);
    const V3 X = v3i(Xi), Y = v3i(Yi), Z = v3i(Zi);
    V3 coord = v3i(ic);
    if (((f >> DF_ROT_SHIFT) & 3u) != VX_ROT_NONE) {
        // CoordSystem::transform, voxels/Block.cpp:47-55
        const V3 na = X * a.x + Y * a.y + Z * a.z;
        const V3 nb = X * b.x + Y * b.y + Z * b.z;
        a = na + v3i(fix);
        b = nb + v3i(fix);
    }
    const V3 center = (a + b) * 0.5f;
    coord = coord - (v3(0.5f, 0.5f, 0.5f) - center);
    const V3 sx = X * size.x, sy = Y * size.y, sz = Z * size.z;
    const V3 nsx = (-X) * size.x, nsy = (-Y) * size.y, nsz = (-Z) * size.z;
    const bool lights = !(f & DF_SHADELESS), ao = (f & DF_AO) != 0;
    V4 tint = splat4(0);
    if (!ao) tint = pick_light(T, ic.x, ic.y, ic.z);
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
        V3 fx, fy, fz;
        int tex;
        switch (k) {
            case 0: fx = sx; fy = sy; fz = sz; tex = 5; break;
            case 1: fx = nsx; fy = sy; fz = nsz; tex = 4; break;
            case 2: fx = sx; fy = nsz; fz = sy; tex = 3; break;
            case 3: fx = nsx; fy = nsz; fz = nsy; tex = 2; break;
            case 4: fx = nsz; fy = sy; fz = sx; tex = 1; break;
            default: fx = sz; fy = sy; fz = nsx; tex = 0; break;
        }
        const UV r = uv_of(W, id, tex);
        if (ao)
            face_ao(wr, T, coord, fx, fy, fz, r, lights);
        else
            face_flat(wr, coord, fx, fy, fz, r, tint, lights);
    }
}

// blockCustomModel, BlocksRenderer.cpp:275-321
__device__ void block_custom(Writer& wr, const DevWorld& W, const Tile& T, I3 ic, uint32_t id,
                             uint32_t f, uint32_t state) {
    I3 Xi, Yi, Zi;
    axes_of(f, state, Xi, Yi, Zi, nullptr);
    const V3 X = v3i(Xi), Y = v3i(Yi), Z = v3i(Zi);
    const V3 coord = v3i(ic);
    const DevGeom g = W.geom[id];
    for (int m = 0; m < g.mesh_count; m++) {
        const vx_model_mesh mesh = W.meshes[g.mesh_begin + m];
        const vx_model_vertex* mv = W.verts + mesh.vertex_begin;
        if (wr.full((uint32_t)VSZ * mesh.vertex_count)) return;
        for (int tri = 0; tri < mesh.vertex_count / 3; tri++) {
            const vx_model_vertex& pp = mv[tri * 3 + (tri % 2) * 2];
            const vx_model_vertex& qq = mv[tri * 3 + 1];
            V3 r = v3(pp.coord[0], pp.coord[1], pp.coord[2]) - v3(qq.coord[0], qq.coord[1], qq.coord[2]);
            r = normalize3(r);
            for (int i = 0; i < 3; i++) {
                const vx_model_vertex& vert = mv[tri * 3 + i];
                const V3 n = vert.normal[0] * X + vert.normal[1] * Y + vert.normal[2] * Z;
                float d = dot3(n, sun());
                d = 0.8f + d * 0.2f;
                const V3 vc = v3(vert.coord[0] - 0.5f, vert.coord[1] - 0.5f, vert.coord[2] - 0.5f);
                const Vtx o = vertex_ao(T, coord + vc.x * X + vc.y * Y + vc.z * Z, vert.uv[0],
                                        vert.uv[1], splat4(d), cross3(r, n), r, n);
                wr.tri_vertex(o);
            }
        }
    }
}

__device__ void draw_voxel(Writer& wr, const DevWorld& W, const Tile& T, uint32_t vox, uint32_t info,
                           int x, int y, int z) {
    const uint32_t id = vox & 0xFFFFu, state = vox >> 16;
    const uint32_t f = __ldg(&W.defs[id].flags);
    switch ((f >> DF_MODEL_SHIFT) & 7u) {
        case VX_MODEL_BLOCK:
            block_cube(wr, W, T, I3 {x, y, z}, id, f, state, (info >> VI_MASK_SHIFT) & 0x3Fu);
            break;
        case VX_MODEL_XSPRITE:
            block_xsprite(wr, W, T, x, y, z, id);
            break;
        case VX_MODEL_AABB:
            block_aabb(wr, W, T, I3 {x, y, z}, id, f, state);
            break;
        case VX_MODEL_CUSTOM:
            block_custom(wr, W, T, I3 {x, y, z}, id, f, state);
            break;
        default:
            break;
    }
}

// ---------------------------------------------------------------------------
// shared front half of both passes: stage the tile, classify the 4096 voxels
// ---------------------------------------------------------------------------
struct TileShared {
    uint16_t id[TSZ];
    uint16_t info[TILE_H * 256];
    uint8_t st[TILE_H * 256];  // low state byte (rotation | segment << 3) of the own voxels
    uint32_t dflags[256];
    uint32_t rowblk[(TILE_H + 2) * TW];  // per (layer, z) row incl. halo: blocking-cell bits
    uint32_t rowsc[(TILE_H + 2) * TW];   // own rows: voxels that are drawable simple cubes (before culling)
    uint32_t rowoth[(TILE_H + 2) * TW];  // own rows: every other drawable candidate (scalar classify)
    uint32_t rankmask[8];      // ranks with any drawable voxel
    uint32_t passmask[2][8];   // ... per pass (0 opaque, 1 translucent)
    int pools[9];
    unsigned long long scan[8], scan2[8];
};

__device__ __forceinline__ Tile make_tile(const DevWorld& W, const TileShared& S, int ty0) {
    return Tile {&W, S.id, nullptr, ty0, S.pools, S.dflags, 0, 0, -1, -1, -1, -1, nullptr};
}
__device__ __forceinline__ uint32_t own_id(const TileShared& S, int vi) {
    return S.id[tidx(vi >> 8, (vi >> 4) & 15, vi & 15)];
}
// id | state << 16 of an own voxel (user bits dropped: the mesher never reads them)
__device__ __forceinline__ uint32_t own_voxel(const TileShared& S, int vi) {
    return (uint32_t)S.id[tidx(vi >> 8, (vi >> 4) & 15, vi & 15)] | ((uint32_t)S.st[vi] << 16);
}

// What the 16 voxels of one row add to the output, per category
// (draw-group rank | translucent << 8): most rows hold one or two categories.
struct RowSummary {
    uint32_t cat[2];       // category | 0x200 when the slot is used
    uint32_t units[2];     // fast emission units: opaque cube faces
    uint32_t slow[2];      // low half: slow emission units (whole voxels of other models); high half: translucent
                           // cube faces (a list and a pass of their own)
    uint32_t floats[2];    // vertex floats (opaque) / de-indexed floats (translucent)
    uint32_t idx[2];       // indices (opaque) / sort entries (translucent)
    bool multi;            // a third category appeared: use the per-voxel walk
    uint32_t drawn;        // bit x: voxel x of the row is drawn (has VI_DRAW)
};
constexpr uint32_t CAT_USED = 0x200u;

// contribution of one classified voxel
template <bool SPLIT>
__device__ __forceinline__ void voxel_contrib(const DevWorld& W, const uint32_t* dflags, uint32_t info,
                                              uint32_t id, uint32_t& cat, uint32_t& nu, uint32_t& f,
                                              uint32_t& i) {
    uint32_t fl, ix;
    voxel_counts(W, info, id, fl, ix);
    const uint32_t mask = (info >> VI_MASK_SHIFT) & 0x3Fu;
    const bool cube = ((flags_of(W, dflags, id) >> DF_MODEL_SHIFT) & 7u) == VX_MODEL_BLOCK;
    const bool tr = (info & VI_TRANSL) != 0;
    // units by kind: opaque cube faces | everything of another model << 8 | translucent cube faces << 16
    // (SPLIT: translucent cube faces are listed apart; otherwise they go with the slow kinds)
    nu = cube ? (tr ? (uint32_t)__popc(mask) << (SPLIT ? 16 : 8) : (uint32_t)__popc(mask)) : 1u << 8;
    cat = (info & VI_RANK) | (tr ? 0x100u : 0u);
    f = tr ? ix * VSZ : fl;
    i = tr ? (fl ? 1u : 0u) : ix;
}

__device__ __forceinline__ void summary_add(RowSummary& rs, uint32_t cat, uint32_t nu, uint32_t f,
                                            uint32_t i) {
    // both slots are updated with selects: choosing a slot first makes the compiler address the
    // arrays through a pointer, which puts the summary in local memory
    const uint32_t c = cat | CAT_USED;
    const bool m0 = rs.cat[0] == c || rs.cat[0] == 0;
    const bool m1 = !m0 && (rs.cat[1] == c || rs.cat[1] == 0);
    rs.cat[0] = m0 ? c : rs.cat[0];
    const uint32_t st = ((nu >> 8) & 0xFFu) | ((nu >> 16) << 16);
    rs.units[0] += m0 ? nu & 0xFFu : 0u;
    rs.slow[0] += m0 ? st : 0u;
    rs.floats[0] += m0 ? f : 0u;
    rs.idx[0] += m0 ? i : 0u;
    rs.cat[1] = m1 ? c : rs.cat[1];
    rs.units[1] += m1 ? nu & 0xFFu : 0u;
    rs.slow[1] += m1 ? st : 0u;
    rs.floats[1] += m1 ? f : 0u;
    rs.idx[1] += m1 ? i : 0u;
    rs.multi = rs.multi || !(m0 || m1);
}

// totals of category `cat` in this thread's row
struct RowTot {
    unsigned long long fi;     // floats << 32 | idx (or entries)
    unsigned long long units;  // fast units | slow units << 21 | translucent cube units << 42 (each < 2^21 per tile)
};
template <bool SPLIT>
__device__ __forceinline__ RowTot row_total(const DevWorld& W, const TileShared& S,
                                            const RowSummary& rs, uint32_t cat);

// Returns false (uniformly) when the tile draws nothing: the caller leaves.
template <bool WITH_LIGHT, bool SPLIT>
__device__ bool stage_and_classify(const DevWorld& W, TileShared& S, const ChunkMeta& m, int tile,
                                   RowSummary& rs) {
    const int tid = threadIdx.x;
    if (tid < 9) S.pools[tid] = W.pool_of(m.cx + tid % 3 - 1, m.cz + tid / 3 - 1);
    if (tid < 8) S.rankmask[tid] = S.passmask[0][tid] = S.passmask[1][tid] = 0;
    {
        const uint32_t f = tid < W.n_defs ? __ldg(&W.defs[tid].flags) : 0u;
        S.dflags[tid] = f;
    }
    __syncthreads();
    const int ty0 = tile * TILE_H;
    static_assert(!WITH_LIGHT, "the staged light plane is gone: emission reads W.lm");
    // ---- row masks from the chunk's mesher planes (k_derive / k_edit keep them): bit x+1 of
    // rowblk[(ly+1)*18 + z+1] = cell (ly, z, x) blocks a simple cube's face; rowsc / rowoth = the
    // row's drawable simple cubes / other drawable candidates.  No voxel is read for this: a tile
    // that turns out to draw nothing (the solid interior of a terrain chunk) costs a few plane words.
    for (int r = tid; r < (TILE_H + 2) * TW; r += 256) {
        const int rl = r / TW - 1, rz = r % TW - 1;  // layer / z of the row; halo rows are -1 or 16
        const int y = ty0 + rl;
        uint32_t bits = 0x3FFFFu, sc = 0, oth = 0;  // void blocks (outside the world, missing chunk)
        if ((unsigned)y < (unsigned)CH) {
            const int zz = rz & 15, sh = (zz & 1) * 16, wi = y * LAYER_WORDS + (zz >> 1);
            const int pz = S.pools[rz < 0 ? 1 : rz > 15 ? 7 : 4];
            if (pz >= 0) bits = (((__ldg(W.mblk_of(pz) + wi) >> sh) & 0xFFFFu) << 1) | 0x20001u;
            if ((unsigned)rz < 16u) {
                // the cells across the x borders; corner cells of halo rows are never looked at
                const int pl = S.pools[3], pr = S.pools[5];
                if (pl >= 0 && !((__ldg(W.mblk_of(pl) + wi) >> (sh + 15)) & 1u)) bits &= ~1u;
                if (pr >= 0 && !((__ldg(W.mblk_of(pr) + wi) >> sh) & 1u)) bits &= ~0x20000u;
                if ((unsigned)rl < (unsigned)TILE_H) {
                    sc = ((__ldg(W.msc_of(pz) + wi) >> sh) & 0xFFFFu) << 1;
                    oth = ((__ldg(W.moth_of(pz) + wi) >> sh) & 0xFFFFu) << 1;
                }
            }
        }
        S.rowblk[r] = bits;
        S.rowsc[r] = sc;
        S.rowoth[r] = oth;
    }
    __syncthreads();
    {
        // does the tile draw anything at all?  (simple cubes with an open face, or candidates of
        // the scalar path)  BlocksRenderer::build only scans [bottom, top) (:625-638)
        const int ly = tid >> 4, z = tid & 15, y = ty0 + ly;
        const int r = (ly + 1) * TW + (z + 1);
        const bool in_range = y >= m.bottom && y < m.top;
        const uint32_t open_any = (~S.rowblk[r + 1] >> 1) | (~S.rowblk[r - 1] >> 1) | (~S.rowblk[r + TW] >> 1) |
                                  (~S.rowblk[r - TW] >> 1) | (~S.rowblk[r] >> 2) | ~S.rowblk[r];
        const bool mine = in_range && ((((S.rowsc[r] >> 1) & open_any) | (S.rowoth[r] >> 1)) & 0xFFFFu);
        const bool scalar = in_range && ((S.rowoth[r] >> 1) & 0xFFFFu);
        // (__syncthreads_or reduces to a truth value, not to the OR of its arguments)
        if (!__syncthreads_or(mine)) return false;
        // block ids (and states) are only needed by the scalar path
        if (__syncthreads_or(scalar)) {
            load_tile<WITH_LIGHT>(W, S.pools, S.dflags, ty0, S.id, nullptr, S.st);
            __syncthreads();
        }
    }
    const Tile T = make_tile(W, S, ty0);
    // ---- thread t classifies row (ly = t >> 4, z = t & 15): BlocksRenderer::build
    // only scans [bottom, top) (:625-638)
    {
        const int ly = tid >> 4, z = tid & 15, y = ty0 + ly;
        const int r = (ly + 1) * TW + (z + 1);
        const bool in_range = y >= m.bottom && y < m.top;
        // open-face masks of the row for simple cubes, blockCube face order (:345-380)
        const uint32_t o0 = ~S.rowblk[r + 1] >> 1, o1 = ~S.rowblk[r - 1] >> 1;    // +Z, -Z
        const uint32_t o2 = ~S.rowblk[r + TW] >> 1, o3 = ~S.rowblk[r - TW] >> 1;  // +Y, -Y
        const uint32_t o4 = ~S.rowblk[r] >> 2, o5 = ~S.rowblk[r];                 // +X, -X
        uint32_t ranks_seen = 0, rank_hi = 0, pass_seen0 = 0, pass_seen1 = 0;
        rs.cat[0] = rs.cat[1] = 0;
        rs.units[0] = rs.units[1] = rs.slow[0] = rs.slow[1] = 0;
        rs.floats[0] = rs.floats[1] = rs.idx[0] = rs.idx[1] = 0;
        rs.multi = false;
        rs.drawn = 0;
        // nothing drawn until shown otherwise: the row's 16 info words as two 16-byte stores
        {
            uint4* irow = reinterpret_cast<uint4*>(S.info + tid * 16);
            irow[0] = irow[1] = make_uint4(0, 0, 0, 0);
        }
        // ---- simple cubes, the whole row at once: a voxel is drawn iff any of its faces is open
        const uint32_t open_any = o0 | o1 | o2 | o3 | o4 | o5;
        uint32_t sc = in_range ? (S.rowsc[r] >> 1) & open_any & 0xFFFFu : 0u;
        rs.drawn = sc;
        if (sc) {
            // all simple cubes share draw group 0, hence one rank and one (opaque) category
            // (draw group 0 is the smallest group there can be: rank 0; no block id needed)
            const uint32_t rank = 0;
            const uint32_t faces = __popc(sc & o0) + __popc(sc & o1) + __popc(sc & o2) + __popc(sc & o3) +
                                   __popc(sc & o4) + __popc(sc & o5);
            summary_add(rs, rank, faces, faces * 4 * VSZ, faces * 6);
            if (rank < 32) {
                ranks_seen |= 1u << rank;
                pass_seen0 |= 1u << rank;
            } else {
                atomicOr(&S.rankmask[rank >> 5], 1u << (rank & 31u));
                atomicOr(&S.passmask[0][rank >> 5], 1u << (rank & 31u));
            }
            while (sc) {
                const int x = __ffs(sc) - 1;
                sc &= sc - 1;
                const uint32_t mask = ((o0 >> x) & 1u) | (((o1 >> x) & 1u) << 1) | (((o2 >> x) & 1u) << 2) |
                                      (((o3 >> x) & 1u) << 3) | (((o4 >> x) & 1u) << 4) | (((o5 >> x) & 1u) << 5);
                S.info[tid * 16 + x] = (uint16_t)(rank | (mask << VI_MASK_SHIFT) | VI_DRAW);
            }
        }
        // ---- every other candidate: the scalar isOpen / model rules
        uint32_t oth = in_range ? (S.rowoth[r] >> 1) & 0xFFFFu : 0u;
        while (oth) {
            const int x = __ffs(oth) - 1;
            oth &= oth - 1;
            const int vi = tid * 16 + x;
            const uint32_t info = classify(W, T, own_voxel(S, vi), x, y, z);
            if (!info) continue;
            S.info[vi] = (uint16_t)info;
            rs.drawn |= 1u << x;
            uint32_t cat, nu, cf, ci;
            voxel_contrib<SPLIT>(W, S.dflags, info, own_id(S, vi), cat, nu, cf, ci);
            summary_add(rs, cat, nu, cf, ci);
            const uint32_t rk = info & VI_RANK;
            const int ps = (info & VI_TRANSL) ? 1 : 0;
            if (rk < 32) {
                ranks_seen |= 1u << rk;
                pass_seen0 |= ps ? 0u : 1u << rk;
                pass_seen1 |= ps ? 1u << rk : 0u;
            } else {
                atomicOr(&S.rankmask[rk >> 5], 1u << (rk & 31u));
                atomicOr(&S.passmask[ps][rk >> 5], 1u << (rk & 31u));
                rank_hi = 1;
            }
        }
        (void)rank_hi;
        if (ranks_seen) atomicOr(&S.rankmask[0], ranks_seen);
        if (pass_seen0) atomicOr(&S.passmask[0][0], pass_seen0);
        if (pass_seen1) atomicOr(&S.passmask[1][0], pass_seen1);
    }
    __syncthreads();
    return true;
}

// Thread t owns the 16 consecutive voxels of row (ly = t >> 4, z = t & 15).
__device__ __forceinline__ int row_voxel(int tid, int k) { return tid * 16 + k; }

template <bool SPLIT>
__device__ __forceinline__ RowTot row_total(const DevWorld& W, const TileShared& S,
                                            const RowSummary& rs, uint32_t cat) {
    RowTot t {0, 0};
    if (!rs.multi) {
#pragma unroll
        for (int k = 0; k < 2; k++)
            if (rs.cat[k] == (cat | CAT_USED)) {
                t.fi += ((unsigned long long)rs.floats[k] << 32) | rs.idx[k];
                t.units += (unsigned long long)rs.units[k] | ((unsigned long long)(rs.slow[k] & 0xFFFFu) << 21) | ((unsigned long long)(rs.slow[k] >> 16) << 42);
            }
        return t;
    }
    for (int k = 0; k < 16; k++) {  // rare: three or more categories in one row
        const int vi = row_voxel(threadIdx.x, k);
        const uint32_t info = S.info[vi];
        if (!(info & VI_DRAW)) continue;
        uint32_t c, nu, f, i;
        voxel_contrib<SPLIT>(W, S.dflags, info, own_id(S, vi), c, nu, f, i);
        if (c == cat) {
            t.fi += ((unsigned long long)f << 32) | i;
            t.units += (unsigned long long)(nu & 0xFFu) | ((unsigned long long)((nu >> 8) & 0xFFu) << 21) | ((unsigned long long)(nu >> 16) << 42);
        }
    }
    return t;
}

// block-wide exclusive scan of a pair of values with one barrier pair
__device__ void block_scan_pair(TileShared& S, unsigned long long a, unsigned long long b,
                                unsigned long long* ex_a, unsigned long long* ex_b,
                                unsigned long long* tot_a, unsigned long long* tot_b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long ta = __shfl_up_sync(0xFFFFFFFFu, ia, o);
        const unsigned long long tb = __shfl_up_sync(0xFFFFFFFFu, ib, o);
        if (lane >= o) {
            ia += ta;
            ib += tb;
        }
    }
    __syncthreads();  // S.scan may still be read from the previous call
    if (lane == 31) {
        S.scan[warp] = ia;
        S.scan2[warp] = ib;
    }
    __syncthreads();
    unsigned long long ba = 0, bb = 0, xa = 0, xb = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const unsigned long long sa = S.scan[w], sb = S.scan2[w];
        if (w < warp) {
            ba += sa;
            bb += sb;
        }
        xa += sa;
        xb += sb;
    }
    *tot_a = xa;
    *tot_b = xb;
    *ex_a = ba + ia - a;
    *ex_b = bb + ib - b;
}

// two block-wide reductions with one barrier pair
__device__ void block_scan2(TileShared& S, unsigned long long a, unsigned long long b,
                            unsigned long long* ta, unsigned long long* tb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
        b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    }
    __syncthreads();
    if (lane == 0) {
        S.scan[warp] = a;
        S.scan2[warp] = b;
    }
    __syncthreads();
    unsigned long long x = 0, y = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        x += S.scan[w];
        y += S.scan2[w];
    }
    *ta = x;
    *tb = y;
}

// ---------------------------------------------------------------------------
// k_mesh_emit: unit-parallel emission.  A unit is one face of a cube-model
// voxel (the common case: one thread computes its 4 vertices from the 3x3
// light cells of the face plane) or one whole voxel of another model (generic
// path).
// final offsets, then processed by all threads.
// ---------------------------------------------------------------------------
constexpr uint32_t U_GENERIC = 1u << 15;
constexpr uint32_t U_FIRST = 1u << 16;   // translucent: writes the voxel's entry
constexpr uint32_t U_TRANSL = 1u << 17;  // unit of the translucent pass
constexpr int U_RANK_SHIFT = 18;         // 8 bits: draw-group rank
constexpr int U_NFACES_SHIFT = 26;       // 3 bits: faces of the voxel (entry size)

// One emission unit, written by k_mesh_classify and consumed by k_mesh_emit.
struct __align__(16) Unit {
    uint32_t code;  // voxel in tile (12) | face (3) << 12 | flags | rank | faces
    uint32_t foff;  // float offset relative to the (tile, category) base
    uint32_t ioff;  // index offset (opaque) / entry index (translucent), same base
    uint32_t where; // chunk index in the batch * 16 + tile
};

#ifndef VX_EMIT_UPT
#define VX_EMIT_UPT 4
#endif
#ifndef VX_EMIT_MINB
#define VX_EMIT_MINB 4
#endif
constexpr int EMIT_UPT = VX_EMIT_UPT;    // units per thread of k_mesh_emit<0> and <2>
constexpr int EMIT_UPT_GENERIC = 1;   // units per thread of k_mesh_emit<1> (few, slow units: more CTAs)
constexpr int STAGE_ROW = 13;  // uint2 per staged face: 12 + 1 pad (odd stride: conflict-free 8-byte rows)
constexpr int STAGE_ROW_T = 19;  // translucent face: six de-indexed vertices = 18 uint2 + 1 pad

struct EmitShared {
    float lut15[16];   // n / 15.0f
    float dtab[6];     // 0.7f + dot(normalize(axis), SUN) * 0.3f for +x,-x,+y,-y,+z,-z
    uint32_t kept_f, kept_i;
    uint32_t mn[3], mx[3];
    int has_pts;
};

__device__ __forceinline__ int dir_index(I3 a) {
    return a.x ? (a.x < 0 ? 1 : 0) : a.y ? (a.y < 0 ? 3 : 2) : (a.z < 0 ? 5 : 4);
}

struct FaceGeom {
    I3 fx, fy, fz;
    int tex;
};
__device__ __forceinline__ FaceGeom face_geom(int k, I3 X, I3 Y, I3 Z) {
    switch (k) {  // blockCube face order, BlocksRenderer.cpp:345-380
        case 0: return FaceGeom {X, Y, Z, 5};
        case 1: return FaceGeom {-X, Y, -Z, 4};
        case 2: return FaceGeom {X, -Z, Y, 3};
        case 3: return FaceGeom {X, Z, -Y, 2};
        case 4: return FaceGeom {-Z, Y, X, 1};
        default: return FaceGeom {Z, Y, -X, 0};
    }
}

// The four vertices of one cube-model face (faceAO / face of blockCube).  The
// axes are unit integer vectors, so normalize() is the identity, vertex
// positions are exact half-integers and the AO sample position of corner
// (sx, sy) is c + fz + (sx > 0) fx + (sy > 0) fy: the 4 corners read the 3x3
// cells of the plane in front of the face.
__device__ __forceinline__ void cube_face(const DevWorld& W, const Tile& T, const EmitShared& E,
                                          uint32_t vox, int x, int y, int z, int k, Vtx (&v)[4]) {
    const uint32_t id = vox & 0xFFFFu, state = vox >> 16;
    const uint32_t f = flags_of(W, T.dflags, id);
    I3 X, Y, Z;
    axes_of(f, state, X, Y, Z, nullptr);
    const FaceGeom g = face_geom(k, X, Y, Z);
    const bool lights = !(f & DF_SHADELESS), ao = (f & DF_AO) != 0;
    const UV r = uv_of(W, id, g.tex);  // issued early: the only load that depends on the block id
    uint32_t l0, l1, l2, l3;
    if (ao && lights) {
        const float d = E.dtab[dir_index(g.fz)];
        // lm of the plane cells (a, b) in {-1,0,1}^2 at c + fz + a fx + b fy
        const I3 c0p {x + g.fz.x, y + g.fz.y, z + g.fz.z};
        uint32_t cell[3][3];
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int a = 0; a < 3; a++)
                cell[b][a] = T.id ? pick_lm(T, c0p.x + (a - 1) * g.fx.x + (b - 1) * g.fy.x,
                                            c0p.y + (a - 1) * g.fx.y + (b - 1) * g.fy.y,
                                            c0p.z + (a - 1) * g.fx.z + (b - 1) * g.fy.z)
                                  : near_lm(T, c0p.x + (a - 1) * g.fx.x + (b - 1) * g.fy.x,
                                            c0p.y + (a - 1) * g.fx.y + (b - 1) * g.fy.y,
                                            c0p.z + (a - 1) * g.fx.z + (b - 1) * g.fy.z);
        // a channel that is zero in all nine cells is zero at all four corners
        uint32_t any = 0;
#pragma unroll
        for (int b = 0; b < 3; b++)
#pragma unroll
            for (int a = 0; a < 3; a++) any |= cell[b][a];
        // channel by channel: the nine cells go through the n / 15 table once, each corner adds
        // four of them in pickSoftLight's order: c, c - right, c - right - up, c - up (:413-420)
        l0 = l1 = l2 = l3 = 0;
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            const int sh = 4 * ch;
            if (!((any >> sh) & 15u)) continue;
            float fl[3][3];
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int a = 0; a < 3; a++) fl[b][a] = E.lut15[(cell[b][a] >> sh) & 15u];
            auto corner = [&](int A, int B) -> uint32_t {
                float v = fl[B + 1][A + 1] + fl[B + 1][A];
                v = v + fl[B][A];
                v = v + fl[B][A + 1];
                v = v * 0.25f;
                v = v * d;
                return ((uint32_t)(long long)(v * 255) & 0xffu) << (24 - 8 * ch);
            };
            l0 |= corner(0, 0);
            l1 |= corner(1, 0);
            l2 |= corner(1, 1);
            l3 |= corner(0, 1);
        }
    } else if (ao) {
        l0 = l1 = l2 = l3 = 0xFFFFFFFFu;  // tint 1, no light sampling (:138-145)
    } else {
        V4 tint = pick_light(T, x + g.fz.x, y + g.fz.y, z + g.fz.z);
        if (lights) tint = tint * E.dtab[dir_index(g.fz)];
        l0 = l1 = l2 = l3 = pack_light(tint);
    }
    // positions last: twelve floats that nothing above needs stay out of the register file until here
    const V3 fc = v3((float)x, (float)y, (float)z);
    const V3 FX = v3i(g.fx), FY = v3i(g.fy), FZ = v3i(g.fz);
    const float s = 0.5f;
    const V3 p0 = fc + (-FX - FY + FZ) * s, p1 = fc + (FX - FY + FZ) * s;
    const V3 p2 = fc + (FX + FY + FZ) * s, p3 = fc + (-FX + FY + FZ) * s;
    v[0] = Vtx {p0.x, p0.y, p0.z, r.u1, r.v1, l0};
    v[1] = Vtx {p1.x, p1.y, p1.z, r.u2, r.v1, l1};
    v[2] = Vtx {p2.x, p2.y, p2.z, r.u2, r.v2, l2};
    v[3] = Vtx {p3.x, p3.y, p3.z, r.u1, r.v2, l3};
}

__device__ __forceinline__ void store_vtx6(uint32_t* dst, const Vtx& a) {
    uint2* d = reinterpret_cast<uint2*>(dst);
    d[0] = make_uint2(__float_as_uint(a.x), __float_as_uint(a.y));
    d[1] = make_uint2(__float_as_uint(a.z), __float_as_uint(a.u));
    d[2] = make_uint2(__float_as_uint(a.v), a.l);
}

struct ArenaCaps {
    unsigned long long c[4];
};

struct GenericOut {
    uint32_t kept_f, kept_i;
    float mn[3], mx[3];
    bool has_pts;
};

// One voxel of a non-cube model (X-sprite, aabb, custom), kept out of line so
// that the cube fast path stays lean.
__device__ __noinline__ void generic_unit(const DevWorld& W, const Tile& T, const ChunkMeta& m,
                                          uint32_t vv, uint32_t info, int x, int y, int z, int pass,
                                          uint32_t foff, uint32_t ioff, uint32_t capacity,
                                          uint32_t* vtx, int32_t* idx, vx_sort_entry* ent, float wx,
                                          float wz, GenericOut& out) {
    Writer wr;
    wr.fpos = foff;
    wr.dead = false;
    wr.has_pts = false;
    out.kept_f = out.kept_i = 0;
    out.has_pts = false;
    if (!pass) {
        wr.vtx = vtx;
        wr.idx = idx;
        wr.ipos = ioff;
        wr.cap = capacity;
        wr.transl = false;
        draw_voxel(wr, W, T, vv, info, x, y, z);
        if (wr.fpos > foff) {  // the cursor only passes kept primitives
            out.kept_f = wr.fpos;
            out.kept_i = wr.ipos;
        }
    } else {
        wr.vtx = vtx;
        wr.idx = nullptr;
        wr.ipos = 0;
        wr.cap = 0;
        wr.transl = true;
        wr.wx = wx;
        wr.wz = wz;
        draw_voxel(wr, W, T, vv, info, x, y, z);
        if (wr.has_pts) {
            out.has_pts = true;
            for (int j = 0; j < 3; j++) {
                out.mn[j] = wr.mn[j];
                out.mx[j] = wr.mx[j];
            }
        }
        if (wr.fpos != foff) {
            vx_sort_entry e;
            e.position[0] = x + m.cx * CW + 0.5f;
            e.position[1] = y + 0.5f;
            e.position[2] = z + m.cz * CD + 0.5f;
            e.first_float = foff;
            e.n_floats = wr.fpos - foff;
            ent[ioff] = e;
        }
    }
}

// Everything process_unit() needs; the accumulators are per thread.
struct EmitCtx {
    const DevWorld* W;
    // the tile is rebuilt from these scalars where it is needed: the out-of-line generic path takes
    // its address, and one shared object would drag the cube path's state into local memory too
    const uint32_t* dflags;
    int own, px, pz, pxz;
    const int* hdr_pools;
    __device__ __forceinline__ Tile tile() const {
        return Tile {W, nullptr, nullptr, ty0, nullptr, dflags, cx, cz, own, px, pz, pxz, hdr_pools};
    }
    const EmitShared* E;
    const vx_voxel* vox;  // the chunk's voxels
    int cx, cz;
    int ty0;
    uint32_t* c_vtx;
    int32_t* c_idx;
    uint32_t* c_tfl;
    vx_sort_entry* c_ent;
    float wx, wz;
    uint32_t capacity;
    uint32_t kept_f, kept_i;
    float mn[3], mx[3];
    bool has_pts;
    // opaque cube faces are not stored by the thread that computed them: the four
    // vertices go to this lane's shared-memory slot and the warp writes all its faces
    // together, consecutive lanes to consecutive 8-byte pieces
    uint2* stage;
    bool st_valid;
    uint32_t* st_vptr;
    int32_t* st_iptr;
    int st_b;

    __device__ __forceinline__ void add_pt(float px, float py, float pz) {
        if (!has_pts) {
            has_pts = true;
            mn[0] = mx[0] = px;
            mn[1] = mx[1] = py;
            mn[2] = mx[2] = pz;
        } else {
            mn[0] = fminf(mn[0], px);
            mx[0] = fmaxf(mx[0], px);
            mn[1] = fminf(mn[1], py);
            mx[1] = fmaxf(mx[1], py);
            mn[2] = fminf(mn[2], pz);
            mx[2] = fmaxf(mx[2], pz);
        }
    }
};

// Emits one unit (a cube face, or a whole voxel of another model) at its final offsets.
// MODE selects which of the three kinds this instantiation handles (the lists are split by kind):
// 0 opaque cube faces, 1 whole voxels of other models, 2 translucent cube faces.
template <int MODE>
__device__ __forceinline__ void process_unit(EmitCtx& C, const Unit un) {
    const DevWorld& W = *C.W;
    const int pass = (un.code & U_TRANSL) ? 1 : 0;
    const int vi = un.code & 0xFFFu;
    const int x = vi & 15, z = (vi >> 4) & 15, y = C.ty0 + (vi >> 8);
    const uint32_t vv = C.vox[y * 256 + z * 16 + x] & 0x00FFFFFFu;  // id | rotation, segment
    const uint32_t info = 0;  // only blockCube reads the face mask, and cubes are never generic units
    if constexpr (MODE == 0) {
        // opaque cube face: staged for the warp's cooperative store
        const int fk = (un.code >> 12) & 7;
        if (un.foff + 4 * VSZ > C.capacity) return;  // overflow (:124)
        Vtx v[4];
        const Tile T = C.tile();
        cube_face(W, T, *C.E, vv, x, y, z, fk, v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            C.stage[3 * j + 0] = make_uint2(__float_as_uint(v[j].x), __float_as_uint(v[j].y));
            C.stage[3 * j + 1] = make_uint2(__float_as_uint(v[j].z), __float_as_uint(v[j].u));
            C.stage[3 * j + 2] = make_uint2(__float_as_uint(v[j].v), v[j].l);
        }
        C.st_valid = true;
        C.st_vptr = C.c_vtx + un.foff;
        C.st_iptr = C.c_idx + un.ioff;
        C.st_b = (int)(un.foff / VSZ);
        C.kept_f = max(C.kept_f, un.foff + 4 * VSZ);
        C.kept_i = max(C.kept_i, un.ioff + 6);
    } else if constexpr (MODE == 2) {
        // translucent cube face (water, glass, ice): six de-indexed world-space vertices, staged for
        // the warp's cooperative store, + the voxel's sort entry
        const int fk = (un.code >> 12) & 7;
        Vtx v[4];
        const Tile T = C.tile();
        cube_face(W, T, *C.E, vv, x, y, z, fk, v);
#pragma unroll
        for (int j = 0; j < 4; j++) C.add_pt(v[j].x, v[j].y, v[j].z);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j].x = v[j].x + C.wx;
            v[j].y = v[j].y + 0.5f;
            v[j].z = v[j].z + C.wz;
        }
        {
            const int order[6] = {0, 1, 2, 0, 2, 3};  // (:560-572)
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const Vtx& a = v[order[j]];
                C.stage[3 * j + 0] = make_uint2(__float_as_uint(a.x), __float_as_uint(a.y));
                C.stage[3 * j + 1] = make_uint2(__float_as_uint(a.z), __float_as_uint(a.u));
                C.stage[3 * j + 2] = make_uint2(__float_as_uint(a.v), a.l);
            }
        }
        C.st_valid = true;
        C.st_vptr = C.c_tfl + un.foff;
        if (un.code & U_FIRST) {
            vx_sort_entry e;
            e.position[0] = x + C.cx * CW + 0.5f;
            e.position[1] = y + 0.5f;
            e.position[2] = z + C.cz * CD + 0.5f;
            e.first_float = un.foff;
            e.n_floats = ((un.code >> U_NFACES_SHIFT) & 7u) * 6 * VSZ;
            C.c_ent[un.ioff] = e;
        }
    } else if (!(un.code & U_GENERIC)) {
        // translucent cube face in a build that lists them with the slow kinds (few of them: a
        // terrain world's water surface): six de-indexed world-space vertices + the voxel's sort entry
        const int fk = (un.code >> 12) & 7;
        Vtx v[4];
        const Tile T = C.tile();
        cube_face(W, T, *C.E, vv, x, y, z, fk, v);
#pragma unroll
        for (int j = 0; j < 4; j++) C.add_pt(v[j].x, v[j].y, v[j].z);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j].x = v[j].x + C.wx;
            v[j].y = v[j].y + 0.5f;
            v[j].z = v[j].z + C.wz;
        }
        uint32_t* dst = C.c_tfl + un.foff;
        store_vtx6(dst, v[0]);
        store_vtx6(dst + VSZ, v[1]);
        store_vtx6(dst + 2 * VSZ, v[2]);
        store_vtx6(dst + 3 * VSZ, v[0]);
        store_vtx6(dst + 4 * VSZ, v[2]);
        store_vtx6(dst + 5 * VSZ, v[3]);
        if (un.code & U_FIRST) {
            vx_sort_entry e;
            e.position[0] = x + C.cx * CW + 0.5f;
            e.position[1] = y + 0.5f;
            e.position[2] = z + C.cz * CD + 0.5f;
            e.first_float = un.foff;
            e.n_floats = ((un.code >> U_NFACES_SHIFT) & 7u) * 6 * VSZ;
            C.c_ent[un.ioff] = e;
        }
    } else {
        // the out-of-line path takes addresses: give it copies, so that the cube path's
        // state is not forced into local memory (that cost 50 M local-store sectors per launch)
        GenericOut go;
        const Tile tg = C.tile();
        ChunkMeta mg;
        mg.cx = C.cx;
        mg.cz = C.cz;
        generic_unit(W, tg, mg, vv, info, x, y, z, pass, un.foff, un.ioff, C.capacity,
                     pass ? C.c_tfl : C.c_vtx, C.c_idx, C.c_ent, C.wx, C.wz, go);
        if (!pass) {
            if (go.kept_f) {
                C.kept_f = max(C.kept_f, go.kept_f);
                C.kept_i = max(C.kept_i, go.kept_i);
            }
        } else if (go.has_pts) {
            C.add_pt(go.mn[0], go.mn[1], go.mn[2]);
            C.add_pt(go.mx[0], go.mx[1], go.mx[2]);
        }
    }
}

// Writes the units of this thread's row that belong to category (rank r, pass)
// to dst[0..]; `fo`, `io` are the row's float / index offsets relative to the
// tile's base for that category.
template <bool SPLIT>
__device__ __forceinline__ void write_row_units(const DevWorld& W, const TileShared& S, Unit* fast, Unit* slow,
                                                Unit* tcube, uint32_t r, int pass, uint32_t fo, uint32_t io,
                                                uint32_t where, uint32_t drawn) {
    // translucent cube faces: a list (and a pass) of their own when SPLIT, with the slow kinds otherwise
    Unit*& cube_dst = pass ? (SPLIT ? tcube : slow) : fast;
    const uint32_t want_tr = pass ? VI_TRANSL : 0u;
    const uint32_t common = (pass ? U_TRANSL : 0u) | (r << U_RANK_SHIFT);
    while (drawn) {  // the row's drawn voxels in x order
        const int k = __ffs(drawn) - 1;
        drawn &= drawn - 1;
        const int vi = row_voxel(threadIdx.x, k);
        const uint32_t info = S.info[vi];
        if ((info & VI_RANK) != r || (info & VI_TRANSL) != want_tr) continue;
        const uint32_t mask = (info >> VI_MASK_SHIFT) & 0x3Fu;
        // a simple cube's record is complete without its block id (the ids are not even staged
        // when the tile holds nothing else)
        const bool simple = (S.rowsc[((threadIdx.x >> 4) + 1) * TW + (threadIdx.x & 15) + 1] >> (k + 1)) & 1u;
        uint32_t fl, ix, model = VX_MODEL_BLOCK;
        if (simple) {
            fl = (uint32_t)__popc(mask) * 4 * VSZ;
            ix = (uint32_t)__popc(mask) * 6;
        } else {
            const uint32_t id = own_id(S, vi);
            voxel_counts(W, info, id, fl, ix);
            model = (flags_of(W, S.dflags, id) >> DF_MODEL_SHIFT) & 7u;
        }
        if (model == VX_MODEL_BLOCK) {
            const uint32_t nf = __popc(mask);
            uint32_t mk = mask, j = 0;
            while (mk) {
                const uint32_t fk = __ffs(mk) - 1;
                mk &= mk - 1;
                Unit un;
                un.code = (uint32_t)vi | (fk << 12) | (j == 0 ? U_FIRST : 0u) | common |
                          (nf << U_NFACES_SHIFT);
                un.foff = fo + j * (pass ? 6 * VSZ : 4 * VSZ);
                un.ioff = pass ? io : io + j * 6;
                un.where = where;
                *reinterpret_cast<uint4*>(cube_dst++) = *reinterpret_cast<const uint4*>(&un);
                j++;
            }
        } else {
            Unit un;
            un.code = (uint32_t)vi | U_GENERIC | U_FIRST | common;
            un.foff = fo;
            un.ioff = io;
            un.where = where;
            *reinterpret_cast<uint4*>(slow++) = *reinterpret_cast<const uint4*>(&un);
        }
        fo += pass ? ix * VSZ : fl;
        io += pass ? (fl ? 1u : 0u) : ix;
    }
}

__device__ __forceinline__ void init_emit_shared(EmitShared& E) {
    const int tid = threadIdx.x;
    if (tid < 16) E.lut15[tid] = __fdiv_rn((float)tid, 15.0f);
    if (tid < 6) {
        const V3 ax[6] = {v3(1, 0, 0), v3(-1, 0, 0), v3(0, 1, 0), v3(0, -1, 0), v3(0, 0, 1), v3(0, 0, -1)};
        float d = dot3(normalize3(ax[tid]), sun());
        E.dtab[tid] = 0.7f + d * 0.3f;
    }
    if (tid == 0) {
        E.kept_f = E.kept_i = 0;
        E.has_pts = 0;
        for (int j = 0; j < 3; j++) {
            E.mn[j] = 0xFFFFFFFFu;
            E.mx[j] = 0;
        }
    }
}

// k_mesh_lm: grid (8, chunks of the batch and their neighbours).  The light the
// mesher reads (Chunks::getVoxels backlight applied, zero where isOpenForLight is
// false, BlocksRenderer.cpp:385-411) as one u16 plane per chunk, for the layers
// any unit of the 3x3 neighbourhood can sample.  A pure stream: k_mesh_emit then
// needs one 2-byte load per light cell, mostly from L2.
__global__ void __launch_bounds__(256) k_mesh_lm(DevWorld W, const int32_t* pools, int air_lp) {
    const int p = pools[blockIdx.y];
    __shared__ uint32_t s_dflags[256];
    __shared__ int s_lo, s_hi;
    const int tid = threadIdx.x;
    s_dflags[tid] = tid < W.n_defs ? __ldg(&W.defs[tid].flags) : 0u;
    if (tid == 0) {
        s_lo = CH;
        s_hi = 0;
    }
    __syncthreads();
    if (tid < 9) {
        const ChunkMeta& m = W.meta[p];
        const int q = W.pool_of(m.cx + tid % 3 - 1, m.cz + tid / 3 - 1);
        if (q >= 0) {
            atomicMin(&s_lo, W.meta[q].bottom);
            atomicMax(&s_hi, W.meta[q].top);
        }
    }
    __syncthreads();
    const int lo = max(0, s_lo - 2), hi = min(CH, s_hi + 2);  // layers [lo, hi)
    const uint4* vox4 = reinterpret_cast<const uint4*>(W.vox_of(p));
    const uint4* l4 = reinterpret_cast<const uint4*>(W.light_of(p));
    uint4* o4 = reinterpret_cast<uint4*>(W.lm_of(p));
    // 8 voxels per thread: two uint4 of voxels, one uint4 of lights -> one uint4 of lm
    for (int u = blockIdx.x * 1024 + tid; u < (blockIdx.x + 1) * 1024; u += 256) {
        const int y = u >> 5;
        if (y < lo || y >= hi) continue;
        const uint4 l = l4[u];
        const uint32_t lv[8] = {l.x & 0xFFFFu, l.x >> 16, l.y & 0xFFFFu, l.y >> 16,
                                l.z & 0xFFFFu, l.z >> 16, l.w & 0xFFFFu, l.w >> 16};
        uint32_t o[8];
        if (air_lp) {
            // air lets light through (core:air, core_defs.cpp:13-20), so "lightPassing or air" is
            // the lp plane: 1 bit per voxel instead of its 4-byte record
            const uint32_t lpb = (__ldg(W.lp_of(p) + (u >> 2)) >> ((u & 3) * 8)) & 0xFFu;
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = ((lpb >> j) & 1u) ? (W.backlight ? backlight(lv[j]) : lv[j]) : 0u;
        } else {
            const uint4 a = vox4[2 * u], b = vox4[2 * u + 1];
            const uint32_t id[8] = {a.x & 0xFFFFu, a.y & 0xFFFFu, a.z & 0xFFFFu, a.w & 0xFFFFu,
                                    b.x & 0xFFFFu, b.y & 0xFFFFu, b.z & 0xFFFFu, b.w & 0xFFFFu};
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = mesher_light(W, flags_of(W, s_dflags, id[j]), id[j], lv[j]);
        }
        o4[u] = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
    }
}

// k_mesh_classify: grid (NTILE, n).  Stages the tile's block ids, classifies
// its 4096 voxels, writes the per (tile, draw group) totals that k_mesh_scan
// turns into offsets, and lists the tile's emission units (with offsets
// relative to the tile's base for their category) in a global unit list whose
// slice is reserved with one atomicAdd.
template <bool SPLIT>
#ifndef VX_CLASSIFY_MINB
#define VX_CLASSIFY_MINB 6
#endif
__global__ void __launch_bounds__(256, VX_CLASSIFY_MINB) k_mesh_classify(const __grid_constant__ DevWorld W, const int32_t* pools,
                                                       uint32_t* tile_cnt, Unit* ulist,
                                                       unsigned long long* cursors,
                                                       unsigned long long ucap, Unit* tlist, unsigned long long tcap) {
    const int tid = threadIdx.x;
    const int p = pools[blockIdx.y];
    ChunkMeta m;
    memset(&m, 0, sizeof(m));
    if (p >= 0) m = W.meta[p];
    // BlocksRenderer::build scans [bottom, top) only (:625-638)
    if (p < 0 || (int)blockIdx.x * TILE_H >= m.top || (int)(blockIdx.x + 1) * TILE_H <= m.bottom) return;
    __shared__ TileShared S;
    __shared__ uint32_t s_ubase, s_sbase, s_tbase, s_ok;
    RowSummary rs;
    if (!stage_and_classify<false, SPLIT>(W, S, m, blockIdx.x, rs)) return;
    // ---- units of the whole tile -> one slice of the fast list (opaque cube faces: the front of
    // `ulist`, growing up) and one of the slow list (everything else: the back, growing down), so
    // that each emission kernel walks a dense range of its own kind
    unsigned long long mine_f = 0, mine_s = 0, mine_t = 0;
    if (!rs.multi) {
        mine_f = rs.units[0] + rs.units[1];
        mine_s = (rs.slow[0] & 0xFFFFu) + (rs.slow[1] & 0xFFFFu);
        mine_t = (rs.slow[0] >> 16) + (rs.slow[1] >> 16);
    } else {
        for (int k = 0; k < 16; k++) {
            const int vi = row_voxel(tid, k);
            const uint32_t info = S.info[vi];
            if (!(info & VI_DRAW)) continue;
            uint32_t c, nu, f, i;
            voxel_contrib<SPLIT>(W, S.dflags, info, own_id(S, vi), c, nu, f, i);
            mine_f += nu & 0xFFu;
            mine_s += (nu >> 8) & 0xFFu;
            mine_t += nu >> 16;
        }
    }
    unsigned long long tot_f_all, tot_s_all;
    block_scan2(S, mine_f | (mine_t << 32), mine_s, &tot_f_all, &tot_s_all);
    const unsigned long long tot_t_all = tot_f_all >> 32;
    tot_f_all &= 0xFFFFFFFFull;
    if (tot_f_all + tot_s_all + tot_t_all == 0) return;  // uniform: nothing to draw in this tile
    if (tid == 0) {
        // cursors[5] / [4] / [7]: fast / slow / translucent-cube units listed so far.  The fast and the
        // slow slices share one list (front, growing up / back, growing down) and stay inside it; the
        // two regions can only overlap in an attempt whose totals exceed the list, and such an
        // attempt is thrown away by the host (and skipped by the emission kernels).  Translucent
        // cube faces have a list of their own.
        // (a tile that holds nothing of a kind does not touch that kind's cursor)
        const unsigned long long uf = tot_f_all ? atomicAdd(&cursors[5], tot_f_all) : 0ull;
        const unsigned long long us = tot_s_all ? atomicAdd(&cursors[4], tot_s_all) : 0ull;
        const unsigned long long ut = (SPLIT && tot_t_all) ? atomicAdd(&cursors[7], tot_t_all) : 0ull;
        const bool ok = uf + tot_f_all <= ucap && us + tot_s_all <= ucap && ut + tot_t_all <= tcap;
        if (!ok) atomicExch(&cursors[6], 1ull);
        s_ubase = (uint32_t)uf;
        s_sbase = ok ? (uint32_t)(ucap - us - tot_s_all) : 0u;
        s_tbase = (uint32_t)ut;
        s_ok = ok ? 1u : 0u;
    }
    __syncthreads();
    Unit* const tile_fast = ulist + s_ubase;
    Unit* const tile_slow = ulist + s_sbase;
    Unit* const tile_t = tlist + s_tbase;
    const bool ok = s_ok != 0;
    uint32_t* out = tile_cnt + ((size_t)blockIdx.y * NTILE + blockIdx.x) * W.n_ranks * 4;
    uint32_t frun = 0, srun = 0, trun = 0;
    for (int rw = 0; rw < 8; rw++) {
        uint32_t rm = S.rankmask[rw];
        while (rm) {
            const uint32_t r = rw * 32 + __ffs(rm) - 1;
            rm &= rm - 1;
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {  // 0: opaque, 1: translucent
                if (!((S.passmask[pass][rw] >> (r & 31u)) & 1u)) continue;  // uniform
                const RowTot t = row_total<SPLIT>(W, S, rs, r | (pass ? 0x100u : 0u));
                unsigned long long ex_fi, ex_u, tot_fi, tot_u;
                block_scan_pair(S, t.fi, t.units, &ex_fi, &ex_u, &tot_fi, &tot_u);
                if (tid == 0) {
                    out[r * 4 + pass * 2 + 0] = (uint32_t)(tot_fi >> 32);
                    out[r * 4 + pass * 2 + 1] = (uint32_t)tot_fi;
                }
                if (tot_u == 0) continue;  // uniform
                constexpr unsigned long long M21 = (1ull << 21) - 1;
                if (ok && t.units)
                    write_row_units<SPLIT>(W, S, tile_fast + frun + (uint32_t)(ex_u & M21),
                                    tile_slow + srun + (uint32_t)((ex_u >> 21) & M21),
                                    tile_t + trun + (uint32_t)(ex_u >> 42),
                                    r, pass, (uint32_t)(ex_fi >> 32), (uint32_t)ex_fi,
                                    blockIdx.y * NTILE + blockIdx.x, rs.drawn);
                frun += (uint32_t)(tot_u & M21);
                srun += (uint32_t)((tot_u >> 21) & M21);
                trun += (uint32_t)(tot_u >> 42);
            }
        }
    }
}

// What one thread of k_mesh_emit needs to know about its chunk, in one 64-byte record
// (four independent 16-byte loads right after the unit itself).
struct __align__(16) EmitHdr {
    int32_t pools[9];  // 3x3 neighbourhood, [dz + 1][dx + 1]
    int32_t cx, cz;
    uint32_t v_off, i_off, tf_off, te_off;
    uint32_t over;     // tot_floats > capacity: kept sizes matter
};
static_assert(sizeof(EmitHdr) == 64, "EmitHdr is four uint4");

// Emission: one thread per emission unit of the whole batch (flat grids over
// the unit lists; k_mesh_emit_cubes / k_mesh_emit_generic below).  A thread computes a cube face (4 vertices from the 3x3
// light cells of the face plane, read straight from HBM / L2) or a whole voxel
// of another model, and writes it at its final offset.  No staging and no
// barrier after the prologue: the kernel is a stream of independent units.
template <int MODE>
__device__ __forceinline__ void emit_body(const DevWorld& W, const EmitHdr* hdrs,
                                          const uint32_t* tile_base, const Unit* ulist,
                                          MeshChunkOut* outs, uint32_t* vertices,
                                          int32_t* indices, uint32_t* tfloats,
                                          vx_sort_entry* entries, uint32_t capacity,
                                          const unsigned long long* totals, const ArenaCaps& caps,
                                          unsigned long long ucap, const Unit* tlist, unsigned long long tcap,
                                          EmitShared& E, uint32_t* s_dflags, int& s_skip, uint2* stage_raw,
                                          unsigned cta) {
    const int tid = threadIdx.x;
    constexpr int UPT = MODE == 1 ? EMIT_UPT_GENERIC : EMIT_UPT;
    constexpr int SROW = MODE == 2 ? STAGE_ROW_T : STAGE_ROW;  // per warp: 32 faces of 12 (opaque) / 18 (translucent) uint2 + 1 pad: conflict-free rows
    // arenas or unit list too small for this batch: emit nothing, the host grows them and launches
    // again (totals[] is {vertex floats, indices, entry floats, entries, slow units, fast units});
    // the grid is sized for more than is listed: a block beyond the listed units leaves before it sets
    // anything up
    const unsigned long long n_mine = totals[MODE == 1 ? 4 : MODE == 0 ? 5 : 7];  // units of this instantiation's kind
    if (totals[0] > caps.c[0] || totals[1] > caps.c[1] || totals[2] > caps.c[2] || totals[3] > caps.c[3] ||
        totals[4] + totals[5] > ucap || totals[7] > tcap || (unsigned long long)cta * (256 * UPT) >= n_mine)
        return;
    (void)s_skip;
    init_emit_shared(E);
    s_dflags[tid] = tid < W.n_defs ? __ldg(&W.defs[tid].flags) : 0u;
    __syncthreads();
    const int lane = tid & 31;
    uint2* const wstage = stage_raw + (MODE == 1 ? 0 : (tid >> 5) * 32 * SROW);
    // the fast kinds (opaque cube faces) fill the list from the front, the slow kinds from the back:
    // each instantiation walks a dense range of its own units
    const unsigned long long cta_base = (unsigned long long)cta * (256 * UPT);
    const unsigned long long list_base = MODE == 1 ? ucap - n_mine : 0ull;
    const Unit* const my_list = MODE == 2 ? tlist : ulist;
    const int n_rounds = UPT;
    // EMIT_UPT units per thread: the table set-up above is paid once per 1024 units
#pragma unroll 1
    for (int it = 0; it < n_rounds; it++) {
    bool st_valid = false;
    uint32_t* st_vptr = nullptr;
    int32_t* st_iptr = nullptr;
    int st_b = 0;
    int a_c = -1;                      // chunk whose translucent AABB this lane extends
    uint32_t a_mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, a_mx[3] = {0, 0, 0};
    int k_c = -1;                      // chunk (one that hit the capacity cut-off) whose kept sizes this lane extends
    uint32_t k_f = 0, k_i = 0;
    const unsigned long long u = cta_base + it * 256 + tid;
    if (u < n_mine) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(my_list) + list_base + u);
    // (opaque cube faces are emitted by one instantiation of this kernel, everything else — other
    // models, translucent faces — by the other: the generic path is an out-of-line call, and a call
    // in the cube path's body costs it spills.)
    // The capacity cut-off (BlocksRenderer.cpp:84,124,162,292) keeps a primitive iff it ends at or
    // below `capacity` floats; offsets grow monotonically, so an opaque unit that *starts* at or
    // beyond it keeps nothing.  In a chunk that overflows (configs[3]: 33 k faces kept of several
    // hundred thousand) nearly every unit is such a unit: drop it on its first loads.
    const uint32_t r = (q.x >> U_RANK_SHIFT) & 0xFFu;
    const int ps = (q.x & U_TRANSL) ? 2 : 0;
    const uint32_t* base = tile_base + (size_t)q.w * W.n_ranks * 4;
    {
    const int c = q.w >> 4, tile = q.w & 15;
    // ---- everything about the chunk in four independent loads (plus the category base)
    const uint4* hp = reinterpret_cast<const uint4*>(hdrs + c);
    const uint4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2), h3 = __ldg(hp + 3);
    const uint32_t bf = __ldg(base + r * 4 + ps), bi = __ldg(base + r * 4 + ps + 1);
    if ((q.x & U_TRANSL) || (unsigned long long)q.y + bf < capacity) {
    // pools[dz + 1][dx + 1]: h0 = {0,1,2,3}, h1 = {4,5,6,7}, h2 = {8, cx, cz, v_off}, h3 = {i_off, tf_off, te_off, over}
    const int own = (int)h1.x;
    const int vx = q.x & 15, vz = (q.x >> 4) & 15;
    const int xm = (int)h0.w, xp = (int)h1.y, zm = (int)h0.y, zp = (int)h1.w;          // (-1,0) (+1,0) (0,-1) (0,+1)
    const int mm = (int)h0.x, pm = (int)h0.z, mp = (int)h1.z, pp = (int)h2.x;          // (dx,dz): (-,-) (+,-) (-,+) (+,+)
    const int px = vx == 0 ? xm : xp, pz = vz == 0 ? zm : zp;  // only used when the voxel touches that border
    const int pxz = vx == 0 ? (vz == 0 ? mm : mp) : (vz == 0 ? pm : pp);
    EmitCtx C;
    C.W = &W; C.E = &E; C.vox = W.vox_of(own); C.ty0 = tile * TILE_H;
    C.dflags = s_dflags;
    C.own = own; C.px = px; C.pz = pz; C.pxz = pxz;
    C.hdr_pools = reinterpret_cast<const int*>(hdrs + c);
    C.cx = (int)h2.y;
    C.cz = (int)h2.z;
    C.c_vtx = vertices + h2.w;
    C.c_idx = indices + h3.x;
    C.c_tfl = tfloats + h3.y;
    C.c_ent = entries + h3.z;
    C.wx = C.cx * CW + 0.5f;
    C.wz = C.cz * CD + 0.5f;
    C.capacity = capacity;
    C.kept_f = C.kept_i = 0;
    C.has_pts = false;
    C.stage = wstage + lane * SROW;
    C.st_valid = false;
    C.st_vptr = nullptr;
    C.st_iptr = nullptr;
    C.st_b = 0;
    Unit un;
    un.code = q.x;
    un.foff = q.y + bf;
    un.ioff = q.z + bi;
    process_unit<MODE>(C, un);
    st_valid = C.st_valid;
    st_vptr = C.st_vptr;
    st_iptr = C.st_iptr;
    st_b = C.st_b;
    // kept sizes only matter for chunks that hit the capacity cut-off (k_mesh_final
    // takes the totals otherwise); the translucent AABB feeds the merge rule
    // (every lane of a warp usually works on the same chunk: the warp folds these before it
    // touches the chunk's record — in a chunk that overflows every kept face would otherwise
    // be two atomics on the same two words)
    if (C.kept_f && h3.w) {
        k_c = c;
        k_f = C.kept_f;
        k_i = C.kept_i;
    }
    if (C.has_pts) {
        a_c = c;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            a_mn[j] = f2o(C.mn[j]);
            a_mx[j] = f2o(C.mx[j]);
        }
    }
    }
    }
    }
    {
        // ---- kept sizes of a chunk that hit the capacity cut-off
        const unsigned have = __ballot_sync(0xFFFFFFFFu, k_c >= 0);
        if (have) {
            const int c0 = __shfl_sync(0xFFFFFFFFu, k_c, __ffs(have) - 1);
            if (__all_sync(0xFFFFFFFFu, k_c < 0 || k_c == c0)) {
                uint32_t mf = k_f, mi = k_i;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    mf = max(mf, __shfl_xor_sync(0xFFFFFFFFu, mf, o));
                    mi = max(mi, __shfl_xor_sync(0xFFFFFFFFu, mi, o));
                }
                if (lane == 0) {
                    atomicMax(&outs[c0].kept_floats, mf);
                    atomicMax(&outs[c0].kept_idx, mi);
                }
            } else if (k_c >= 0) {
                atomicMax(&outs[k_c].kept_floats, k_f);
                atomicMax(&outs[k_c].kept_idx, k_i);
            }
        }
    }
    if constexpr (MODE != 0) {
        // ---- translucent AABB of the chunk (feeds the merge rule of k_mesh_final)
        const unsigned have = __ballot_sync(0xFFFFFFFFu, a_c >= 0);
        if (have) {
            const int c0 = __shfl_sync(0xFFFFFFFFu, a_c, __ffs(have) - 1);
            if (__all_sync(0xFFFFFFFFu, a_c < 0 || a_c == c0)) {
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    uint32_t mn = a_mn[j], mx = a_mx[j];
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
                        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
                    }
                    if (lane == 0) {
                        atomicMin(&outs[c0].aabb_min[j], mn);
                        atomicMax(&outs[c0].aabb_max[j], mx);
                    }
                }
            } else if (a_c >= 0) {
                for (int j = 0; j < 3; j++) {
                    atomicMin(&outs[a_c].aabb_min[j], a_mn[j]);
                    atomicMax(&outs[a_c].aabb_max[j], a_mx[j]);
                }
            }
        }
    }
    if constexpr (MODE == 1) continue;
    if constexpr (MODE == 2) {
        // ---- the warp stores its staged faces: 32 x 144 bytes, lane i of step j writes piece 32 j + i
        __syncwarp();
        const unsigned tmask = __ballot_sync(0xFFFFFFFFu, st_valid);
        if (tmask) {
#pragma unroll
            for (int j = 0; j < 18; j++) {
                const int i = j * 32 + lane, f = i / 18, part = i - f * 18;
                const unsigned long long vp = __shfl_sync(0xFFFFFFFFu, (unsigned long long)st_vptr, f);
                if ((tmask >> f) & 1u) reinterpret_cast<uint2*>(vp)[part] = wstage[f * SROW + part];
            }
        }
        __syncwarp();
        continue;
    }
    // ---- the warp stores its staged faces: lane i of step j writes the 8-byte piece
    // j * 32 + i of the warp's 32 x 96 bytes, so faces that are adjacent in the arena
    // (the usual case) leave the SM as full 32-byte sectors
    __syncwarp();
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, st_valid);
    if (mask) {
        // piece i = 32 j + lane belongs to face i / 12; stepping j adds 32 = 2 * 12 + 8
        int f = lane / 12, part = lane - f * 12;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const unsigned long long vp = __shfl_sync(0xFFFFFFFFu, (unsigned long long)st_vptr, f);
            if ((mask >> f) & 1u) reinterpret_cast<uint2*>(vp)[part] = wstage[f * STAGE_ROW + part];
            part += 8;
            const int wrap = part >= 12;
            part -= wrap ? 12 : 0;
            f += 2 + wrap;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int i = j * 32 + lane, f = i / 3, part = i - f * 3;
            const unsigned long long ip = __shfl_sync(0xFFFFFFFFu, (unsigned long long)st_iptr, f);
            const int b = __shfl_sync(0xFFFFFFFFu, st_b, f);  // first vertex of the face within its chunk
            if ((mask >> f) & 1u) {
                // indices 0,1,2,0,2,3 from there (:150)
                const uint2 o = part == 0 ? make_uint2(b, b + 1) : part == 1 ? make_uint2(b + 2, b) : make_uint2(b + 2, b + 3);
                if (!(ip & 7ull)) {
                    reinterpret_cast<uint2*>(ip)[part] = o;
                } else {  // an odd number of custom-model triangles earlier in the chunk
                    reinterpret_cast<uint32_t*>(ip)[2 * part] = o.x;
                    reinterpret_cast<uint32_t*>(ip)[2 * part + 1] = o.y;
                }
            }
        }
    }
    __syncwarp();  // the staging rows are reused by the next unit
    }
}

// Opaque cube faces (blocks [0, fast_ctas)) and translucent cube faces (the blocks after them) in
// one launch: the two kinds stage their vertices the same way, and the few translucent blocks of a
// terrain world ride in the tail of the opaque ones instead of paying a launch of their own.
template <bool WITH_T>
__global__ void __launch_bounds__(256, VX_EMIT_MINB) k_mesh_emit_cubes(const __grid_constant__ DevWorld W, const EmitHdr* hdrs,
                                                         const uint32_t* tile_base, const Unit* ulist,
                                                         MeshChunkOut* outs, uint32_t* vertices,
                                                         int32_t* indices, uint32_t* tfloats,
                                                         vx_sort_entry* entries, uint32_t capacity,
                                                         const unsigned long long* totals, ArenaCaps caps,
                                                         unsigned long long ucap, const Unit* tlist,
                                                         unsigned long long tcap, unsigned fast_ctas) {
    __shared__ EmitShared E;
    __shared__ uint32_t s_dflags[256];
    __shared__ int s_skip;
    __shared__ uint2 s_stage[8 * 32 * (WITH_T ? STAGE_ROW_T : STAGE_ROW)];
    if (!WITH_T || blockIdx.x < fast_ctas)
        emit_body<0>(W, hdrs, tile_base, ulist, outs, vertices, indices, tfloats, entries, capacity, totals, caps, ucap,
                     tlist, tcap, E, s_dflags, s_skip, s_stage, blockIdx.x);
    else if constexpr (WITH_T)
        emit_body<2>(W, hdrs, tile_base, ulist, outs, vertices, indices, tfloats, entries, capacity, totals, caps, ucap,
                     tlist, tcap, E, s_dflags, s_skip, s_stage, blockIdx.x - fast_ctas);
}

// Whole voxels of the other models (X-sprites, aabb, custom): few, long units.
__global__ void __launch_bounds__(256, 4) k_mesh_emit_generic(const __grid_constant__ DevWorld W, const EmitHdr* hdrs,
                                                           const uint32_t* tile_base, const Unit* ulist,
                                                           MeshChunkOut* outs, uint32_t* vertices,
                                                           int32_t* indices, uint32_t* tfloats,
                                                           vx_sort_entry* entries, uint32_t capacity,
                                                           const unsigned long long* totals, ArenaCaps caps,
                                                           unsigned long long ucap, const Unit* tlist,
                                                           unsigned long long tcap) {
    __shared__ EmitShared E;
    __shared__ uint32_t s_dflags[256];
    __shared__ int s_skip;
    __shared__ uint2 s_stage[32 * STAGE_ROW];
    emit_body<1>(W, hdrs, tile_base, ulist, outs, vertices, indices, tfloats, entries, capacity, totals, caps, ucap, tlist,
                 tcap, E, s_dflags, s_skip, s_stage, blockIdx.x);
}

// k_mesh_scan: one warp per chunk.  Turns the per (tile, group) counts into
// exclusive offsets in reference order (group-major, tile-minor) and fills the
// chunk's summary; slice sizes are what can survive the capacity cut-off.
__global__ void __launch_bounds__(32) k_mesh_scan(const int32_t* pools, uint32_t* tile_cnt,
                                                  MeshChunkOut* outs, int R, uint32_t capacity) {
    const int c = blockIdx.x, lane = threadIdx.x;
    uint32_t* cnt = tile_cnt + (size_t)c * NTILE * R * 4;
    uint32_t run[4] = {0, 0, 0, 0};
    for (int base = 0; base < R * NTILE; base += 32) {
        const int e = base + lane;
        const bool ok = e < R * NTILE;
        uint32_t* p = cnt + ((size_t)(e % NTILE) * R + e / NTILE) * 4;
        uint32_t v[4] = {0, 0, 0, 0};
        if (ok) {
            const uint4 q = *reinterpret_cast<const uint4*>(p);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
        uint32_t ex[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t inc = v[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += t;
            }
            ex[k] = run[k] + inc - v[k];
            run[k] += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (ok) *reinterpret_cast<uint4*>(p) = make_uint4(ex[0], ex[1], ex[2], ex[3]);
    }
    if (lane == 0) {
        MeshChunkOut o;
        memset(&o, 0, sizeof(o));
        o.pool = pools[c];
        o.cancelled = o.pool < 0;  // BlocksRenderer::build :617-622
        for (int j = 0; j < 3; j++) {
            o.aabb_min[j] = 0xFFFFFFFFu;
            o.aabb_max[j] = 0;
        }
        o.tot_floats = run[0];
        o.tot_idx = run[1];
        o.t_floats = run[2];
        o.t_entries = run[3];
        // slice sizes, parked in the offset fields until k_mesh_offsets scans them
        const uint32_t vf = min(run[0], capacity);
        const uint32_t ix = run[0] <= capacity ? run[1] : min(run[1], capacity / 4 + 8);
        o.v_off = (vf + 3u) & ~3u;
        o.i_off = (ix + 3u) & ~3u;
        o.tf_off = (run[2] + 3u) & ~3u;
        o.te_off = run[3];
        outs[c] = o;
    }
}

// k_mesh_offsets: exclusive scan of the slice sizes over the chunks of the
// batch (one block) -> arena offsets; totals[4] = arena sizes in elements.
__global__ void __launch_bounds__(1024) k_mesh_offsets(DevWorld W, MeshChunkOut* outs, int n,
                                                       unsigned long long* totals, EmitHdr* hdrs,
                                                       uint32_t capacity) {
    __shared__ unsigned long long wsum[32][4];
    __shared__ unsigned long long carry[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) carry[tid] = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int c = base + tid;
        unsigned long long v[4] = {0, 0, 0, 0};
        if (c < n) {
            v[0] = outs[c].v_off;
            v[1] = outs[c].i_off;
            v[2] = outs[c].tf_off;
            v[3] = outs[c].te_off;
        }
        unsigned long long inc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            inc[k] = v[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc[k], o);
                if (lane >= o) inc[k] += t;
            }
            if (lane == 31) wsum[warp][k] = inc[k];
        }
        __syncthreads();
        unsigned long long off[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned long long b = carry[k];
            for (int w = 0; w < warp; w++) b += wsum[w][k];
            off[k] = b + inc[k] - v[k];
        }
        if (c < n) {
            // offsets are 32-bit: the host rejects batches whose totals do not fit
            outs[c].v_off = (uint32_t)off[0];
            outs[c].i_off = (uint32_t)off[1];
            outs[c].tf_off = (uint32_t)off[2];
            outs[c].te_off = (uint32_t)off[3];
            EmitHdr h;
            const int p = outs[c].pool;
            h.cx = h.cz = 0;
            for (int k = 0; k < 9; k++) h.pools[k] = -1;
            if (p >= 0) {
                h.cx = W.meta[p].cx;
                h.cz = W.meta[p].cz;
                for (int k = 0; k < 9; k++) h.pools[k] = W.pool_of(h.cx + k % 3 - 1, h.cz + k / 3 - 1);
            }
            h.v_off = (uint32_t)off[0];
            h.i_off = (uint32_t)off[1];
            h.tf_off = (uint32_t)off[2];
            h.te_off = (uint32_t)off[3];
            h.over = outs[c].tot_floats > capacity;
            hdrs[c] = h;
        }
        __syncthreads();
        if (tid < 4) {
            unsigned long long b = carry[tid];
            for (int w = 0; w < 32; w++) b += wsum[w][tid];
            carry[tid] = b;
        }
        __syncthreads();
    }
    if (tid < 4) totals[tid] = carry[tid];
}

// "additional powerful optimization" of renderTranslucent (:588-606): when the
// AABB of all translucent vertices is thinner than 0.01 on an axis, the entries
// collapse into the first one.
__global__ void k_mesh_final(MeshChunkOut* outs, vx_sort_entry* entries, int n,
                             const unsigned long long* totals, ArenaCaps caps, uint32_t capacity) {
    if (totals[0] > caps.c[0] || totals[1] > caps.c[1] || totals[2] > caps.c[2] ||
        totals[3] > caps.c[3])
        return;  // the emit pass was skipped (see k_mesh_emit)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    MeshChunkOut* co = outs + i;
    co->n_entries_final = co->t_entries;
    if (co->tot_floats <= capacity) {  // nothing was cut off
        co->kept_floats = co->tot_floats;
        co->kept_idx = co->tot_idx;
    }
    if (co->cancelled || co->t_entries <= 1) return;
    const float sx = fabsf(o2f(co->aabb_max[0]) - o2f(co->aabb_min[0]));
    const float sy = fabsf(o2f(co->aabb_max[1]) - o2f(co->aabb_min[1]));
    const float sz = fabsf(o2f(co->aabb_max[2]) - o2f(co->aabb_min[2]));
    if (sy < 0.01f || sx < 0.01f || sz < 0.01f) {
        vx_sort_entry* e = entries + co->te_off;
        e->first_float = 0;
        e->n_floats = co->t_floats;
        co->n_entries_final = 1;
    }
}

template <class T>
int grow(vx_ctx* ctx, T** ptr, size_t* cap, size_t need) {
    if (need <= *cap && *ptr) return 0;
    size_t ncap = std::max<size_t>(need + need / 4, 1024);
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    VX_CUDA(cudaMalloc(reinterpret_cast<void**>(ptr), ncap * sizeof(T)));
    *cap = ncap;
    return 0;
}

void init_orient_table(Orient (*t)[8]) {
    // voxels/Block.cpp:57-92; CoordSystem ctor :39-45 (fix = -sum of negative axes)
    memset(t, 0, sizeof(Orient) * 3 * 8);
    auto set = [&](int prof, int v, int x0, int x1, int x2, int y0, int y1, int y2, int z0, int z1,
                   int z2) {
        Orient& o = t[prof][v];
        const int a[3][3] = {{x0, x1, x2}, {y0, y1, y2}, {z0, z1, z2}};
        for (int i = 0; i < 3; i++) {
            bool neg = false;
            for (int j = 0; j < 3; j++) {
                o.ax[i][j] = (int8_t)a[i][j];
                if (a[i][j] < 0) neg = true;
            }
            if (neg)
                for (int j = 0; j < 3; j++) o.fix[j] = (int8_t)(o.fix[j] - a[i][j]);
        }
    };
    for (int v = 0; v < 6; v++) set(VX_ROT_NONE, v, 1, 0, 0, 0, 1, 0, 0, 0, 1);
    set(VX_ROT_PIPE, 0, 1, 0, 0, 0, 0, 1, 0, -1, 0);
    set(VX_ROT_PIPE, 1, 0, 0, 1, -1, 0, 0, 0, -1, 0);
    set(VX_ROT_PIPE, 2, -1, 0, 0, 0, 0, -1, 0, -1, 0);
    set(VX_ROT_PIPE, 3, 0, 0, -1, 1, 0, 0, 0, -1, 0);
    set(VX_ROT_PIPE, 4, 1, 0, 0, 0, 1, 0, 0, 0, 1);
    set(VX_ROT_PIPE, 5, 1, 0, 0, 0, -1, 0, 0, 0, -1);
    set(VX_ROT_PANE, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1);
    set(VX_ROT_PANE, 1, 0, 0, -1, 0, 1, 0, 1, 0, 0);
    set(VX_ROT_PANE, 2, -1, 0, 0, 0, 1, 0, 0, 0, -1);
    set(VX_ROT_PANE, 3, 0, 0, 1, 0, 1, 0, -1, 0, 0);
}

}  // namespace

int vx_mesh_init(vx_ctx* ctx) {
    Orient t[3][8];
    init_orient_table(t);
    VX_CUDA(cudaMemcpyToSymbol(c_orient, t, sizeof(t)));
    return 0;
}

extern "C" {

int vx_mesh_build(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, uint64_t capacity) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    ctx->mesh_out.clear();
    ctx->mesh_keys.assign(keys, keys + n);
    ctx->mesh_capacity = capacity;
    if (n <= 0) return 0;
    if (capacity < 1024 || capacity > 0xFFFFFF00ull) {
        ctx->fail("vx_mesh_build: capacity out of range [1024, 2^32)");
        return -1;
    }
    cudaStream_t st = ctx->stream;
    const int R = ctx->W.n_ranks;
    // the slot look-ups below are host work during which the device idles: reuse them while the
    // same keys are meshed over an unchanged slot map (the usual frame-to-frame case)
    const bool same = ctx->mesh_prep_epoch == ctx->slot_epoch && ctx->mesh_prep_keys.size() == (size_t)n &&
                      std::memcmp(ctx->mesh_prep_keys.data(), keys, sizeof(vx_chunk_key) * n) == 0;
    if (!same) {
        ctx->mesh_prep_pools.resize(n);
        for (int i = 0; i < n; i++) ctx->mesh_prep_pools[i] = ctx->pool_of(keys[i].cx, keys[i].cz);
        // chunks whose light the batch can sample: the meshed ones and their 8 neighbours
        ctx->mesh_prep_lm_pools.clear();
        std::vector<uint8_t> mark(ctx->pool_cap, 0);
        for (int i = 0; i < n; i++) {
            if (ctx->mesh_prep_pools[i] < 0) continue;
            for (int dz = -1; dz <= 1; dz++)
                for (int dx = -1; dx <= 1; dx++) {
                    const int q = ctx->pool_of(keys[i].cx + dx, keys[i].cz + dz);
                    if (q >= 0 && !mark[q]) {
                        mark[q] = 1;
                        ctx->mesh_prep_lm_pools.push_back(q);
                    }
                }
        }
        std::sort(ctx->mesh_prep_lm_pools.begin(), ctx->mesh_prep_lm_pools.end());
        ctx->mesh_prep_keys.assign(keys, keys + n);
        ctx->mesh_prep_epoch = ctx->slot_epoch;
    }
    const std::vector<int>& pools = ctx->mesh_prep_pools;
    const std::vector<int>& lm_pools = ctx->mesh_prep_lm_pools;
    if (grow(ctx, &ctx->d_mesh_pools, &ctx->mesh_pools_cap, (size_t)n)) return -1;
    if (grow(ctx, &ctx->d_mesh_out, &ctx->mesh_out_cap, (size_t)n)) return -1;
    if (grow(ctx, &ctx->d_emit_hdr, &ctx->emit_hdr_cap, (size_t)n * 4)) return -1;
    if (grow(ctx, &ctx->d_tile_cnt, &ctx->tile_cnt_cap, (size_t)n * NTILE * R * 4)) return -1;
    if (ctx->mesh_pools_cache != pools) {
        VX_CUDA(cudaMemcpyAsync(ctx->d_mesh_pools, pools.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
        ctx->mesh_pools_cache = pools;
    }
    const size_t cnt_words = (size_t)n * NTILE * R * 4;
    VX_CUDA(cudaEventRecord(ctx->ev0, st));
    if (!lm_pools.empty()) {
        if (vx_materialize(ctx, lm_pools)) return -1;
        if (vx_upload_list(ctx, lm_pools, 4)) return -1;
        VX_LAUNCH(ctx, KID_LM, k_mesh_lm<<<dim3(8, (unsigned)lm_pools.size()), 256, 0, st>>>(
                                   ctx->W, ctx->d_list + 4 * (size_t)ctx->pool_cap, (int)ctx->defs[0].light_passing));
    }
    ctx->mesh_out.assign(n, MeshChunkOut {});
    // Everything is queued against the unit list and the arenas we already have;
    // only when the batch turns out not to fit (first call, or a bigger batch)
    // are they grown and the passes queued again.  A steady-state call has one
    // host synchronisation.
    // (a cold call takes three passes: the unit list grows, the arenas grow, the build runs; a fourth
    // covers a batch that outgrows both by different amounts)
    bool built = false;
    for (int attempt = 0; attempt < 5; attempt++) {
        ArenaCaps caps;
        caps.c[0] = ctx->vertices_cap;
        caps.c[1] = ctx->indices_cap;
        caps.c[2] = ctx->tfloats_cap;
        caps.c[3] = ctx->entries_cap;
        const bool split = ctx->split_transl;
        const unsigned long long ucap = ctx->units_cap, tcap = split ? ctx->units_t_cap : 0;
        unsigned long long slow_covered = 0;
        VX_CUDA(cudaMemsetAsync(ctx->d_tile_cnt, 0, cnt_words * 4, st));
        VX_CUDA(cudaMemsetAsync(ctx->d_mesh_totals, 0, 8 * sizeof(uint64_t), st));
        auto classify = split ? k_mesh_classify<true> : k_mesh_classify<false>;
        VX_LAUNCH(ctx, KID_CLASSIFY,
                  classify<<<dim3(NTILE, n), 256, 0, st>>>(
                      ctx->W, ctx->d_mesh_pools, ctx->d_tile_cnt, reinterpret_cast<Unit*>(ctx->d_units),
                      ctx->d_mesh_totals, ucap, reinterpret_cast<Unit*>(ctx->d_units_t), tcap));
        VX_LAUNCH(ctx, KID_SCAN, k_mesh_scan<<<n, 32, 0, st>>>(ctx->d_mesh_pools, ctx->d_tile_cnt,
                                                               ctx->d_mesh_out, R, (uint32_t)capacity));
        VX_LAUNCH(ctx, KID_OFFSETS,
                  k_mesh_offsets<<<1, 1024, 0, st>>>(ctx->W, ctx->d_mesh_out, n, ctx->d_mesh_totals,
                                                     reinterpret_cast<EmitHdr*>(ctx->d_emit_hdr), (uint32_t)capacity));
        {
            const unsigned fast_ctas = (unsigned)((ucap + 256 * EMIT_UPT - 1) / (256 * EMIT_UPT));
            const unsigned t_ctas = (unsigned)((tcap + 256 * EMIT_UPT - 1) / (256 * EMIT_UPT));
            auto emit_cubes = split ? k_mesh_emit_cubes<true> : k_mesh_emit_cubes<false>;
            if (fast_ctas + t_ctas)
                VX_LAUNCH(ctx, KID_EMIT, emit_cubes<<<fast_ctas + t_ctas, 256, 0, st>>>(
                              ctx->W, reinterpret_cast<const EmitHdr*>(ctx->d_emit_hdr), ctx->d_tile_cnt,
                              reinterpret_cast<const Unit*>(ctx->d_units), ctx->d_mesh_out,
                              reinterpret_cast<uint32_t*>(ctx->d_vertices), ctx->d_indices,
                              reinterpret_cast<uint32_t*>(ctx->d_tfloats), ctx->d_entries, (uint32_t)capacity,
                              ctx->d_mesh_totals, caps, ucap, reinterpret_cast<const Unit*>(ctx->d_units_t), tcap, fast_ctas));
            // (the slow kinds are a small part of the list: the grid covers what the last build listed, plus a
            // margin; a build that lists more than was covered is queued again, below)
            static const bool no_hint = getenv("VXGPU_NO_SLOW_HINT") != nullptr;
            slow_covered = (ctx->slow_units_hint && !no_hint) ? std::min<unsigned long long>(ucap, ctx->slow_units_hint) : ucap;
            if (slow_covered)
                VX_LAUNCH(ctx, KID_EMIT_GENERIC,
                          k_mesh_emit_generic<<<(unsigned)((slow_covered + 256 * EMIT_UPT_GENERIC - 1) / (256 * EMIT_UPT_GENERIC)), 256, 0, st>>>(
                              ctx->W, reinterpret_cast<const EmitHdr*>(ctx->d_emit_hdr), ctx->d_tile_cnt,
                              reinterpret_cast<const Unit*>(ctx->d_units), ctx->d_mesh_out,
                              reinterpret_cast<uint32_t*>(ctx->d_vertices), ctx->d_indices,
                              reinterpret_cast<uint32_t*>(ctx->d_tfloats), ctx->d_entries, (uint32_t)capacity,
                              ctx->d_mesh_totals, caps, ucap, reinterpret_cast<const Unit*>(ctx->d_units_t), tcap));
        }
        VX_LAUNCH(ctx, KID_FINAL,
                  k_mesh_final<<<(n + 127) / 128, 128, 0, st>>>(ctx->d_mesh_out, ctx->d_entries, n,
                                                                ctx->d_mesh_totals, caps, (uint32_t)capacity));
        VX_CUDA(cudaGetLastError());
        VX_CUDA(cudaMemcpyAsync(ctx->mesh_out.data(), ctx->d_mesh_out, n * sizeof(MeshChunkOut),
                                cudaMemcpyDeviceToHost, st));
        VX_CUDA(cudaMemcpyAsync(ctx->h_mesh_totals, ctx->d_mesh_totals, 8 * sizeof(uint64_t),
                                cudaMemcpyDeviceToHost, st));
        VX_CUDA(cudaEventRecord(ctx->ev1, st));
        VX_CUDA(cudaStreamSynchronize(st));
        if (vx_light_finish(ctx)) {  // (a light build queued before this one: its verdict is in by now)
            ctx->mesh_out.clear();
            return -1;
        }
        const uint64_t* t = ctx->h_mesh_totals;  // 0-3 arena totals, 4 slow units, 5 fast units, 6 unit-list overflow, 7 translucent cube units
        if (t[4] + t[5] > ucap || (split && t[7] > tcap)) {
            // a unit list was too small: some tiles were not listed, nothing below is valid
            if (t[4] + t[5] > ucap) {
                if (grow(ctx, &ctx->d_units, &ctx->units_words, (size_t)(t[4] + t[5]) * 4)) return -1;  // 4 words per unit
                ctx->units_cap = ctx->units_words / 4;
            }
            if (split && t[7] > tcap) {
                if (grow(ctx, &ctx->d_units_t, &ctx->units_t_words, (size_t)t[7] * 4)) return -1;
                ctx->units_t_cap = ctx->units_t_words / 4;
            }
            continue;
        }
        {
            const unsigned long long hint_was = ctx->slow_units_hint;
            // (blocks beyond the listed units cost next to nothing: twice the last count, and no sudden drop)
            ctx->slow_units_hint = std::max<unsigned long long>(2 * t[4] + 16384, hint_was - hint_was / 8);
            if (hint_was && t[4] > slow_covered) continue;  // more whole-voxel units than the grid covered
        }
        if (t[0] >= 0xFFFFFFFFull || t[1] >= 0xFFFFFFFFull || t[2] >= 0xFFFFFFFFull) {
            ctx->fail("vx_mesh_build: batch output exceeds 2^32 words; split the batch");
            ctx->mesh_out.clear();
            return -1;
        }
        ctx->arena_sizes[0] = t[0];
        ctx->arena_sizes[1] = t[1];
        ctx->arena_sizes[2] = t[3];
        ctx->arena_sizes[3] = t[2];
        if (t[0] <= caps.c[0] && t[1] <= caps.c[1] && t[2] <= caps.c[2] && t[3] <= caps.c[3]) {
            built = true;
            break;
        }
        if (grow(ctx, &ctx->d_vertices, &ctx->vertices_cap, (size_t)t[0])) return -1;
        if (grow(ctx, &ctx->d_indices, &ctx->indices_cap, (size_t)t[1])) return -1;
        if (grow(ctx, &ctx->d_tfloats, &ctx->tfloats_cap, (size_t)t[2])) return -1;
        if (grow(ctx, &ctx->d_entries, &ctx->entries_cap, (size_t)t[3])) return -1;
        // kept_* / aabb accumulators were not touched by the skipped pass
    }
    {
        // Where translucent cube faces are a large share of the output (seas of glass: the worst-case
        // world has 18 M of them per 1,024 chunks) they get a list and an emission pass of their own,
        // staged and stored like the opaque faces; where they are few (a terrain world's water surface)
        // that pass and the third list cost more than they return, and the faces ride with the slow
        // kinds.  Decided from the previous build of this context; the results are the same either way.
        static const char* force = getenv("VXGPU_SPLIT_TRANSL");
        const uint64_t* t = ctx->h_mesh_totals;
        ctx->split_transl = force ? force[0] == '1' : (t[2] > (1u << 20) && t[2] * 4 > t[0]);
    }
    if (!built) {  // the totals changed from pass to pass: nothing in the arenas is valid
        ctx->fail("vx_mesh_build: arena sizing failed");
        ctx->mesh_out.clear();
        return -1;
    }
    cudaEventElapsedTime(&ctx->last_mesh_ms, ctx->ev0, ctx->ev1);
    return 0;
}

int vx_mesh_get_info(vx_ctx* ctx, int32_t i, vx_mesh_info* out) {
    if (!ctx || !out) return -1;
    if (i < 0 || i >= (int)ctx->mesh_out.size()) {
        ctx->fail("vx_mesh_get_info: index out of range");
        return -1;
    }
    const MeshChunkOut& o = ctx->mesh_out[i];
    memset(out, 0, sizeof(*out));
    out->cx = ctx->mesh_keys[i].cx;
    out->cz = ctx->mesh_keys[i].cz;
    out->cancelled = o.cancelled;
    if (o.cancelled) return 0;
    out->n_vertex_floats = o.kept_floats;
    out->n_indices = o.kept_idx;
    out->n_entries = o.n_entries_final;
    out->n_entry_floats = o.t_floats;
    return 0;
}

int vx_mesh_get_slice(vx_ctx* ctx, int32_t i, vx_mesh_slice* out) {
    if (!ctx || !out) return -1;
    if (i < 0 || i >= (int)ctx->mesh_out.size()) {
        ctx->fail("vx_mesh_get_slice: index out of range");
        return -1;
    }
    const MeshChunkOut& o = ctx->mesh_out[i];
    out->vertex_off = o.v_off;
    out->index_off = o.i_off;
    out->entry_off = o.te_off;
    out->entry_float_off = o.tf_off;
    return 0;
}

int vx_mesh_arena_sizes(vx_ctx* ctx, uint64_t sizes[4]) {
    if (!ctx || !sizes) return -1;
    for (int k = 0; k < 4; k++) sizes[k] = ctx->mesh_out.empty() ? 0 : ctx->arena_sizes[k];
    return 0;
}

int vx_mesh_read_arena(vx_ctx* ctx, float* vertices, int32_t* indices, vx_sort_entry* entries,
                       float* entry_floats) {
    if (!ctx) return -1;
    if (ctx->mesh_out.empty()) return 0;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const uint64_t* sz = ctx->arena_sizes;
    if (vertices && sz[0])
        VX_CUDA(cudaMemcpyAsync(vertices, ctx->d_vertices, sz[0] * 4, cudaMemcpyDeviceToHost, st));
    if (indices && sz[1])
        VX_CUDA(cudaMemcpyAsync(indices, ctx->d_indices, sz[1] * 4, cudaMemcpyDeviceToHost, st));
    if (entries && sz[2])
        VX_CUDA(cudaMemcpyAsync(entries, ctx->d_entries, sz[2] * sizeof(vx_sort_entry),
                                cudaMemcpyDeviceToHost, st));
    if (entry_floats && sz[3])
        VX_CUDA(cudaMemcpyAsync(entry_floats, ctx->d_tfloats, sz[3] * 4, cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int vx_mesh_read(vx_ctx* ctx, int32_t i, float* vertices, int32_t* indices, vx_sort_entry* entries,
                 float* entry_floats) {
    if (!ctx) return -1;
    if (i < 0 || i >= (int)ctx->mesh_out.size()) {
        ctx->fail("vx_mesh_read: index out of range");
        return -1;
    }
    cudaSetDevice(ctx->device);
    const MeshChunkOut& o = ctx->mesh_out[i];
    if (o.cancelled) {
        ctx->fail("vx_mesh_read: chunk was cancelled");
        return -1;
    }
    cudaStream_t st = ctx->stream;
    if (vertices && o.kept_floats)
        VX_CUDA(cudaMemcpyAsync(vertices, ctx->d_vertices + o.v_off, (size_t)o.kept_floats * 4,
                                cudaMemcpyDeviceToHost, st));
    if (indices && o.kept_idx)
        VX_CUDA(cudaMemcpyAsync(indices, ctx->d_indices + o.i_off, (size_t)o.kept_idx * 4,
                                cudaMemcpyDeviceToHost, st));
    if (entries && o.n_entries_final)
        VX_CUDA(cudaMemcpyAsync(entries, ctx->d_entries + o.te_off,
                                (size_t)o.n_entries_final * sizeof(vx_sort_entry),
                                cudaMemcpyDeviceToHost, st));
    if (entry_floats && o.t_floats)
        VX_CUDA(cudaMemcpyAsync(entry_floats, ctx->d_tfloats + o.tf_off, (size_t)o.t_floats * 4,
                                cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
