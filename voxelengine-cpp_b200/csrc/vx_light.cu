// libvxgpu: lighting.  Replaces Lighting::{prebuildSkyLight, buildSkyLight,
// onChunkLoaded, clear} and LightSolver's add phase (lighting/Lighting.cpp:28-161,
// lighting/LightSolver.cpp:18-35,97-125) with a batched, activation-based
// wavefront solver.
//
// What the reference computes for a full relight: each channel is
// the least fixpoint of
//     u activated, L(u) >= 2, v in N6(u), lightPassing(v), L(v) < L(u)-1
//         =>  L(v) := L(u)-1, v activated
// started from the seed set of Lighting::buildSkyLight / onChunkLoaded.  The
// fixpoint does not depend on queue or chunk order, so all chunks of a batch
// and all four channels (packed in the 16-bit light_t) are solved together:
//   k_ystar   per column the row where buildSkyLight starts seeding
//   k_seed    emitter seeding + the activated bits of every chunk border plane
//             and y-slab boundary layer
//   k_solve   one CTA per dirty 16x32x16 y-slab: light slab + 1-voxel halo of
//             "propagating" light P (= light of activated voxels) staged in
//             shared memory; a dense first sweep, then sparse wavefront rounds
//             driven by a changed-voxel bit mask (dilate -> candidate words ->
//             ballot compaction) until the slab is at its local fixpoint; then
//             the border planes are published and the neighbour slabs whose
//             halo changed are marked dirty for the next launch.
// The host relaunches k_solve until no slab is dirty.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cooperative_groups.h>

#include "vx_ctx.hpp"

using namespace vx;
namespace cg = cooperative_groups;

namespace {

// ---------------------------------------------------------------------------
// nibble SIMD: a light_t holds four 4-bit channels.  E() spreads them to the
// low nibbles of four bytes so that byte-wise compares have 4 guard bits.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nE(uint32_t a) { return (a | (a << 12)) & 0x0F0F0F0Fu; }
__device__ __forceinline__ uint32_t nC(uint32_t e) { return (e | (e >> 12)) & 0xFFFFu; }
// bytes of a >= bytes of b -> 0x0F, else 0
__device__ __forceinline__ uint32_t nGE(uint32_t a, uint32_t b) {
    return ((((a | 0x10101010u) - b) & 0x10101010u) >> 4) * 15u;
}
__device__ __forceinline__ uint32_t nMAX(uint32_t a, uint32_t b) {
    uint32_t m = nGE(a, b);
    return b ^ ((a ^ b) & m);
}
// saturating byte-wise decrement
__device__ __forceinline__ uint32_t nDEC(uint32_t e) {
    return e - (((e + 0x0F0F0F0Fu) & 0x10101010u) >> 4);
}
// 4 bits -> 16-bit nibble mask (bit k -> nibble k = 0xF)
__device__ __forceinline__ uint32_t bits_to_mask(uint32_t a) {
    uint32_t s = (a | (a << 3) | (a << 6) | (a << 9)) & 0x1111u;
    return s * 15u;
}
// 16-bit light -> 4 bits (nibble k != 0)
__device__ __forceinline__ uint32_t nonzero_bits(uint32_t l) {
    uint32_t t = (l | (l >> 1) | (l >> 2) | (l >> 3)) & 0x1111u;
    return (t | (t >> 3) | (t >> 6) | (t >> 9)) & 0xFu;
}


// ---------------------------------------------------------------------------
// k_derive: per-chunk tables derived from the voxel array at put / edit time.
// Also Chunk::updateHeights (voxels/Chunk.cpp:21-34).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_derive(DevWorld W, const int32_t* pools) {
    const int p = pools[blockIdx.x];
    __shared__ uint32_t s_nslp[PLANE_WORDS];
    __shared__ int s_first, s_last, s_skmin, s_skmax;
    __shared__ int s_lay[CH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_first = PLANE_WORDS;
        s_last = -1;
        s_skmin = CH;
        s_skmax = -1;
    }
    if (tid < CH) s_lay[tid] = 0;
    __syncthreads();
    const vx_voxel* vox = W.vox_of(p);
    uint32_t* lp = W.lp_of(p);
    uint32_t* em = W.emit_of(p);
    int first = PLANE_WORDS, last = -1;
    for (int w = warp; w < PLANE_WORDS; w += 32) {
        const uint32_t vv = vox[w * 32 + lane];
        uint32_t id = vv & 0xFFFFu;
        uint32_t f = id < (uint32_t)W.n_defs ? __ldg(&W.defs[id].flags) : 0u;
        uint32_t b_lp = __ballot_sync(0xFFFFFFFFu, f & DF_LP);
        uint32_t b_em = __ballot_sync(0xFFFFFFFFu, f & DF_EMISSIVE);
        uint32_t b_ns = __ballot_sync(0xFFFFFFFFu, !(f & DF_SLP));
        uint32_t b_na = __ballot_sync(0xFFFFFFFFu, id != 0);
        const uint32_t mb = mesher_bits(vv, f, (uint32_t)W.n_defs);
        uint32_t b_blk = __ballot_sync(0xFFFFFFFFu, mb & 1u);
        uint32_t b_sc = __ballot_sync(0xFFFFFFFFu, mb & 2u);
        uint32_t b_oth = __ballot_sync(0xFFFFFFFFu, mb & 4u);
        if (lane == 0) {
            lp[w] = b_lp;
            em[w] = b_em;
            s_nslp[w] = b_ns;
            W.mblk_of(p)[w] = b_blk;
            W.msc_of(p)[w] = b_sc;
            W.moth_of(p)[w] = b_oth;
        }
        if (b_na) {
            first = min(first, w);
            last = max(last, w);
            if (lane == 0) atomicAdd(&s_lay[w / LAYER_WORDS], __popc(b_na));
        }
    }
    if (lane == 0) {
        if (first < PLANE_WORDS) atomicMin(&s_first, first);
        if (last >= 0) atomicMax(&s_last, last);
    }
    __syncthreads();
    if (tid < 256) {  // column (x = tid & 15, z = tid >> 4)
        const int x = tid & 15, z = tid >> 4;
        const int bit = (z & 1) * 16 + x, wz = z >> 1;
        int sc = -1;
        for (int y = CH - 1; y >= 0; y--) {
            if ((s_nslp[y * LAYER_WORDS + wz] >> bit) & 1u) {
                sc = y;
                break;
            }
        }
        W.skycol_of(p)[tid] = (int16_t)sc;
        // On a pristine lightmap (S = 15 exactly above skycol) Lighting::buildSkyLight (:75-91) starts
        // seeding a column at the first light-passing voxel at or below skycol, or at y = 0: static.
        int ys0 = sc < 0 ? -1 : 0;
        for (int y = sc; y > 0; y--) {
            if ((lp[y * LAYER_WORDS + wz] >> bit) & 1u) {  // written above by this block
                ys0 = y;
                break;
            }
        }
        W.ystar0_of(p)[tid] = (int16_t)ys0;
        W.laycnt_of(p)[tid] = (int16_t)s_lay[tid];  // non-air voxels per layer (Chunk::updateHeights)
        atomicMin(&s_skmin, sc);
        atomicMax(&s_skmax, sc);
    }
    __syncthreads();
    if (tid == 0) {
        W.meta[p].sky_min = s_skmin;
        W.meta[p].sky_max = s_skmax;
        // word index -> y; exact first/last voxel is inside that word's layer
        ChunkMeta* m = &W.meta[p];
        if (s_first < PLANE_WORDS) m->bottom = s_first / LAYER_WORDS;
        if (s_last >= 0) m->top = s_last / LAYER_WORDS + 1;
    }
}

// Rebuilds faceL (the border-plane copy of `light`) from the light array.
__global__ void __launch_bounds__(256) k_faces_refresh(DevWorld W, const int32_t* pools) {
    const int p = pools[blockIdx.x];
    const vx_light* L = W.light_of(p);
    for (int t = threadIdx.x; t < 4 * CH * 16; t += blockDim.x) {
        const int f = t >> 12, y = (t >> 4) & 255, i = t & 15;
        int x, z;
        if (f == FACE_XM) { x = 0; z = i; }
        else if (f == FACE_XP) { x = 15; z = i; }
        else if (f == FACE_ZM) { x = i; z = 0; }
        else { x = i; z = 15; }
        W.faceL_of(p, f)[y * 16 + i] = L[vidx(x, y, z)];
    }
}

// ---------------------------------------------------------------------------
// Lazy lightmaps.  Lighting::clear (:28-39) and prebuildSkyLight (:41-63) of a lightmap nothing
// has read yet only move ChunkMeta::lstate (ZERO -> PRISTINE) and set ChunkMeta::lvirt: the
// 128 KiB array and its border planes then hold *nothing*, and what they would hold follows
// from the state and the sky columns.  k_seed writes the real thing when the chunk enters a
// build (it has to write the lightmap once anyway); every other reader goes through
// vx_materialize first.  A full rebuild thus writes each lightmap once instead of three times.
// ---------------------------------------------------------------------------
// the 8 voxels of uint4 unit u (layer ly = u >> 5 of a slab starting at y0) of a virtual lightmap
__device__ __forceinline__ uint4 virt_light8(const int16_t* sc, bool pristine, int y, int u) {
    uint32_t m[4] = {0, 0, 0, 0};
    if (pristine) {
        const int z = (u >> 1) & 15, x0 = (u & 1) * 8;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (y > sc[z * 16 + x0 + j]) m[j >> 1] |= 0xF000u << ((j & 1) * 16);
    }
    return make_uint4(m[0], m[1], m[2], m[3]);
}

// grid (NSLAB, n): writes out the virtual lightmaps among `pools`
__global__ void __launch_bounds__(256) k_materialize(DevWorld W, const int32_t* pools) {
    const int p = pools[blockIdx.y];
    if (!W.meta[p].lvirt) return;
    const int s = blockIdx.x, tid = threadIdx.x, y0 = s * SLAB_H;
    __shared__ int16_t sc[256];
    sc[tid] = W.skycol_of(p)[tid];
    const bool pristine = W.meta[p].lstate == LS_PRISTINE;
    __syncthreads();
    uint4* L4 = reinterpret_cast<uint4*>(W.light_of(p)) + s * 1024;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int u = k * 256 + tid;
        L4[u] = virt_light8(sc, pristine, y0 + (u >> 5), u);
    }
    for (int t = tid; t < 4 * SLAB_H * 16; t += 256) {
        const int f = t >> 9, y = y0 + ((t >> 4) & 31), i = t & 15;
        int x, z;
        if (f == FACE_XM) { x = 0; z = i; }
        else if (f == FACE_XP) { x = 15; z = i; }
        else if (f == FACE_ZM) { x = i; z = 0; }
        else { x = i; z = 15; }
        W.faceL_of(p, f)[y * 16 + i] = (vx_light)((pristine && y > sc[z * 16 + x]) ? 0xF000u : 0u);
    }
}

__global__ void k_devirt(DevWorld W, const int32_t* pools, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) W.meta[pools[t]].lvirt = 0;
}

// Lighting::prebuildSkyLight of lightmaps that are all virtual (the host knows): state only
__global__ void k_prebuild_virtual(DevWorld W, const int32_t* pools, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    ChunkMeta* m = &W.meta[pools[t]];
    int h = m->sky_max < 0 ? 0 : m->sky_max;  // `highest` starts at 0 (:43)
    if (h < CH - 1) h++;
    m->highest = h;
    m->lstate = LS_PRISTINE;
}

// Lighting::clear without touching the arrays
__global__ void k_clear_virtual(DevWorld W, const int32_t* pools, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        ChunkMeta* m = &W.meta[pools[t]];
        m->lstate = LS_ZERO;
        m->lvirt = 1;
    }
}

// ---------------------------------------------------------------------------
// k_prebuild: Lighting::prebuildSkyLight (lighting/Lighting.cpp:41-63).
// S := 15 above the highest non-skyLightPassing voxel of every column.
// grid (NSLAB, n)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_prebuild(DevWorld W, const int32_t* pools, int track) {
    const int p = pools[blockIdx.y];
    const int s = blockIdx.x, tid = threadIdx.x;
    if (W.meta[p].lvirt) {
        // nothing has read this lightmap since it was cleared: it stays unwritten, prebuilt in state only
        // (no block changes lvirt here, so all eight agree)
        if (s == 0 && tid == 0) {
            ChunkMeta* m = &W.meta[p];
            int h = m->sky_max < 0 ? 0 : m->sky_max;  // `highest` starts at 0 (:43)
            if (h < CH - 1) h++;
            m->highest = h;
            m->lstate = LS_PRISTINE;
        }
        return;
    }
    __shared__ int16_t sc[256];
    __shared__ int s_max;
    if (tid == 0) s_max = 0;
    sc[tid] = W.skycol_of(p)[tid];
    // block 0 moves the state on at its very end; a block that already sees PRISTINE just takes
    // the read-modify-write path, which is always right
    const bool zero = W.meta[p].lstate == LS_ZERO;
    __syncthreads();
    if (s == 0) {
        atomicMax(&s_max, (int)sc[tid]);
    }
    uint4* L4 = reinterpret_cast<uint4*>(W.light_of(p));
    // 8 voxels per uint4; unit u: y = u >> 5, z = (u >> 1) & 15, x0 = (u & 1) * 8
    for (int k = 0; k < 4; k++) {
        const int u = s * 1024 + k * 256 + tid;
        const int y = u >> 5, z = (u >> 1) & 15, x0 = (u & 1) * 8;
        uint32_t m[4] = {0, 0, 0, 0};
        bool any = false;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (y > sc[z * 16 + x0 + j]) {
                m[j >> 1] |= 0xF000u << ((j & 1) * 16);
                any = true;
            }
        }
        if (any) {
            // a lightmap known to be all zero (Lighting::clear, or a chunk put without light) is
            // written without being read
            uint4 v = zero ? make_uint4(0, 0, 0, 0) : L4[u];
            v.x |= m[0];
            v.y |= m[1];
            v.z |= m[2];
            v.w |= m[3];
            L4[u] = v;
        }
    }
    // border planes
    for (int t = tid; t < 4 * SLAB_H * 16; t += 256) {
        const int f = t >> 9, y = s * SLAB_H + ((t >> 4) & 31), i = t & 15;
        int x, z;
        if (f == FACE_XM) { x = 0; z = i; }
        else if (f == FACE_XP) { x = 15; z = i; }
        else if (f == FACE_ZM) { x = i; z = 0; }
        else { x = i; z = 15; }
        if (y > sc[z * 16 + x]) {
            vx_light* fl = W.faceL_of(p, f) + y * 16 + i;
            *fl = zero ? (vx_light)0xF000u : (vx_light)(*fl | 0xF000u);
        }
    }
    if (s == 0) {
        __syncthreads();
        if (tid == 0) {
            int h = s_max < 0 ? 0 : s_max;  // `highest` starts at 0
            if (h < CH - 1) h++;
            W.meta[p].highest = h;
            const int ls = W.meta[p].lstate;
            W.meta[p].lstate = (track && (ls == LS_ZERO || ls == LS_PRISTINE)) ? LS_PRISTINE : LS_UNKNOWN;
        }
    }
}

// Lighting::clear (lighting/Lighting.cpp:28-39) for a list of chunks.
__global__ void __launch_bounds__(256) k_clear(DevWorld W, const int32_t* pools) {
    const int p = pools[blockIdx.y];
    uint4* L4 = reinterpret_cast<uint4*>(W.light_of(p));
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    for (int u = blockIdx.x * 1024 + threadIdx.x; u < (blockIdx.x + 1) * 1024; u += 256) L4[u] = z4;
    uint4* F4 = reinterpret_cast<uint4*>(W.faceL_of(p, 0));
    for (int u = blockIdx.x * 256 + threadIdx.x; u < (blockIdx.x + 1) * 256; u += 256) F4[u] = z4;
    if (blockIdx.x == 0 && threadIdx.x == 0) W.meta[p].lstate = LS_ZERO;
}

// ---------------------------------------------------------------------------
// light build
// ---------------------------------------------------------------------------
__global__ void k_light_begin(DevWorld W, const int32_t* lit, int n_lit, const int32_t* solve,
                              int n_solve) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_solve) {
        const int p = solve[t];
        ChunkMeta* m = &W.meta[p];
        m->any_act = 0;
        m->pad0 = 1;  // in_set
        m->pad1 = 0;  // faces that hold a seed (activated voxel)
        // what the eight k_seed blocks of the chunk would otherwise each look up again
        const int nx[4] = {m->cx - 1, m->cx + 1, m->cx, m->cx}, nz[4] = {m->cz, m->cz, m->cz - 1, m->cz + 1};
        bool pristine = m->lstate == LS_PRISTINE;
        int sky_hi = m->sky_max;
        for (int f = 0; f < 4; f++) {
            const int qn = W.pool_of(nx[f], nz[f]);
            m->nbp[f] = qn;
            if (qn < 0) continue;
            if (W.meta[qn].lstate != LS_PRISTINE) pristine = false;
            sky_hi = max(sky_hi, W.meta[qn].sky_max);
        }
        m->nb_pristine = pristine;
        m->nb_sky_hi = sky_hi;
    }
    if (t < n_lit) W.meta[lit[t]].lit_now = 1;
}

// k_ystar: the row y* at which Lighting::buildSkyLight (Lighting.cpp:75-91)
// finds a column's first non-15 sky value: scanning down from highestPoint,
// non-lightPassing voxels are skipped (except at y == 0) and the first
// remaining voxel with S != 15 is y*; -1 when the scan runs off the bottom.
__global__ void __launch_bounds__(1024) k_ystar(DevWorld W, const int32_t* lit) {
    const int p = lit[blockIdx.x];
    __shared__ uint32_t cand[PLANE_WORDS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int highest = W.meta[p].highest;
    const vx_light* L = W.light_of(p);
    const uint32_t* lp = W.lp_of(p);
    const int nw = (highest + 1) * LAYER_WORDS;
    if (W.meta[p].lstate == LS_PRISTINE) {
        // prebuilt-only lightmap: the answer is the static per-column table k_derive / k_edit keep
        if (tid < 256) W.ystar_of(p)[tid] = W.ystar0_of(p)[tid];
        return;
    } else {
        const bool virt = W.meta[p].lvirt != 0;  // (and not pristine: all zero)
        for (int w = warp; w < nw; w += 32) {
            uint32_t l = virt ? 0u : L[w * 32 + lane];
            uint32_t s15 = __ballot_sync(0xFFFFFFFFu, (l >> 12) == 15u);
            if (lane == 0) cand[w] = (lp[w] | (w < LAYER_WORDS ? 0xFFFFFFFFu : 0u)) & ~s15;
        }
    }
    __syncthreads();
    if (tid < 256) {
        const int x = tid & 15, z = tid >> 4;
        const int bit = (z & 1) * 16 + x, wz = z >> 1;
        int ys = -1;
        for (int y = highest; y >= 0; y--) {
            if ((cand[y * LAYER_WORDS + wz] >> bit) & 1u) {
                ys = y;
                break;
            }
        }
        W.ystar_of(p)[tid] = (int16_t)ys;
    }
}

// flags.loadedLights chunks are lit without Lighting::buildSkyLight: no sky seeds
__global__ void k_ystar_none(DevWorld W, const int32_t* lit) {
    W.ystar_of(lit[blockIdx.x])[threadIdx.x] = -1;
}

// 18x18 tile of y* around a chunk: entry (z+1)*18 + (x+1) for local column
// (x, z) in [-1, 16]; -1 where the column's chunk is not being lit.
__device__ __forceinline__ void load_ystar_tile(const DevWorld& W, int p, int nxm, int nxp, int nzm, int nzp,
                                                int16_t* ys) {
    for (int t = threadIdx.x; t < 18 * 18; t += blockDim.x) {
        const int lx = t % 18 - 1, lz = t / 18 - 1;
        const bool ox = (unsigned)lx > 15u, oz = (unsigned)lz > 15u;
        // the corners are never read: a column's seeds depend on its four side neighbours only
        const int q = ox ? (oz ? -1 : (lx < 0 ? nxm : nxp)) : (oz ? (lz < 0 ? nzm : nzp) : p);
        int16_t v = -1;
        if (q >= 0 && W.meta[q].lit_now) v = W.ystar_of(q)[(lz & 15) * 16 + (lx & 15)];
        ys[t] = v;
    }
}

// k_seed: grid (NSLAB, n_solve).  Per y-slab: (a) emitter seeding of lit chunks
// (light := max(light, emission) where LightSolver::add accepts it), (b) the
// seed bits of every voxel -> nibble plane W.a0 (only written when non-zero),
// (c) the activated bits of the border planes / slab boundary layers, so that
// the first k_solve launch already sees its halo, (d) per-slab activity flags.
constexpr uint32_t SF_ANY = 1u;       // some voxel of the slab is a seed
constexpr uint32_t SF_FACE0 = 2u;     // << face: the face plane holds a seed
constexpr uint32_t SF_BOTTOM = 32u;   // bottom layer holds a seed
constexpr uint32_t SF_TOP = 64u;      // top layer holds a seed

#ifndef VX_SEED_MINB
#define VX_SEED_MINB 6
#endif
#ifndef VX_SOLVE_PER_SM
#define VX_SOLVE_PER_SM 99
#endif
__global__ void __launch_bounds__(256, VX_SEED_MINB) k_seed(DevWorld W, const int32_t* solve, int expand) {
    const int p = solve[blockIdx.y];
    const int s = blockIdx.x, y0 = s * SLAB_H;
    const int tid = threadIdx.x;
    const ChunkMeta m = W.meta[p];
    const bool lit = m.lit_now != 0;
    // the chunk's lightmap exists as a state only (lazy clear / prebuild): this kernel writes it out
    const bool virt = m.lvirt != 0, vpristine = m.lstate == LS_PRISTINE;
    __shared__ int16_t scol[256];
    scol[tid] = virt ? W.skycol_of(p)[tid] : (int16_t)0;
    __shared__ __align__(16) uint16_t Ls[SLAB_H * 256];
    __shared__ __align__(16) uint8_t a0[SLAB_H * 256];
    __shared__ int16_t ys[18 * 18];
    __shared__ uint32_t emw[SLAB_WORDS], lpw[SLAB_WORDS];
    __shared__ uint32_t pl_sky[SLAB_WORDS], pl_any[SLAB_WORDS], pl_rai[SLAB_WORDS];
    __shared__ uint16_t hm[4][SLAB_H];
    __shared__ uint32_t hy[2][8];
    __shared__ uint32_t s_flags;
    const int lane = tid & 31, warp = tid >> 5;
    vx_light* Lg = W.light_of(p) + (size_t)y0 * 256;
    const vx_voxel* vox = W.vox_of(p) + (size_t)y0 * 256;
    // ---- pristine fast path: the chunk and its four side neighbours hold exactly
    // the prebuilt sky (S = 15 above skycol, nothing else), so a slab that lies
    // entirely above every sky column around it holds only inert sky seeds, and a
    // slab entirely at or below its own sky columns holds no light at all; with
    // no emitter in the slab neither needs its lightmap read.
    {
        const bool pristine = m.nb_pristine != 0;  // k_light_begin looked at the neighbours
        const int sky_hi = m.nb_sky_hi;
        const uint32_t ew = lit ? W.emit_of(p)[s * SLAB_WORDS + tid] : 0u;
        const int has_emit = __syncthreads_or(ew != 0);
        const bool above = y0 - 1 > sky_hi, below = y0 + SLAB_H - 1 <= m.sky_min;
        if (pristine && !has_emit && (above || below)) {
            if (virt) {
                // all sky (above every column) or all dark (below every column): written out here
                const uint32_t w = above ? 0xF000F000u : 0u;
                const uint4 v4 = make_uint4(w, w, w, w);
                uint4* L4 = reinterpret_cast<uint4*>(Lg);
#pragma unroll
                for (int k = 0; k < 4; k++) L4[tid + 256 * k] = v4;
                // border planes: 4 faces x 32 rows x 32 bytes
                reinterpret_cast<uint4*>(W.faceL_of(p, tid >> 6) + (y0 + ((tid >> 1) & 31)) * 16)[tid & 1] = v4;
            }
            if (tid < 4 * SLAB_H) W.faceA_of(p, tid >> 5)[y0 + (tid & 31)] = 0ull;
            if (tid < 2 * 16) W.layerA_of(p, 2 * s + (tid >> 4))[tid & 15] = 0ull;
            if (tid == 0) {
                W.sflag_of(p)[s] = 0;
                if (above && lit && expand) {  // the border planes are (inert) sky seeds
                    atomicOr(&W.meta[p].any_act, 1);
                    atomicOr(&W.meta[p].pad1, 0xF);
                }
            }
            return;
        }
    }
    {
        // every load of the slab and of its halo up front.  The halo tests below look at sky
        // nibbles only, which the emitter seeding of other blocks (R, G, B) never touches.
        const uint4* src = reinterpret_cast<const uint4*>(Lg);
        uint4 l4[SLAB_H * 256 / 8 / 256];
#pragma unroll
        for (int k = 0; k < SLAB_H * 256 / 8 / 256; k++) {
            const int u = tid + 256 * k;
            l4[k] = virt ? virt_light8(scol, vpristine, y0 + (u >> 5), u) : src[u];
        }
        const uint32_t emv = lit ? W.emit_of(p)[s * SLAB_WORDS + tid] : 0u;
        const uint32_t lpv = W.lp_of(p)[s * SLAB_WORDS + tid];
        // "raisable by a 15" bits of the halo for the inert-seed filter: neighbour chunks' border
        // planes (light-passing unknown there: assume it, which only keeps more seeds) ...
        uint32_t hl[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int r = (tid >> 4) + 16 * k;  // face r >> 5 == k >> 1: a constant per unrolled step
            const int qf = m.nbp[k >> 1];
            const int of = (k >> 1) ^ 1, yy = y0 + (r & 31), i = tid & 15;
            hl[k] = 0xF000u;
            if (qf >= 0) {
                const ChunkMeta* mq = W.meta + qf;
                if (!mq->lvirt) {
                    hl[k] = W.faceL_of(qf, of)[yy * 16 + i];
                } else {
                    // the neighbour's border plane is not in memory either: its sky columns say what it holds
                    // (its own k_seed block may be writing it out right now; the flag only falls afterwards)
                    const int col = of == FACE_XM ? i * 16 : of == FACE_XP ? i * 16 + 15 : of == FACE_ZM ? i : 15 * 16 + i;
                    hl[k] = (mq->lstate == LS_PRISTINE && yy > W.skycol_of(qf)[col]) ? 0xF000u : 0u;
                }
            }
        }
        // ... and the layers just below / above the slab
        const vx_light* Lall = W.light_of(p);
        const uint32_t* lpg = W.lp_of(p);
        bool ray[2];
#pragma unroll
        for (int up = 0; up < 2; up++) {
            const int y = up ? y0 + SLAB_H : y0 - 1;
            ray[up] = false;
            if (y >= 0 && y < CH) {
                const uint32_t sv = virt ? ((vpristine && y > scol[tid]) ? 15u : 0u) : (uint32_t)(Lall[y * 256 + tid] >> 12);
                ray[up] = ((lpg[y * 8 + warp] >> lane) & 1u) && sv < 14u;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(Ls);
#pragma unroll
        for (int k = 0; k < SLAB_H * 256 / 8 / 256; k++) dst[tid + 256 * k] = l4[k];
        emw[tid] = emv;
        lpw[tid] = lpv;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int r = (tid >> 4) + 16 * k;
            const int f = r >> 5, ly = r & 31;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, (hl[k] >> 12) < 14u);
            if ((tid & 15) == 0) hm[f][ly] = (uint16_t)(b >> (lane & 16));
        }
#pragma unroll
        for (int up = 0; up < 2; up++) {
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, ray[up]);
            if (lane == 0) hy[up][warp] = b;
        }
        if (tid == 0) s_flags = 0;
    }
    load_ystar_tile(W, p, m.nbp[FACE_XM], m.nbp[FACE_XP], m.nbp[FACE_ZM], m.nbp[FACE_ZP], ys);
    __syncthreads();
    // ---- emitters (word tid)
    if (emw[tid]) {
        uint32_t bits = emw[tid];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int v = tid * 32 + b;
            const uint32_t e = __ldg(&W.defs[vox[v] & 0xFFFFu].emission);
            const uint32_t l = Ls[v];
            uint32_t nl = l;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                uint32_t ec = (e >> (4 * ch)) & 0xFu;
                if (ec > 1u && ec >= ((l >> (4 * ch)) & 0xFu))
                    nl = (nl & ~(0xFu << (4 * ch))) | (ec << (4 * ch));
            }
            if (nl != l) {
                Ls[v] = (uint16_t)nl;
                Lg[v] = (vx_light)nl;
                const int x = v & 15, z = (v >> 4) & 15, y = y0 + (v >> 8);
                if (x == 0) W.faceL_of(p, FACE_XM)[y * 16 + z] = (vx_light)nl;
                if (x == 15) W.faceL_of(p, FACE_XP)[y * 16 + z] = (vx_light)nl;
                if (z == 0) W.faceL_of(p, FACE_ZM)[y * 16 + x] = (vx_light)nl;
                if (z == 15) W.faceL_of(p, FACE_ZP)[y * 16 + x] = (vx_light)nl;
            }
        }
    }
    __syncthreads();
    // ---- seed bits of the 32 voxels of column (x, z) = (tid & 15, tid >> 4).
    // Lane l of warp w handles voxel k*256 + w*32 + l = bit l of plane word k*8+w.
    uint32_t any_all = 0, face_seed = 0;
    {
        const int x = tid & 15, z = tid >> 4;
        const int c = (z + 1) * 18 + (x + 1);
        // Lighting::buildSkyLight (:82-89): (x, y*+1, z) and, for y <= y* of a lit
        // column, its four side neighbours
        const int smax = max(max((int)ys[c - 1], (int)ys[c + 1]), max((int)ys[c - 18], (int)ys[c + 18]));
        const int yown1 = ys[c] >= 0 ? ys[c] + 1 : -100;
        const bool border = lit && expand && (x == 0 || x == 15 || z == 0 || z == 15);
        uint32_t fmask = 0;
        if (x == 0) fmask |= 1u << FACE_XM;
        if (x == 15) fmask |= 1u << FACE_XP;
        if (z == 0) fmask |= 1u << FACE_ZM;
        if (z == 15) fmask |= 1u << FACE_ZP;
        for (int k = 0; k < SLAB_H; k++) {
            const int v = k * 256 + tid, y = y0 + k;
            const uint32_t l = Ls[v];
            uint32_t bits = 0;
            if ((l >> 12) > 1u && (y <= smax || y == yown1)) bits |= 8u;
            if (border && l) {  // onChunkLoaded expand (:126-155); LightSolver::add needs > 1
                if ((l & 0xFu) > 1u) bits |= 1u;
                if (((l >> 4) & 0xFu) > 1u) bits |= 2u;
                if (((l >> 8) & 0xFu) > 1u) bits |= 4u;
                if ((l >> 12) > 1u) bits |= 8u;
            }
            if ((emw[v >> 5] >> (v & 31)) & 1u) {  // onChunkLoaded emitters (:110-124)
                const uint32_t e = __ldg(&W.defs[vox[v] & 0xFFFFu].emission);
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    uint32_t ec = (e >> (4 * ch)) & 0xFu;
                    if (ec > 1u && ec >= ((l >> (4 * ch)) & 0xFu)) bits |= 1u << ch;
                }
            }
            a0[v] = (uint8_t)bits;
            if (bits) {
                any_all = 1;
                face_seed |= fmask;
            }
            // planes for the inert-seed filter below
            const int w = k * 8 + warp;
            const uint32_t b_sky = __ballot_sync(0xFFFFFFFFu, bits == 8u && (l >> 12) == 15u);
            const uint32_t b_any = __ballot_sync(0xFFFFFFFFu, bits != 0);
            const uint32_t b_rai = __ballot_sync(0xFFFFFFFFu, ((lpw[w] >> lane) & 1u) && (l >> 12) < 14u);
            if (lane == 0) {
                pl_sky[w] = b_sky;
                pl_any[w] = b_any;
                pl_rai[w] = b_rai;
            }
        }
    }
    // ---- drop inert sky seeds.  A seed u can only ever raise a light-passing
    // neighbour v with L(v) < L(u) - 1, and light never decreases during a build,
    // so a seed whose neighbours are all at or above that stays inert (if u is
    // raised later, that raise re-activates it).  Checked here for the common
    // case only — a pure sky seed with S = 15, inert iff no light-passing
    // neighbour has S < 14 — as a bit-plane dilation; every other seed is kept.
    // Inert seeds still count for Chunk::flags.modified (any_act / face seeds):
    // open-sky border planes then cost nothing in k_solve.
    {
        // (hm / hy, the halo's "raisable by a 15" bits, were computed with the first loads)
    }
    __syncthreads();
    uint32_t any_act;
    {
        const int ly = tid >> 3, zp = tid & 7;
        const uint32_t c = pl_rai[tid];
        const uint32_t up = ly < SLAB_H - 1 ? pl_rai[tid + 8] : hy[1][zp];
        const uint32_t dn = ly > 0 ? pl_rai[tid - 8] : hy[0][zp];
        const uint32_t nzm = (c << 16) | (zp > 0 ? pl_rai[tid - 1] >> 16 : (uint32_t)hm[FACE_ZM][ly]);
        const uint32_t nzp = (c >> 16) | (zp < 7 ? pl_rai[tid + 1] << 16 : (uint32_t)hm[FACE_ZP][ly] << 16);
        const uint32_t hxm = hm[FACE_XM][ly], hxp = hm[FACE_XP][ly];
        const uint32_t xl = ((c << 1) & 0xFFFEFFFEu) | ((hxm >> (2 * zp)) & 1u) |
                            (((hxm >> (2 * zp + 1)) & 1u) << 16);
        const uint32_t xr = ((c >> 1) & 0x7FFF7FFFu) | (((hxp >> (2 * zp)) & 1u) << 15) |
                            (((hxp >> (2 * zp + 1)) & 1u) << 31);
        const uint32_t live = up | dn | nzm | nzp | xl | xr;
        uint32_t inert = pl_sky[tid] & ~live;
        // ... and, whatever the channel, a seed none of whose six neighbours lets light through
        // (an emitter embedded in rock: LightSolver.cpp:116 never passes) — cells across a chunk
        // border count as light-passing here, which only keeps more seeds
        {
            const uint32_t lc = lpw[tid];
            const int yu = y0 + SLAB_H, yd = y0 - 1;
            const uint32_t lup = ly < SLAB_H - 1 ? lpw[tid + 8] : (yu < CH ? __ldg(W.lp_of(p) + yu * 8 + zp) : 0u);
            const uint32_t ldn = ly > 0 ? lpw[tid - 8] : (yd >= 0 ? __ldg(W.lp_of(p) + yd * 8 + zp) : 0u);
            const uint32_t lzm = (lc << 16) | (zp > 0 ? lpw[tid - 1] >> 16 : 0xFFFFu);
            const uint32_t lzp = (lc >> 16) | (zp < 7 ? lpw[tid + 1] << 16 : 0xFFFF0000u);
            const uint32_t lxl = ((lc << 1) & 0xFFFEFFFEu) | 0x00010001u;
            const uint32_t lxr = ((lc >> 1) & 0x7FFF7FFFu) | 0x80008000u;
            inert |= pl_any[tid] & ~(lup | ldn | lzm | lzp | lxl | lxr);
        }
        any_act = pl_any[tid] & ~inert;
        while (inert) {
            const int b = __ffs(inert) - 1;
            inert &= inert - 1;
            a0[tid * 32 + b] = 0;
        }
    }
    const uint32_t any = __syncthreads_or(any_act != 0);  // a0[] final and visible
    any_all = __syncthreads_or(any_all != 0);
    uint32_t flags = any ? SF_ANY : 0u;
    // ---- border planes: 4 faces x 32 rows of 16 voxels
    const int i16 = tid & 15;
    for (int r = tid >> 4; r < 4 * SLAB_H; r += 16) {
        const int f = r >> 5, ly = r & 31;
        int x, z;
        if (f == FACE_XM) { x = 0; z = i16; }
        else if (f == FACE_XP) { x = 15; z = i16; }
        else if (f == FACE_ZM) { x = i16; z = 0; }
        else { x = i16; z = 15; }
        unsigned long long a = any ? (unsigned long long)a0[ly * 256 + z * 16 + x] << (4 * i16) : 0ull;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) a |= __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (i16 == 0) {
            W.faceA_of(p, f)[y0 + ly] = a;
            if (a) flags |= SF_FACE0 << f;
        }
    }
    // ---- bottom / top layer of the slab: 2 x 16 z-rows
    for (int r = tid >> 4; r < 2 * 16; r += 16) {
        const int top = r >> 4, z = r & 15, x = i16;
        const int ly = top ? SLAB_H - 1 : 0;
        unsigned long long a = any ? (unsigned long long)a0[ly * 256 + z * 16 + x] << (4 * x) : 0ull;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) a |= __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (x == 0) {
            W.layerA_of(p, 2 * s + top)[z] = a;
            if (a) flags |= top ? SF_TOP : SF_BOTTOM;
        }
    }
    if (flags) atomicOr(&s_flags, flags);
    // ---- the seed nibble plane (2 voxels per byte), only when it is non-zero
    if (any) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(W.a0_of(p) + (size_t)s * (SLAB_H * 128));
        for (int u = tid; u < SLAB_H * 256 / 8; u += 256) {  // 8 voxels -> one u32
            const uint2 b = *reinterpret_cast<const uint2*>(a0 + u * 8);
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                w |= ((b.x >> (8 * j)) & 0xFu) << (4 * j);
                w |= ((b.y >> (8 * j)) & 0xFu) << (16 + 4 * j);
            }
            dst[u] = w;
        }
    }
    if (virt) {
        // the slab as it stands (sky columns + emitters) and its rows of the border planes
        uint4* dst = reinterpret_cast<uint4*>(Lg);
        const uint4* lsrc = reinterpret_cast<const uint4*>(Ls);
#pragma unroll
        for (int k = 0; k < SLAB_H * 256 / 8 / 256; k++) dst[tid + 256 * k] = lsrc[tid + 256 * k];
        for (int t = tid; t < 4 * SLAB_H * 16; t += 256) {
            const int f = t >> 9, ly = (t >> 4) & 31, i = t & 15;
            int x, z;
            if (f == FACE_XM) { x = 0; z = i; }
            else if (f == FACE_XP) { x = 15; z = i; }
            else if (f == FACE_ZM) { x = i; z = 0; }
            else { x = i; z = 15; }
            W.faceL_of(p, f)[(y0 + ly) * 16 + i] = Ls[ly * 256 + z * 16 + x];
        }
    }
    __syncthreads();
    if (face_seed) atomicOr(&W.meta[p].pad1, (int)face_seed);
    if (tid == 0) {
        W.sflag_of(p)[s] = (uint8_t)s_flags;
        if (any_all) atomicOr(&W.meta[p].any_act, 1);
    }
}

// ---------------------------------------------------------------------------
// k_solve: grid (NSLAB, n_solve), 256 threads.
// ---------------------------------------------------------------------------
constexpr int PW = 18;              // padded row
constexpr int PL = PW * PW;         // padded layer (324)
constexpr int PSZ = (SLAB_H + 2) * PL;

__device__ __forceinline__ int pidx(int ly, int z, int x) {
    return (ly + 1) * PL + (z + 1) * PW + (x + 1);
}

// One relaxation of voxel (light `own`, cell c of the padded P plane): returns
// the new light (== own when nothing rises) and the raised-channel nibble mask.
__device__ __forceinline__ uint32_t relax(const uint16_t* P, int c, uint32_t own, uint32_t* raised) {
    const uint32_t a = P[c - 1], b = P[c + 1], d = P[c - PW], e = P[c + PW], f = P[c - PL],
                   g = P[c + PL];
    if (((a | b | d | e | f | g | own) & 0x0FFFu) == 0) {
        // sky channel only: the packed values order like their S nibbles
        const uint32_t mx = max(max(max(a, b), max(d, e)), max(f, g));
        const uint32_t m1 = mx >= 0x1000u ? mx - 0x1000u : 0u;
        *raised = m1 > own ? 0xF000u : 0u;
        return m1 > own ? m1 : own;
    }
    uint32_t x = nMAX(nE(a), nE(b));
    x = nMAX(x, nMAX(nE(d), nE(e)));
    x = nMAX(x, nMAX(nE(f), nE(g)));
    const uint32_t m1 = nDEC(x);
    const uint32_t gt = nGE(m1, nE(own) + 0x01010101u);  // m1 > own
    const uint32_t rm = nC(gt);
    *raised = rm;
    return (own & ~rm) | (nC(m1) & rm);
}

// position of the n-th (0-based) set bit of m; m has more than n bits set
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const uint32_t low = m & ((1u << w) - 1u);
        const int c = __popc(low);
        if (n >= c) {
            n -= c;
            m >>= w;
            pos += w;
        } else {
            m = low;
        }
    }
    return pos;
}

// The solver's work queue (below) and the per-build counter block.
struct WorkLists {
    int32_t* list;    // ring of 2 * stride slots; slot value = item + 1 (item = pool * NSLAB + slab), 0 = empty
    int32_t* count;   // the build's counter block (Q_* below)
    int32_t* dirty;   // [pool_cap * NSLAB] state of every slab (ST_*)
    int32_t stride;   // pool_cap * NSLAB
};

// k_solve_async.  The fixpoint does not depend on the order in which slabs are
// solved, so nothing forces level-synchronous passes ("solve every dirty slab,
// grid barrier, repeat": every CTA then waits at each barrier for the slowest
// slab, a fifth of the solve on a terrain world).  Instead: one ring queue
// (wl.list[0 .. 2 * stride), slot value = item + 1, 0 = empty) that running slabs
// append to, and a four-state word per slab (wl.dirty) that keeps a slab from
// being queued twice or solved by two CTAs at once:
//   IDLE -mark-> QUEUED -pop-> RUNNING -mark-> RUNNING_DIRTY -finish-> QUEUED (again)
//                              RUNNING -finish-> IDLE
// count[Q_PENDING] = slabs queued or running; a CTA that finds the queue empty
// leaves when it is 0.  A slab's publications are fenced before its marks, and a
// consumer fences between taking RUNNING and loading its halo (through L2), so
// either the consumer sees the publication or the producer sees RUNNING and
// marks it dirty (store-buffering with sequentially consistent fences).
constexpr int Q_HEAD = 8, Q_TAIL = 9, Q_PENDING = 10, Q_ERR = 11, Q_DONE = 12;
constexpr int ST_IDLE = 0, ST_QUEUED = 1, ST_RUNNING = 2, ST_RUNNING_DIRTY = 3;

__device__ __forceinline__ void q_push(const WorkLists& wl, int item) {
    atomicAdd(&wl.count[Q_PENDING], 1);
    const int at = atomicAdd(&wl.count[Q_TAIL], 1);
    atomicExch(&wl.list[at % (2 * wl.stride)], item + 1);
}

__device__ __forceinline__ void mark_dirty(const WorkLists& wl, int item) {
    int* st = &wl.dirty[item];
    int s = *reinterpret_cast<volatile int*>(st);
    while (true) {
        if (s == ST_QUEUED || s == ST_RUNNING_DIRTY) return;
        const int old = atomicCAS(st, s, s == ST_IDLE ? ST_QUEUED : ST_RUNNING_DIRTY);
        if (old == s) {
            if (s == ST_IDLE) q_push(wl, item);
            return;
        }
        s = old;
    }
}

// thread 0 of a CTA: the next slab to solve, or -1 when the build is complete.
// A consumer takes a ticket with one fetch-and-add and waits for *its* ring slot:
// no compare-and-swap on the head (hundreds of CTAs retrying one serialised the
// pops at one per memory round trip, ~0.7 us) and every waiter polls its own
// word.  The ring is 2 * stride slots: queued slabs (each at most once, so at
// most stride) plus the tickets of waiting CTAs never wrap onto each other.
__device__ int q_pop(const WorkLists& wl) {
    const int ticket = atomicAdd(&wl.count[Q_HEAD], 1);
    volatile int* slot = wl.list + ticket % (2 * wl.stride);
    const long long t0 = clock64();
    unsigned backoff = 32;
    for (int spin = 0;; spin++) {
        const int v = *slot;
        if (v) {
            *slot = 0;
            return v - 1;
        }
        // a push raises `pending` before it writes the slot and nothing lowers it until the
        // slab has been solved, so pending == 0 here means no slab will ever reach this slot
        if ((spin & 3) == 3) {
            if (*reinterpret_cast<volatile int*>(&wl.count[Q_PENDING]) == 0) return -1;
            if (*reinterpret_cast<volatile int*>(&wl.count[Q_ERR])) return -1;
        }
        __nanosleep(backoff);
        if (backoff < 1024) backoff *= 2;
        if (clock64() - t0 > 4000000000ll) {  // ~2 s without work while slabs are pending: a bug, not a wait
            atomicExch(&wl.count[Q_ERR], 1);
            return -1;
        }
    }
}

// first work list: a slab is active iff it holds a (non-inert) seed or one
// touches it from a neighbour slab
__global__ void k_worklist0(DevWorld W, const int32_t* solve, int n_solve, WorkLists wl) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_solve * NSLAB) return;
    const int p = solve[t / NSLAB], s = t % NSLAB;
    const ChunkMeta m = W.meta[p];
    const uint8_t* sf = W.sflag_of(p);
    bool act = sf[s] & SF_ANY;
    if (s > 0 && (sf[s - 1] & SF_TOP)) act = true;
    if (s < NSLAB - 1 && (sf[s + 1] & SF_BOTTOM)) act = true;
    const int qx[4] = {m.cx - 1, m.cx + 1, m.cx, m.cx}, qz[4] = {m.cz, m.cz, m.cz - 1, m.cz + 1};
    int nb[4];
    for (int f = 0; f < 4; f++) {
        const int q = W.pool_of(qx[f], qz[f]);
        nb[f] = q >= 0 && W.meta[q].pad0 ? q : -1;  // chunks outside the solve set do not take part
        if (nb[f] >= 0 && (W.sflag_of(q)[s] & (SF_FACE0 << (f ^ 1)))) act = true;
    }
    if (s == 0) W.nbr[p] = make_int4(nb[0], nb[1], nb[2], nb[3]);  // read by every solve of the chunk's slabs
#ifdef VX_SOLVE_STATS
    if (act) atomicAdd(&wl.count[16 + s], 1);
    if (sf[s] & SF_ANY) atomicAdd(&wl.count[48 + s], 1);
#endif
    if (act) mark_dirty(wl, p * NSLAB + s);
}

struct SolveShared {
    uint16_t Ls[SLAB_H * 256];
    uint16_t Pp[PSZ];
    uint32_t lpw[SLAB_WORDS], candw[SLAB_WORDS], cinit[SLAB_WORDS];
    uint32_t chg[2][SLAB_WORDS];
    uint16_t list[SLAB_WORDS];
    int nlist[2];
    uint32_t layers_changed;
    uint32_t marks;  // neighbours to queue once everything is published: 4 faces, below, above
    uint64_t bar;    // mbarrier of the slab's bulk copy (one phase per solve)
};

#ifdef VX_SOLVE_PROFILE
#define VX_PHASE(k)                                                                                     \
    do {                                                                                                \
        if (threadIdx.x == 0) {                                                                         \
            const long long now_ = clock64();                                                           \
            atomicAdd(reinterpret_cast<unsigned long long*>(wl.count + 24 + 2 * (k)), (unsigned long long)(now_ - tph_)); \
            tph_ = now_;                                                                                \
        }                                                                                               \
    } while (0)
#else
#define VX_PHASE(k) do {} while (0)
#endif

__device__ void solve_slab(const DevWorld& W, SolveShared& S, const WorkLists& wl, int p, int s, uint32_t bar_parity,
                           bool slab_in_flight = false) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef VX_SOLVE_PROFILE
    long long tph_ = clock64();
#endif
    // the four side neighbours taking part in this build (k_worklist0): one load instead of the
    // chunk record, four slot-table look-ups and four in-set flags
    const int4 nb4 = __ldg(W.nbr + p);
    int q[4];
    q[FACE_XM] = nb4.x;
    q[FACE_XP] = nb4.y;
    q[FACE_ZM] = nb4.z;
    q[FACE_ZP] = nb4.w;
    // the slab's own seeds are taken by its first solve (the flag byte is cleared below);
    // later solves restart from the halo only
    const uint32_t own_flags = __ldcg(W.sflag_of(p) + s);
    uint16_t* const Ls = S.Ls;
    uint16_t* const Pp = S.Pp;
    uint32_t* const lpw = S.lpw;
    uint32_t* const candw = S.candw;
    uint32_t* const cinit = S.cinit;
    uint32_t (*const chg)[SLAB_WORDS] = S.chg;
    uint16_t* const list = S.list;
    int* const nlist = S.nlist;
    uint32_t& layers_changed = S.layers_changed;
    // scratch that is only live outside the rounds, aliased onto round-loop buffers
    unsigned long long* const rowA = reinterpret_cast<unsigned long long*>(S.candw);  // 128 rows
    unsigned long long* const layA = reinterpret_cast<unsigned long long*>(S.list);   // 32 rows

    const int y0 = s * SLAB_H;
    vx_light* Lg = W.light_of(p) + (size_t)y0 * 256;
    {
        // every load of the phase is issued before anything waits on one: the slab's light (4 x 16
        // bytes per thread), its lp word, and one activation row of the halo per thread — 128
        // border rows of the four neighbours and 32 rows of the slabs below / above (candw / list
        // are free until the first round).  Other SMs write those planes while this kernel runs:
        // they are read from L2, never from L1.
        // the slab's light: 16 KiB contiguous, one bulk copy (TMA) issued by one thread straight into
        // shared memory, through L2 (other SMs solved this slab before); the caller's barrier
        // guarantees nobody still reads Ls
        if (tid == 0 && !slab_in_flight) bulk_load(Ls, Lg, SLAB_H * 256 * sizeof(vx_light), &S.bar);
        const uint32_t lpv = W.lp_of(p)[s * SLAB_WORDS + tid];
        unsigned long long rowv = 0ull;
        if (tid < 4 * SLAB_H) {
            const int f = tid >> 5, qq = q[f];
            // (faceA of a chunk outside this build's solve set is left over from an earlier build:
            // such neighbours are -1 in the table)
            if (qq >= 0) rowv = __ldcg(W.faceA_of(qq, f ^ 1) + y0 + (tid & 31));
        } else if (tid < 4 * SLAB_H + 32) {
            const int r = tid - 4 * SLAB_H, up = r >> 4, z = r & 15;
            const int y = up ? y0 + SLAB_H : y0 - 1;
            // boundary layer of the neighbouring slab: its top (li odd) when below us
            const int li = up ? 2 * (s + 1) : 2 * (s - 1) + 1;
            if (y >= 0 && y < CH) rowv = __ldcg(W.layerA_of(p, li) + z);
        }
        uint32_t* pz = reinterpret_cast<uint32_t*>(Pp);
        for (int u = tid; u < PSZ / 2; u += 256) pz[u] = 0;
        chg[0][tid] = 0;
        chg[1][tid] = 0;
        cinit[tid] = 0;
        lpw[tid] = lpv;
        if (tid < 4 * SLAB_H) rowA[tid] = rowv;
        else if (tid < 4 * SLAB_H + 32) layA[tid - 4 * SLAB_H] = rowv;
        if (tid == 0) {
            layers_changed = 0;
            nlist[0] = 0;
            nlist[1] = 0;
            S.marks = 0;
            if (own_flags) W.sflag_of(p)[s] = 0;
        }
        mbar_wait(&S.bar, bar_parity);  // the slab has landed
    }
    __syncthreads();
    VX_PHASE(0);  // load

    // ---- halo of propagating light; own voxels next to a lit halo cell become
    // the first candidates
    for (int t = tid; t < 4 * SLAB_H * 16; t += 256) {
        const int f = t >> 9, ly = (t >> 4) & 31, i = t & 15;
        const uint32_t a = (uint32_t)(rowA[t >> 4] >> (4 * i)) & 0xFu;
        if (!a) continue;
        const int qq = q[f];
        const int of = f ^ 1;  // the neighbour's face that touches ours
        const int y = y0 + ly;
        const uint32_t pv = __ldcg(W.faceL_of(qq, of) + y * 16 + i) & bits_to_mask(a);
        if (!pv) continue;
        int x, z, ox, oz;  // halo cell and the own voxel it touches
        if (f == FACE_XM) { x = -1; z = i; ox = 0; oz = i; }
        else if (f == FACE_XP) { x = 16; z = i; ox = 15; oz = i; }
        else if (f == FACE_ZM) { x = i; z = -1; ox = i; oz = 0; }
        else { x = i; z = 16; ox = i; oz = 15; }
        Pp[pidx(ly, z, x)] = (uint16_t)pv;
        // a first candidate only if this cell can actually raise the voxel it touches
        const uint32_t ol = Ls[ly * 256 + oz * 16 + ox];
        if (~nGE(nE(ol), nDEC(nE(pv))) & 0x0F0F0F0Fu)
            atomicOr(&cinit[ly * 8 + (oz >> 1)], 1u << ((oz & 1) * 16 + ox));
    }
    for (int t = tid; t < 2 * 256; t += 256) {
        const int up = t >> 8, z = (t >> 4) & 15, x = t & 15;
        const int y = up ? y0 + SLAB_H : y0 - 1;
        const uint32_t a = (uint32_t)(layA[up * 16 + z] >> (4 * x)) & 0xFu;
        if (!a) continue;
        const uint32_t pv = __ldcg(W.light_of(p) + vidx(x, y, z)) & bits_to_mask(a);
        if (!pv) continue;
        Pp[pidx(up ? SLAB_H : -1, z, x)] = (uint16_t)pv;
        const int ly = up ? SLAB_H - 1 : 0;
        const uint32_t ol = Ls[ly * 256 + z * 16 + x];
        if (~nGE(nE(ol), nDEC(nE(pv))) & 0x0F0F0F0Fu)
            atomicOr(&cinit[ly * 8 + (z >> 1)], 1u << ((z & 1) * 16 + x));
    }
    // ---- own seeds (first launch only; later launches restart from the halo):
    // thread t owns word t = voxels 32t .. 32t+31 = one uint4 of seed nibbles
    if (own_flags & SF_ANY) {
        const uint4 a4 = reinterpret_cast<const uint4*>(W.a0_of(p) + (size_t)s * (SLAB_H * 128))[tid];
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
        uint32_t cm = 0;
        if (a4.x | a4.y | a4.z | a4.w) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const uint32_t a = (aw[j >> 3] >> (4 * (j & 7))) & 0xFu;
                if (a) {
                    const int v = tid * 32 + j;
                    const uint32_t pv = Ls[v] & bits_to_mask(a);
                    if (pv) {
                        Pp[pidx(v >> 8, (v >> 4) & 15, v & 15)] = (uint16_t)pv;
                        cm |= 1u << j;
                    }
                }
            }
        }
        chg[0][tid] = cm;
    }

    VX_PHASE(1);  // halo + seeds
    int cur = 0;
    uint32_t any_change = 0;
    uint32_t lc_acc = 0;  // layers this thread changed; merged into layers_changed after the rounds
    bool first = true;
    // ---- sparse wavefront rounds
    while (true) {
        __syncthreads();  // previous round's writes (light, P, chg) are visible
        {
            if (tid == 0) nlist[cur ^ 1] = 0;  // counter of the next round
            const uint32_t* cc = chg[cur];
            const int ly = tid >> 3, zp = tid & 7;
            const uint32_t c = cc[tid];
            const uint32_t up = ly < SLAB_H - 1 ? cc[tid + 8] : 0u;
            const uint32_t dn = ly > 0 ? cc[tid - 8] : 0u;
            const uint32_t nzm = (c << 16) | (zp > 0 ? cc[tid - 1] >> 16 : 0u);
            const uint32_t nzp = (c >> 16) | (zp < 7 ? cc[tid + 1] << 16 : 0u);
            const uint32_t xl = (c << 1) & 0xFFFEFFFEu;
            const uint32_t xr = (c >> 1) & 0x7FFF7FFFu;
            uint32_t cand = up | dn | nzm | nzp | xl | xr;
            if (first) cand |= cinit[tid];
            cand &= lpw[tid];
            candw[tid] = cand;
            chg[cur ^ 1][tid] = 0;
            const uint32_t nzb = __ballot_sync(0xFFFFFFFFu, cand != 0);
            int base = 0;
            if (lane == 0 && nzb) base = atomicAdd(&nlist[cur], __popc(nzb));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (cand) list[base + __popc(nzb & ((1u << lane) - 1u))] = (uint16_t)tid;
        }
        first = false;
#ifdef VX_SOLVE_PROFILE
        if (tid == 0) atomicAdd(&wl.count[40], 1);
#endif
        __syncthreads();
        const int n = nlist[cur];
        if (n == 0) break;
        // A round's candidates are few (configs[1]: 28 words, 130 voxels on average), and a word
        // holds 4-5 of them: one word per warp step would leave 85 % of the lanes idle and cost
        // 3-4 dependent steps per round.  Instead the warp packs the candidate voxels of all its
        // words (list[warp], list[warp + 8], ...: lane i owns the i-th) into dense lanes: a scan
        // of the per-word counts, then lane j finds its word by binary search over the scan and
        // its voxel as the k-th set bit of that word's mask.
        {
            const int nw = (n - warp + 7) >> 3;
            int myw = 0;
            uint32_t mycand = 0;
            if (lane < nw) {
                myw = list[warp + 8 * lane];
                mycand = candw[myw];
            }
            const int cnt = __popc(mycand);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            for (int j0 = 0; j0 < total; j0 += 32) {
                const int j = min(j0 + lane, total - 1);
                int wi = 0;  // smallest lane whose inclusive count exceeds j
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int val = __shfl_sync(0xFFFFFFFFu, incl, wi + step - 1);
                    if (val <= j) wi += step;
                }
                const int w = __shfl_sync(0xFFFFFFFFu, myw, wi);
                const uint32_t cd = __shfl_sync(0xFFFFFFFFu, mycand, wi);
                const int before = __shfl_sync(0xFFFFFFFFu, incl - cnt, wi);
                if (j0 + lane < total) {
                    const int bit = nth_set_bit(cd, j - before);
                    const int v = w * 32 + bit;
                    const int c = pidx(v >> 8, (v >> 4) & 15, v & 15);
                    const uint32_t own = Ls[v];
                    uint32_t rm;
                    const uint32_t nl = relax(Pp, c, own, &rm);
                    if (rm) {
                        Ls[v] = (uint16_t)nl;
                        Pp[c] = (uint16_t)((Pp[c] & ~rm) | (nl & rm));
                        atomicOr(&chg[cur ^ 1][w], 1u << bit);
                        lc_acc |= 1u << (w >> 3);
                        any_change = 1;
                    }
                }
            }
        }
        cur ^= 1;
    }
    // `nlist == 0` was read after a barrier by every thread: uniform exit.
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) lc_acc |= __shfl_xor_sync(0xFFFFFFFFu, lc_acc, o);
    if (lane == 0 && lc_acc) atomicOr(&layers_changed, lc_acc);
    any_change = __syncthreads_or(any_change != 0);
    VX_PHASE(2);  // rounds
    if (!any_change) return;
    if (tid == 0) atomicOr(&W.meta[p].any_act, 1);  // raised voxels are activated

    // ---- publish: light layers that changed
    const uint32_t lc = layers_changed;
    {
        uint4* dst = reinterpret_cast<uint4*>(Lg);
        const uint4* src = reinterpret_cast<const uint4*>(Ls);
        for (int u = tid; u < SLAB_H * 256 / 8; u += 256)
            if ((lc >> (u >> 5)) & 1u) dst[u] = src[u];
    }
    // border planes + activated bits.  The neighbour slab is re-queued only when a
    // row changed AND some voxel of it can still raise its neighbour across the
    // border (activated light at least 2 above what the neighbour holds): two
    // slabs that reach the same light on both sides of a border (open water, sky
    // shafts) would otherwise keep re-solving each other for nothing.
    // Everything this needs from HBM is fetched up front, one latency instead of one per row:
    // the slab's own activation rows and the in-set flags (shared memory: the round buffers
    // are dead), and the eight border-light values each thread compares against.
    // Rows of layers in which nothing was raised have nothing to publish: their light is what
    // the border planes hold and their activated voxels (seeds) are in faceA since k_seed.
    uint32_t oldl[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int r = (tid >> 4) + 16 * k;
        oldl[k] = ((lc >> (r & 31)) & 1u) ? __ldcg(W.faceL_of(p, r >> 5) + (y0 + (r & 31)) * 16 + (tid & 15)) : 0u;
    }
    if (tid < 4 * SLAB_H) rowA[tid] = ((lc >> (tid & 31)) & 1u) ? __ldcg(W.faceA_of(p, tid >> 5) + y0 + (tid & 31)) : 0ull;
    if (tid < 4) reinterpret_cast<int*>(layA)[tid] = q[tid] >= 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int r = (tid >> 4) + 16 * k;
        const int f = r >> 5, ly = r & 31, i = tid & 15;
        int x, z;
        if (f == FACE_XM) { x = 0; z = i; }
        else if (f == FACE_XP) { x = 15; z = i; }
        else if (f == FACE_ZM) { x = i; z = 0; }
        else { x = i; z = 15; }
        const int y = y0 + ly;
        const bool rowchg = (lc >> ly) & 1u;
        const uint32_t nl = Ls[ly * 256 + z * 16 + x];
        const uint32_t pv = rowchg ? Pp[pidx(ly, z, x)] : 0u;
        bool diff = false;
        if (rowchg && oldl[k] != nl) {
            W.faceL_of(p, f)[y * 16 + i] = (vx_light)nl;
            diff = true;
        }
        unsigned long long a = (unsigned long long)nonzero_bits(pv) << (4 * i);
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) a |= __shfl_xor_sync(0xFFFFFFFFu, a, o);
        const unsigned long long olda = rowA[r];  // same value for the 16 lanes of the row
        if ((a | olda) != olda) {
            diff = true;
            if (i == 0) atomicOr(W.faceA_of(p, f) + y, a);  // never drops bits an earlier solve on another SM set
        }
        const bool in_set = reinterpret_cast<const int*>(layA)[f] != 0;
        bool live = false;
        if (in_set && pv) {
            // the voxel across the border: it must let light through (a third of all re-solves
            // used to be for solid neighbours and raised nothing), and hold less than this voxel
            // can give it (a stale read is lower: conservative)
            const int nx = f == FACE_XM ? 15 : f == FACE_XP ? 0 : x, nz = f == FACE_ZM ? 15 : f == FACE_ZP ? 0 : z;
            const int ni = vidx(nx, y, nz);
            if ((__ldg(W.lp_of(q[f]) + (ni >> 5)) >> (ni & 31)) & 1u) {
                const uint32_t hl = __ldcg(W.faceL_of(q[f], f ^ 1) + y * 16 + i);
                live = (~nGE(nE(hl), nDEC(nE(pv))) & 0x0F0F0F0Fu) != 0;
            }
        }
        const uint32_t grp = 0xFFFFu << (lane & 16);
        const bool rowdiff = __ballot_sync(0xFFFFFFFFu, diff) & grp;
        const bool rowlive = __ballot_sync(0xFFFFFFFFu, live) & grp;
        if (rowdiff && rowlive && i == 0) atomicOr(&S.marks, 1u << f);
    }
    // slab boundary layers: same rule towards the slab below / above
    for (int r = tid >> 4; r < 2 * 16; r += 16) {
        const int top = r >> 4, z = r & 15, x = tid & 15;
        const int ly = top ? SLAB_H - 1 : 0;
        const uint32_t pv = Pp[pidx(ly, z, x)];
        unsigned long long a = (unsigned long long)nonzero_bits(pv) << (4 * x);
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) a |= __shfl_xor_sync(0xFFFFFFFFu, a, o);
        if (x == 0 && a) atomicOr(W.layerA_of(p, 2 * s + top) + z, a);
        const int ny = top ? y0 + SLAB_H : y0 - 1;
        bool live = false;
        if (pv && ny >= 0 && ny < CH && ((lc >> ly) & 1u)) {
            const int ni = ny * 256 + z * 16 + x;
            if ((W.lp_of(p)[ni >> 5] >> (ni & 31)) & 1u) {
                const uint32_t hl = __ldcg(W.light_of(p) + ni);
                live = (~nGE(nE(hl), nDEC(nE(pv))) & 0x0F0F0F0Fu) != 0;
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, live) && lane == 0) atomicOr(&S.marks, top ? 32u : 16u);
    }
    VX_PHASE(3);  // write back + publish
    // ---- queue the neighbours, after everything above is visible to them: the block's writes
    // are ordered before the barrier, the marking thread's fence after it carries them to the
    // device (fences are cumulative; this is how a cooperative-groups grid barrier publishes)
    __syncthreads();
    if (tid < 6 && ((S.marks >> tid) & 1u)) {
        __threadfence();
        mark_dirty(wl, tid < 4 ? q[tid] * NSLAB + s : p * NSLAB + (tid == 5 ? s + 1 : s - 1));
    }
    VX_PHASE(4);  // fence + marks
}

// The whole solve in one launch without iterations (see WorkLists).  Cooperative
// launch only because spinning CTAs must all be resident.
__global__ void __launch_bounds__(256, 5) k_solve_async(const __grid_constant__ DevWorld W, WorkLists wl) {
    __shared__ __align__(16) SolveShared S;
    __shared__ int s_item;
    if (threadIdx.x == 0) mbar_init(&S.bar, 1);
    uint32_t parity = 0;
    while (true) {
        __syncthreads();  // the previous slab's shared state (and s_item) is dead
        if (threadIdx.x == 0) {
            const int item = q_pop(wl);
            if (item >= 0) {
                atomicExch(&wl.dirty[item], ST_RUNNING);
                __threadfence();
                // the slab starts moving before the other threads even know which one it is
                bulk_load(S.Ls, W.light_of(item / NSLAB) + (size_t)(item % NSLAB) * SLAB_H * 256,
                          SLAB_H * 256 * sizeof(vx_light), &S.bar);
            }
            s_item = item;
        }
        __syncthreads();
        const int item = s_item;
        if (item < 0) break;
#ifdef VX_SOLVE_STATS
        if (threadIdx.x == 0) atomicAdd(&wl.count[56 + item % NSLAB], 1);
#endif
        solve_slab(W, S, wl, item / NSLAB, item % NSLAB, parity, true);
        parity ^= 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicCAS(&wl.dirty[item], ST_RUNNING, ST_IDLE) == ST_RUNNING_DIRTY) {
                atomicExch(&wl.dirty[item], ST_QUEUED);  // its halo changed while it ran: once more
                q_push(wl, item);
            }
            atomicAdd(&wl.count[Q_DONE], 1);
            atomicSub(&wl.count[Q_PENDING], 1);
        }
    }
}

// k_light_end: Chunk::flags.modified as LightSolver leaves it: add() marks the
// seed's chunk, and every popped entry marks the chunks of its six neighbours
// (LightSolver.cpp:29,111).  Every activated voxel is popped once, so a chunk is
// modified iff it holds an activated voxel or a neighbour's facing border plane
// does.  Also flags.lighted for the lit chunks (ChunksController.cpp:127).
__global__ void __launch_bounds__(256) k_light_end(DevWorld W, const int32_t* solve) {
    const int p = solve[blockIdx.x];
    const int tid = threadIdx.x;
    ChunkMeta* m = &W.meta[p];
    int q[4];
    q[FACE_XM] = W.pool_of(m->cx - 1, m->cz);
    q[FACE_XP] = W.pool_of(m->cx + 1, m->cz);
    q[FACE_ZM] = W.pool_of(m->cx, m->cz - 1);
    q[FACE_ZP] = W.pool_of(m->cx, m->cz + 1);
    int any = 0;
    for (int f = 0; f < 4; f++) {
        if (q[f] < 0 || !W.meta[q[f]].pad0) continue;
        if ((W.meta[q[f]].pad1 >> (f ^ 1)) & 1) any = 1;
        if (W.faceA_of(q[f], f ^ 1)[tid] != 0ull) any = 1;
    }
    any = __syncthreads_or(any);
    if (tid == 0) {
        if (any || m->any_act) m->modified = 1;
        if (m->lit_now) m->lighted = 1;
    }
}

__global__ void k_light_reset(DevWorld W, const int32_t* solve, int n_solve) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_solve) {
        ChunkMeta* m = &W.meta[solve[t]];
        m->lit_now = 0;
        m->pad0 = 0;
        m->pad1 = 0;
        m->any_act = 0;
        m->lstate = LS_UNKNOWN;  // the build may have written any chunk of the solve set
        m->lvirt = 0;            // k_seed wrote out every lightmap of the solve set
    }
}

__global__ void k_clear_modified(DevWorld W) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < W.pool_cap) W.meta[t].modified = 0;
}

std::vector<int> pools_of(vx_ctx* ctx, const vx_chunk_key* keys, int n) {
    std::vector<int> out;
    out.reserve(n);
    for (int i = 0; i < n; i++) {
        int p = ctx->pool_of(keys[i].cx, keys[i].cz);
        if (p >= 0) out.push_back(p);
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

}  // namespace

int vx_derive_chunks(vx_ctx* ctx, const std::vector<int>& pools, bool light_given) {
    if (pools.empty()) return 0;
    if (vx_upload_list(ctx, pools, 0)) return -1;
    VX_LAUNCH(ctx, KID_DERIVE, k_derive<<<(unsigned)pools.size(), 1024, 0, ctx->stream>>>(ctx->W, ctx->d_list));
    if (light_given)
        VX_LAUNCH(ctx, KID_FACES,
                  k_faces_refresh<<<(unsigned)pools.size(), 256, 0, ctx->stream>>>(ctx->W, ctx->d_list));
    VX_CUDA(cudaGetLastError());
    return 0;
}

int vx_light_init(vx_ctx* ctx) {
    // the solver is one cooperative launch: CTAs that wait for work must all be resident
    ctx->coop_grid_async = 0;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
    int per_sm = 0;
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve_async, 256, 0) == cudaSuccess && per_sm > 0)
        ctx->coop_grid_async = std::min(per_sm, VX_SOLVE_PER_SM) * ctx->n_sm;
    cudaGetLastError();
    if (ctx->coop_grid_async <= 0) {
        ctx->fail("vxgpu: the device does not support cooperative launches (needed by the light solver)");
        return -1;
    }
    return 0;
}

int vx_materialize(vx_ctx* ctx, const std::vector<int>& pools) {
    std::vector<int> need;
    for (int p : pools)
        if (p >= 0 && ctx->h_meta[p].lvirt) need.push_back(p);
    if (need.empty()) return 0;
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
    if (vx_upload_list(ctx, need, 0)) return -1;
    const unsigned n = (unsigned)need.size();
    VX_LAUNCH(ctx, KID_CLEAR, k_materialize<<<dim3(NSLAB, n), 256, 0, ctx->stream>>>(ctx->W, ctx->d_list));
    VX_LAUNCH(ctx, KID_RESET, k_devirt<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->W, ctx->d_list, (int)n));
    VX_CUDA(cudaGetLastError());
    for (int p : need) ctx->h_meta[p].lvirt = 0;
    return 0;  // stream-ordered (list 0 is only rewritten by calls that synchronise or queue behind this)
}

// The host side of the end of a light build (see vx_light_build): waits for it, takes its device time and the
// verdict of the solver's queue.  Called by whatever synchronises next or needs the answer; 0 when nothing is
// pending.
int vx_light_finish(vx_ctx* ctx) {
    if (!ctx->light_pending) return 0;
    ctx->light_pending = false;
    cudaStream_t st = ctx->stream;
    VX_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->last_light_ms, ctx->ev_l0, ctx->ev_l1);
    const int32_t* v = ctx->h_light_verdict;
    if (getenv("VXGPU_DEBUG")) {
        fprintf(stderr, "[vxgpu] async solve: %d slab solves\n", v[Q_DONE]);
#ifdef VX_SOLVE_STATS
        {
            int32_t h[64];
            cudaMemcpy(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[vxgpu]   per slab index: first queue");
            for (int k = 0; k < NSLAB; k++) fprintf(stderr, " %d", h[16 + k]);
            fprintf(stderr, " | own seeds");
            for (int k = 0; k < NSLAB; k++) fprintf(stderr, " %d", h[48 + k]);
            fprintf(stderr, " | solves");
            for (int k = 0; k < NSLAB; k++) fprintf(stderr, " %d", h[56 + k]);
            fprintf(stderr, "\n");
        }
#endif
#ifdef VX_SOLVE_PROFILE
        int32_t h[48];
        cudaMemcpy(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[5] = {"load", "halo+seeds", "rounds", "publish", "fence"};
        for (int k = 0; k < 5; k++)
            fprintf(stderr, "[vxgpu]   %-10s %8.1f kcycles per solve\n", names[k],
                    *(unsigned long long*)(h + 24 + 2 * k) / 1e3 / std::max(1, v[Q_DONE]));
        fprintf(stderr, "[vxgpu]   rounds per solve %.1f\n", (double)h[40] / std::max(1, v[Q_DONE]));
#endif
    }
    if (v[Q_ERR] || v[Q_PENDING] != 0) {
        // leave the queue clean for the next build; the lightmaps are not complete
        const size_t stride = (size_t)ctx->light_pending_stride;
        cudaMemsetAsync(ctx->d_wl, 0, stride * 3 * sizeof(int32_t), st);
        cudaMemsetAsync(ctx->d_wl_dirty, 0, stride * 3 * sizeof(int32_t), st);
        cudaStreamSynchronize(st);
        ctx->fail("vx_light_build: solver queue stalled (watchdog)");
        return -1;
    }
    return 0;
}

int vx_zero_light(vx_ctx* ctx, const std::vector<int>& pools) {
    if (pools.empty()) return 0;
    if (vx_upload_list(ctx, pools, 0)) return -1;
    VX_LAUNCH(ctx, KID_CLEAR,
              k_clear<<<dim3(NSLAB, (unsigned)pools.size()), 256, 0, ctx->stream>>>(ctx->W, ctx->d_list));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaStreamSynchronize(ctx->stream));  // list 0 is reused by the caller
    return 0;
}

extern "C" {

int vx_light_prebuild(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    std::vector<int> pools = pools_of(ctx, keys, n);
    if (pools.empty()) return 0;
    if (vx_upload_list(ctx, pools, 2)) return -1;
    bool all_virtual = ctx->virt_on;
    for (int p : pools) all_virtual = all_virtual && ctx->h_meta[p].lvirt;
    if (all_virtual) {
        const unsigned nv = (unsigned)pools.size();
        VX_LAUNCH(ctx, KID_PREBUILD, k_prebuild_virtual<<<(nv + 255) / 256, 256, 0, ctx->stream>>>(
                                         ctx->W, ctx->d_list + 2 * (size_t)ctx->pool_cap, (int)nv));
        VX_CUDA(cudaGetLastError());
        return 0;
    }
    static const int track = getenv("VXGPU_NO_PRISTINE") ? 0 : 1;  // debugging: never take the pristine short cuts
    VX_LAUNCH(ctx, KID_PREBUILD, k_prebuild<<<dim3(NSLAB, (unsigned)pools.size()), 256, 0, ctx->stream>>>(
                                     ctx->W, ctx->d_list + 2 * (size_t)ctx->pool_cap, track));
    VX_CUDA(cudaGetLastError());
    return 0;  // stream-ordered; the next call that returns data synchronises
}

int vx_light_clear(vx_ctx* ctx) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    std::vector<int> pools;
    for (int p = 0; p < ctx->pool_cap; p++)
        if (ctx->h_meta[p].present) pools.push_back(p);
    if (pools.empty()) return 0;
    if (vx_upload_list(ctx, pools, 3)) return -1;
    if (ctx->virt_on) {
        const unsigned n = (unsigned)pools.size();
        VX_LAUNCH(ctx, KID_CLEAR, k_clear_virtual<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
                                      ctx->W, ctx->d_list + 3 * (size_t)ctx->pool_cap, (int)n));
        for (int p : pools) ctx->h_meta[p].lvirt = 1;
        VX_CUDA(cudaGetLastError());
        return 0;
    }
    VX_LAUNCH(ctx, KID_CLEAR, k_clear<<<dim3(NSLAB, (unsigned)pools.size()), 256, 0, ctx->stream>>>(
                                  ctx->W, ctx->d_list + 3 * (size_t)ctx->pool_cap));
    VX_CUDA(cudaGetLastError());
    return 0;  // stream-ordered
}

int vx_chunks_clear_modified(vx_ctx* ctx) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    if (ctx->pool_cap == 0) return 0;
    VX_LAUNCH(ctx, KID_CLEARMOD, k_clear_modified<<<(ctx->pool_cap + 255) / 256, 256, 0, ctx->stream>>>(ctx->W));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int vx_light_build(vx_ctx* ctx, const vx_chunk_key* keys, int32_t n, int32_t expand) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    std::vector<int> lit = pools_of(ctx, keys, n);
    if (lit.empty()) return 0;
    // solve set: the lit chunks plus their present 8-neighbours (light travels
    // at most 14 voxels from a seed, so it never leaves that ring)
    std::vector<uint8_t> mark(ctx->pool_cap, 0);
    std::vector<int> solve;
    for (int p : lit) {
        const ChunkMeta& m = ctx->h_meta[p];
        for (int dz = -1; dz <= 1; dz++)
            for (int dx = -1; dx <= 1; dx++) {
                int q = ctx->pool_of(m.cx + dx, m.cz + dz);
                if (q >= 0 && !mark[q]) {
                    mark[q] = 1;
                    solve.push_back(q);
                }
            }
    }
    std::sort(solve.begin(), solve.end());
    for (int p : solve) ctx->h_meta[p].lvirt = 0;  // k_seed writes out every lightmap of the solve set
    if (vx_upload_list(ctx, lit, 0)) return -1;
    if (vx_upload_list(ctx, solve, 1)) return -1;
    const int32_t* d_lit = ctx->d_list;
    const int32_t* d_solve = ctx->d_list + ctx->pool_cap;
    const unsigned n_lit = (unsigned)lit.size(), n_solve = (unsigned)solve.size();
    cudaStream_t st = ctx->stream;
    VX_CUDA(cudaEventRecord(ctx->ev_l0, st));
    VX_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 64 * sizeof(int32_t), st));
    VX_LAUNCH(ctx, KID_BEGIN, k_light_begin<<<(n_solve + 255) / 256, 256, 0, st>>>(
                                  ctx->W, d_lit, (int)n_lit, d_solve, (int)n_solve));
    if (expand & VX_LIGHT_NO_SKY)
        VX_LAUNCH(ctx, KID_YSTAR, k_ystar_none<<<n_lit, 256, 0, st>>>(ctx->W, d_lit));
    else
        VX_LAUNCH(ctx, KID_YSTAR, k_ystar<<<n_lit, 1024, 0, st>>>(ctx->W, d_lit));
    VX_LAUNCH(ctx, KID_SEED, k_seed<<<dim3(NSLAB, n_solve), 256, 0, st>>>(ctx->W, d_solve, expand & VX_LIGHT_EXPAND));
    WorkLists wl;
    wl.list = ctx->d_wl;
    wl.count = ctx->d_counters;
    wl.dirty = ctx->d_wl_dirty;
    wl.stride = ctx->pool_cap * NSLAB;
    VX_LAUNCH(ctx, KID_WL0, k_worklist0<<<(n_solve * NSLAB + 255) / 256, 256, 0, st>>>(
                                ctx->W, d_solve, (int)n_solve, wl));
    {
        // one launch, no iterations: slabs are solved as their halos change
        // never more CTAs than slabs: the ring holds the queued slabs (<= stride) plus one ticket per
        // waiting CTA (<= grid) in 2 * stride slots
        const int grid = std::min(ctx->coop_grid_async, std::max(1, std::min(wl.stride, (int)n_solve * NSLAB)));
        void* args[] = {(void*)&ctx->W, (void*)&wl};
        VX_LAUNCH(ctx, KID_SOLVE,
                  VX_CUDA(cudaLaunchCooperativeKernel((void*)k_solve_async, dim3(grid), dim3(256), args, 0, st)));
        // the queue's verdict is read with the call's single synchronisation at the end
        VX_CUDA(cudaMemcpyAsync(ctx->h_light_verdict, ctx->d_counters, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    VX_LAUNCH(ctx, KID_END, k_light_end<<<n_solve, 256, 0, st>>>(ctx->W, d_solve));
    VX_LAUNCH(ctx, KID_RESET,
              k_light_reset<<<(n_solve + 255) / 256, 256, 0, st>>>(ctx->W, d_solve, (int)n_solve));
    VX_CUDA(cudaEventRecord(ctx->ev_l1, st));
    VX_CUDA(cudaGetLastError());
    // No host synchronisation here: the kernels of whatever comes next (usually the mesh build) queue up
    // behind these.  What the solver's queue said is read by vx_light_finish.
    ctx->light_pending = true;
    ctx->light_pending_stride = wl.stride;
    static const bool debug = getenv("VXGPU_DEBUG") != nullptr;
    if (debug) return vx_light_finish(ctx);
    return 0;
}

}  // extern "C"
