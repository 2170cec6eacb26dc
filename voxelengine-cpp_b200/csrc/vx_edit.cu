// libvxgpu: block edits.  Replaces blocks_agent::set (voxels/blocks_agent.cpp:
// 10-77) + Lighting::onBlockSet (lighting/Lighting.cpp:163-215) with
// LightSolver::remove / solve (lighting/LightSolver.cpp:37-126).
//
// Unlike a full relight, an edit is NOT the from-scratch fixpoint (SURVEY §7
// hard part 1): the removal rules are reproduced literally.  One CTA runs one
// edit: the removal BFS is processed level-synchronously (all queue entries of
// light level l, then l-1, ...; empirically identical to the reference FIFO,
// SURVEY §8c probe 4), the add BFS in rounds; R, G, B and S entries share the
// queues (they touch different nibbles).  Light words are updated with 32-bit
// CAS, so the entries of a level can be processed by all threads at once.
//
// Where the light lives while the edit runs: a 33 x 64 x 33 box of the world around the edit
// (16 voxels either side in x / z, 47 below and 16 above) is copied into shared memory, the
// rounds run there — a round costs a barrier instead of three dependent trips to L2 — and the
// voxels that changed are written back.  An edit that reaches outside its box (a sky column
// falling further than 47 voxels, light flooding into a region an earlier edit left under-lit)
// is found out before anything was written: the attempt is dropped and the edit runs again on
// the lightmaps in global memory, which has no such bound.
//
// Edits whose regions of influence cannot overlap run concurrently: the host orders a batch
// (an edit comes after every earlier edit whose column is within |dx| + |dz| <= 64 voxels) and
// a CTA waits for exactly those predecessors.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "vx_ctx.hpp"

using namespace vx;

namespace {

constexpr int QCAP = 32768;       // entries per queue
constexpr int MAX_CONC = 256;     // CTAs (edits) per launch
constexpr int ET = 1024;          // threads per edit: a level of the flood fill holds up to ~2,700 entries
constexpr int CONFLICT = 64;      // see header comment
constexpr int SQ = 3072;          // queue entries held in shared memory (the rest: global scratch)

// the box of a local attempt
constexpr int BX = 33, BZ = 33, BY = 64, BOXN = BX * BZ * BY;
constexpr int BOX_BELOW = 47;                 // layers under the edit
constexpr int BOX_WORDS = BOXN / 2;           // two voxels per 32-bit word, as in the lightmaps
constexpr int BOX_ROWS = BY * BZ * 3;          // 16-voxel rows of the world the box touches: (layer, z, chunk column)
constexpr int BOX_PLANE = BOX_ROWS / 2;       // words of a one-bit-per-voxel plane: one 16-bit half per row
static_assert(BOXN % 2 == 0 && BOX_ROWS % 2 == 0, "whole words");
constexpr int BOX_DIRTY = (BOXN + 31) / 32;   // words of the changed-voxel plane (one bit per box voxel, light order)
// dynamic shared memory of k_edit, in words: box light | light-passing plane | emissive plane | changed-voxel
// plane | the first SQ entries of the three queues | the tables below.  (Everything an edit's loops touch
// sits at a fixed offset: no pointer to it has to live in a register, or in local memory.)
constexpr int BOX_DIRTY_AT = BOX_WORDS + 2 * BOX_PLANE;
constexpr int Q_AT = BOX_DIRTY_AT + BOX_DIRTY;
constexpr int POOLS_AT = Q_AT + 3 * SQ;   // int[25]: pool index of the 5x5 chunks around the edit's chunk
constexpr int FLAGS_AT = POOLS_AT + 25;   // int[25], per chunk of that table: faces to refresh | 16: light changed
                                          // | 32: a voxel of the chunk was looked at (Chunk::flags.modified)
constexpr int EMIS_AT = FLAGS_AT + 25;    // uint16[256]: DevDef::emission of block ids < 256
constexpr int EDIT_WORDS = EMIS_AT + 128;
constexpr size_t EDIT_SMEM = (size_t)EDIT_WORDS * sizeof(uint32_t);

extern __shared__ __align__(16) uint32_t edit_dyn[];
__device__ __forceinline__ int* edit_pools() { return reinterpret_cast<int*>(edit_dyn + POOLS_AT); }
__device__ __forceinline__ int* edit_flags() { return reinterpret_cast<int*>(edit_dyn + FLAGS_AT); }
__device__ __forceinline__ uint16_t* edit_emis() { return reinterpret_cast<uint16_t*>(edit_dyn + EMIS_AT); }

// queue entry: dx+32 (6) | dz+32 (6) | y (8) | light (4) | channel (2)
__device__ __forceinline__ uint32_t pack(int dx, int dz, int y, uint32_t light, int ch) {
    return (uint32_t)(dx + 32) | ((uint32_t)(dz + 32) << 6) | ((uint32_t)y << 12) | (light << 20) |
           ((uint32_t)ch << 24);
}

struct Ed {
    const DevWorld& W;   // the kernel's __grid_constant__ parameter: no per-thread copy of its 200+ bytes
    int ex, ez;          // the edit's column: origin of the packed coordinates
    int y0;              // first layer of the box of a local attempt,
    int ylo;             // the first one that is filled (the rest lie deeper than this edit can reach)
    uint32_t* q0;        // global scratch behind the three queues (QCAP entries each)
    int* err;
    int aborted;         // this thread asked for a voxel outside the box
    uint32_t vis;        // local attempt: chunks (bit = index in the 5x5 table) this thread's entries looked at
    // the 5x5 chunk table's corner (an edit reaches 31 voxels in x / z), the box's corner
    __device__ __forceinline__ int bx() const { return (ex >> 4) - 2; }
    __device__ __forceinline__ int bz() const { return (ez >> 4) - 2; }
    __device__ __forceinline__ int x0() const { return ex - 16; }
    __device__ __forceinline__ int z0() const { return ez - 16; }
};

struct Loc {
    int p, i;  // pool; voxel index in the chunk (global) or in the box (local attempt)
    int b;     // local attempt: bit index of the voxel in the box planes (row * 16 + x & 15)
};

__device__ __forceinline__ bool locate(const DevWorld& W, int x, int y, int z, Loc& l) {
    if (y < 0 || y >= CH) return false;  // Chunks::getChunkByVoxel, Chunks.cpp:142-151
    l.p = W.pool_of(x >> 4, z >> 4);
    if (l.p < 0) return false;
    l.i = vidx(x & 15, y, z & 15);
    return true;
}
// The light store an edit runs on: LOCAL = the box in shared memory, otherwise the lightmaps.
// the same through the edit's 5x5 table: no slot-table load on the way to every voxel
template <bool LOCAL>
__device__ __forceinline__ bool locate(Ed& E, int x, int y, int z, Loc& l) {
    if (y < 0 || y >= CH) return false;
    const unsigned cx = (unsigned)((x >> 4) - E.bx()), cz = (unsigned)((z >> 4) - E.bz());
    l.p = (cx < 5u && cz < 5u) ? edit_pools()[cz * 5 + cx] : E.W.pool_of(x >> 4, z >> 4);
    if (l.p < 0) return false;
    if constexpr (LOCAL) {
        const unsigned ux = (unsigned)(x - E.x0()), uy = (unsigned)(y - E.y0), uz = (unsigned)(z - E.z0());
        if (ux >= (unsigned)BX || uy >= (unsigned)BY || uz >= (unsigned)BZ || y < E.ylo) {
            E.aborted = 1;
            return false;
        }
        l.i = (int)((uy * BZ + uz) * BX + ux);
        l.b = (int)(((uy * BZ + uz) * 3 + (unsigned)((x >> 4) - (E.x0() >> 4))) * 16 + (unsigned)(x & 15));
    } else {
        l.i = vidx(x & 15, y, z & 15);
    }
    return true;
}


// a box voxel was written: the write-back visits exactly the marked voxels
__device__ __forceinline__ void box_mark(int i) { atomicOr(edit_dyn + BOX_DIRTY_AT + (i >> 5), 1u << (i & 31)); }

template <bool LOCAL>
__device__ __forceinline__ uint32_t* light_word(const Ed& E, const Loc& l) {
    if constexpr (LOCAL) return edit_dyn + (l.i >> 1);
    else return reinterpret_cast<uint32_t*>(E.W.light_of(l.p)) + (l.i >> 1);
}
__device__ __forceinline__ int nib_shift(const Loc& l, int ch) { return (l.i & 1) * 16 + ch * 4; }

template <bool LOCAL>
__device__ __forceinline__ uint32_t read_word(const Ed& E, const Loc& l) {
    return *reinterpret_cast<volatile uint32_t*>(light_word<LOCAL>(E, l));
}
template <bool LOCAL>
__device__ __forceinline__ uint32_t read_nib(const Ed& E, const Loc& l, int ch) {
    return (read_word<LOCAL>(E, l) >> nib_shift(l, ch)) & 0xFu;
}
template <bool LOCAL>
__device__ __forceinline__ bool light_passing(const Ed& E, const Loc& l) {
    if constexpr (LOCAL) return (edit_dyn[BOX_WORDS + (l.b >> 5)] >> (l.b & 31)) & 1u;
    else return (E.W.lp_of(l.p)[l.i >> 5] >> (l.i & 31)) & 1u;
}
// emission nibble of the block at a located voxel (0 for air and for unknown ids)
template <bool LOCAL>
__device__ __forceinline__ uint32_t emission_at(const Ed& E, const Loc& l, int x, int y, int z, int ch) {
    uint32_t id;
    if constexpr (LOCAL) {
        if (!((edit_dyn[BOX_WORDS + BOX_PLANE + (l.b >> 5)] >> (l.b & 31)) & 1u)) return 0;
        id = E.W.vox_of(l.p)[vidx(x & 15, y, z & 15)] & 0xFFFFu;
    } else {
        id = E.W.vox_of(l.p)[l.i] & 0xFFFFu;
    }
    if (id == 0 || id >= (uint32_t)E.W.n_defs) return 0;
    return ((id < 256u ? (uint32_t)edit_emis()[id] : __ldg(&E.W.defs[id].emission)) >> (4 * ch)) & 0xFu;
}

__device__ __forceinline__ void chunk_flag(const Ed& E, int x, int z, int p, int bits) {
    const unsigned cx = (unsigned)((x >> 4) - E.bx()), cz = (unsigned)((z >> 4) - E.bz());
    if (cx < 5u && cz < 5u) {
        int* f = edit_flags() + cz * 5 + cx;  // flushed to ChunkMeta when the edit is done
        if ((*reinterpret_cast<volatile int*>(f) & bits) != bits) atomicOr(f, bits);
    } else {  // (cannot happen: an edit reaches two chunks at most)
        if (bits & 15) atomicOr(&E.W.meta[p].face_dirty, bits & 15);
        if (bits & 16) E.W.meta[p].lstate = LS_UNKNOWN;
        if (bits & 32) E.W.meta[p].modified = 1;
    }
}
// `chunk->flags.modified = true` of the solver loops (LightSolver.cpp:74, 107)
__device__ __forceinline__ void note_visit(const Ed& E, int x, int z, const Loc& l) { chunk_flag(E, x, z, l.p, 32); }

__device__ __forceinline__ int border_bits(int x, int z) {
    const int lx = x & 15, lz = z & 15;
    int bits = 16;
    if (lx == 0) bits |= 1 << FACE_XM;
    if (lx == 15) bits |= 1 << FACE_XP;
    if (lz == 0) bits |= 1 << FACE_ZM;
    if (lz == 15) bits |= 1 << FACE_ZP;
    return bits;
}
// faceL (the border-plane copy of light) is not written here: concurrent entries of a level can
// change different nibbles of one voxel, and a copy updated nibble by nibble next to the CAS on
// the light word can end up stale.  A changed border voxel only marks its chunk's face dirty;
// k_faces_refresh_dirty copies those faces from the lightmap once the batch is done (nothing
// reads faceL in between: it only feeds the seeding of later builds).
// (A local attempt finds its changes when it writes the box back.)
template <bool LOCAL>
__device__ __forceinline__ void note_change(Ed& E, int x, int z, const Loc& l) {
    if constexpr (LOCAL) box_mark(l.i);
    else chunk_flag(E, x, z, l.p, border_bits(x, z));
}

// nibble := v if pred(old nibble); returns true (and the old nibble) on success
template <bool LOCAL, class Pred>
__device__ __forceinline__ bool cas_nib(Ed& E, int x, int z, const Loc& l, int ch, uint32_t v,
                                        Pred pred, uint32_t* old_out) {
    uint32_t* wp = light_word<LOCAL>(E, l);
    const int sh = nib_shift(l, ch);
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
    while (true) {
        const uint32_t on = (old >> sh) & 0xFu;
        if (!pred(on)) return false;
        const uint32_t nw = (old & ~(0xFu << sh)) | (v << sh);
        const uint32_t got = atomicCAS(wp, old, nw);
        if (got == old) {
            *old_out = on;
            if (on != v) note_change<LOCAL>(E, x, z, l);
            return true;
        }
        old = got;
    }
}

// queues: the first SQ entries live in shared memory, the rest in the CTA's global scratch
__device__ __forceinline__ void qput(const Ed& E, int which, int at, uint32_t entry) {
    if (at < SQ)
        edit_dyn[Q_AT + which * SQ + at] = entry;
    else if (at < QCAP)
        E.q0[which * QCAP + at] = entry;
    else
        *E.err = 1;
}
__device__ __forceinline__ void push(const Ed& E, int which, int* count, uint32_t entry) {
    qput(E, which, atomicAdd(count, 1), entry);
}
__device__ __forceinline__ uint32_t qget(const Ed& E, int which, int k) {
    return k < SQ ? edit_dyn[Q_AT + which * SQ + k] : E.q0[which * QCAP + k];
}
// ---- the box's own loops.  One thread takes one neighbour of one queue entry: a warp running by
// itself issues an instruction every five cycles or so, and six neighbours taken in turn by one
// thread made a level of a few entries as long as a level of a thousand.  A neighbour sits at a
// fixed offset from its entry (light: 33 voxels a row, planes: 48); an entry whose neighbours could
// lie outside what was filled (|dx| + |dz| >= 16, beyond the box's floor or ceiling unless that is
// the world's) gives the attempt up.
__constant__ int c_nb[6][3] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {1, 0, 0}, {-1, 0, 0}};

// one entry per lane that has one, one atomic on the queue's counter per warp.  Every lane calls this.
__device__ __forceinline__ void warp_push(const Ed& E, int which, int* count, bool have, uint32_t entry) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, have);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (have) qput(E, which, base + __popc(m & ((1u << lane) - 1u)), entry);
}

struct BoxVisit {
    int ni, nb;      // the neighbour: light index, plane bit index
    int sh;          // shift of the entry's channel in the neighbour's light word
    uint32_t el;     // the entry's light level
    uint32_t ne;     // the entry the neighbour would get, light field still the popped entry's
};

// neighbour k of entry e; false when there is none (outside the world) or the attempt is given up
__device__ __forceinline__ bool box_visit(Ed& E, uint32_t e, int k, BoxVisit& V) {
    const int ux = (int)(e & 63u) - 16, uz = (int)((e >> 6) & 63u) - 16, ey = (int)((e >> 12) & 255u);
    const int dx = c_nb[k][0], dy = c_nb[k][1], dz = c_nb[k][2];
    V.el = (e >> 20) & 0xFu;
    const int ny = ey + dy, nuy = ny - E.y0;
    if (abs(ux - 16) + abs(uz - 16) >= 16 || ey < E.ylo || ey >= E.y0 + BY) {
        E.aborted = 1;
        return false;
    }
    if (ny < E.ylo || ny >= E.y0 + BY) {
        if ((unsigned)ny < (unsigned)CH) E.aborted = 1;
        return false;
    }
    const int nux = ux + dx, nuz = uz + dz, t = nuy * BZ + nuz;
    V.ni = t * BX + nux;
    V.nb = t * 48 + nux + (E.x0() & 15);
    V.sh = (V.ni & 1) * 16 + (int)((e >> 24) & 3u) * 4;
    V.ne = e + (uint32_t)(dz * 64 + dy * 4096 + dx);
    // Chunk::flags.modified of the chunk the neighbour lies in
    E.vis |= 1u << ((((E.z0() + nuz) >> 4) - E.bz()) * 5 + ((E.x0() + nux) >> 4) - E.bx());
    return true;
}

// nibble (sh) of box voxel ni := v while pred(nibble); true (the voxel is marked) when this thread wrote it
template <class Pred>
__device__ __forceinline__ bool box_cas(int ni, int sh, uint32_t v, Pred pred) {
    uint32_t* wp = edit_dyn + (ni >> 1);
    uint32_t w = *reinterpret_cast<volatile uint32_t*>(wp);
    while (true) {
        if (!pred((w >> sh) & 0xFu)) return false;
        const uint32_t got = atomicCAS(wp, w, (w & ~(0xFu << sh)) | (v << sh));
        if (got == w) {
            box_mark(ni);
            return true;
        }
        w = got;
    }
}

// LightSolver::solve add loop body (LightSolver.cpp:97-125) in the box, one neighbour.  Every lane calls this.
__device__ __forceinline__ void add_visit_box(Ed& E, bool valid, uint32_t e, int k, int nextq, int* n_next) {
    BoxVisit V;
    bool raised = false;
    if (valid && box_visit(E, e, k, V) && V.el >= 2 && ((edit_dyn[BOX_WORDS + (V.nb >> 5)] >> (V.nb & 31)) & 1u))
        raised = box_cas(V.ni, V.sh, V.el - 1, [&](uint32_t o) { return o + 2 <= V.el; });
    warp_push(E, nextq, n_next, raised, V.ne - (1u << 20));
}

// LightSolver::solve removal loop body (LightSolver.cpp:60-95) in the box, one neighbour.  Every lane calls this.
__device__ __forceinline__ void removal_visit_box(Ed& E, bool valid, uint32_t e, int k, int nextq, int* n_next, int addq,
                                                  int* n_add) {
    BoxVisit V;
    bool have_next = false, have_add = false;
    uint32_t light = 0;
    if (valid && box_visit(E, e, k, V)) {
        light = (*reinterpret_cast<volatile uint32_t*>(edit_dyn + (V.ni >> 1)) >> V.sh) & 0xFu;
        if (V.el >= 2 && light == V.el - 1) {  // (not 0: el - 1 >= 1)
            uint32_t emission = 0;
            if ((edit_dyn[BOX_WORDS + BOX_PLANE + (V.nb >> 5)] >> (V.nb & 31)) & 1u) {  // an emitter keeps its emission
                const int wx = E.ex + (int)(V.ne & 63u) - 32, wy = (int)((V.ne >> 12) & 255u), wz = E.ez + (int)((V.ne >> 6) & 63u) - 32;
                Loc l;
                if (locate<true>(E, wx, wy, wz, l)) emission = emission_at<true>(E, l, wx, wy, wz, (int)((e >> 24) & 3u));
            }
            if (box_cas(V.ni, V.sh, emission, [&](uint32_t o) { return o == V.el - 1; })) {
                have_next = true;
                have_add = emission != 0;
                light = emission;
            } else {
                // lost the race: what the winner left, as the second pop of the reference's queue reads it
                light = (*reinterpret_cast<volatile uint32_t*>(edit_dyn + (V.ni >> 1)) >> V.sh) & 0xFu;
                have_add = light >= V.el;
            }
        } else {
            have_add = light >= V.el;
        }
    }
    warp_push(E, nextq, n_next, have_next, V.ne - (1u << 20));
    warp_push(E, addq, n_add, have_add, (V.ne & ~(0xFu << 20)) | (light << 20));
}


// LightSolver::solve removal loop body for one popped entry (LightSolver.cpp:60-95) on the lightmaps
// in global memory.  Everything the six neighbours need — block id, light word — is fetched first,
// then the six compare-and-swaps are issued back to back, then their results are looked at: three
// memory round trips per entry instead of a dozen dependent ones.  A CAS that lost a race (another
// entry of the level changed the word) takes the one-at-a-time path with the fresh value.  A nibble
// that does not hold level - 1 is not written by anyone during this level, so the value fetched up
// front is what a later read would return.
__device__ void removal_entry(Ed& E, bool valid, uint32_t e, int nextq, int* n_next, int addq, int* n_add) {
    const int x = E.ex + (int)(e & 63u) - 32, z = E.ez + (int)((e >> 6) & 63u) - 32;
    const int y = (e >> 12) & 255u, ch = (e >> 24) & 3u;
    const uint32_t el = (e >> 20) & 0xFu;
    if (!valid) return;
    Loc l[6];
    bool ok[6];
    uint32_t word[6], emission[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int nx = x + c_nb[k][0], ny = y + c_nb[k][1], nz = z + c_nb[k][2];
        ok[k] = locate<false>(E, nx, ny, nz, l[k]);
        word[k] = 0;
        emission[k] = 0;
        if (ok[k]) {
            word[k] = read_word<false>(E, l[k]);
            emission[k] = emission_at<false>(E, l[k], nx, ny, nz, ch);
        }
    }
    bool want[6];
    uint32_t got[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int sh = nib_shift(l[k], ch);
        const uint32_t on = (word[k] >> sh) & 0xFu;
        want[k] = ok[k] && el >= 2 && on != 0 && on == el - 1;
        got[k] = 0;
        if (want[k])
            got[k] = atomicCAS(light_word<false>(E, l[k]), word[k], (word[k] & ~(0xFu << sh)) | (emission[k] << sh));
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (!ok[k]) continue;
        const int nx = x + c_nb[k][0], ny = y + c_nb[k][1], nz = z + c_nb[k][2];
        note_visit(E, nx, nz, l[k]);
        const int sh = nib_shift(l[k], ch);
        bool cleared = false;
        uint32_t old = (word[k] >> sh) & 0xFu;
        if (want[k]) {
            if (got[k] == word[k]) {
                cleared = true;
                if (old != emission[k]) note_change<false>(E, nx, nz, l[k]);
            } else {
                cleared = cas_nib<false>(E, nx, nz, l[k], ch, emission[k], [&](uint32_t on) { return on != 0 && on == el - 1; }, &old);
            }
        }
        if (cleared) {
            if (emission[k]) push(E, addq, n_add, pack(nx - E.ex, nz - E.ez, ny, emission[k], ch));
            push(E, nextq, n_next, pack(nx - E.ex, nz - E.ez, ny, old, ch));
        } else {
            // (after a lost race: what the winner left — 0 or the block's emission — exactly as the
            // second pop of the reference's queue would read it)
            const uint32_t light = want[k] ? read_nib<false>(E, l[k], ch) : old;
            if (light >= el) push(E, addq, n_add, pack(nx - E.ex, nz - E.ez, ny, light, ch));
        }
    }
}

// LightSolver::solve add loop body for one popped entry (LightSolver.cpp:97-125); same shape.
// Light only rises during the add rounds, so a neighbour that already holds el - 1 or more when
// it is fetched never needs this entry.
__device__ void add_entry(Ed& E, bool valid, uint32_t e, int nextq, int* n_next) {
    const int x = E.ex + (int)(e & 63u) - 32, z = E.ez + (int)((e >> 6) & 63u) - 32;
    const int y = (e >> 12) & 255u, ch = (e >> 24) & 3u;
    const uint32_t el = (e >> 20) & 0xFu;
    if (!valid) return;
    Loc l[6];
    bool ok[6], lp[6];
    uint32_t word[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ok[k] = locate<false>(E, x + c_nb[k][0], y + c_nb[k][1], z + c_nb[k][2], l[k]);
        lp[k] = false;
        word[k] = 0;
        if (ok[k]) {
            lp[k] = light_passing<false>(E, l[k]);
            word[k] = read_word<false>(E, l[k]);
        }
    }
    bool want[6];
    uint32_t got[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int sh = nib_shift(l[k], ch);
        want[k] = ok[k] && el >= 2 && lp[k] && ((word[k] >> sh) & 0xFu) + 2 <= el;
        got[k] = 0;
        if (want[k])
            got[k] = atomicCAS(light_word<false>(E, l[k]), word[k], (word[k] & ~(0xFu << sh)) | ((el - 1) << sh));
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (!ok[k]) continue;
        const int nx = x + c_nb[k][0], nz = z + c_nb[k][2];
        note_visit(E, nx, nz, l[k]);
        if (!want[k]) continue;
        bool raised = got[k] == word[k];
        if (raised) {
            note_change<false>(E, nx, nz, l[k]);
        } else {
            uint32_t old;
            raised = cas_nib<false>(E, nx, nz, l[k], ch, el - 1, [&](uint32_t on) { return on + 2 <= el; }, &old);
        }
        if (raised) push(E, nextq, n_next, pack(nx - E.ex, nz - E.ez, y + c_nb[k][1], el - 1, ch));
    }
}

struct EdShared {
    uint32_t seeds[272];   // removal seeds (LightSolver::remove)
    int n_seed;
    uint32_t seed_levels;  // bit l: some seed holds light level l
    int n[3];              // fill counts of the add rounds' queues (queue 2: what the removal hands over)
    int lc[3];             // fill counts of the removal levels (level k of a phase: lc[k % 3], queue k & 1)
};

// LightSolver::remove (LightSolver.cpp:37-48)
template <bool LOCAL>
__device__ __forceinline__ void ls_remove(Ed& E, EdShared& S, int ch, int x, int y, int z) {
    Loc l;
    if (!locate<LOCAL>(E, x, y, z, l)) return;
    uint32_t old;
    if (cas_nib<LOCAL>(E, x, z, l, ch, 0, [](uint32_t on) { return on != 0; }, &old)) {
        const int at = atomicAdd(&S.n_seed, 1);
        S.seeds[at] = pack(x - E.ex, z - E.ez, y, old, ch);
        atomicOr(&S.seed_levels, 1u << old);
    }
}

// LightSolver::add(x,y,z,emission) (LightSolver.cpp:18-31); entries go to queue 2
template <bool LOCAL>
__device__ __forceinline__ void ls_add(Ed& E, EdShared& S, int ch, int x, int y, int z, uint32_t emission) {
    if (emission <= 1) return;
    Loc l;
    if (!locate<LOCAL>(E, x, y, z, l)) return;
    uint32_t old;
    if (cas_nib<LOCAL>(E, x, z, l, ch, emission, [&](uint32_t on) { return emission >= on; }, &old)) {
        push(E, 2, &S.n[2], pack(x - E.ex, z - E.ez, y, emission, ch));
        note_visit(E, x, z, l);
    }
}

__device__ void reset_queues(EdShared& S) {
    __syncthreads();
    if (threadIdx.x == 0) {
        S.n_seed = 0;
        S.seed_levels = 0;
        S.n[0] = S.n[1] = S.n[2] = 0;
        S.lc[0] = S.lc[1] = S.lc[2] = 0;
    }
    __syncthreads();
}

// -DVX_EDIT_STATS: nanoseconds per phase, summed over the edits of a call, in stats[2 ..] (see vx_light_edit)
#ifdef VX_EDIT_STATS
__device__ unsigned long long g_edit_cyc[8];
__device__ unsigned long long g_edit_hist[16];  // removal levels by size class: ns, count
__device__ unsigned long long g_edit_rounds[6];  // removal levels, their entries, add rounds, their entries, ns in each
#define EDIT_ROUND(k, n)                                                            \
    do {                                                                            \
        if (threadIdx.x == 0) {                                                     \
            atomicAdd(&g_edit_rounds[k], 1ull);                                     \
            atomicAdd(&g_edit_rounds[(k) + 1], (unsigned long long)(n));            \
        }                                                                           \
    } while (0)
__device__ __forceinline__ unsigned long long edit_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define EDIT_MARK(k)                                                           \
    do {                                                                       \
        if (threadIdx.x == 0) {                                                \
            const unsigned long long now_ = edit_now();                        \
            atomicAdd(stats + 2 + (k), (int)(now_ - mark_));                   \
            mark_ = now_;                                                      \
        }                                                                      \
    } while (0)
#else
#define EDIT_MARK(k) do {} while (0)
#define EDIT_ROUND(k, n) do {} while (0)
#endif

// The removal half of LightSolver::solve for every channel at once: the seeds in S.seeds, level
// by level; what borders the cleared region goes to queue 2.  One barrier per level: the fill
// counts rotate over three slots (this level's, the next level's, and the one being zeroed for
// the level after).  Returns false when a local attempt left its box.
template <bool LOCAL>
__device__ bool removal_phase(Ed& E, EdShared& S) {
    const int tid = threadIdx.x;
    if (__syncthreads_or(E.aborted)) return false;
#ifdef VX_EDIT_STATS
    const unsigned long long t_in = edit_now();
    unsigned long long t_lv = t_in;
    int prev_n = -1;
    long long lv_cyc = 0;
#endif
    const uint32_t levels = S.seed_levels;
    const int n_seed = S.n_seed;
    if (!levels) return true;
    const int top = 31 - __clz(levels);
    for (int k = tid; k < n_seed; k += ET)
        if ((int)((S.seeds[k] >> 20) & 0xFu) == top) push(E, 0, &S.lc[0], S.seeds[k]);
    bool ok = true;
    int it = 0;
    for (int level = top; level >= 1; level--, it++) {
#ifdef VX_EDIT_STATS
        const long long bar_in = clock64();
#endif
        const bool gave_up = __syncthreads_or(E.aborted);
#ifdef VX_EDIT_STATS
        if (tid == 0) atomicAdd(&g_edit_cyc[5], (unsigned long long)(clock64() - bar_in));  // thread 0 at the barrier
#endif
        if (gave_up) {
            ok = false;
            break;
        }
        const int cur = it & 1, c = it % 3, cn = (it + 1) % 3, cz = (it + 2) % 3;
        const int n = min(S.lc[c], QCAP);
#ifdef VX_EDIT_STATS
        if (tid == 0) {
            const unsigned long long now = edit_now();
            if (prev_n >= 0) {
                const int cls = prev_n < 64 ? 0 : prev_n < 512 ? 1 : prev_n <= 1024 ? 2 : 3;
                atomicAdd(&g_edit_hist[cls * 2], now - t_lv);
                atomicAdd(&g_edit_hist[cls * 2 + 1], 1ull);
            } else {
                atomicAdd(&g_edit_hist[8], now - t_in);  // preamble
                atomicAdd(&g_edit_hist[9], 1ull);
            }
            t_lv = now;
            prev_n = n;
        }
#endif
        if (tid == 0) S.lc[cz] = 0;  // (read by everyone a level ago, filled again a level from now)
        if (n == 0 && !(levels & ((1u << level) - 1u))) break;
        EDIT_ROUND(0, n);
#ifdef VX_EDIT_STATS
        if (tid == 0) {
            const long long c_ = clock64();
            if (lv_cyc) atomicAdd(&g_edit_cyc[4], (unsigned long long)(c_ - lv_cyc));  // whole level, barrier to barrier
            lv_cyc = c_;
        }
#endif
        if (level >= 2 && ((levels >> (level - 1)) & 1u))  // the next level's seeds ride along
            for (int k = tid; k < n_seed; k += ET)
                if ((int)((S.seeds[k] >> 20) & 0xFu) == level - 1) push(E, cur ^ 1, &S.lc[cn], S.seeds[k]);
        if constexpr (LOCAL) {
            const int items = 6 * n;  // (entry, neighbour) pairs, neighbour-major: a warp mostly shares its neighbour
            for (int base = 0; base < items; base += ET) {
                const int item = base + tid;
                const int k = (item >= n) + (item >= 2 * n) + (item >= 3 * n) + (item >= 4 * n) + (item >= 5 * n), idx = item - k * n;
                const bool valid = item < items;
                removal_visit_box(E, valid, valid ? qget(E, cur, idx) : 0u, k, cur ^ 1, &S.lc[cn], 2, &S.n[2]);
            }
        } else {
            for (int base = 0; base < n; base += ET) {
                const int k = base + tid;
                removal_entry(E, k < n, k < n ? qget(E, cur, k) : 0u, cur ^ 1, &S.lc[cn], 2, &S.n[2]);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        S.n_seed = 0;
        S.seed_levels = 0;
        S.lc[0] = S.lc[1] = S.lc[2] = 0;
#ifdef VX_EDIT_STATS
        atomicAdd(&g_edit_rounds[4], edit_now() - t_in);
#endif
    }
    return ok;
}

// The add half: queue 2 -> 0 -> 1 -> 2 ..., one barrier per round (counts rotate as above).
template <bool LOCAL>
__device__ bool add_phase(Ed& E, EdShared& S) {
    const int tid = threadIdx.x;
    int a = 2, b = 0, c = 1;
    bool ok = true;
#ifdef VX_EDIT_STATS
    const unsigned long long t_in = edit_now();
#endif
    while (true) {
        if (__syncthreads_or(E.aborted)) {
            ok = false;
            break;
        }
        const int n = min(S.n[a], QCAP);
        if (tid == 0) S.n[c] = 0;
        if (n == 0) break;
        EDIT_ROUND(2, n);
        if constexpr (LOCAL) {
            const int items = 6 * n;
            for (int base = 0; base < items; base += ET) {
                const int item = base + tid;
                const int k = (item >= n) + (item >= 2 * n) + (item >= 3 * n) + (item >= 4 * n) + (item >= 5 * n), idx = item - k * n;
                const bool valid = item < items;
                add_visit_box(E, valid, valid ? qget(E, a, idx) : 0u, k, b, &S.n[b]);
            }
        } else {
            for (int base = 0; base < n; base += ET) {
                const int k = base + tid;
                add_entry(E, k < n, k < n ? qget(E, a, k) : 0u, b, &S.n[b]);
            }
        }
        const int t = a;
        a = b;
        b = c;
        c = t;
    }
    __syncthreads();
    if (tid == 0) S.n[0] = S.n[1] = S.n[2] = 0;
#ifdef VX_EDIT_STATS
    if (tid == 0) atomicAdd(&g_edit_rounds[5], edit_now() - t_in);
#endif
    return ok;
}

__device__ __forceinline__ bool nonair_at(const DevWorld& W, int p, int lx, int y, int lz) {
    return (W.vox_of(p)[vidx(lx, y, lz)] & 0xFFFFu) != 0;
}

// Lighting::onBlockSet (Lighting.cpp:163-215) on one of the two light stores.  The reference runs
// the solvers twice when a block is broken or an emitter placed (removal + re-fill, then the new
// light); here the second run's seeds join the first run's add rounds: those rounds are a monotone
// relaxation whose result does not depend on the order entries are taken in, and a seed pushed
// before a voxel was raised is pushed again, with the raised value, when it is.
// Returns false when a local attempt left its box (nothing global has been written by then).
template <bool LOCAL>
__device__ __forceinline__ bool on_block_set(Ed& E, EdShared& S, int* s_hi, int x, int y, int z, uint32_t id, uint32_t bf,
                                          uint32_t bem, int p) {
    const DevWorld& W = E.W;
    const int tid = threadIdx.x;
    const int lx = x & 15, lz = z & 15;
    if (tid < 3) ls_remove<LOCAL>(E, S, tid, x, y, z);
    if (id == 0) {
        if (!removal_phase<LOCAL>(E, S)) return false;  // solverR/G/B->solve(), removal half
        Loc up;
        const bool sky_above = locate<LOCAL>(E, x, y + 1, z, up) && read_nib<LOCAL>(E, up, 3) == 0xFu;
        if (sky_above) {
            // walk down from y while the voxel is air (the test reads the *air*
            // def's skyLightPassing, Lighting.cpp:173-180)
            int stop = -1;  // highest i <= y that ends the walk
            if (bf & DF_SLP) {
                if (tid <= y && nonair_at(W, p, lx, tid, lz)) atomicMax(s_hi, tid);
                __syncthreads();
                stop = *s_hi;
            }
            if (tid > stop && tid <= y) ls_add<LOCAL>(E, S, 3, x, tid, z, 0xFu);
        }
        if (tid < 24) {
            const int k = tid >> 2, ch = tid & 3;
            const int nx = x + c_nb[k][0], ny = y + c_nb[k][1], nz = z + c_nb[k][2];
            Loc l;
            if (locate<LOCAL>(E, nx, ny, nz, l))  // LightSolver::add(x,y,z), :33-35
                ls_add<LOCAL>(E, S, ch, nx, ny, nz, read_nib<LOCAL>(E, l, ch));
        }
        return add_phase<LOCAL>(E, S);
    }
    if (!(bf & DF_SLP)) {
        if (tid == 0) ls_remove<LOCAL>(E, S, 3, x, y, z);
        // column below: i = y-1 .. b, b = first i with i == 0 or voxel(i-1) non-air
        if (y >= 1) {
            if (tid <= y - 2 && nonair_at(W, p, lx, tid, lz)) atomicMax(s_hi, tid);
            __syncthreads();
            const int b = *s_hi + 1;
            if (tid >= b && tid <= y - 1) ls_remove<LOCAL>(E, S, 3, x, tid, z);
        }
    }
    if (!removal_phase<LOCAL>(E, S)) return false;
    if ((bem & 0x0FFFu) && tid < 3) ls_add<LOCAL>(E, S, tid, x, y, z, (bem >> (4 * tid)) & 0xFu);
    return add_phase<LOCAL>(E, S);
}

// Work item w of a sweep over the box = one 16-voxel row of the world: (layer uy, chunk column seg, line uz),
// uz fastest, so that a warp's rows are consecutive lines of one chunk's layer (consecutive 32-byte pieces
// of its lightmap, one sector of each plane).  `row` is where the row's plane halves are kept
// ((uy * BZ + uz) * 3 + seg, see Loc::b).  Returns the pool of the row's chunk, or -1 when no voxel of the
// row lies within |dx| + |dz| <= 16 of the edit's column (nothing further out can be looked at, see
// box_visit) or the chunk is absent.
__device__ __forceinline__ int box_row(const Ed& E, int w, int& row, int& uy, int& uz, int& ux0) {
    uz = w % BZ;
    const int t = w / BZ, seg = t % 3;
    uy = t / 3;
    row = (uy * BZ + uz) * 3 + seg;
    ux0 = ((E.x0() >> 4) + seg) * 16 - E.x0();  // box x of the row's first voxel
    const int lo_x = max(ux0, 0), hi_x = min(ux0 + 15, BX - 1);
    const int near_x = lo_x > 16 ? lo_x - 16 : hi_x < 16 ? 16 - hi_x : 0;
    if (near_x + abs(uz - 16) > 16) return -1;
    return edit_pools()[(((E.z0() + uz) >> 4) - E.bz()) * 5 + ((E.x0() >> 4) + seg - E.bx())];
}

constexpr int FILL_ROWS = 4;  // rows of the box a thread has in flight while it is filled
constexpr int MAXD = 8;  // predecessors an edit can wait for inside one launch (more: the launch is cut before it)

// grid: one CTA per edit of the launch.  deps[b][MAXD] = the earlier edits of this launch (block
// indices, -1 padded) whose reach overlaps this one's: the CTA starts when their `done` flags are up
// (cooperative launch: every CTA is resident, and a predecessor always has a lower block index).
__global__ void __launch_bounds__(ET) k_edit(const __grid_constant__ DevWorld W, const vx_edit* edits, uint32_t* scratch,
                                             int* err, const int32_t* deps, int32_t* done, int try_local, int* stats) {
    uint32_t* const s_box = edit_dyn;
    uint32_t* const s_lp = s_box + BOX_WORDS;
    uint32_t* const s_em = s_lp + BOX_PLANE;
    int* const s_pools5 = edit_pools();
    int* const s_chunk_flags = edit_flags();
    uint16_t* const s_emis = edit_emis();
    __shared__ EdShared S;
    __shared__ int s_hi, s_lo, s_bot, s_top, s_smin, s_smax, s_ground;
    const int tid = threadIdx.x;
#ifdef VX_EDIT_STATS
    unsigned long long mark_ = edit_now();
#endif
    const vx_edit ed = edits[blockIdx.x];
    const int x = ed.x, y = ed.y, z = ed.z;
    const uint32_t id = ed.id;
    if (tid < 25) {
        s_pools5[tid] = W.pool_of((x >> 4) - 2 + tid % 5, (z >> 4) - 2 + tid / 5);
        s_chunk_flags[tid] = 0;
    }
    if (tid < 256) s_emis[tid] = tid < W.n_defs ? (uint16_t)__ldg(&W.defs[tid].emission) : (uint16_t)0;
    if (tid == 0) {
        S.n_seed = 0;
        S.seed_levels = 0;
        S.n[0] = S.n[1] = S.n[2] = 0;
        S.lc[0] = S.lc[1] = S.lc[2] = 0;
        s_hi = -1;
        s_bot = CH;
        s_top = -1;
        s_smin = CH;
        s_smax = -1;
        s_ground = 0;
    }
    __syncthreads();
    Ed E {W, x, z, min(max(y - BOX_BELOW, 0), CH - BY), 0, scratch + (size_t)blockIdx.x * 3 * QCAP, err, 0, 0u};
    // (the box's own loops do not look chunks up: all nine the box can touch have to be there)
    bool all_nine = true;
#pragma unroll
    for (int k = 0; k < 9; k++) all_nine = all_nine && s_pools5[(1 + k / 3) * 5 + 1 + k % 3] >= 0;
    const bool local = try_local && all_nine;
    if (local) {
        // while the predecessors finish: the box's rows on their way into L2, the changed-voxel plane zeroed
        for (int w = tid; w < BOX_ROWS; w += ET) {
            int row, uy, uz, ux0;
            const int q = box_row(E, w, row, uy, uz, ux0);
            if (q >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(W.light_of(q) + vidx(0, E.y0 + uy, (E.z0() + uz) & 15)));
        }
        for (int k = tid; k < BOX_DIRTY; k += ET) s_box[BOX_DIRTY_AT + k] = 0u;
    }
    if (tid < CH && s_pools5[12] >= 0 && (unsigned)y < (unsigned)CH)  // the edited column, for the summaries below
        asm volatile("prefetch.global.L2 [%0];" ::"l"(W.vox_of(s_pools5[12]) + vidx(x & 15, tid, z & 15)));
    if (tid < MAXD) {
        const int d = deps[blockIdx.x * MAXD + tid];
        if (d >= 0) {
            unsigned backoff = 32;
            while (*reinterpret_cast<volatile int32_t*>(done + d) == 0) {
                __nanosleep(backoff);
                if (backoff < 512) backoff *= 2;
            }
        }
        __threadfence();  // what the predecessors wrote is visible to everything below
    }
    __syncthreads();
    EDIT_MARK(0);  // waiting for predecessors
    Loc here;
    if (!locate(W, x, y, z, here) || id >= (uint32_t)W.n_defs) {  // blocks_agent.cpp:18-27
        if (tid == 0) atomicExch(done + blockIdx.x, 1);
        return;
    }
    const int p = here.p, lx = x & 15, lz = z & 15;
    const uint32_t bf = __ldg(&W.defs[id].flags);
    const uint32_t bem = __ldg(&W.defs[id].emission);

    // ---- blocks_agent::set
    if (tid == 0) {
        vx_voxel* vp = W.vox_of(p) + here.i;
        const uint32_t oldid = *reinterpret_cast<volatile vx_voxel*>(vp) & 0xFFFFu;
        *vp = VX_VOXEL(id, ed.state);
        const uint32_t bit = 1u << (here.i & 31);
        uint32_t* lpw = W.lp_of(p) + (here.i >> 5);
        uint32_t* emw = W.emit_of(p) + (here.i >> 5);
        if (bf & DF_LP) atomicOr(lpw, bit); else atomicAnd(lpw, ~bit);
        if (bf & DF_EMISSIVE) atomicOr(emw, bit); else atomicAnd(emw, ~bit);
        {
            const uint32_t mb = mesher_bits(VX_VOXEL(id, ed.state), bf, (uint32_t)W.n_defs);
            uint32_t* pl[3] = {W.mblk_of(p) + (here.i >> 5), W.msc_of(p) + (here.i >> 5), W.moth_of(p) + (here.i >> 5)};
            for (int k = 0; k < 3; k++) {
                if ((mb >> k) & 1u) atomicOr(pl[k], bit);
                else atomicAnd(pl[k], ~bit);
            }
        }
        volatile int16_t* lc = W.laycnt_of(p);
        if (oldid != 0 && id == 0) lc[y] = lc[y] - 1;
        if (oldid == 0 && id != 0) lc[y] = lc[y] + 1;
        W.meta[p].modified = 1;
        int q;
        if (lx == 0 && (q = s_pools5[2 * 5 + 1]) >= 0) W.meta[q].modified = 1;
        if (lz == 0 && (q = s_pools5[1 * 5 + 2]) >= 0) W.meta[q].modified = 1;
        if (lx == 15 && (q = s_pools5[2 * 5 + 3]) >= 0) W.meta[q].modified = 1;
        if (lz == 15 && (q = s_pools5[3 * 5 + 2]) >= 0) W.meta[q].modified = 1;
    }
    __syncthreads();
    // ---- the column's and the chunk's summaries: Chunk::updateHeights (layer counts), skycol of the
    // edited column (only Lighting::prebuildSkyLight reads it) and where buildSkyLight would start on
    // a pristine lightmap (see k_derive).  Threads 0 .. 255 are the column's voxels / the chunk's
    // layers / its columns in turn.
    int my_sc = 0;
    if (tid < CH) {
        const uint32_t vid = *reinterpret_cast<volatile vx_voxel*>(W.vox_of(p) + vidx(lx, tid, lz)) & 0xFFFFu;
        const uint32_t f = vid < (uint32_t)W.n_defs ? __ldg(&W.defs[vid].flags) : 0u;
        if (!(f & DF_SLP)) atomicMax(&s_hi, tid);
        if (vid != 0 && tid < y) atomicMax(&s_ground, tid);  // what a sky column through the edit comes down on
        if (reinterpret_cast<volatile int16_t*>(W.laycnt_of(p))[tid] > 0) {
            atomicMin(&s_bot, tid);
            atomicMax(&s_top, tid + 1);
        }
        my_sc = W.skycol_of(p)[tid];
    }
    __syncthreads();
    const int sky = s_hi;
    if (tid == 0) s_lo = sky < 0 ? -1 : 0;
    __syncthreads();
    if (tid < CH) {
        const int ni = vidx(lx, tid, lz);
        if (tid <= sky && ((__ldcg(W.lp_of(p) + (ni >> 5)) >> (ni & 31)) & 1u)) atomicMax(&s_lo, tid);
        const int sc = tid == lz * 16 + lx ? sky : my_sc;
        atomicMax(&s_smax, sc);
        atomicMin(&s_smin, sc);
    }
    __syncthreads();
    if (tid == 0) {
        W.skycol_of(p)[lz * 16 + lx] = (int16_t)sky;
        W.ystar0_of(p)[lz * 16 + lx] = (int16_t)s_lo;
        ChunkMeta* m = &W.meta[p];
        m->sky_max = s_smax;
        m->sky_min = s_smin;
        if (y < m->bottom) {
            m->bottom = y;
        } else if (y + 1 > m->top) {
            m->top = y + 1;
        } else if (id == 0) {  // Chunk::updateHeights
            if (s_bot < CH) m->bottom = s_bot;
            if (s_top >= 0) m->top = s_top;
        }
        s_hi = -1;
    }
    __syncthreads();

    // block light reaches 15 layers below the edit, a sky column through it 15 below where it lands
    E.ylo = max(E.y0, s_ground - 16);
    const int first_row = (E.ylo - E.y0) * (BZ * 3);
    EDIT_MARK(1);  // blocks_agent::set and the column / chunk summaries
    // ---- Lighting::onBlockSet: in the box first, on the lightmaps themselves if that fell short
    bool finished = false;
    if (local) {
        // the box is filled a 16-voxel row of the world at a time (32 bytes of a lightmap, half a word
        // of each plane), FILL_ROWS rows per thread in flight
        uint16_t* const box16 = reinterpret_cast<uint16_t*>(s_box);
        uint16_t* const lp16 = reinterpret_cast<uint16_t*>(s_lp);
        uint16_t* const em16 = reinterpret_cast<uint16_t*>(s_em);
#pragma unroll 1
        for (int base = first_row; base < BOX_ROWS; base += FILL_ROWS * ET) {
            uint4 lo[FILL_ROWS], hi[FILL_ROWS];
            uint32_t wl[FILL_ROWS], we[FILL_ROWS];
            int q[FILL_ROWS], row[FILL_ROWS], uy[FILL_ROWS], uz[FILL_ROWS], ux0[FILL_ROWS];
#pragma unroll
            for (int u = 0; u < FILL_ROWS; u++) {
                const int w = base + u * ET + tid;
                q[u] = w < BOX_ROWS ? box_row(E, w, row[u], uy[u], uz[u], ux0[u]) : -1;
                if (q[u] >= 0) {
                    const int wz = E.z0() + uz[u], vi = vidx(0, E.y0 + uy[u], wz & 15);
                    const uint4* src = reinterpret_cast<const uint4*>(W.light_of(q[u]) + vi);
                    lo[u] = __ldcg(src);
                    hi[u] = __ldcg(src + 1);
                    wl[u] = __ldcg(W.lp_of(q[u]) + (vi >> 5)) >> ((wz & 1) * 16);
                    we[u] = __ldcg(W.emit_of(q[u]) + (vi >> 5)) >> ((wz & 1) * 16);
                }
            }
#pragma unroll
            for (int u = 0; u < FILL_ROWS; u++) {
                if (q[u] < 0) continue;
                const int t = uy[u] * BZ + uz[u];
                lp16[row[u]] = (uint16_t)wl[u];
                em16[row[u]] = (uint16_t)we[u];
                const uint32_t w8[8] = {lo[u].x, lo[u].y, lo[u].z, lo[u].w, hi[u].x, hi[u].y, hi[u].z, hi[u].w};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const unsigned ux = (unsigned)(ux0[u] + i);
                    if (ux < (unsigned)BX) box16[t * BX + ux] = (uint16_t)(w8[i >> 1] >> ((i & 1) * 16));
                }
            }
        }
        __syncthreads();
        EDIT_MARK(2);  // box filled
        const bool ok = on_block_set<true>(E, S, &s_hi, x, y, z, id, bf, bem, p);
        EDIT_MARK(3);  // onBlockSet in the box
        if (ok && !__syncthreads_or(E.aborted)) {
            for (uint32_t v = E.vis; v; v &= v - 1) {
                int* f = s_chunk_flags + (__ffs(v) - 1);
                if (!(*reinterpret_cast<volatile int*>(f) & 32)) atomicOr(f, 32);
            }
            // what was written goes back to the lightmaps a row at a time: a row that lies wholly within reach
            // (no other edit running now can touch any of it) as 32 bytes, the marked voxels of the others one
            // by one; a marked voxel on a chunk border marks that face
            for (int w = first_row + tid; w < BOX_ROWS; w += ET) {
                int row, uy, uz, ux0;
                const int q = box_row(E, w, row, uy, uz, ux0);
                if (q < 0) continue;
                const int t = uy * BZ + uz, i_lo = max(0, -ux0), i_hi = min(15, BX - 1 - ux0);
                const int d = t * BX + ux0 + i_lo;  // first voxel of the row that is in the box
                const uint32_t* dp = s_box + BOX_DIRTY_AT + (d >> 5);
                const unsigned long long two = dp[0] | ((unsigned long long)((d >> 5) + 1 < BOX_DIRTY ? dp[1] : 0u) << 32);
                const uint32_t marked = ((uint32_t)(two >> (d & 31)) & ((2u << (i_hi - i_lo)) - 1u)) << i_lo;
                if (!marked) continue;
                const int wz = E.z0() + uz, dzz = abs(uz - 16);
                vx_light* dst = W.light_of(q) + vidx(0, E.y0 + uy, wz & 15);
                const uint16_t* src = box16 + t * BX + ux0;
                if (i_lo == 0 && i_hi == 15 && abs(ux0 - 16) + dzz <= 16 && abs(ux0 + 15 - 16) + dzz <= 16) {
                    uint32_t w8[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) w8[j] = (uint32_t)src[2 * j] | ((uint32_t)src[2 * j + 1] << 16);
                    reinterpret_cast<uint4*>(dst)[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
                    reinterpret_cast<uint4*>(dst)[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
                } else {
                    for (uint32_t m = marked; m; m &= m - 1) {
                        const int i = __ffs(m) - 1;
                        dst[i] = src[i];
                    }
                }
                int bits = 16;
                if ((wz & 15) == 0) bits |= 1 << FACE_ZM;
                if ((wz & 15) == 15) bits |= 1 << FACE_ZP;
                if (marked & 1u) bits |= 1 << FACE_XM;
                if (marked & 0x8000u) bits |= 1 << FACE_XP;
                chunk_flag(E, E.x0() + ux0 + i_lo, wz, q, bits);
            }
            finished = true;
            EDIT_MARK(4);  // box written back
        } else {
            reset_queues(S);
            if (tid < 25) s_chunk_flags[tid] = 0;
            if (tid == 0) s_hi = -1;
            E.aborted = 0;
            E.vis = 0u;
            __syncthreads();
        }
    }
    if (!finished) {
        if (tid == 0 && stats) atomicAdd(stats, 1);
        on_block_set<false>(E, S, &s_hi, x, y, z, id, bf, bem, p);
        EDIT_MARK(5);  // onBlockSet on the global lightmaps
    }
    // what the edit changed, per chunk: faces for k_faces_refresh_dirty, the lightmap is no longer
    // what prebuildSkyLight left, Chunk::flags.modified of every chunk the solvers looked at
    __syncthreads();
    if (tid < 25 && s_chunk_flags[tid] && s_pools5[tid] >= 0) {
        ChunkMeta* m = &W.meta[s_pools5[tid]];
        if (s_chunk_flags[tid] & 15) atomicOr(&m->face_dirty, s_chunk_flags[tid] & 15);
        if (s_chunk_flags[tid] & 16) m->lstate = LS_UNKNOWN;
        if (s_chunk_flags[tid] & 32) m->modified = 1;
    }
    // everything this edit wrote is out before its successors are let go (the barrier orders the CTA's
    // writes before thread 0's fence, the fence before the flag: the pattern of a grid-wide barrier)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicExch(done + blockIdx.x, 1);
    }
    EDIT_MARK(6);  // flags out, fence
}

// Copies the faces an edit batch marked dirty from the lightmap into faceL.  grid: touched chunks.
__global__ void __launch_bounds__(256) k_faces_refresh_dirty(const __grid_constant__ DevWorld W, const int32_t* pools) {
    const int p = pools[blockIdx.x];
    const int dirty = W.meta[p].face_dirty;
    if (!dirty) return;
    const vx_light* L = W.light_of(p);
    for (int f = 0; f < 4; f++) {
        if (!((dirty >> f) & 1)) continue;
        for (int t = threadIdx.x; t < CH * 16; t += blockDim.x) {
            const int yy = t >> 4, i = t & 15;
            int xx, zz;
            if (f == FACE_XM) { xx = 0; zz = i; }
            else if (f == FACE_XP) { xx = 15; zz = i; }
            else if (f == FACE_ZM) { xx = i; zz = 0; }
            else { xx = i; zz = 15; }
            W.faceL_of(p, f)[yy * 16 + i] = L[vidx(xx, yy, zz)];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) W.meta[p].face_dirty = 0;
}

}  // namespace

extern "C" int vx_light_edit(vx_ctx* ctx, const vx_edit* edits, int32_t n) {
    if (!ctx) return -1;
    if (n <= 0) return 0;
    cudaSetDevice(ctx->device);
    if (vx_light_finish(ctx)) return -1;
    if (ctx->pool_cap == 0) {
        ctx->fail("vx_light_edit: configure the window first");
        return -1;
    }
    // ---- waves: edit i runs after every earlier edit within CONFLICT voxels.  (Light moves one voxel
    // per step along an axis: an edit's reads and writes stay within |dx| + |dz| <= 32 of its column —
    // any y: a sky column can fall all the way down.)  preds of edit i: pred[pred_at[i] .. pred_at[i + 1])
    std::vector<int> wave(n, 0), pred, pred_at(n + 1, 0);
    auto conflicts = [&](int i, int j) {
        return std::abs(edits[j].x - edits[i].x) + std::abs(edits[j].z - edits[i].z) <= CONFLICT;
    };
    if (n <= 256) {  // a frame's worth of edits: every pair is looked at
        for (int i = 0; i < n; i++) {
            int w = 0;
            for (int j = 0; j < i; j++)
                if (conflicts(i, j)) {
                    w = std::max(w, wave[j] + 1);
                    pred.push_back(j);
                }
            wave[i] = w;
            pred_at[i + 1] = (int)pred.size();
        }
    } else {  // 64-voxel cells
        std::unordered_map<long long, std::vector<int>> cells;
        auto key = [](int cx, int cz) { return ((long long)cx << 32) ^ (unsigned int)cz; };
        for (int i = 0; i < n; i++) {
            const int cx = edits[i].x >> 6, cz = edits[i].z >> 6;
            int w = 0;
            for (int dz = -1; dz <= 1; dz++)
                for (int dx = -1; dx <= 1; dx++) {
                    auto it = cells.find(key(cx + dx, cz + dz));
                    if (it == cells.end()) continue;
                    for (int j : it->second)
                        if (conflicts(i, j)) {
                            w = std::max(w, wave[j] + 1);
                            pred.push_back(j);
                        }
                }
            wave[i] = w;
            pred_at[i + 1] = (int)pred.size();
            cells[key(cx, cz)].push_back(i);
        }
    }
    int n_waves = 0;
    for (int i = 0; i < n; i++) n_waves = std::max(n_waves, wave[i] + 1);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return wave[a] < wave[b]; });
    std::vector<vx_edit> sorted(n);
    for (int i = 0; i < n; i++) sorted[i] = edits[order[i]];

    // the chunks the batch can read or write (31 voxels = two chunks around each edit): their lightmaps
    // have to be there, their border planes are refreshed afterwards
    std::vector<int> touched;
    touched.reserve((size_t)n * 9);
    if (ctx->edit_seen.size() != (size_t)ctx->pool_cap) ctx->edit_seen.assign(ctx->pool_cap, 0u);
    if (++ctx->edit_epoch == 0u) {
        std::fill(ctx->edit_seen.begin(), ctx->edit_seen.end(), 0u);
        ctx->edit_epoch = 1u;
    }
    for (int i = 0; i < n; i++)
        for (int dz = -2; dz <= 2; dz++)
            for (int dx = -2; dx <= 2; dx++) {
                const int q = ctx->pool_of((edits[i].x >> 4) + dx, (edits[i].z >> 4) + dz);
                if (q >= 0 && ctx->edit_seen[q] != ctx->edit_epoch) {
                    ctx->edit_seen[q] = ctx->edit_epoch;
                    touched.push_back(q);
                }
            }
    if (vx_materialize(ctx, touched)) return -1;
    cudaStream_t st = ctx->stream;
    if (!ctx->d_edit_scratch) {
        VX_CUDA(cudaFuncSetAttribute(k_edit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EDIT_SMEM));
        int per_sm = 0;
        VX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_edit, ET, EDIT_SMEM));
        if (per_sm < 1) {
            ctx->fail("vx_light_edit: k_edit does not fit an SM");
            return -1;
        }
        ctx->edit_resident = per_sm * ctx->n_sm;
        VX_CUDA(cudaMalloc((void**)&ctx->d_edit_scratch, (size_t)MAX_CONC * 3 * QCAP * sizeof(uint32_t)));
    }
    // VXGPU_EDIT_GLOBAL=1: every edit runs on the lightmaps in global memory (the path an edit takes
    // when it leaves its shared-memory box), for the tests
    const char* force_global = std::getenv("VXGPU_EDIT_GLOBAL");
    int try_local = !(force_global && force_global[0] == '1');
    // Launches: edits in wave order (a predecessor always comes first), as many per launch as can be
    // resident at once; inside a launch an edit waits for exactly the earlier edits it overlaps with,
    // not for its whole wave.  Predecessors in an earlier launch are done by stream order.  An edit
    // with more than MAXD predecessors inside the running launch starts a new one.
    std::vector<int> pos(n);
    for (int k = 0; k < n; k++) pos[order[k]] = k;
    const int per_launch = std::max(1, std::min(MAX_CONC, ctx->edit_resident));
    // everything the call's kernels read goes up in one copy from pinned memory:
    // counters[16] (zero) | done[n] (zero) | deps[n][MAXD] | touched[m] | edits[n]
    static_assert(sizeof(vx_edit) % sizeof(int32_t) == 0, "vx_edit is a whole number of words");
    const size_t at_done = 16, at_deps = at_done + (size_t)n, at_touched = at_deps + (size_t)n * MAXD,
                 at_edits = at_touched + touched.size(),
                 words = at_edits + (size_t)n * (sizeof(vx_edit) / sizeof(int32_t));
    if (words > ctx->edit_buf_cap) {
        if (ctx->h_edit_buf) cudaFreeHost(ctx->h_edit_buf);
        cudaFree(ctx->d_edit_buf);
        ctx->h_edit_buf = ctx->d_edit_buf = nullptr;
        ctx->edit_buf_cap = 0;
        const size_t cap = words + words / 2;
        VX_CUDA(cudaMallocHost((void**)&ctx->h_edit_buf, cap * sizeof(int32_t)));
        VX_CUDA(cudaMalloc((void**)&ctx->d_edit_buf, cap * sizeof(int32_t)));
        ctx->edit_buf_cap = cap;
    }
    int32_t* hb = ctx->h_edit_buf;
    std::memset(hb, 0, at_deps * sizeof(int32_t));
    std::fill(hb + at_deps, hb + at_touched, -1);
    std::vector<int> launch_at;  // first sorted index of every launch
    int start = 0;
    launch_at.push_back(0);
    std::vector<int> mine;
    for (int k = 0; k < n; k++) {
        mine.clear();
        for (int at = pred_at[order[k]]; at < pred_at[order[k] + 1]; at++)
            if (pos[pred[at]] >= start) mine.push_back(pos[pred[at]] - start);
        if (k - start >= per_launch || (int)mine.size() > MAXD) {
            start = k;
            launch_at.push_back(k);
            mine.clear();
        }
        for (size_t q = 0; q < mine.size(); q++) hb[at_deps + (size_t)k * MAXD + q] = mine[q];
    }
    launch_at.push_back(n);
    if (!touched.empty()) std::memcpy(hb + at_touched, touched.data(), touched.size() * sizeof(int32_t));
    std::memcpy(hb + at_edits, sorted.data(), (size_t)n * sizeof(vx_edit));
    int32_t* db = ctx->d_edit_buf;
    VX_CUDA(cudaMemcpyAsync(db, hb, words * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VX_CUDA(cudaEventRecord(ctx->ev0, st));
    int n_launches = 0;
    for (size_t li = 0; li + 1 < launch_at.size(); li++) {
        const int at = launch_at[li], cnt = launch_at[li + 1] - at;
        const vx_edit* ed = reinterpret_cast<const vx_edit*>(db + at_edits) + at;
        const int32_t* dp = db + at_deps + (size_t)at * MAXD;
        int32_t* dn = db + at_done + at;
        int* errp = db;
        int* statp = db + 1;  // edits that left their box
        void* args[] = {(void*)&ctx->W, (void*)&ed, (void*)&ctx->d_edit_scratch, (void*)&errp, (void*)&dp, (void*)&dn,
                        (void*)&try_local, (void*)&statp};
        VX_LAUNCH(ctx, KID_EDIT,
                  VX_CUDA(cudaLaunchCooperativeKernel((void*)k_edit, dim3(cnt), dim3(ET), args, EDIT_SMEM, st)));
        n_launches++;
    }
    if (!touched.empty())
        VX_LAUNCH(ctx, KID_FACES, k_faces_refresh_dirty<<<(unsigned)touched.size(), 256, 0, st>>>(ctx->W, db + at_touched));
    VX_CUDA(cudaEventRecord(ctx->ev1, st));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(ctx->h_counters, db, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->last_edit_ms, ctx->ev0, ctx->ev1);
    ctx->last_edit_waves = n_waves;
    ctx->last_edit_launches = n_launches;
    ctx->last_edit_global = ctx->h_counters[1];
#ifdef VX_EDIT_STATS
    {
        static long long acc[8] = {0};
        static int calls = 0;
        for (int k = 0; k < 7; k++) acc[k] += ctx->h_counters[3 + k];
        acc[7] += n;
        if (++calls % 50 == 0) {
            unsigned long long hh[16];
            cudaMemcpyFromSymbol(hh, g_edit_hist, sizeof(hh));
            fprintf(stderr, "[edit stats] removal levels: n<64: %llu x %.2f us; n<512: %llu x %.2f us; n<=1024: %llu x %.2f us; more: %llu x %.2f us; preamble %llu x %.2f us\n",
                    hh[1], hh[0] / 1e3 / (hh[1] + 1e-9), hh[3], hh[2] / 1e3 / (hh[3] + 1e-9), hh[5], hh[4] / 1e3 / (hh[5] + 1e-9),
                    hh[7], hh[6] / 1e3 / (hh[7] + 1e-9), hh[9], hh[8] / 1e3 / (hh[9] + 1e-9));
            unsigned long long cy[8];
            cudaMemcpyFromSymbol(cy, g_edit_cyc, sizeof(cy));
            unsigned long long lv = hh[1] + hh[3] + hh[5] + hh[7];
            fprintf(stderr, "[edit stats] thread 0, cycles per removal level: %.0f, of which at the barrier %.0f\n",
                    (double)cy[4] / lv, (double)cy[5] / lv);
            unsigned long long r[6];
            cudaMemcpyFromSymbol(r, g_edit_rounds, sizeof(r));
            fprintf(stderr, "[edit stats] per edit: %.1f removal levels with %.0f entries in %.1f us, %.1f add rounds with %.0f entries in %.1f us\n",
                    (double)r[0] / acc[7], (double)r[1] / acc[7], r[4] / 1e3 / acc[7], (double)r[2] / acc[7], (double)r[3] / acc[7],
                    r[5] / 1e3 / acc[7]);
        }
        if (calls % 50 == 0)
            fprintf(stderr, "[edit stats] %lld edits, us per edit: wait %.1f set %.1f fill %.1f box %.1f back %.1f global %.1f end %.1f\n",
                    acc[7], acc[0] / 1e3 / acc[7], acc[1] / 1e3 / acc[7], acc[2] / 1e3 / acc[7], acc[3] / 1e3 / acc[7],
                    acc[4] / 1e3 / acc[7], acc[5] / 1e3 / acc[7], acc[6] / 1e3 / acc[7]);
    }
#endif
    if (ctx->h_counters[0]) {
        ctx->fail("vx_light_edit: light queue overflow (an edit touched more than 32768 voxels per level)");
        return -1;
    }
    return 0;
}

extern "C" int vx_last_edit_stats(vx_ctx* ctx, float* ms, int32_t* waves, int32_t* launches) {
    if (!ctx) return -1;
    if (ms) *ms = ctx->last_edit_ms;
    if (waves) *waves = ctx->last_edit_waves;
    if (launches) *launches = ctx->last_edit_launches;
    return 0;
}

extern "C" int32_t vx_last_edit_global(vx_ctx* ctx) { return ctx ? ctx->last_edit_global : -1; }
