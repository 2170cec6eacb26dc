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
// Edits whose regions of influence cannot overlap run concurrently: the host
// groups a batch into waves (an edit's wave is one more than the latest wave of
// any earlier edit whose column is within |dx| + |dz| <= 64 voxels), and each wave is one launch.
#include <algorithm>
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

// queue entry: dx+32 (6) | dz+32 (6) | y (8) | light (4) | channel (2)
__device__ __forceinline__ uint32_t pack(int dx, int dz, int y, uint32_t light, int ch) {
    return (uint32_t)(dx + 32) | ((uint32_t)(dz + 32) << 6) | ((uint32_t)y << 12) | (light << 20) |
           ((uint32_t)ch << 24);
}

struct Ed {
    const DevWorld& W;   // the kernel's __grid_constant__ parameter: no per-thread copy of its 200+ bytes
    int ex, ez;          // origin of the packed coordinates
    uint32_t* q[3];
    int* err;
    const int* pools5;   // shared memory: pool index of the 5x5 chunks around the edit's chunk (an edit
    int bx, bz;          // reaches 31 voxels in x / z), chunk coordinates of its corner
    uint32_t (*sq)[3072];        // shared memory: the first SQ entries of each queue
    const uint16_t* emis;        // shared memory: DevDef::emission of block ids < 256
    int* chunk_flags;            // shared memory, per chunk of the 5x5 table: faces to refresh | 16: light changed
};
constexpr int SQ = 3072;

struct Loc {
    int p, i;
};

__device__ __forceinline__ bool locate(const DevWorld& W, int x, int y, int z, Loc& l) {
    if (y < 0 || y >= CH) return false;  // Chunks::getChunkByVoxel, Chunks.cpp:142-151
    l.p = W.pool_of(x >> 4, z >> 4);
    if (l.p < 0) return false;
    l.i = vidx(x & 15, y, z & 15);
    return true;
}
// the same through the edit's 5x5 table: no slot-table load on the way to every voxel
__device__ __forceinline__ bool locate(const Ed& E, int x, int y, int z, Loc& l) {
    if (y < 0 || y >= CH) return false;
    const unsigned cx = (unsigned)((x >> 4) - E.bx), cz = (unsigned)((z >> 4) - E.bz);
    l.p = (cx < 5u && cz < 5u) ? E.pools5[cz * 5 + cx] : E.W.pool_of(x >> 4, z >> 4);
    if (l.p < 0) return false;
    l.i = vidx(x & 15, y, z & 15);
    return true;
}

__device__ __forceinline__ uint32_t* light_word(const DevWorld& W, const Loc& l) {
    return reinterpret_cast<uint32_t*>(W.light_of(l.p)) + (l.i >> 1);
}
__device__ __forceinline__ int nib_shift(const Loc& l, int ch) { return (l.i & 1) * 16 + ch * 4; }

__device__ __forceinline__ uint32_t read_nib(const DevWorld& W, const Loc& l, int ch) {
    return (*reinterpret_cast<volatile uint32_t*>(light_word(W, l)) >> nib_shift(l, ch)) & 0xFu;
}

// faceL (the border-plane copy of light) is not written here: concurrent entries of a level can
// change different nibbles of one voxel, and a copy updated nibble by nibble next to the CAS on
// the light word can end up stale.  A changed border voxel only marks its chunk's face dirty;
// k_faces_refresh_dirty copies those faces from the lightmap once the batch's waves are done
// (nothing reads faceL in between: it only feeds the seeding of later builds).
__device__ __forceinline__ void note_change(const Ed& E, int x, int z, const Loc& l) {
    const int lx = l.i & 15, lz = (l.i >> 4) & 15;
    int bits = 16;
    if (lx == 0) bits |= 1 << FACE_XM;
    if (lx == 15) bits |= 1 << FACE_XP;
    if (lz == 0) bits |= 1 << FACE_ZM;
    if (lz == 15) bits |= 1 << FACE_ZP;
    const unsigned cx = (unsigned)((x >> 4) - E.bx), cz = (unsigned)((z >> 4) - E.bz);
    if (cx < 5u && cz < 5u) {
        int* f = E.chunk_flags + cz * 5 + cx;  // flushed to ChunkMeta when the edit is done
        if ((*reinterpret_cast<volatile int*>(f) & bits) != bits) atomicOr(f, bits);
    } else {  // (cannot happen: an edit reaches two chunks at most)
        if (bits & 15) atomicOr(&E.W.meta[l.p].face_dirty, bits & 15);
        E.W.meta[l.p].lstate = LS_UNKNOWN;
    }
}

// nibble := v if pred(old nibble); returns true (and the old nibble) on success
template <class Pred>
__device__ __forceinline__ bool cas_nib(const Ed& E, int x, int z, const Loc& l, int ch, uint32_t v,
                                        Pred pred, uint32_t* old_out) {
    const DevWorld& W = E.W;
    uint32_t* wp = light_word(W, l);
    const int sh = nib_shift(l, ch);
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
    while (true) {
        const uint32_t on = (old >> sh) & 0xFu;
        if (!pred(on)) return false;
        const uint32_t nw = (old & ~(0xFu << sh)) | (v << sh);
        const uint32_t got = atomicCAS(wp, old, nw);
        if (got == old) {
            *old_out = on;
            if (on != v) note_change(E, x, z, l);
            return true;
        }
        old = got;
    }
}

// queues: the first SQ entries live in shared memory, the rest in the CTA's global scratch
__device__ __forceinline__ void push(const Ed& E, int which, int* count, uint32_t entry) {
    const int at = atomicAdd(count, 1);
    if (at < SQ)
        E.sq[which][at] = entry;
    else if (at < QCAP)
        E.q[which][at] = entry;
    else
        *E.err = 1;
}
__device__ __forceinline__ uint32_t qget(const Ed& E, int which, int k) {
    return k < SQ ? E.sq[which][k] : E.q[which][k];
}

__constant__ int c_nb[6][3] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {1, 0, 0}, {-1, 0, 0}};

// LightSolver::solve removal loop body for one popped entry (LightSolver.cpp:60-95).
// Everything the six neighbours need — block id, light word — is fetched first, then the six
// compare-and-swaps are issued back to back, then their results are looked at: three memory
// round trips per entry instead of a dozen dependent ones.  A CAS that lost a race (another entry
// of the level changed the word) takes the original one-at-a-time path with the fresh value.
// A nibble that does not hold level - 1 is not written by anyone during this level, so the value
// fetched up front is what a later read would return.
__device__ void removal_entry(const Ed& E, uint32_t e, int nextq, int* n_next, int addq, int* n_add) {
    const DevWorld& W = E.W;
    const int x = E.ex + (int)(e & 63u) - 32, z = E.ez + (int)((e >> 6) & 63u) - 32;
    const int y = (e >> 12) & 255u, ch = (e >> 24) & 3u;
    const uint32_t el = (e >> 20) & 0xFu;
    Loc l[6];
    bool ok[6];
    uint32_t id[6], word[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ok[k] = locate(E, x + c_nb[k][0], y + c_nb[k][1], z + c_nb[k][2], l[k]);
        id[k] = 0;
        word[k] = 0;
        if (ok[k]) {
            id[k] = W.vox_of(l[k].p)[l[k].i] & 0xFFFFu;
            word[k] = *reinterpret_cast<volatile uint32_t*>(light_word(W, l[k]));
        }
    }
    uint32_t emission[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        emission[k] = 0;
        if (ok[k] && id[k] != 0 && id[k] < (uint32_t)W.n_defs)
            emission[k] = ((id[k] < 256u ? (uint32_t)E.emis[id[k]] : __ldg(&W.defs[id[k]].emission)) >> (4 * ch)) & 0xFu;
    }
    bool want[6];
    uint32_t got[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int sh = nib_shift(l[k], ch);
        const uint32_t on = (word[k] >> sh) & 0xFu;
        want[k] = ok[k] && el >= 2 && on != 0 && on == el - 1;
        got[k] = 0;
        if (want[k]) got[k] = atomicCAS(light_word(W, l[k]), word[k], (word[k] & ~(0xFu << sh)) | (emission[k] << sh));
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (!ok[k]) continue;
        const int nx = x + c_nb[k][0], ny = y + c_nb[k][1], nz = z + c_nb[k][2];
        W.meta[l[k].p].modified = 1;
        const int sh = nib_shift(l[k], ch);
        bool cleared = false;
        uint32_t old = (word[k] >> sh) & 0xFu;
        if (want[k]) {
            if (got[k] == word[k]) {
                cleared = true;
                if (old != emission[k]) note_change(E, nx, nz, l[k]);
            } else {
                cleared = cas_nib(E, nx, nz, l[k], ch, emission[k], [&](uint32_t on) { return on != 0 && on == el - 1; }, &old);
            }
        }
        if (cleared) {
            if (emission[k]) push(E, addq, n_add, pack(nx - E.ex, nz - E.ez, ny, emission[k], ch));
            push(E, nextq, n_next, pack(nx - E.ex, nz - E.ez, ny, old, ch));
        } else {
            // (after a lost race: what the winner left — 0 or the block's emission — exactly as the
            // second pop of the reference's queue would read it)
            const uint32_t light = want[k] ? read_nib(W, l[k], ch) : old;
            if (light >= el) push(E, addq, n_add, pack(nx - E.ex, nz - E.ez, ny, light, ch));
        }
    }
}

// LightSolver::solve add loop body for one popped entry (LightSolver.cpp:97-125); same shape.
// Light only rises during the add rounds, so a neighbour that already holds el - 1 or more when
// it is fetched never needs this entry.
__device__ void add_entry(const Ed& E, uint32_t e, int nextq, int* n_next) {
    const DevWorld& W = E.W;
    const int x = E.ex + (int)(e & 63u) - 32, z = E.ez + (int)((e >> 6) & 63u) - 32;
    const int y = (e >> 12) & 255u, ch = (e >> 24) & 3u;
    const uint32_t el = (e >> 20) & 0xFu;
    Loc l[6];
    bool ok[6];
    uint32_t lpw[6], word[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ok[k] = locate(E, x + c_nb[k][0], y + c_nb[k][1], z + c_nb[k][2], l[k]);
        lpw[k] = 0;
        word[k] = 0;
        if (ok[k]) {
            lpw[k] = W.lp_of(l[k].p)[l[k].i >> 5];
            word[k] = *reinterpret_cast<volatile uint32_t*>(light_word(W, l[k]));
        }
    }
    bool want[6];
    uint32_t got[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int sh = nib_shift(l[k], ch);
        want[k] = ok[k] && el >= 2 && ((lpw[k] >> (l[k].i & 31)) & 1u) && ((word[k] >> sh) & 0xFu) + 2 <= el;
        got[k] = 0;
        if (want[k]) got[k] = atomicCAS(light_word(W, l[k]), word[k], (word[k] & ~(0xFu << sh)) | ((el - 1) << sh));
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (!ok[k]) continue;
        W.meta[l[k].p].modified = 1;
        if (!want[k]) continue;
        bool raised = got[k] == word[k];
        if (raised) {
            note_change(E, x + c_nb[k][0], z + c_nb[k][2], l[k]);
        } else {
            uint32_t old;
            raised = cas_nib(E, x + c_nb[k][0], z + c_nb[k][2], l[k], ch, el - 1, [&](uint32_t on) { return on + 2 <= el; }, &old);
        }
        if (raised) push(E, nextq, n_next, pack(x + c_nb[k][0] - E.ex, z + c_nb[k][2] - E.ez, y + c_nb[k][1], el - 1, ch));
    }
}

struct EdShared {
    uint32_t seeds[272];  // removal seeds of this phase (LightSolver::remove)
    int n_seed;
    int n[3];             // queue fill counts
    int red;
};

// LightSolver::remove (LightSolver.cpp:37-48)
__device__ __forceinline__ void ls_remove(const Ed& E, EdShared& S, int ch, int x, int y, int z) {
    Loc l;
    if (!locate(E, x, y, z, l)) return;
    uint32_t old;
    if (cas_nib(E, x, z, l, ch, 0, [](uint32_t on) { return on != 0; }, &old)) {
        const int at = atomicAdd(&S.n_seed, 1);
        S.seeds[at] = pack(x - E.ex, z - E.ez, y, old, ch);
    }
}

// LightSolver::add(x,y,z,emission) (LightSolver.cpp:18-31); entries go to queue `addq`
__device__ __forceinline__ void ls_add(const Ed& E, EdShared& S, int addq, int ch, int x, int y, int z,
                                       uint32_t emission) {
    if (emission <= 1) return;
    Loc l;
    if (!locate(E, x, y, z, l)) return;
    uint32_t old;
    if (cas_nib(E, x, z, l, ch, emission, [&](uint32_t on) { return emission >= on; }, &old)) {
        push(E, addq, &S.n[addq], pack(x - E.ex, z - E.ez, y, emission, ch));
        E.W.meta[l.p].modified = 1;
    }
}

// Runs LightSolver::solve for every channel that has queued work: the removal
// seeds in S.seeds, the add entries already in queue 2.
__device__ void solve_queues(const Ed& E, EdShared& S) {
    const int tid = threadIdx.x;
    __syncthreads();
    // ---- removal, level-synchronous.  queues 0/1 ping-pong, boundary -> queue 2
    if (S.n_seed > 0) {
        int cur = 0;
        if (tid == 0) S.n[0] = S.n[1] = 0;
        __syncthreads();
        for (int level = 15; level >= 1; level--) {
            for (int k = tid; k < S.n_seed; k += ET)
                if (((S.seeds[k] >> 20) & 0xFu) == (uint32_t)level) push(E, cur, &S.n[cur], S.seeds[k]);
            __syncthreads();
            const int n = min(S.n[cur], QCAP);
            for (int k = tid; k < n; k += ET)
                removal_entry(E, qget(E, cur, k), cur ^ 1, &S.n[cur ^ 1], 2, &S.n[2]);
            __syncthreads();
            if (tid == 0) S.n[cur] = 0;
            cur ^= 1;
            __syncthreads();
        }
        if (tid == 0) S.n_seed = 0;
        __syncthreads();
    }
    // ---- add rounds: queue 2 -> 0 -> 2 ...
    int cur = 2, nxt = 0;
    if (tid == 0) S.n[0] = S.n[1] = 0;
    __syncthreads();
    while (true) {
        const int n = min(S.n[cur], QCAP);
        if (n == 0) break;
        for (int k = tid; k < n; k += ET) add_entry(E, qget(E, cur, k), nxt, &S.n[nxt]);
        __syncthreads();
        if (tid == 0) S.n[cur] = 0;
        const int t = cur;
        cur = nxt;
        nxt = t;
        __syncthreads();
    }
    // leave queue 2 as the (empty) add queue for the next phase
    if (cur != 2 && tid == 0) S.n[2] = 0;
    __syncthreads();
}

__device__ __forceinline__ bool nonair_at(const DevWorld& W, int p, int lx, int y, int lz) {
    return (W.vox_of(p)[vidx(lx, y, lz)] & 0xFFFFu) != 0;
}

// grid: one CTA per edit of the wave
constexpr int MAXD = 8;  // predecessors an edit can wait for inside one launch (more: the launch is cut before it)

// grid: one CTA per edit of the launch.  deps[b][MAXD] = the earlier edits of this launch (block
// indices, -1 padded) whose reach overlaps this one's: the CTA starts when their `done` flags are up
// (cooperative launch: every CTA is resident, and a predecessor always has a lower block index).
__global__ void __launch_bounds__(ET) k_edit(const __grid_constant__ DevWorld W, const vx_edit* edits, uint32_t* scratch,
                                             int* err, const int32_t* deps, int32_t* done) {
    __shared__ EdShared S;
    __shared__ int s_hi, s_lo;
    const int tid = threadIdx.x;
    const vx_edit ed = edits[blockIdx.x];
    const int x = ed.x, y = ed.y, z = ed.z;
    const uint32_t id = ed.id;
    if (threadIdx.x < MAXD) {
        const int d = deps[blockIdx.x * MAXD + threadIdx.x];
        if (d >= 0) {
            unsigned backoff = 64;
            while (*reinterpret_cast<volatile int32_t*>(done + d) == 0) {
                __nanosleep(backoff);
                if (backoff < 2048) backoff *= 2;
            }
        }
        __threadfence();  // what the predecessors wrote is visible to everything below
    }
    __syncthreads();
    __shared__ int s_pools5[25], s_chunk_flags[25];
    __shared__ uint32_t s_q[3][SQ];
    __shared__ uint16_t s_emis[256];
    Ed E {W, x, z, {nullptr, nullptr, nullptr}, nullptr, s_pools5, (x >> 4) - 2, (z >> 4) - 2, s_q, s_emis, s_chunk_flags};
    if (tid < 25) {
        s_pools5[tid] = W.pool_of((x >> 4) - 2 + tid % 5, (z >> 4) - 2 + tid / 5);
        s_chunk_flags[tid] = 0;
    }
    if (tid < 256) s_emis[tid] = tid < W.n_defs ? (uint16_t)__ldg(&W.defs[tid].emission) : (uint16_t)0;
    __syncthreads();
    E.q[0] = scratch + (size_t)blockIdx.x * 3 * QCAP;
    E.q[1] = E.q[0] + QCAP;
    E.q[2] = E.q[1] + QCAP;
    E.err = err;
    Loc here;
    if (!locate(W, x, y, z, here) || id >= (uint32_t)W.n_defs) {  // blocks_agent.cpp:18-27
        if (tid == 0) done[blockIdx.x] = 1;
        return;
    }
    const int p = here.p, lx = x & 15, lz = z & 15;
    const uint32_t bf = __ldg(&W.defs[id].flags);
    const uint32_t bem = __ldg(&W.defs[id].emission);
    if (tid == 0) {
        S.n_seed = 0;
        S.n[0] = S.n[1] = S.n[2] = 0;
        s_hi = -1;
    }
    __syncthreads();

    // ---- blocks_agent::set
    if (tid == 0) {
        vx_voxel* vp = W.vox_of(p) + here.i;
        const uint32_t oldid = *vp & 0xFFFFu;
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
        int16_t* lc = W.laycnt_of(p);
        if (oldid != 0 && id == 0) lc[y]--;
        if (oldid == 0 && id != 0) lc[y]++;
        ChunkMeta* m = &W.meta[p];
        m->modified = 1;
        if (y < m->bottom) {
            m->bottom = y;
        } else if (y + 1 > m->top) {
            m->top = y + 1;
        } else if (id == 0) {  // Chunk::updateHeights
            for (int k = 0; k < CH; k++)
                if (lc[k] > 0) {
                    m->bottom = k;
                    break;
                }
            for (int k = CH - 1; k >= 0; k--)
                if (lc[k] > 0) {
                    m->top = k + 1;
                    break;
                }
        }
        int q;
        if (lx == 0 && (q = W.pool_of((x >> 4) - 1, z >> 4)) >= 0) W.meta[q].modified = 1;
        if (lz == 0 && (q = W.pool_of(x >> 4, (z >> 4) - 1)) >= 0) W.meta[q].modified = 1;
        if (lx == 15 && (q = W.pool_of((x >> 4) + 1, z >> 4)) >= 0) W.meta[q].modified = 1;
        if (lz == 15 && (q = W.pool_of(x >> 4, (z >> 4) + 1)) >= 0) W.meta[q].modified = 1;
    }
    __syncthreads();
    // skycol of the edited column (only Lighting::prebuildSkyLight reads it)
    {
        // (threads 0 .. 255 are the column's 256 voxels here and in the column walks below)
        if (tid < CH) {
            const uint32_t vid = W.vox_of(p)[vidx(lx, tid, lz)] & 0xFFFFu;
            const uint32_t f = vid < (uint32_t)W.n_defs ? __ldg(&W.defs[vid].flags) : 0u;
            if (!(f & DF_SLP)) atomicMax(&s_hi, tid);
        }
        __syncthreads();
        // ... and where buildSkyLight would start on a pristine lightmap (see k_derive)
        const int sky = s_hi;
        __syncthreads();
        if (tid == 0) s_lo = sky < 0 ? -1 : 0;
        __syncthreads();
        {
            const int ni = vidx(lx, min(tid, CH - 1), lz);
            if (tid < CH && tid <= sky && ((W.lp_of(p)[ni >> 5] >> (ni & 31)) & 1u)) atomicMax(&s_lo, tid);
        }
        __syncthreads();
        if (tid == 0) {
            W.skycol_of(p)[lz * 16 + lx] = (int16_t)sky;
            W.ystar0_of(p)[lz * 16 + lx] = (int16_t)s_lo;
            s_hi = -1;
            s_lo = CH;
        }
        __syncthreads();
        // sky_min / sky_max of the chunk follow
        if (tid < 256) {
            const int sc = W.skycol_of(p)[tid];
            atomicMax(&s_hi, sc);
            atomicMin(&s_lo, sc);
        }
        __syncthreads();
        if (tid == 0) {
            W.meta[p].sky_max = s_hi;
            W.meta[p].sky_min = s_lo;
            s_hi = -1;
        }
        __syncthreads();
    }

    // ---- Lighting::onBlockSet
    if (tid < 3) ls_remove(E, S, tid, x, y, z);
    if (id == 0) {
        solve_queues(E, S);  // solverR/G/B->solve()
        Loc up;
        const bool sky_above = locate(W, x, y + 1, z, up) && read_nib(W, up, 3) == 0xFu;
        if (sky_above) {
            // walk down from y while the voxel is air (the test reads the *air*
            // def's skyLightPassing, Lighting.cpp:173-180)
            int stop = -1;  // highest i <= y that ends the walk
            if (bf & DF_SLP) {
                if (tid <= y && nonair_at(W, p, lx, tid, lz)) atomicMax(&s_hi, tid);
                __syncthreads();
                stop = s_hi;
            }
            if (tid > stop && tid <= y) ls_add(E, S, 2, 3, x, tid, z, 0xFu);
        }
        if (tid < 24) {
            const int k = tid >> 2, ch = tid & 3;
            const int nx = x + c_nb[k][0], ny = y + c_nb[k][1], nz = z + c_nb[k][2];
            Loc l;
            if (locate(W, nx, ny, nz, l))  // LightSolver::add(x,y,z), :33-35
                ls_add(E, S, 2, ch, nx, ny, nz, read_nib(W, l, ch));
        }
        solve_queues(E, S);
    } else {
        if (!(bf & DF_SLP)) {
            if (tid == 0) ls_remove(E, S, 3, x, y, z);
            // column below: i = y-1 .. b, b = first i with i == 0 or voxel(i-1) non-air
            if (y >= 1) {
                if (tid <= y - 2 && nonair_at(W, p, lx, tid, lz)) atomicMax(&s_hi, tid);
                __syncthreads();
                const int b = s_hi + 1;
                if (tid >= b && tid <= y - 1) ls_remove(E, S, 3, x, tid, z);
            }
        }
        solve_queues(E, S);
        if (bem & 0x0FFFu) {
            if (tid < 3) ls_add(E, S, 2, tid, x, y, z, (bem >> (4 * tid)) & 0xFu);
            solve_queues(E, S);
        }
    }
    // what the edit changed, per chunk: faces for k_faces_refresh_dirty, and the lightmap is no
    // longer what prebuildSkyLight left
    __syncthreads();
    if (tid < 25 && s_chunk_flags[tid] && s_pools5[tid] >= 0) {
        ChunkMeta* m = &W.meta[s_pools5[tid]];
        if (s_chunk_flags[tid] & 15) atomicOr(&m->face_dirty, s_chunk_flags[tid] & 15);
        m->lstate = LS_UNKNOWN;
    }
    // everything this edit wrote is out before its successors are let go
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(done + blockIdx.x, 1);
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
    if (ctx->pool_cap == 0) {
        ctx->fail("vx_light_edit: configure the window first");
        return -1;
    }
    // ---- waves: edit i runs after every earlier edit within CONFLICT voxels
    std::vector<int> wave(n, 0);
    std::vector<std::vector<int>> pred(n);  // earlier edits whose reach overlaps edit i's
    {
        struct Rec { int x, z, wave, idx; };
        std::unordered_map<long long, std::vector<Rec>> cells;
        auto key = [](int cx, int cz) { return ((long long)cx << 32) ^ (unsigned int)cz; };
        for (int i = 0; i < n; i++) {
            const int cx = edits[i].x >> 6, cz = edits[i].z >> 6;  // 64-voxel cells
            int w = 0;
            for (int dz = -1; dz <= 1; dz++)
                for (int dx = -1; dx <= 1; dx++) {
                    auto it = cells.find(key(cx + dx, cz + dz));
                    if (it == cells.end()) continue;
                    for (const Rec& r : it->second)
                        // light moves one voxel per step along an axis: an edit's reads and writes stay within
                        // |dx| + |dz| <= 32 of its column (any y: a sky column can fall all the way down)
                        if (std::abs(r.x - edits[i].x) + std::abs(r.z - edits[i].z) <= CONFLICT) {
                            w = std::max(w, r.wave + 1);
                            pred[i].push_back(r.idx);
                        }
                }
            wave[i] = w;
            cells[key(cx, cz)].push_back(Rec {edits[i].x, edits[i].z, w, i});
        }
    }
    int n_waves = 0;
    for (int i = 0; i < n; i++) n_waves = std::max(n_waves, wave[i] + 1);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return wave[a] < wave[b]; });
    std::vector<vx_edit> sorted(n);
    for (int i = 0; i < n; i++) sorted[i] = edits[order[i]];

    {
        // an edit reads and writes light up to 15 voxels away: its chunk and the eight around it
        std::vector<int> near;
        for (int i = 0; i < n; i++)
            for (int dz = -1; dz <= 1; dz++)
                for (int dx = -1; dx <= 1; dx++) near.push_back(ctx->pool_of((edits[i].x >> 4) + dx, (edits[i].z >> 4) + dz));
        std::sort(near.begin(), near.end());
        near.erase(std::unique(near.begin(), near.end()), near.end());
        if (vx_materialize(ctx, near)) return -1;
    }
    cudaStream_t st = ctx->stream;
    if (!ctx->d_edit_scratch) {
        VX_CUDA(cudaMalloc((void**)&ctx->d_edit_scratch, (size_t)MAX_CONC * 3 * QCAP * sizeof(uint32_t)));
        int per_sm = 0;
        VX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_edit, ET, 0));
        ctx->edit_resident = std::max(1, per_sm) * ctx->n_sm;
    }
    if ((size_t)n > ctx->edits_cap) {
        if (ctx->d_edits) cudaFree(ctx->d_edits);
        ctx->d_edits = nullptr;
        ctx->edits_cap = 0;
        VX_CUDA(cudaMalloc((void**)&ctx->d_edits, (size_t)n * sizeof(vx_edit)));
        ctx->edits_cap = n;
    }
    VX_CUDA(cudaMemcpyAsync(ctx->d_edits, sorted.data(), (size_t)n * sizeof(vx_edit),
                            cudaMemcpyHostToDevice, st));
    VX_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(int32_t), st));
    VX_CUDA(cudaEventRecord(ctx->ev0, st));
    // Launches: edits in wave order (a predecessor always comes first), as many per launch as can be
    // resident at once; inside a launch an edit waits for exactly the earlier edits it overlaps with,
    // not for its whole wave.  Predecessors in an earlier launch are done by stream order.  An edit
    // with more than MAXD predecessors inside the running launch starts a new one.
    std::vector<int> pos(n);
    for (int k = 0; k < n; k++) pos[order[k]] = k;
    const int per_launch = std::max(1, std::min(MAX_CONC, ctx->edit_resident));
    std::vector<int32_t> h_deps((size_t)n * MAXD, -1);
    std::vector<int> launch_at;  // first sorted index of every launch
    int start = 0;
    launch_at.push_back(0);
    for (int k = 0; k < n; k++) {
        std::vector<int> mine;
        for (int j : pred[order[k]])
            if (pos[j] >= start) mine.push_back(pos[j] - start);
        if (k - start >= per_launch || (int)mine.size() > MAXD) {
            start = k;
            launch_at.push_back(k);
            mine.clear();
        }
        for (size_t q = 0; q < mine.size(); q++) h_deps[(size_t)k * MAXD + q] = mine[q];
    }
    launch_at.push_back(n);
    if ((size_t)n > ctx->edit_deps_cap) {
        cudaFree(ctx->d_edit_deps);
        cudaFree(ctx->d_edit_done);
        ctx->d_edit_deps = ctx->d_edit_done = nullptr;
        ctx->edit_deps_cap = 0;
        VX_CUDA(cudaMalloc((void**)&ctx->d_edit_deps, (size_t)n * MAXD * sizeof(int32_t)));
        VX_CUDA(cudaMalloc((void**)&ctx->d_edit_done, (size_t)n * sizeof(int32_t)));
        ctx->edit_deps_cap = n;
    }
    VX_CUDA(cudaMemcpyAsync(ctx->d_edit_deps, h_deps.data(), h_deps.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    VX_CUDA(cudaMemsetAsync(ctx->d_edit_done, 0, (size_t)n * sizeof(int32_t), st));
    int n_launches = 0;
    for (size_t li = 0; li + 1 < launch_at.size(); li++) {
        const int at = launch_at[li], cnt = launch_at[li + 1] - at;
        const vx_edit* ed = ctx->d_edits + at;
        const int32_t* dp = ctx->d_edit_deps + (size_t)at * MAXD;
        int32_t* dn = ctx->d_edit_done + at;
        int* errp = ctx->d_counters;
        void* args[] = {(void*)&ctx->W, (void*)&ed, (void*)&ctx->d_edit_scratch, (void*)&errp, (void*)&dp, (void*)&dn};
        VX_LAUNCH(ctx, KID_EDIT, VX_CUDA(cudaLaunchCooperativeKernel((void*)k_edit, dim3(cnt), dim3(ET), args, 0, st)));
        n_launches++;
    }
    {
        // border planes of the chunks the batch may have written (31 voxels = two chunks around each edit)
        std::vector<int> touched;
        for (int i = 0; i < n; i++)
            for (int dz = -2; dz <= 2; dz++)
                for (int dx = -2; dx <= 2; dx++) {
                    const int q = ctx->pool_of((edits[i].x >> 4) + dx, (edits[i].z >> 4) + dz);
                    if (q >= 0) touched.push_back(q);
                }
        std::sort(touched.begin(), touched.end());
        touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
        if (!touched.empty()) {
            if (vx_upload_list(ctx, touched, 0)) return -1;
            VX_LAUNCH(ctx, KID_FACES, k_faces_refresh_dirty<<<(unsigned)touched.size(), 256, 0, st>>>(ctx->W, ctx->d_list));
        }
    }
    VX_CUDA(cudaEventRecord(ctx->ev1, st));
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VX_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->last_edit_ms, ctx->ev0, ctx->ev1);
    ctx->last_edit_waves = n_waves;
    ctx->last_edit_launches = n_launches;
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
