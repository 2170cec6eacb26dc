// oracle/libvxoracle.so — a from-scratch CPU restatement of the reference's
// chunk light + mesh path, behind the C API of oracle/vxoracle.h.
//
// TEST INFRASTRUCTURE ONLY: the checker the CUDA path is compared against when
// oracle/_ref/libvxref.so (the reference's own objects) is not present, and
// the thing validated against that library by tests/test_oracle_vs_reference.py.
// Nothing under voxelengine-cpp_b200/ links, loads or calls this file.
//
// Every function cites the reference code whose behaviour it restates (paths
// relative to /root/reference/src).  It is written sequentially and for
// clarity, not speed.  Parity status: pinned against oracle/_ref on seeded
// worlds here; the reference itself has no tests or golden vectors on this
// path, and GLM is replaced by its published scalar definitions.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <set>
#include <thread>
#include <vector>

#include "vxoracle.h"

namespace {

constexpr int CW = VX_CHUNK_W, CH = VX_CHUNK_H, CD = VX_CHUNK_D, CVOL = VX_CHUNK_VOL;

// ---------------------------------------------------------------------------
// small fp32 vector helpers with the evaluation order GLM's scalar path uses
// ---------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
struct V4 {
    float r, g, b, a;
};
struct I3 {
    int x, y, z;
};

inline V3 v3(float x, float y, float z) { return V3 {x, y, z}; }
inline V3 v3(I3 i) { return V3 {(float)i.x, (float)i.y, (float)i.z}; }
inline I3 trunc3(V3 v) { return I3 {(int)v.x, (int)v.y, (int)v.z}; }
inline V3 operator+(V3 a, V3 b) { return V3 {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3 {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3 {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return V3 {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return V3 {s * a.x, s * a.y, s * a.z}; }
inline I3 operator+(I3 a, I3 b) { return I3 {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline I3 operator-(I3 a, I3 b) { return I3 {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline I3 operator-(I3 a) { return I3 {-a.x, -a.y, -a.z}; }
inline V4 operator+(V4 a, V4 b) { return V4 {a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a}; }
inline V4 operator*(V4 a, V4 b) { return V4 {a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a}; }
inline V4 operator*(V4 a, float s) { return V4 {a.r * s, a.g * s, a.b * s, a.a * s}; }
inline V4 splat4(float s) { return V4 {s, s, s, s}; }

// glm::dot (vec3): tmp = a*b; tmp.x + tmp.y + tmp.z
inline float dot3(V3 a, V3 b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return px + py + pz;
}
// glm::normalize: v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
inline V3 normalize3(V3 v) {
    float inv = 1.0f / std::sqrt(dot3(v, v));
    return v * inv;
}
// glm::cross
inline V3 cross3(V3 x, V3 y) {
    return V3 {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}

// graphics/render/BlocksRenderer.cpp:12
const V3 SUN = {0.2275f, 0.9388f, -0.1005f};

// voxels/Block.cpp:57-92 — axes of every BlockRotProfile variant; slots past
// variantsCount are value-initialised (all zero) exactly as in the reference.
struct Orient {
    I3 ax[3];
    I3 fix;
};
Orient make_orient(I3 x, I3 y, I3 z) {
    // CoordSystem::CoordSystem, voxels/Block.cpp:39-45
    Orient o {{x, y, z}, {0, 0, 0}};
    for (I3 a : {x, y, z}) {
        if (a.x < 0 || a.y < 0 || a.z < 0) o.fix = o.fix - a;
    }
    return o;
}
struct RotProfile {
    Orient v[8];
};
RotProfile make_profile(int which) {
    RotProfile p {};
    I3 X {1, 0, 0}, Y {0, 1, 0}, Z {0, 0, 1};
    if (which == VX_ROT_PIPE) {
        p.v[0] = make_orient({1, 0, 0}, {0, 0, 1}, {0, -1, 0});
        p.v[1] = make_orient({0, 0, 1}, {-1, 0, 0}, {0, -1, 0});
        p.v[2] = make_orient({-1, 0, 0}, {0, 0, -1}, {0, -1, 0});
        p.v[3] = make_orient({0, 0, -1}, {1, 0, 0}, {0, -1, 0});
        p.v[4] = make_orient({1, 0, 0}, {0, 1, 0}, {0, 0, 1});
        p.v[5] = make_orient({1, 0, 0}, {0, -1, 0}, {0, 0, -1});
    } else if (which == VX_ROT_PANE) {
        p.v[0] = make_orient({1, 0, 0}, {0, 1, 0}, {0, 0, 1});
        p.v[1] = make_orient({0, 0, -1}, {0, 1, 0}, {1, 0, 0});
        p.v[2] = make_orient({-1, 0, 0}, {0, 1, 0}, {0, 0, -1});
        p.v[3] = make_orient({0, 0, 1}, {0, 1, 0}, {-1, 0, 0});
    } else {
        for (int i = 0; i < 6; i++) p.v[i] = make_orient(X, Y, Z);
    }
    return p;
}

// maths/util.hpp:29-48,81-87 — util::PseudoRandom
struct PseudoRandom {
    uint16_t seed = 0;
    int rand() {
        seed = (uint16_t)(seed + 0x7ed5 + (seed << 6));
        seed = (uint16_t)(seed ^ 0xc23c ^ (seed >> 9));
        seed = (uint16_t)(seed + 0x1656 + (seed << 3));
        seed = (uint16_t)((seed + 0xa264) ^ (seed << 4));
        seed = (uint16_t)(seed + 0xfd70 - (seed << 3));
        seed = (uint16_t)(seed ^ 0xba49 ^ (seed >> 8));
        return (int)seed;
    }
    static uint64_t shuffle_bits_step(uint64_t x, uint64_t m, unsigned shift) {
        uint64_t t = ((x >> shift) ^ x) & m;
        return (x ^ t) ^ (t << shift);
    }
    void setSeed(long number) {
        uint64_t n = (uint64_t)number;
        n = shuffle_bits_step(n, 0x2222222222222222ull, 1);
        n = shuffle_bits_step(n, 0x0c0c0c0c0c0c0c0cull, 2);
        n = shuffle_bits_step(n, 0x00f000f000f000f0ull, 4);
        seed = (uint16_t)n;
        rand();
    }
};

struct Chunk {
    int cx, cz;
    int bottom = 0, top = CH;  // voxels/Chunk.cpp:16-19
    std::vector<vx_voxel> vox;
    std::vector<vx_light> light;
    int highest = 0;
    bool modified = false, lighted = false;
    Chunk(int x, int z) : cx(x), cz(z), vox(CVOL, 0), light(CVOL, 0) {}
};

inline int vidx(int x, int y, int z) { return (y * CD + z) * CW + x; }
inline int floordiv16(int a) { return a >> 4; }  // maths/voxmaths.hpp floordiv<16>

struct LightEntry {
    int x, y, z;
    uint8_t light;
};

struct MeshOut {
    bool cancelled = true;
    std::vector<float> vertices;
    std::vector<int32_t> indices;
    std::vector<vx_sort_entry> entries;
    std::vector<float> entry_floats;
};

}  // namespace

struct vxo_world {
    std::vector<vx_block_def> defs;
    std::vector<float> uv;
    std::vector<vx_model_mesh> meshes;
    std::vector<vx_model_vertex> verts;
    std::vector<RotProfile> profiles;  // indexed by vx_rotation_profile
    std::vector<uint8_t> drawGroups;   // sorted unique (Content::drawGroups)
    vx_settings settings;
    int ox, oz, w, d;
    std::vector<std::unique_ptr<Chunk>> slots;
    std::deque<LightEntry> addq[4], remq[4];
    MeshOut last;

    Chunk* chunk(int cx, int cz) const {  // Chunks::getChunk, AreaMap2D::getIf
        int lx = cx - ox, lz = cz - oz;
        if (lx < 0 || lz < 0 || lx >= w || lz >= d) return nullptr;
        return slots[(size_t)lz * w + lx].get();
    }
    Chunk* chunkByVoxel(int x, int y, int z) const {  // Chunks.cpp:142-151
        if (y < 0 || y >= CH) return nullptr;
        return chunk(floordiv16(x), floordiv16(z));
    }
    int getLight(int x, int y, int z, int ch) const {  // Chunks.cpp:104-121
        Chunk* c = chunkByVoxel(x, y, z);
        if (!c) return 0;
        return (c->light[vidx(x - c->cx * CW, y, z - c->cz * CD)] >> (ch * 4)) & 0xF;
    }
    const vx_voxel* getVoxel(int x, int y, int z) const {
        Chunk* c = chunkByVoxel(x, y, z);
        if (!c) return nullptr;
        return &c->vox[vidx(x - c->cx * CW, y, z - c->cz * CD)];
    }
};

namespace {

inline int nib(vx_light l, int ch) { return (l >> (ch * 4)) & 0xF; }
inline void setnib(vx_light& l, int ch, int v) {
    l = (vx_light)((l & ~(0xF << (ch * 4))) | (v << (ch * 4)));
}

// ---------------------------------------------------------------------------
// LightSolver (lighting/LightSolver.cpp)
// ---------------------------------------------------------------------------

// LightSolver::add(x,y,z,emission), LightSolver.cpp:18-31
void ls_add(vxo_world& w, int ch, int x, int y, int z, int emission) {
    if (emission <= 1) return;
    Chunk* c = w.chunkByVoxel(x, y, z);
    if (!c) return;
    vx_light& cell = c->light[vidx(x - c->cx * CW, y, z - c->cz * CD)];
    if (emission < nib(cell, ch)) return;
    w.addq[ch].push_back({x, y, z, (uint8_t)emission});
    c->modified = true;
    setnib(cell, ch, emission);
}

// LightSolver::add(x,y,z), LightSolver.cpp:33-35
void ls_add_current(vxo_world& w, int ch, int x, int y, int z) {
    ls_add(w, ch, x, y, z, w.getLight(x, y, z, ch));
}

// LightSolver::remove, LightSolver.cpp:37-48
void ls_remove(vxo_world& w, int ch, int x, int y, int z) {
    Chunk* c = w.chunkByVoxel(x, y, z);
    if (!c) return;
    vx_light& cell = c->light[vidx(x - c->cx * CW, y, z - c->cz * CD)];
    int light = nib(cell, ch);
    if (light == 0) return;
    w.remq[ch].push_back({x, y, z, (uint8_t)light});
    setnib(cell, ch, 0);
}

// LightSolver::solve, LightSolver.cpp:50-126
void ls_solve(vxo_world& w, int ch) {
    static const int NB[6][3] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0},
                                 {0, -1, 0}, {1, 0, 0}, {-1, 0, 0}};
    auto& remq = w.remq[ch];
    auto& addq = w.addq[ch];
    while (!remq.empty()) {  // removal phase, :60-95
        LightEntry e = remq.front();
        remq.pop_front();
        for (auto& n : NB) {
            int x = e.x + n[0], y = e.y + n[1], z = e.z + n[2];
            Chunk* c = w.chunkByVoxel(x, y, z);
            if (!c) continue;
            int i = vidx(x - c->cx * CW, y, z - c->cz * CD);
            c->modified = true;
            int light = nib(c->light[i], ch);
            if (light != 0 && light == e.light - 1) {
                uint16_t id = VX_VOXEL_ID(c->vox[i]);
                int emission = id != 0 ? w.defs[id].emission[ch] : 0;
                if (emission) {
                    addq.push_back({x, y, z, (uint8_t)emission});
                    setnib(c->light[i], ch, emission);
                } else {
                    setnib(c->light[i], ch, 0);
                }
                remq.push_back({x, y, z, (uint8_t)light});
            } else if (light >= e.light) {
                addq.push_back({x, y, z, (uint8_t)light});
            }
        }
    }
    while (!addq.empty()) {  // add phase, :97-125
        LightEntry e = addq.front();
        addq.pop_front();
        for (auto& n : NB) {
            int x = e.x + n[0], y = e.y + n[1], z = e.z + n[2];
            Chunk* c = w.chunkByVoxel(x, y, z);
            if (!c) continue;
            int i = vidx(x - c->cx * CW, y, z - c->cz * CD);
            c->modified = true;
            int light = nib(c->light[i], ch);
            const vx_block_def& def = w.defs[VX_VOXEL_ID(c->vox[i])];
            if (def.light_passing && light + 2 <= e.light) {
                setnib(c->light[i], ch, e.light - 1);
                addq.push_back({x, y, z, (uint8_t)(e.light - 1)});
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Lighting (lighting/Lighting.cpp)
// ---------------------------------------------------------------------------

// Lighting::prebuildSkyLight, Lighting.cpp:41-63
void prebuild_sky(vxo_world& w, Chunk& c) {
    int highest = 0;
    for (int z = 0; z < CD; z++) {
        for (int x = 0; x < CW; x++) {
            for (int y = CH - 1; y >= 0; y--) {
                int i = vidx(x, y, z);
                if (!w.defs[VX_VOXEL_ID(c.vox[i])].sky_light_passing) {
                    highest = std::max(highest, y);
                    break;
                }
                setnib(c.light[i], 3, 15);
            }
        }
    }
    if (highest < CH - 1) highest++;
    c.highest = highest;
}

// Lighting::buildSkyLight, Lighting.cpp:65-94
void build_sky(vxo_world& w, int cx, int cz) {
    Chunk* c = w.chunk(cx, cz);
    if (!c) return;
    for (int z = 0; z < CD; z++) {
        for (int x = 0; x < CW; x++) {
            int gx = x + cx * CW, gz = z + cz * CD;
            for (int y = c->highest; y >= 0; y--) {
                while (y > 0 &&
                       !w.defs[VX_VOXEL_ID(c->vox[vidx(x, y, z)])].light_passing) {
                    y--;
                }
                if (nib(c->light[vidx(x, y, z)], 3) != 15) {
                    ls_add_current(w, 3, gx, y + 1, gz);
                    for (; y >= 0; y--) {
                        ls_add_current(w, 3, gx + 1, y, gz);
                        ls_add_current(w, 3, gx - 1, y, gz);
                        ls_add_current(w, 3, gx, y, gz + 1);
                        ls_add_current(w, 3, gx, y, gz - 1);
                    }
                }
            }
        }
    }
    ls_solve(w, 3);
}

// Lighting::onChunkLoaded, Lighting.cpp:97-161
void on_chunk_loaded(vxo_world& w, int cx, int cz, bool expand) {
    Chunk* c = w.chunk(cx, cz);
    if (!c) return;
    for (int y = 0; y < CH; y++) {
        for (int z = 0; z < CD; z++) {
            for (int x = 0; x < CW; x++) {
                const vx_block_def& def = w.defs[VX_VOXEL_ID(c->vox[vidx(x, y, z)])];
                bool emissive = def.emission[0] | def.emission[1] | def.emission[2] |
                                def.emission[3];  // ContentBuilder.cpp:30
                if (emissive) {
                    int gx = x + cx * CW, gz = z + cz * CD;
                    ls_add(w, 0, gx, y, gz, def.emission[0]);
                    ls_add(w, 1, gx, y, gz, def.emission[1]);
                    ls_add(w, 2, gx, y, gz, def.emission[2]);
                }
            }
        }
    }
    if (expand) {
        auto reseed = [&](int x, int y, int z) {
            int rgbs = c->light[vidx(x, y, z)];
            if (rgbs) {
                int gx = x + cx * CW, gz = z + cz * CD;
                for (int ch = 0; ch < 4; ch++) ls_add(w, ch, gx, y, gz, nib(rgbs, ch));
            }
        };
        for (int x = 0; x < CW; x += CW - 1)
            for (int y = 0; y < CH; y++)
                for (int z = 0; z < CD; z++) reseed(x, y, z);
        for (int z = 0; z < CD; z += CD - 1)
            for (int y = 0; y < CH; y++)
                for (int x = 0; x < CW; x++) reseed(x, y, z);
    }
    ls_solve(w, 0);
    ls_solve(w, 1);
    ls_solve(w, 2);
    ls_solve(w, 3);
}

// Chunk::updateHeights, voxels/Chunk.cpp:21-34
void update_heights(Chunk& c) {
    for (int i = 0; i < CVOL; i++) {
        if (VX_VOXEL_ID(c.vox[i]) != 0) {
            c.bottom = i / (CD * CW);
            break;
        }
    }
    for (int i = CVOL - 1; i >= 0; i--) {
        if (VX_VOXEL_ID(c.vox[i]) != 0) {
            c.top = i / (CD * CW) + 1;
            break;
        }
    }
}

// blocks_agent::set (voxels/blocks_agent.cpp:10-77), minus inventories,
// metadata and extended-block segments which the light/mesh path never reads.
void set_block(vxo_world& w, int x, int y, int z, uint16_t id, uint16_t state) {
    if (y < 0 || y >= CH) return;
    int cx = floordiv16(x), cz = floordiv16(z);
    Chunk* c = w.chunk(cx, cz);
    if (!c) return;
    int lx = x - cx * CW, lz = z - cz * CD;
    c->vox[vidx(lx, y, lz)] = VX_VOXEL(id, state);
    c->modified = true;
    if (y < c->bottom)
        c->bottom = y;
    else if (y + 1 > c->top)
        c->top = y + 1;
    else if (id == 0)
        update_heights(*c);
    Chunk* n;
    if (lx == 0 && (n = w.chunk(cx - 1, cz))) n->modified = true;
    if (lz == 0 && (n = w.chunk(cx, cz - 1))) n->modified = true;
    if (lx == CW - 1 && (n = w.chunk(cx + 1, cz))) n->modified = true;
    if (lz == CD - 1 && (n = w.chunk(cx, cz + 1))) n->modified = true;
}

// Lighting::onBlockSet, Lighting.cpp:163-215
void on_block_set(vxo_world& w, int x, int y, int z, uint16_t id) {
    const vx_block_def& block = w.defs.at(id);
    for (int ch = 0; ch < 3; ch++) ls_remove(w, ch, x, y, z);
    if (id == 0) {
        for (int ch = 0; ch < 3; ch++) ls_solve(w, ch);
        if (w.getLight(x, y + 1, z, 3) == 0xF) {
            for (int i = y; i >= 0; i--) {
                const vx_voxel* v = w.getVoxel(x, i, z);
                if ((v == nullptr || VX_VOXEL_ID(*v) != 0) && block.sky_light_passing)
                    break;
                ls_add(w, 3, x, i, z, 0xF);
            }
        }
        static const int NB[6][3] = {{0, 1, 0}, {0, -1, 0}, {1, 0, 0},
                                     {-1, 0, 0}, {0, 0, 1}, {0, 0, -1}};
        for (auto& n : NB)
            for (int ch = 0; ch < 4; ch++)
                ls_add_current(w, ch, x + n[0], y + n[1], z + n[2]);
        for (int ch = 0; ch < 4; ch++) ls_solve(w, ch);
    } else {
        if (!block.sky_light_passing) {
            ls_remove(w, 3, x, y, z);
            for (int i = y - 1; i >= 0; i--) {
                ls_remove(w, 3, x, i, z);
                // reference dereferences chunks.get(x,i-1,z) unconditionally
                // after the i == 0 test; same short-circuit here
                if (i == 0) break;
                const vx_voxel* v = w.getVoxel(x, i - 1, z);
                if (v == nullptr || VX_VOXEL_ID(*v) != 0) break;
            }
            ls_solve(w, 3);
        }
        for (int ch = 0; ch < 3; ch++) ls_solve(w, ch);
        if (block.emission[0] || block.emission[1] || block.emission[2]) {
            for (int ch = 0; ch < 3; ch++) ls_add(w, ch, x, y, z, block.emission[ch]);
            for (int ch = 0; ch < 3; ch++) ls_solve(w, ch);
        }
    }
}

// ---------------------------------------------------------------------------
// BlocksRenderer (graphics/render/BlocksRenderer.cpp)
// ---------------------------------------------------------------------------
struct UV {
    float u1, v1, u2, v2;
};

struct Mesher {
    const vxo_world& w;
    size_t capacity;
    const Chunk* chunk = nullptr;
    std::vector<float> vb;
    std::vector<int32_t> ib;
    size_t indexOffset = 0;
    bool overflow = false;
    // VoxelsVolume 20x256x20 positioned at (cx*16-2, 0, cz*16-2)
    static constexpr int PAD = 2, VW = CW + 2 * PAD, VD = CD + 2 * PAD;
    std::vector<uint16_t> vid;
    std::vector<vx_light> vlight;
    int vx0 = 0, vz0 = 0;
    PseudoRandom randomizer;

    Mesher(const vxo_world& world, size_t cap)
        : w(world), capacity(cap), vid((size_t)VW * CH * VD), vlight((size_t)VW * CH * VD) {}

    // Chunks::getVoxels, voxels/Chunks.cpp:336-418
    void gather(bool backlight) {
        for (int y = 0; y < CH; y++) {
            for (int lz = 0; lz < VD; lz++) {
                for (int lx = 0; lx < VW; lx++) {
                    int gx = vx0 + lx, gz = vz0 + lz;
                    size_t o = ((size_t)y * VD + lz) * VW + lx;
                    const Chunk* c = w.chunk(floordiv16(gx), floordiv16(gz));
                    if (!c) {
                        vid[o] = VX_BLOCK_VOID;
                        vlight[o] = 0;
                        continue;
                    }
                    int i = vidx(gx - c->cx * CW, y, gz - c->cz * CD);
                    uint16_t id = VX_VOXEL_ID(c->vox[i]);
                    vx_light l = c->light[i];
                    if (backlight && id < w.defs.size() && w.defs[id].light_passing) {
                        l = (vx_light)(std::min(15, nib(l, 0) + 1) |
                                       (std::min(15, nib(l, 1) + 1) << 4) |
                                       (std::min(15, nib(l, 2) + 1) << 8) |
                                       (nib(l, 3) << 12));
                    }
                    vid[o] = id;
                    vlight[o] = l;
                }
            }
        }
    }
    // VoxelsVolume::pickBlockId / pickLight, voxels/VoxelsVolume.hpp:51-65
    // (arguments are chunk-local; the reference adds chunk->x*16 then subtracts
    // the volume origin, which cancels)
    bool inside(int x, int y, int z) const {
        return x >= -PAD && z >= -PAD && x < CW + PAD && z < CD + PAD && y >= 0 && y < CH;
    }
    uint16_t pickId(int x, int y, int z) const {
        if (!inside(x, y, z)) return VX_BLOCK_VOID;
        return vid[((size_t)y * VD + (z + PAD)) * VW + (x + PAD)];
    }
    vx_light pickLightRaw(int x, int y, int z) const {
        if (!inside(x, y, z)) return 0;
        return vlight[((size_t)y * VD + (z + PAD)) * VW + (x + PAD)];
    }

    // BlocksRenderer::isOpen, BlocksRenderer.hpp:121-139
    bool isOpen(I3 p, const vx_block_def& def, uint16_t defId) const {
        uint16_t id = pickId(p.x, p.y, p.z);
        if (id == VX_BLOCK_VOID) return false;
        const vx_block_def& b = w.defs[id];
        if ((b.draw_group != def.draw_group && b.draw_group) || !b.solid) return true;
        if ((def.culling == VX_CULLING_DISABLED ||
             (def.culling == VX_CULLING_OPTIONAL && w.settings.dense_render)) &&
            id == defId)
            return true;
        return !id;
    }
    // isOpenForLight + pickLight, BlocksRenderer.cpp:385-411
    V4 pickLight(int x, int y, int z) const {
        uint16_t id = pickId(x, y, z);
        bool open = id != VX_BLOCK_VOID && (w.defs[id].light_passing || !id);
        if (!open) return splat4(0.0f);
        vx_light l = pickLightRaw(x, y, z);
        return V4 {(float)nib(l, 0) / 15.0f, (float)nib(l, 1) / 15.0f,
                   (float)nib(l, 2) / 15.0f, (float)nib(l, 3) / 15.0f};
    }
    V4 pickLight(I3 c) const { return pickLight(c.x, c.y, c.z); }
    // pickSoftLight, BlocksRenderer.cpp:413-420
    V4 pickSoftLight(I3 c, I3 right, I3 up) const {
        return (pickLight(c) + pickLight(c - right) + pickLight(c - right - up) +
                pickLight(c - up)) *
               0.25f;
    }

    // BlocksRenderer::vertex, BlocksRenderer.cpp:40-61
    void vertex(V3 c, float u, float v, V4 light) {
        vb.push_back(c.x);
        vb.push_back(c.y);
        vb.push_back(c.z);
        vb.push_back(u);
        vb.push_back(v);
        uint32_t packed = ((uint32_t)(light.r * 255) & 0xff) << 24;
        packed |= ((uint32_t)(light.g * 255) & 0xff) << 16;
        packed |= ((uint32_t)(light.b * 255) & 0xff) << 8;
        packed |= ((uint32_t)(light.a * 255) & 0xff);
        float f;
        std::memcpy(&f, &packed, 4);
        vb.push_back(f);
    }
    // BlocksRenderer::index, BlocksRenderer.cpp:63-71
    void index(int a, int b, int c, int d, int e, int f) {
        for (int k : {a, b, c, d, e, f}) ib.push_back((int32_t)(indexOffset + k));
        indexOffset += 4;
    }
    bool full(size_t floats) {
        if (vb.size() + floats > capacity) {
            overflow = true;
            return true;
        }
        return false;
    }

    // face with precalculated lights, BlocksRenderer.cpp:74-97
    void faceLights(V3 coord, float fw, float fh, float fd, V3 ax, V3 ay, V3 az,
                    const UV& r, const V4 (&lights)[4], V4 tint) {
        if (full(VX_VERTEX_SIZE * 4)) return;
        V3 X = ax * fw, Y = ay * fh, Z = az * fd;
        float s = 0.5f;
        vertex(coord + (-X - Y + Z) * s, r.u1, r.v1, lights[0] * tint);
        vertex(coord + (X - Y + Z) * s, r.u2, r.v1, lights[1] * tint);
        vertex(coord + (X + Y + Z) * s, r.u2, r.v2, lights[2] * tint);
        vertex(coord + (-X + Y + Z) * s, r.u1, r.v2, lights[3] * tint);
        index(0, 1, 3, 1, 2, 3);
    }
    // vertexAO, BlocksRenderer.cpp:99-114
    void vertexAO(V3 coord, float u, float v, V4 tint, V3 ax, V3 ay, V3 az) {
        V3 pos = coord + az * 0.5f + (ax + ay) * 0.5f;
        I3 ip {(int)std::round(pos.x), (int)std::round(pos.y), (int)std::round(pos.z)};
        V4 light = pickSoftLight(ip, trunc3(ax), trunc3(ay));
        vertex(coord, u, v, light * tint);
    }
    // faceAO, BlocksRenderer.cpp:116-151
    void faceAO(V3 coord, V3 X, V3 Y, V3 Z, const UV& r, bool lights) {
        if (full(VX_VERTEX_SIZE * 4)) return;
        float s = 0.5f;
        if (lights) {
            float d = dot3(normalize3(Z), SUN);
            d = 0.7f + d * 0.3f;
            V3 ax = normalize3(X), ay = normalize3(Y), az = normalize3(Z);
            V4 tint = splat4(d);
            vertexAO(coord + (-X - Y + Z) * s, r.u1, r.v1, tint, ax, ay, az);
            vertexAO(coord + (X - Y + Z) * s, r.u2, r.v1, tint, ax, ay, az);
            vertexAO(coord + (X + Y + Z) * s, r.u2, r.v2, tint, ax, ay, az);
            vertexAO(coord + (-X + Y + Z) * s, r.u1, r.v2, tint, ax, ay, az);
        } else {
            V4 tint = splat4(1.0f);
            vertex(coord + (-X - Y + Z) * s, r.u1, r.v1, tint);
            vertex(coord + (X - Y + Z) * s, r.u2, r.v1, tint);
            vertex(coord + (X + Y + Z) * s, r.u2, r.v2, tint);
            vertex(coord + (-X + Y + Z) * s, r.u1, r.v2, tint);
        }
        index(0, 1, 2, 0, 2, 3);
    }
    // flat face, BlocksRenderer.cpp:153-178
    void faceFlat(V3 coord, V3 X, V3 Y, V3 Z, const UV& r, V4 tint, bool lights) {
        if (full(VX_VERTEX_SIZE * 4)) return;
        float s = 0.5f;
        if (lights) {
            float d = dot3(normalize3(Z), SUN);
            d = 0.7f + d * 0.3f;
            tint = tint * d;
        }
        vertex(coord + (-X - Y + Z) * s, r.u1, r.v1, tint);
        vertex(coord + (X - Y + Z) * s, r.u2, r.v1, tint);
        vertex(coord + (X + Y + Z) * s, r.u2, r.v2, tint);
        vertex(coord + (-X + Y + Z) * s, r.u1, r.v2, tint);
        index(0, 1, 2, 0, 2, 3);
    }

    const Orient& orient(const vx_block_def& def, int rotation) const {
        return w.profiles[def.rotation_profile].v[rotation];
    }

    // blockCube, BlocksRenderer.cpp:324-383
    void blockCube(I3 c, const UV* tex, const vx_block_def& def, uint16_t id,
                   uint16_t state, bool lights, bool ao) {
        I3 X {1, 0, 0}, Y {0, 1, 0}, Z {0, 0, 1};
        if (def.rotation_profile != VX_ROT_NONE) {
            const Orient& o = orient(def, VX_STATE_ROTATION(state));
            X = o.ax[0];
            Y = o.ax[1];
            Z = o.ax[2];
        }
        V3 fc = v3(c);
        struct F {
            I3 probe, fx, fy, fz;
            int tex;
        } faces[6] = {
            {c + Z, X, Y, Z, 5},   {c - Z, -X, Y, -Z, 4}, {c + Y, X, -Z, Y, 3},
            {c - Y, X, Z, -Y, 2},  {c + X, -Z, Y, X, 1},  {c - X, Z, Y, -X, 0},
        };
        for (const F& f : faces) {
            if (!isOpen(f.probe, def, id)) continue;
            if (ao) {
                faceAO(fc, v3(f.fx), v3(f.fy), v3(f.fz), tex[f.tex], lights);
            } else {
                faceFlat(fc, v3(f.fx), v3(f.fy), v3(f.fz), tex[f.tex], pickLight(f.probe),
                         lights);
            }
        }
    }

    // blockXSprite, BlocksRenderer.cpp:180-217
    void blockXSprite(int x, int y, int z, V3 size, const UV& tex1, const UV& tex2,
                      float spread) {
        I3 px {1, 0, 0}, nx {-1, 0, 0}, up {0, 1, 0};
        V4 l1[4] = {pickSoftLight({x, y + 1, z}, px, up), pickSoftLight({x + 1, y + 1, z}, px, up),
                    pickSoftLight({x + 1, y + 1, z}, px, up), pickSoftLight({x, y + 1, z}, px, up)};
        V4 l2[4] = {pickSoftLight({x, y + 1, z}, nx, up), pickSoftLight({x - 1, y + 1, z}, nx, up),
                    pickSoftLight({x - 1, y + 1, z}, nx, up), pickSoftLight({x, y + 1, z}, nx, up)};
        randomizer.setSeed((long)((x * 52321) ^ (z * 389) ^ y));
        // rand32() = (rand() << 16) | rand() (maths/util.hpp:46-48); only the
        // low 16 bits survive the `short` conversion.  Operand evaluation order
        // is unspecified in C++; the g++ build of oracle/_ref evaluates the
        // left operand first, so the low half is the SECOND value drawn
        // (pinned by tests/test_oracle_vs_reference.py).
        randomizer.rand();
        int lo = randomizer.rand();
        short rnd = (short)lo;
        float xs = ((float)(signed char)rnd / 512) * spread;
        float zs = ((float)(signed char)(rnd >> 8) / 512) * spread;
        const float fw = size.x / 1.41f;
        V4 tint = splat4(0.8f);
        V3 c = v3(x + xs, (float)y, z + zs);
        V3 Y1 = v3(0, 1, 0), Z0 = v3(0, 0, 0);
        faceLights(c, fw, size.y, 0, v3(-1, 0, 1), Y1, Z0, tex1, l2, tint);
        faceLights(c, fw, size.y, 0, v3(1, 0, 1), Y1, Z0, tex1, l1, tint);
        faceLights(c, fw, size.y, 0, v3(-1, 0, -1), Y1, Z0, tex2, l2, tint);
        faceLights(c, fw, size.y, 0, v3(1, 0, -1), Y1, Z0, tex2, l1, tint);
    }

    // blockAABB, BlocksRenderer.cpp:222-273
    void blockAABB(I3 ic, const UV* tex, const vx_block_def& def, int rotation,
                   bool lights, bool ao) {
        if (!def.has_hitbox) return;
        V3 a = v3(def.hitbox_a[0], def.hitbox_a[1], def.hitbox_a[2]);
        V3 b = v3(def.hitbox_b[0], def.hitbox_b[1], def.hitbox_b[2]);
        V3 size = v3(std::fabs(b.x - a.x), std::fabs(b.y - a.y), std::fabs(b.z - a.z));
        V3 X = v3(1, 0, 0), Y = v3(0, 1, 0), Z = v3(0, 0, 1);
        V3 coord = v3(ic);
        if (def.rotation_profile != VX_ROT_NONE) {
            const Orient& o = orient(def, rotation);
            X = v3(o.ax[0]);
            Y = v3(o.ax[1]);
            Z = v3(o.ax[2]);
            // CoordSystem::transform, voxels/Block.cpp:47-55
            V3 na = X * a.x + Y * a.y + Z * a.z;
            V3 nb = X * b.x + Y * b.y + Z * b.z;
            a = na + v3(o.fix);
            b = nb + v3(o.fix);
        }
        V3 center = (a + b) * 0.5f;
        coord = coord - (v3(0.5f, 0.5f, 0.5f) - center);
        V3 sx = X * size.x, sy = Y * size.y, sz = Z * size.z;
        V3 nsx = (-X) * size.x, nsy = (-Y) * size.y, nsz = (-Z) * size.z;
        struct F {
            V3 fx, fy, fz;
            int tex;
        } faces[6] = {
            {sx, sy, sz, 5},   {nsx, sy, nsz, 4}, {sx, nsz, sy, 3},
            {nsx, nsz, nsy, 2}, {nsz, sy, sx, 1},  {sz, sy, nsx, 0},
        };
        if (ao) {
            for (const F& f : faces) faceAO(coord, f.fx, f.fy, f.fz, tex[f.tex], lights);
        } else {
            V4 tint = pickLight(ic);
            for (const F& f : faces) faceFlat(coord, f.fx, f.fy, f.fz, tex[f.tex], tint, lights);
        }
    }

    // blockCustomModel, BlocksRenderer.cpp:275-321
    void blockCustomModel(I3 ic, const vx_block_def& def, int rotation) {
        V3 X = v3(1, 0, 0), Y = v3(0, 1, 0), Z = v3(0, 0, 1);
        V3 coord = v3(ic);
        if (def.rotation_profile != VX_ROT_NONE) {
            const Orient& o = orient(def, rotation);
            X = v3(o.ax[0]);
            Y = v3(o.ax[1]);
            Z = v3(o.ax[2]);
        }
        for (int m = 0; m < def.model_mesh_count; m++) {
            const vx_model_mesh& mesh = w.meshes[def.model_mesh_begin + m];
            const vx_model_vertex* mv = w.verts.data() + mesh.vertex_begin;
            if (full((size_t)VX_VERTEX_SIZE * mesh.vertex_count)) return;
            for (int tri = 0; tri < mesh.vertex_count / 3; tri++) {
                const vx_model_vertex& p = mv[tri * 3 + (tri % 2) * 2];
                const vx_model_vertex& q = mv[tri * 3 + 1];
                V3 r = v3(p.coord[0], p.coord[1], p.coord[2]) -
                       v3(q.coord[0], q.coord[1], q.coord[2]);
                r = normalize3(r);
                for (int i = 0; i < 3; i++) {
                    const vx_model_vertex& vert = mv[tri * 3 + i];
                    V3 n = vert.normal[0] * X + vert.normal[1] * Y + vert.normal[2] * Z;
                    float d = dot3(n, SUN);
                    d = 0.8f + d * 0.2f;
                    V3 vc = v3(vert.coord[0] - 0.5f, vert.coord[1] - 0.5f, vert.coord[2] - 0.5f);
                    vertexAO(coord + vc.x * X + vc.y * Y + vc.z * Z, vert.uv[0], vert.uv[1],
                             splat4(d), cross3(r, n), r, n);
                    ib.push_back((int32_t)(indexOffset++));
                }
            }
        }
    }

    void drawVoxel(int i, const vx_block_def& def, uint16_t id, uint16_t state) {
        UV tex[6];
        std::memcpy(tex, w.uv.data() + (size_t)id * 24, sizeof(tex));
        int x = i % CW, y = i / (CD * CW), z = (i / CD) % CW;
        switch (def.model) {
            case VX_MODEL_BLOCK:
                blockCube({x, y, z}, tex, def, id, state, !def.shadeless,
                          def.ambient_occlusion);
                break;
            case VX_MODEL_XSPRITE:
                blockXSprite(x, y, z, v3(1, 1, 1), tex[0 /*FACE_MX*/], tex[4 /*FACE_MZ*/], 1.0f);
                break;
            case VX_MODEL_AABB:
                blockAABB({x, y, z}, tex, def, VX_STATE_ROTATION(state), !def.shadeless,
                          def.ambient_occlusion);
                break;
            case VX_MODEL_CUSTOM:
                blockCustomModel({x, y, z}, def, VX_STATE_ROTATION(state));
                break;
            default:
                break;
        }
    }

    void resetCursors() {
        overflow = false;
        vb.clear();
        ib.clear();
        indexOffset = 0;
    }

    // build, BlocksRenderer.cpp:610-652 (+ render :435-491, renderTranslucent
    // :493-608, createMesh :654-664)
    void build(const Chunk* c, MeshOut& out) {
        chunk = c;
        vx0 = c->cx * CW - PAD;
        vz0 = c->cz * CD - PAD;
        gather(w.settings.backlight != 0);
        out = MeshOut {};
        if (pickId(0, 0, 0) == VX_BLOCK_VOID) {
            out.cancelled = true;
            return;
        }
        out.cancelled = false;
        int begin = c->bottom * (CW * CD), end = c->top * (CW * CD);
        // per draw group: [first, last] voxel index of that group in range
        int first[256], last[256];
        std::fill(first, first + 256, -1);
        for (int i = begin; i < end; i++) {
            uint8_t g = w.defs[VX_VOXEL_ID(c->vox[i])].draw_group;
            if (first[g] < 0) first[g] = i;
            last[g] = i;
        }
        // ---- translucent pass
        resetCursors();
        float amin[3] = {0, 0, 0}, amax[3] = {0, 0, 0};
        bool aabbInit = false;
        for (uint8_t g : w.drawGroups) {
            if (first[g] < 0) continue;
            for (int i = first[g]; i <= last[g]; i++) {
                uint16_t id = VX_VOXEL_ID(c->vox[i]), state = VX_VOXEL_STATE(c->vox[i]);
                const vx_block_def& def = w.defs[id];
                if (id == 0 || def.draw_group != g || VX_STATE_SEGMENT(state)) continue;
                if (!def.translucent) continue;
                drawVoxel(i, def, id, state);
                if (vb.empty()) continue;
                int x = i % CW, y = i / (CD * CW), z = (i / CD) % CW;
                vx_sort_entry e;
                e.position[0] = x + c->cx * CW + 0.5f;
                e.position[1] = y + 0.5f;
                e.position[2] = z + c->cz * CD + 0.5f;
                e.first_float = (uint32_t)out.entry_floats.size();
                e.n_floats = (uint32_t)ib.size() * VX_VERTEX_SIZE;
                for (int32_t idx : ib) {
                    const float* src = vb.data() + (size_t)idx * VX_VERTEX_SIZE;
                    float p[3] = {src[0], src[1], src[2]};
                    if (!aabbInit) {
                        aabbInit = true;
                        for (int k = 0; k < 3; k++) amin[k] = amax[k] = p[k];
                    } else {
                        for (int k = 0; k < 3; k++) {
                            amin[k] = std::min(amin[k], p[k]);  // AABB::addPoint
                            amax[k] = std::max(amax[k], p[k]);
                        }
                    }
                    out.entry_floats.push_back(p[0] + (c->cx * CW + 0.5f));
                    out.entry_floats.push_back(p[1] + 0.5f);
                    out.entry_floats.push_back(p[2] + (c->cz * CD + 0.5f));
                    out.entry_floats.push_back(src[3]);
                    out.entry_floats.push_back(src[4]);
                    out.entry_floats.push_back(src[5]);
                }
                out.entries.push_back(e);
                vb.clear();
                ib.clear();
                indexOffset = 0;
            }
        }
        // "additional powerful optimization", BlocksRenderer.cpp:588-606; a
        // default AABB is a=(0,0,0) b=(1,1,1) when nothing was added
        float sx = aabbInit ? std::fabs(amax[0] - amin[0]) : 1.0f;
        float sy = aabbInit ? std::fabs(amax[1] - amin[1]) : 1.0f;
        float sz = aabbInit ? std::fabs(amax[2] - amin[2]) : 1.0f;
        if ((sy < 0.01f || sx < 0.01f || sz < 0.01f) && out.entries.size() > 1) {
            vx_sort_entry merged = out.entries[0];
            merged.first_float = 0;
            merged.n_floats = (uint32_t)out.entry_floats.size();
            out.entries.assign(1, merged);
        }
        // ---- opaque pass
        resetCursors();
        bool stop = false;
        for (uint8_t g : w.drawGroups) {
            if (first[g] < 0 || stop) continue;
            for (int i = first[g]; i <= last[g]; i++) {
                uint16_t id = VX_VOXEL_ID(c->vox[i]), state = VX_VOXEL_STATE(c->vox[i]);
                const vx_block_def& def = w.defs[id];
                if (id == 0 || def.draw_group != g || VX_STATE_SEGMENT(state)) continue;
                if (def.translucent) continue;
                drawVoxel(i, def, id, state);
                if (overflow) {
                    stop = true;
                    break;
                }
            }
        }
        out.vertices = vb;
        out.indices = ib;
    }
};

void fill_info(const MeshOut& m, int cx, int cz, vx_mesh_info* info) {
    std::memset(info, 0, sizeof(*info));
    info->cx = cx;
    info->cz = cz;
    info->cancelled = m.cancelled;
    if (m.cancelled) return;
    info->n_vertex_floats = (uint32_t)m.vertices.size();
    info->n_indices = (uint32_t)m.indices.size();
    info->n_entries = (uint32_t)m.entries.size();
    info->n_entry_floats = (uint32_t)m.entry_floats.size();
}

}  // namespace

extern "C" {

const char* vxo_kind(void) {
    return "port";
}

vxo_world* vxo_create(const vx_content* content, const vx_settings* settings,
                      int32_t ox, int32_t oz, int32_t w, int32_t d) {
    auto world = new vxo_world();
    world->defs.assign(content->defs, content->defs + content->n_defs);
    world->uv.assign(content->uv_regions, content->uv_regions + (size_t)content->n_defs * 24);
    world->meshes.assign(content->meshes, content->meshes + content->n_meshes);
    world->verts.assign(content->vertices, content->vertices + content->n_vertices);
    for (int p = 0; p < 3; p++) world->profiles.push_back(make_profile(p));
    std::set<uint8_t> groups;  // Content::drawGroups is a std::set<ubyte>
    for (auto& def : world->defs) groups.insert(def.draw_group);
    world->drawGroups.assign(groups.begin(), groups.end());
    world->settings = *settings;
    world->ox = ox;
    world->oz = oz;
    world->w = w;
    world->d = d;
    world->slots.resize((size_t)w * d);
    return world;
}

void vxo_destroy(vxo_world* world) {
    delete world;
}

int vxo_put_chunk(vxo_world* world, int32_t cx, int32_t cz, const vx_voxel* voxels,
                  const vx_light* light) {
    int lx = cx - world->ox, lz = cz - world->oz;
    if (lx < 0 || lz < 0 || lx >= world->w || lz >= world->d) return -1;
    auto c = std::make_unique<Chunk>(cx, cz);
    std::memcpy(c->vox.data(), voxels, sizeof(vx_voxel) * CVOL);
    if (light) std::memcpy(c->light.data(), light, sizeof(vx_light) * CVOL);
    update_heights(*c);
    world->slots[(size_t)lz * world->w + lx] = std::move(c);
    return 0;
}

// Chunk::decode, voxels/Chunk.cpp:88-97; Lightmap::decode, lighting/Lightmap.cpp:26-34
int vxo_put_chunk_encoded(vxo_world* world, int32_t cx, int32_t cz, const uint8_t* voxels,
                          const uint8_t* lights) {
    std::vector<vx_voxel> vox(CVOL);
    for (int i = 0; i < CVOL; i++) {
        uint16_t id = (uint16_t)(voxels[2 * i] | (voxels[2 * i + 1] << 8));
        uint16_t st = (uint16_t)(voxels[2 * (CVOL + i)] | (voxels[2 * (CVOL + i) + 1] << 8));
        vox[i] = VX_VOXEL(id, st);
    }
    if (!lights) return vxo_put_chunk(world, cx, cz, vox.data(), nullptr);
    std::vector<vx_light> light(CVOL);
    for (int i = 0; i < CVOL; i += 2) {
        uint8_t b = lights[i / 2];
        light[i] = (vx_light)((b & 0xF) << 12);
        light[i + 1] = (vx_light)((b & 0xF0) << 8);
    }
    return vxo_put_chunk(world, cx, cz, vox.data(), light.data());
}

// Chunk::encode, voxels/Chunk.cpp:78-86; Lightmap::encode, lighting/Lightmap.cpp:18-24
int vxo_encode_chunk(vxo_world* world, int32_t cx, int32_t cz, uint8_t* voxels_out,
                     uint8_t* lights_out) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    if (voxels_out) {
        for (int i = 0; i < CVOL; i++) {
            uint16_t id = VX_VOXEL_ID(c->vox[i]), st = VX_VOXEL_STATE(c->vox[i]);
            voxels_out[2 * i] = (uint8_t)id;
            voxels_out[2 * i + 1] = (uint8_t)(id >> 8);
            voxels_out[2 * (CVOL + i)] = (uint8_t)st;
            voxels_out[2 * (CVOL + i) + 1] = (uint8_t)(st >> 8);
        }
    }
    if (lights_out) {
        for (int i = 0; i < CVOL; i += 2)
            lights_out[i / 2] = (uint8_t)(((c->light[i] >> 12) & 0xF) | ((c->light[i + 1] >> 8) & 0xF0));
    }
    return 0;
}

// extrle::encode16 (coders/rle.cpp:160-205) and extrle::encode (:108-131): a run is
// closed when the value changes or its counter reaches max_sequence(16); header =
// counter in 6 (7) bits + extension byte when it does not fit, bit 6 = 16-bit value.
uint64_t vxo_extrle_encode(const uint8_t* src, uint64_t n, uint8_t* dst, int32_t wide) {
    const uint64_t count = wide ? n / 2 : n;
    if (count == 0) return 0;
    auto value = [&](uint64_t i) -> uint32_t {
        return wide ? (uint32_t)(src[2 * i] | (src[2 * i + 1] << 8)) : src[i];
    };
    const uint32_t max_seq = wide ? 0x3FFFu : 0x7FFFu;
    uint64_t o = 0;
    auto flush = [&](uint32_t counter, uint32_t c) {
        if (wide) {
            const uint32_t big = c > 255 ? 0x40u : 0u;
            if (counter >= 0x40) {
                dst[o++] = (uint8_t)(0x80u | big | (counter & 0x3Fu));
                dst[o++] = (uint8_t)(counter >> 6);
            } else {
                dst[o++] = (uint8_t)(counter | big);
            }
            dst[o++] = (uint8_t)c;
            if (big) dst[o++] = (uint8_t)(c >> 8);
        } else {
            if (counter >= 0x80) {
                dst[o++] = (uint8_t)(0x80u | (counter & 0x7Fu));
                dst[o++] = (uint8_t)(counter >> 7);
            } else {
                dst[o++] = (uint8_t)counter;
            }
            dst[o++] = (uint8_t)c;
        }
    };
    uint32_t counter = 0, c = value(0);
    for (uint64_t i = 1; i < count; i++) {
        const uint32_t next = value(i);
        if (next != c || counter == max_seq) {
            flush(counter, c);
            c = next;
            counter = 0;
        } else {
            counter++;
        }
    }
    flush(counter, c);
    return o;
}

// extrle::decode16 (coders/rle.cpp:133-158) and extrle::decode (:93-106)
uint64_t vxo_extrle_decode(const uint8_t* src, uint64_t n, uint8_t* dst, int32_t wide) {
    uint64_t o = 0;
    for (uint64_t i = 0; i < n;) {
        uint32_t len = src[i++];
        bool big = false;
        if (wide) {
            big = (len & 0x40u) != 0;
            if (len & 0x80u) len = (len & 0x3Fu) | ((uint32_t)src[i++] << 6);
            else len &= 0x3Fu;
        } else if (len & 0x80u) {
            len = (len & 0x7Fu) | ((uint32_t)src[i++] << 7);
        }
        uint32_t c = src[i++];
        if (big) c |= (uint32_t)src[i++] << 8;
        for (uint32_t j = 0; j <= len; j++) {
            if (wide) {
                dst[2 * o] = (uint8_t)c;
                dst[2 * o + 1] = (uint8_t)(c >> 8);
            } else {
                dst[o] = (uint8_t)c;
            }
            o++;
        }
    }
    return wide ? o * 2 : o;
}

// ---- terrain fill: WorldGenerator::generate without placements ---------------
namespace {

// generate_pole, world/generator/WorldGenerator.cpp:87-116 (unsigned arithmetic kept)
void generate_pole(const vx_gen_layer* layers, int count, uint32_t last_layers_height, int top, int bottom,
                   uint32_t sea_level, vx_voxel* voxels, int x, int z) {
    uint32_t y = (uint32_t)top;
    uint32_t extension = 0;
    for (int k = 0; k < count; k++) {
        const vx_gen_layer& layer = layers[k];
        if (y < sea_level && !layer.below_sea_level) {
            extension = (uint32_t)std::max(0, (int)layer.height);
            continue;
        }
        int h = layer.height;
        if (h == -1) h = (int)(y - last_layers_height - (uint32_t)bottom + 1u);
        else h += (int)extension;
        h = (int)std::min((uint32_t)h, y + 1u);
        for (uint32_t i = 0; i < (uint32_t)h; i++, y--) {
            if (y < (uint32_t)CH)  // the reference writes out of bounds for rows >= CHUNK_H
                voxels[(y * CD + z) * CW + x] = VX_VOXEL(layer.block, VX_VOXEL_STATE(voxels[(y * CD + z) * CW + x]));
        }
        extension = 0;
    }
}

// util::PseudoRandom seeded with two numbers, maths/util.hpp:72-77, and randFloat, :61-63
struct PlantsRandom {
    PseudoRandom r;
    void seed(int a, int b) {
        r.seed = (uint16_t)(((uint16_t)(a * 23729) | (uint16_t)(b % 16786)) ^ (uint16_t)(b * a));
        r.rand();
    }
    float rand_float() {
        uint32_t hi = (uint32_t)r.rand();  // g++ evaluates (rand() << 16) | rand() left to right
        uint32_t lo = (uint32_t)r.rand();
        return (float)((hi << 16) | lo) / 4294967296.0f;
    }
};

}  // namespace

int vxo_terrain_fill(vxo_world* world, const vx_gen_def* def, const vx_gen_chunk* chunks, int32_t n) {
    std::vector<float> sums(def->n_biomes, 0.0f);  // BiomeElementList ctor, GeneratorDef.hpp:82-87
    for (int b = 0; b < def->n_biomes; b++)
        for (int k = 0; k < def->biomes[b].plants_count; k++)
            sums[b] += def->plants[def->biomes[b].plants_begin + k].weight;
    for (int c = 0; c < n; c++) {
        const vx_gen_chunk& g = chunks[c];
        std::vector<vx_voxel> voxels(CVOL, 0);  // :427
        // generateLand, :403-417 (same loop as in generate, :429-441)
        for (int z = 0; z < CD; z++)
            for (int x = 0; x < CW; x++) {
                const vx_gen_biome& B = def->biomes[std::min((int)g.biomes[z * CW + x], def->n_biomes - 1)];
                int height = (int)(g.heights[z * CW + x] * (float)CH);
                height = std::max(0, height);
                generate_pole(def->layers + B.sea_begin, B.sea_count, B.sea_last_layers_height, (int)def->sea_level,
                              height, def->sea_level, voxels.data(), x, z);
                generate_pole(def->layers + B.ground_begin, B.ground_count, B.ground_last_layers_height, height, 0,
                              def->sea_level, voxels.data(), x, z);
            }
        // generatePlants, :362-401
        PlantsRandom rnd;
        rnd.seed(g.cx, g.cz);
        for (int z = 0; z < CD; z++)
            for (int x = 0; x < CW; x++) {
                const int b = std::min((int)g.biomes[z * CW + x], def->n_biomes - 1);
                const vx_gen_biome& B = def->biomes[b];
                int height = (int)(g.heights[z * CW + x] * (float)CH);
                height = std::min(std::max(0, height), CH - 1);
                if (!(height + 1 > (int)def->sea_level && height + 1 < CH)) continue;
                float r = rnd.rand_float();
                uint32_t plant = 0;  // BiomeElementList::choose, GeneratorDef.hpp:92-105
                if (!(B.plants_count == 0 || r > B.plants_chance || B.plants_chance < 1e-6f)) {
                    r = r / B.plants_chance;
                    r *= sums[b];
                    plant = def->plants[B.plants_begin + B.plants_count - 1].block;
                    for (int k = 0; k < B.plants_count; k++) {
                        r -= def->plants[B.plants_begin + k].weight;
                        if (r <= 0.0f) {
                            plant = def->plants[B.plants_begin + k].block;
                            break;
                        }
                    }
                }
                if (!plant) continue;
                vx_voxel& v = voxels[((height + 1) * CD + z) * CW + x];
                if (VX_VOXEL_ID(v)) continue;
                const uint32_t ground = VX_VOXEL_ID(voxels[(height * CD + z) * CW + x]);
                if (ground >= world->defs.size() || !world->defs[ground].solid) continue;
                uint32_t state = 0;
                if (plant < world->defs.size() && world->defs[plant].rotation_profile != VX_ROT_NONE)
                    state = (uint32_t)rnd.r.rand() % (world->defs[plant].rotation_profile == VX_ROT_PIPE ? 6u : 4u);
                v = VX_VOXEL(plant, state);
            }
        for (auto& v : voxels)  // :452-456
            if (VX_VOXEL_ID(v) == 2) v = VX_VOXEL(0, VX_VOXEL_STATE(v));
        if (vxo_put_chunk(world, g.cx, g.cz, voxels.data(), nullptr)) return -1;
    }
    return 0;
}

int vxo_prebuild_sky(vxo_world* world, int32_t cx, int32_t cz) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    prebuild_sky(*world, *c);
    return 0;
}

int vxo_light_chunk(vxo_world* world, int32_t cx, int32_t cz, int32_t expand) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    // logic/ChunksController.cpp:116-123: no buildSkyLight for cached lights
    if (!(expand & VX_LIGHT_NO_SKY)) build_sky(*world, cx, cz);
    on_chunk_loaded(*world, cx, cz, (expand & VX_LIGHT_EXPAND) != 0);
    c->lighted = true;
    return 0;
}

int vxo_light_clear(vxo_world* world) {  // Lighting::clear, Lighting.cpp:28-39
    for (auto& c : world->slots)
        if (c) std::fill(c->light.begin(), c->light.end(), 0);
    return 0;
}

int vxo_set_block(vxo_world* world, int32_t x, int32_t y, int32_t z, uint16_t id,
                  uint16_t state) {
    set_block(*world, x, y, z, id, state);
    on_block_set(*world, x, y, z, id);
    return 0;
}

int vxo_get_light(vxo_world* world, int32_t cx, int32_t cz, vx_light* out) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    std::memcpy(out, c->light.data(), sizeof(vx_light) * CVOL);
    return 0;
}

int vxo_get_voxels(vxo_world* world, int32_t cx, int32_t cz, vx_voxel* out) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    std::memcpy(out, c->vox.data(), sizeof(vx_voxel) * CVOL);
    return 0;
}

int vxo_get_info(vxo_world* world, int32_t cx, int32_t cz, vx_chunk_info* out) {
    std::memset(out, 0, sizeof(*out));
    Chunk* c = world->chunk(cx, cz);
    if (!c) return 0;
    out->present = 1;
    out->bottom = c->bottom;
    out->top = c->top;
    out->highest_point = c->highest;
    out->modified = c->modified;
    out->lighted = c->lighted;
    return 0;
}

int vxo_clear_modified(vxo_world* world) {
    for (auto& c : world->slots)
        if (c) c->modified = false;
    return 0;
}

int vxo_mark_modified(vxo_world* world, int32_t cx, int32_t cz) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    c->modified = true;
    return 0;
}

int vxo_mesh_chunk(vxo_world* world, int32_t cx, int32_t cz, uint64_t capacity,
                   vx_mesh_info* info) {
    Chunk* c = world->chunk(cx, cz);
    if (!c) return -1;
    Mesher m(*world, (size_t)capacity);
    m.build(c, world->last);
    fill_info(world->last, cx, cz, info);
    return 0;
}

int vxo_mesh_read(vxo_world* world, float* vertices, int32_t* indices,
                  vx_sort_entry* entries, float* entry_floats) {
    const MeshOut& m = world->last;
    if (m.cancelled) return -1;
    if (vertices) std::memcpy(vertices, m.vertices.data(), m.vertices.size() * 4);
    if (indices) std::memcpy(indices, m.indices.data(), m.indices.size() * 4);
    if (entries) std::memcpy(entries, m.entries.data(), m.entries.size() * sizeof(vx_sort_entry));
    if (entry_floats) std::memcpy(entry_floats, m.entry_floats.data(), m.entry_floats.size() * 4);
    return 0;
}

double vxo_mesh_batch(vxo_world* world, const vx_chunk_key* keys, int32_t n,
                      uint64_t capacity, int32_t threads, uint64_t* faces_out) {
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int> next {0};
    std::atomic<uint64_t> indices {0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&]() {
            Mesher m(*world, (size_t)capacity);
            MeshOut out;
            for (int i; (i = next.fetch_add(1)) < n;) {
                Chunk* c = world->chunk(keys[i].cx, keys[i].cz);
                if (!c) continue;
                m.build(c, out);
                indices += out.indices.size();
            }
        });
    }
    for (auto& t : pool) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (faces_out) *faces_out = indices / 6;
    return std::chrono::duration<double>(t1 - t0).count();
}

double vxo_light_batch(vxo_world* world, const vx_chunk_key* keys, int32_t n,
                       int32_t expand) {
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) vxo_light_chunk(world, keys[i].cx, keys[i].cz, expand);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
