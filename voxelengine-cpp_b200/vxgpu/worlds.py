"""Deterministic synthetic worlds for parity tests and benchmarks (numpy).

The reference's terrain generator needs LuaJIT (absent here), so the terrain
below is a native restatement of the *kind* of world `base:demo` produces
(res/content/base/generators/demo.files/biomes.toml: grass_block / dirt /
stone / bazalt layers, water to sea level 64, grass + flower plants) over a
hash value-noise heightmap, plus caves, trees and emitters so that light has
work to do.  All inputs are labelled "synthetic" wherever they are reported.

A chunk is a uint32[65536] array of vx_voxel = id | state << 16, index
(y*16+z)*16+x (src/constants.hpp:55-57).
"""
import numpy as np

from . import abi
from . import content as ct

W, H, D, VOL = abi.CHUNK_W, abi.CHUNK_H, abi.CHUNK_D, abi.CHUNK_VOL
SEA_LEVEL = 64

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xC2B2AE3D27D4EB4F)
_M3 = np.uint64(0x165667B19E3779F9)
_M4 = np.uint64(0xFF51AFD7ED558CCD)
_M5 = np.uint64(0xC4CEB9FE1A85EC53)


def _mix(h):
    with np.errstate(over="ignore"):
        h = h ^ (h >> np.uint64(33))
        h = h * _M4
        h = h ^ (h >> np.uint64(33))
        h = h * _M5
        h = h ^ (h >> np.uint64(33))
    return h


def hash3(x, y, z, seed):
    """uint64 hash of integer lattice points (broadcasting numpy arrays)."""
    with np.errstate(over="ignore"):
        h = (np.asarray(x).astype(np.int64).astype(np.uint64) * _M1
             + np.asarray(y).astype(np.int64).astype(np.uint64) * _M2
             + np.asarray(z).astype(np.int64).astype(np.uint64) * _M3
             + np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x27D4EB2F165667C5))
    return _mix(h)


def rand01(x, y, z, seed):
    return (hash3(x, y, z, seed) >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def _smooth(t):
    return t * t * (3.0 - 2.0 * t)


def noise2(gx, gz, period, seed):
    fx = gx / period
    fz = gz / period
    x0 = np.floor(fx).astype(np.int64)
    z0 = np.floor(fz).astype(np.int64)
    tx = _smooth(fx - x0)
    tz = _smooth(fz - z0)
    a = rand01(x0, 0, z0, seed)
    b = rand01(x0 + 1, 0, z0, seed)
    c = rand01(x0, 0, z0 + 1, seed)
    d = rand01(x0 + 1, 0, z0 + 1, seed)
    return (a * (1 - tx) + b * tx) * (1 - tz) + (c * (1 - tx) + d * tx) * tz


def noise3(gx, gy, gz, period, seed):
    fx, fy, fz = gx / period, gy / period, gz / period
    x0 = np.floor(fx).astype(np.int64)
    y0 = np.floor(fy).astype(np.int64)
    z0 = np.floor(fz).astype(np.int64)
    tx, ty, tz = _smooth(fx - x0), _smooth(fy - y0), _smooth(fz - z0)
    out = 0.0
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = ((tx if dx else 1 - tx) * (ty if dy else 1 - ty)
                     * (tz if dz else 1 - tz))
                out = out + w * rand01(x0 + dx, y0 + dy, z0 + dz, seed)
    return out


def vox(ids, states=None):
    ids = np.asarray(ids, dtype=np.uint32)
    if states is None:
        return ids.copy()
    return ids | (np.asarray(states, dtype=np.uint32) << np.uint32(16))


def empty_chunk():
    return np.zeros(VOL, dtype=np.uint32)


def flat_chunk(cx, cz, seed=0):
    """The literal core:default generator output: one obstacle layer at y=0
    (res/generators/default.toml)."""
    v = np.zeros((H, D, W), dtype=np.uint32)
    v[0] = ct.OBSTACLE
    return v.reshape(-1)


def heightmap(gx, gz, seed):
    n = (0.55 * noise2(gx, gz, 48.0, seed) + 0.30 * noise2(gx, gz, 20.0, seed + 1)
         + 0.15 * noise2(gx, gz, 7.0, seed + 2))
    return (40 + n * 80).astype(np.int64)


def terrain_chunk(cx, cz, seed=2019, caves=True, trees=True, emitters=True):
    """Demo-like terrain, deterministic in (cx,cz,seed)."""
    pad = 3  # trees rooted in neighbour chunks reach 2 voxels in
    xs = np.arange(-pad, W + pad) + cx * W
    zs = np.arange(-pad, D + pad) + cz * D
    gz, gx = np.meshgrid(zs, xs, indexing="ij")  # [z][x]
    hm = heightmap(gx, gz, seed)
    ys = np.arange(H)[:, None, None]
    h3 = hm[None, :, :]
    ids = np.zeros((H, D + 2 * pad, W + 2 * pad), dtype=np.uint32)
    dirt_depth = 5 + (hash3(gx, 1, gz, seed) % np.uint64(3)).astype(np.int64)[None]
    underwater = h3 < SEA_LEVEL
    ids[ys < h3 - dirt_depth] = 0
    solid = ys <= h3
    ids = np.where(solid, ct.STONE, ids)
    ids = np.where(solid & (ys > h3 - dirt_depth), ct.DIRT, ids)
    ids = np.where((ys == h3) & ~underwater, ct.GRASS_BLOCK, ids)
    ids = np.where((ys <= h3) & (ys > h3 - 3) & underwater, ct.SAND, ids)
    ids = np.where((ys > h3) & (ys <= SEA_LEVEL), ct.WATER, ids)
    ids = np.where(ys == 0, ct.BAZALT, ids)
    ids = ids.astype(np.uint32)
    states = np.zeros_like(ids)
    gx3 = np.broadcast_to(gx[None], ids.shape)
    gz3 = np.broadcast_to(gz[None], ids.shape)
    gy3 = np.broadcast_to(ys, ids.shape)
    if caves:
        n = (0.6 * noise3(gx3, gy3 * 1.3, gz3, 14.0, seed + 7)
             + 0.4 * noise3(gx3, gy3 * 1.3, gz3, 6.0, seed + 8))
        carve = (n > 0.62) & (gy3 > 2) & (ids != ct.WATER) & (ids != 0)
        # keep a lid under water so that seas do not drain into caves
        carve &= ~((h3 < SEA_LEVEL + 1) & (gy3 > h3 - 4))
        ids = np.where(carve, 0, ids).astype(np.uint32)
        if emitters:
            r = hash3(gx3, gy3, gz3, seed + 11) % np.uint64(4096)
            floor_below = np.zeros_like(carve)
            floor_below[1:] = (ids[:-1] == ct.STONE)
            torch = carve & floor_below & (r < np.uint64(24))
            ids = np.where(torch, ct.TORCH, ids).astype(np.uint32)
            states = np.where(torch, 4, states)  # pipe profile "up"
            lamp_pick = (r >> np.uint64(3)) % np.uint64(4)
            lamp_ids = np.array([ct.LAMP, ct.RED_LAMP, ct.GREEN_LAMP, ct.BLUE_LAMP],
                                dtype=np.uint32)[lamp_pick.astype(np.int64)]
            lamp = (ids == ct.STONE) & (r >= np.uint64(4090))
            ids = np.where(lamp, lamp_ids, ids).astype(np.uint32)
    # plants on dry grass
    pr = rand01(gx, 2, gz, seed + 3)
    top = np.clip(hm + 1, 0, H - 1)
    zz, xx = np.meshgrid(np.arange(D + 2 * pad), np.arange(W + 2 * pad), indexing="ij")
    dry = (hm >= SEA_LEVEL) & (ids[np.clip(hm, 0, H - 1), zz, xx] == ct.GRASS_BLOCK)
    plant = dry & (pr < 0.35)
    pid = np.where(pr < 0.08, ct.FLOWER, ct.GRASS).astype(np.uint32)
    cur = ids[top, zz, xx]
    ids[top, zz, xx] = np.where(plant & (cur == 0), pid, cur)
    if trees:
        tr = hash3(gx, 3, gz, seed + 5) % np.uint64(97)
        roots = np.argwhere(dry & (tr == np.uint64(0)))
        for (rz, rx) in roots:
            base = int(hm[rz, rx]) + 1
            height = 4 + int(hash3(int(gx[rz, rx]), 4, int(gz[rz, rx]), seed) % np.uint64(3))
            topy = base + height
            if topy + 3 >= H:
                continue
            for dy in range(-2, 3):
                rad = 2 if dy < 1 else 1
                z0, z1 = max(rz - rad, 0), min(rz + rad + 1, D + 2 * pad)
                x0, x1 = max(rx - rad, 0), min(rx + rad + 1, W + 2 * pad)
                blk = ids[topy + dy, z0:z1, x0:x1]
                ids[topy + dy, z0:z1, x0:x1] = np.where(
                    (blk == 0) | (blk == ct.GRASS) | (blk == ct.FLOWER), ct.LEAVES, blk)
            ids[base:topy, rz, rx] = ct.WOOD
            states[base:topy, rz, rx] = 4
    out = vox(ids[:, pad:pad + D, pad:pad + W], states[:, pad:pad + D, pad:pad + W])
    return np.ascontiguousarray(out).reshape(-1)


def checkerboard_chunk(cx, cz, seed=0):
    """3-D checkerboard of stone/air over y in [0,255] (config 3a)."""
    y, z, x = np.meshgrid(np.arange(H), np.arange(D), np.arange(W), indexing="ij")
    v = np.where(((x + y + z) & 1) == 0, ct.STONE, 0).astype(np.uint32)
    return v.reshape(-1)


# (id, weight) — config 3b palette
WORST_PALETTE = [
    (ct.STONE, 70), (ct.GLASS, 8), (ct.WATER, 4), (ct.LEAVES, 6), (ct.LAMP, 1),
    (ct.TORCH, 1), (ct.RED_LAMP, 1), (ct.GREEN_LAMP, 1), (ct.BLUE_LAMP, 1),
    (ct.PANE, 2), (ct.PIPE, 2), (ct.GRASS, 2), (ct.CUSTOM, 1),
]


def random_chunk(cx, cz, seed=7, fill=0.5, palette=None, table=None, ymax=H):
    """iid fill with a weighted palette and random valid rotations (config 3b)."""
    palette = palette or WORST_PALETTE
    table = table or ct.ContentTable()
    rng = np.random.default_rng([seed & 0xFFFFFFFF, cx & 0xFFFFFFFF, cz & 0xFFFFFFFF])
    pid = np.array([p[0] for p in palette], dtype=np.uint32)
    w = np.array([p[1] for p in palette], dtype=np.float64)
    pick = rng.choice(len(pid), size=VOL, p=w / w.sum())
    ids = np.where(rng.random(VOL) < fill, pid[pick], 0).astype(np.uint32)
    if ymax < H:
        ids[ymax * W * D:] = 0
    variants = table.rot_variants()[ids]
    rot = (rng.integers(0, 8, size=VOL) % variants).astype(np.uint32)
    userbits = np.where(ids != 0, rng.integers(0, 256, size=VOL), 0).astype(np.uint32)
    return vox(ids, rot | (userbits << np.uint32(8)))


def sprinkle(voxels, rng, table, n_emitters=32, n_special=64, ymin=1, ymax=140):
    """Config-1 extras: random emitters and aabb / X / translucent / custom
    blocks written over whatever is there."""
    v = voxels.copy()
    emit = [ct.LAMP, ct.RED_LAMP, ct.GREEN_LAMP, ct.BLUE_LAMP, ct.TORCH, ct.LIGHTBULB]
    special = [ct.PANE, ct.PIPE, ct.LIGHTBULB, ct.TORCH, ct.WOODEN_DOOR, ct.GRASS,
               ct.FLOWER, ct.GLASS, ct.WATER, ct.ICE, ct.LEAVES, ct.CUSTOM, ct.WOOD,
               ct.STRUCT_AIR]
    variants = table.rot_variants()
    for pool, n in ((emit, n_emitters), (special, n_special)):
        for _ in range(n):
            x, z = int(rng.integers(0, W)), int(rng.integers(0, D))
            y = int(rng.integers(ymin, ymax))
            bid = int(pool[int(rng.integers(0, len(pool)))])
            rot = int(rng.integers(0, variants[bid]))
            v[(y * D + z) * W + x] = bid | (rot << 16)
    return v


class World:
    """A rectangular window of chunks kept as numpy arrays."""

    def __init__(self, ox, oz, w, d):
        self.ox, self.oz, self.w, self.d = ox, oz, w, d
        self.chunks = {}

    def keys(self):
        return sorted(self.chunks.keys(), key=lambda k: (k[1], k[0]))

    def inner_keys(self):
        """Chunks whose 8 neighbours are all present (the ones the reference
        would light, logic/ChunksController.cpp:106-128)."""
        out = []
        for (cx, cz) in self.keys():
            if all((cx + dx, cz + dz) in self.chunks
                   for dx in (-1, 0, 1) for dz in (-1, 0, 1)):
                out.append((cx, cz))
        return out


def build_world(kind, ox, oz, w, d, seed=2019, **kw):
    gen = {
        "flat": flat_chunk, "terrain": terrain_chunk,
        "checker": checkerboard_chunk, "random": random_chunk,
    }[kind]
    world = World(ox, oz, w, d)
    for cz in range(oz, oz + d):
        for cx in range(ox, ox + w):
            world.chunks[(cx, cz)] = gen(cx, cz, seed, **kw)
    return world


def edit_script(n, seed, table=None, lo=2 * 16, hi=62 * 16, terrain_seed=2019):
    """configs[3]: n place / break edits of emissive blocks, positions uniform in
    [lo, hi)^2 voxels (the inner 60x60 chunks of a 64x64 window), y in [surface - 8, surface + 24);
    half of the edits break a block placed earlier.  Returns [(x, y, z, id, state)]."""
    table = table or ct.ContentTable()
    rng = np.random.default_rng(seed)
    pool = [ct.LAMP, ct.TORCH, ct.RED_LAMP, ct.GREEN_LAMP, ct.BLUE_LAMP, ct.LIGHTBULB]
    variants = table.rot_variants()
    placed, edits = [], []
    for _ in range(n):
        if placed and rng.random() < 0.5:
            x, y, z = placed.pop(int(rng.integers(0, len(placed))))
            edits.append((x, y, z, 0, 0))
        else:
            x = int(rng.integers(lo, hi))
            z = int(rng.integers(lo, hi))
            h = int(heightmap(np.array([x]), np.array([z]), terrain_seed)[0])
            y = int(np.clip(rng.integers(h - 8, h + 24), 1, 254))
            bid = int(pool[int(rng.integers(0, len(pool)))])
            edits.append((x, y, z, bid, int(rng.integers(0, variants[bid]))))
            placed.append((x, y, z))
    return edits
