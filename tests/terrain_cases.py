"""Seeded inputs for the terrain-fill tests and their golden digests."""
import numpy as np

from vxgpu import content as ct
from vxgpu import terrain, worlds


def demo_inputs(keys, seed=2019):
    """Heightmaps in [30, 130] / 256 from the value noise the synthetic worlds use, biomes from a second noise."""
    items = []
    for (cx, cz) in keys:
        h = np.empty(256, np.float32)
        b = np.empty(256, np.uint8)
        for z in range(16):
            for x in range(16):
                gx, gz = cx * 16 + x, cz * 16 + z
                n = worlds.noise2(gx, gz, 20, seed) * 0.75 + worlds.noise2(gx, gz, 6, seed + 1) * 0.25
                h[z * 16 + x] = np.float32((30.0 + 100.0 * n) / 256.0)
                b[z * 16 + x] = int(worlds.noise2(gx, gz, 14, seed + 2) * 2.999)
        items.append((cx, cz, h, b))
    return items


def odd_tables():
    """Layer lists that exercise every branch of generate_pole: extension of skipped layers, a resizeable
    layer in the middle, fixed layers taller than the column, a sea list with fixed layers, struct_air,
    plants with rotation profiles, a zero-chance and an always-chance list."""
    biomes = [
        # 0: everything skipped under the sea extends the next layer; bottom layers after the resizeable one
        dict(ground=[(ct.GRASS_BLOCK, 1, False), (ct.DIRT, 3, False), (ct.SAND, 2, True), (ct.STONE, -1, True),
                     (ct.COAL_ORE, 4, True), (ct.BAZALT, 2, True)],
             sea=[(ct.ICE, 1, True), (ct.WATER, -1, True)],
             plants=[(ct.GRASS, 1.0), (ct.FLOWER, 0.5), (ct.TORCH, 0.25), (ct.PANE, 0.25)], plants_chance=0.9),
        # 1: no resizeable layer: fixed layers only, taller than low columns; struct_air on top
        dict(ground=[(ct.STRUCT_AIR, 2, True), (ct.PLANKS, 30, True), (ct.BRICK, 500, True)],
             sea=[(ct.WATER, 3, True), (ct.GLASS, 2, True)],
             plants=[(ct.LIGHTBULB, 1.0)], plants_chance=1.0),
        # 2: sea list that is skipped entirely, ground with a skipped last layer, chance below MIN_CHANCE
        dict(ground=[(ct.STONE, -1, True), (ct.METAL, 3, False)],
             sea=[], plants=[(ct.GRASS, 1.0)], plants_chance=1e-7),
        # 3: empty ground (nothing but the sea), plants on non-solid ground never placed
        dict(ground=[], sea=[(ct.WATER, -1, True)], plants=[(ct.FLOWER, 1.0)], plants_chance=1.0),
        # 4: eight layers each
        dict(ground=[(ct.GRASS_BLOCK, 1, False), (ct.DIRT, 1, True), (ct.SAND, 1, False), (ct.RUST, 1, True),
                     (ct.STONE, -1, True), (ct.BRICK, 1, True), (ct.METAL, 1, True), (ct.BAZALT, 1, True)],
             sea=[(ct.ICE, 1, True), (ct.GLASS, 1, True), (ct.WATER, -1, True), (ct.SAND, 1, True),
                  (ct.DIRT, 1, True), (ct.STONE, 1, True), (ct.RUST, 1, True), (ct.BAZALT, 1, True)],
             plants=[(ct.PIPE, 2.0), (ct.WOODEN_DOOR, 1.0), (ct.GRASS, 1.0)], plants_chance=0.5),
    ]
    return terrain.Tables(biomes, sea_level=70)


def odd_inputs(keys, seed):
    rng = np.random.default_rng(seed)
    items = []
    for i, (cx, cz) in enumerate(keys):
        # heights over the whole range incl. 0, just below 1, the sea level and its neighbours
        h = rng.random(256).astype(np.float32) * np.float32(0.999)
        h[:8] = np.array([0.0, 1e-6, 69 / 256, 70 / 256, 71 / 256, 254.5 / 256, 255 / 256, 0.9999], np.float32)
        if i % 3 == 1:
            h = (rng.integers(60, 80, 256) / 256.0).astype(np.float32)   # around the sea level
        b = rng.integers(0, 5, 256).astype(np.uint8)
        if i % 4 == 2:
            b[:] = i % 5
        items.append((cx, cz, h, b))
    return items
