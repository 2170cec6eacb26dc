"""Generator tables for vx_terrain_fill (include/vxgpu.h): plain-Python biome
descriptions -> the vx_gen_def / vx_gen_chunk structs, plus the land rules of
the reference's `base:demo` generator restated from
res/content/base/generators/demo.files/biomes.toml and demo.toml (sea-level 64).
Structures (trees, towers) are placements and out of scope for the fill."""
import ctypes as C

import numpy as np

from . import abi
from . import content as ct


class Tables:
    """biomes: list of dicts {ground: [(block, height, below_sea_level)], sea: [...],
    plants: [(block, weight)], plants_chance: float}.  Plants are sorted by weight,
    descending, as the reference's loader does (content/loading/GeneratorLoader.cpp)."""

    def __init__(self, biomes, sea_level):
        layers, plants, recs = [], [], []
        for b in biomes:
            rec = abi.GenBiome()
            for name, key in (("ground", "ground"), ("sea", "sea")):
                lst = b.get(key, [])
                setattr(rec, name + "_begin", len(layers))
                setattr(rec, name + "_count", len(lst))
                # BlocksLayers::lastLayersHeight: heights after the resizeable layer
                last, seen = 0, False
                for (block, height, below) in lst:
                    if seen:
                        last += height
                    if height == -1:
                        seen = True
                    layers.append((block, height, below))
                setattr(rec, name + "_last_layers_height", last)
            pl = sorted(b.get("plants", []), key=lambda e: -e[1])
            rec.plants_begin, rec.plants_count = len(plants), len(pl)
            rec.plants_chance = b.get("plants_chance", 0.0)
            plants.extend(pl)
            recs.append(rec)
        self.biomes = (abi.GenBiome * len(recs))(*recs)
        self.layers = (abi.GenLayer * max(1, len(layers)))()
        for i, (block, height, below) in enumerate(layers):
            self.layers[i].block, self.layers[i].height, self.layers[i].below_sea_level = block, height, int(below)
        self.plants = (abi.GenPlant * max(1, len(plants)))()
        for i, (block, weight) in enumerate(plants):
            self.plants[i].block, self.plants[i].weight = block, weight
        self.gen = abi.GenDef()
        self.gen.biomes = C.cast(self.biomes, C.c_void_p)
        self.gen.n_biomes = len(recs)
        self.gen.layers = C.cast(self.layers, C.c_void_p)
        self.gen.n_layers = len(layers)
        self.gen.plants = C.cast(self.plants, C.c_void_p)
        self.gen.n_plants = len(plants)
        self.gen.sea_level = sea_level

    def chunks(self, items):
        """items: (cx, cz, heights float32[256], biomes uint8[256]) -> (ctypes array, keep-alive list)."""
        items = list(items)
        arr = (abi.GenChunk * max(1, len(items)))()
        keep = []
        for i, (cx, cz, heights, biomes) in enumerate(items):
            h = np.ascontiguousarray(heights, dtype=np.float32).reshape(-1)
            b = np.ascontiguousarray(biomes, dtype=np.uint8).reshape(-1)
            assert h.size == 256 and b.size == 256
            keep += [h, b]
            arr[i].cx, arr[i].cz = cx, cz
            arr[i].heights, arr[i].biomes = h.ctypes.data, b.ctypes.data
        return arr, keep


def demo_tables():
    """forest / desert / plains of base:demo (land + plants fields)."""
    water = [(ct.WATER, -1, True)]
    plants = [(ct.GRASS, 1.0), (ct.FLOWER, 0.03)]
    forest = dict(ground=[(ct.GRASS_BLOCK, 1, False), (ct.DIRT, 7, False), (ct.STONE, -1, True), (ct.BAZALT, 1, True)],
                  sea=water, plants=plants, plants_chance=0.4)
    desert = dict(ground=[(ct.SAND, 6, True), (ct.STONE, -1, True), (ct.BAZALT, 1, True)], sea=water)
    plains = dict(ground=[(ct.GRASS_BLOCK, 1, False), (ct.DIRT, 5, False), (ct.STONE, -1, True), (ct.BAZALT, 1, True)],
                  sea=water, plants=plants, plants_chance=0.3)
    return Tables([forest, desert, plains], sea_level=64)
