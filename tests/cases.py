"""Seeded parity cases shared by the golden generator, the oracle tests and the
GPU tests.  A case is a small script over a backend exposing the same methods
as oracle/pyoracle.OracleWorld (put_world, prebuild_sky, light_chunk / light_build,
set_block, get_light, get_info, mesh_chunk): the reference oracle, the port
oracle and vxgpu.gpu.GpuWorld all qualify.

Inputs follow SURVEY §8d: base-pack block table, synthetic UV grid, backlight
and dense-render on, capacity 800 000 unless stated.
"""
import hashlib

import numpy as np

from vxgpu import content as ct
from vxgpu import worlds

_TABLE = None


def table():
    global _TABLE
    if _TABLE is None:
        _TABLE = ct.ContentTable()
    return _TABLE


class Case:
    def __init__(self, name, world, light_keys, mesh_keys, capacity=800000,
                 backlight=True, dense_render=True, edits=(), prebuild=True,
                 light_twice=False, expand=True):
        self.name = name
        self.world = world
        self.light_keys = light_keys
        self.mesh_keys = mesh_keys
        self.capacity = capacity
        self.backlight = backlight
        self.dense_render = dense_render
        self.edits = list(edits)
        self.prebuild = prebuild
        self.light_twice = light_twice
        self.expand = expand

    def settings(self):
        return ct.default_settings(self.backlight, self.dense_render)


def _c1(seed):
    w = worlds.build_world("terrain", -1, -1, 3, 3, seed=seed)
    rng = np.random.default_rng(seed)
    w.chunks[(0, 0)] = worlds.sprinkle(w.chunks[(0, 0)], rng, table())
    return Case("c1_s%d" % seed, w, [(0, 0)], [(0, 0)])


def _multi(seed, n=5):
    w = worlds.build_world("terrain", 0, 0, n, n, seed=seed)
    rng = np.random.default_rng(seed + 100)
    for k in w.inner_keys():
        w.chunks[k] = worlds.sprinkle(w.chunks[k], rng, table(), 12, 24)
    inner = w.inner_keys()
    return Case("multi_s%d" % seed, w, inner, inner)


def _random(seed, ymax, capacity=800000, fill=0.5, name=None):
    w = worlds.build_world("random", -1, -1, 3, 3, seed=seed, fill=fill,
                           table=table(), ymax=ymax)
    return Case(name or "random_s%d_y%d_c%d" % (seed, ymax, capacity), w,
                [(0, 0)], [(0, 0)], capacity=capacity)


def _random_edits(world, rng, n, keys, ylo, yhi, place_pool):
    """Alternating place / break edits in the given chunks (config 4 style)."""
    placed = []
    edits = []
    variants = table().rot_variants()
    for i in range(n):
        if placed and rng.random() < 0.5:
            x, y, z = placed.pop(int(rng.integers(0, len(placed))))
            edits.append((x, y, z, 0, 0))
        else:
            cx, cz = keys[int(rng.integers(0, len(keys)))]
            x = cx * 16 + int(rng.integers(0, 16))
            z = cz * 16 + int(rng.integers(0, 16))
            y = int(rng.integers(ylo, yhi))
            bid = int(place_pool[int(rng.integers(0, len(place_pool)))])
            rot = int(rng.integers(0, variants[bid]))
            edits.append((x, y, z, bid, rot))
            placed.append((x, y, z))
    return edits


def _edits(seed, n=120):
    w = worlds.build_world("terrain", 0, 0, 5, 5, seed=seed)
    inner = w.inner_keys()
    rng = np.random.default_rng(seed + 500)
    pool = [ct.LAMP, ct.TORCH, ct.RED_LAMP, ct.GREEN_LAMP, ct.BLUE_LAMP, ct.LIGHTBULB,
            ct.STONE, ct.GLASS, ct.WATER, ct.LEAVES, ct.PANE]
    hm = worlds.heightmap(np.array([40]), np.array([40]), seed)[0]
    edits = _random_edits(w, rng, n, inner, max(1, int(hm) - 30), min(250, int(hm) + 24),
                          pool)
    return Case("edits_s%d" % seed, w, inner, [(2, 2), (1, 2)], edits=edits)


def named_cases():
    """name -> factory; heavier cases are built lazily."""
    out = {}
    for seed in range(1, 17):
        out["c1_s%d" % seed] = (lambda s=seed: _c1(s))
    for seed in (21, 22):
        out["multi_s%d" % seed] = (lambda s=seed: _multi(s))
    out["random_s7_y40"] = lambda: _random(7, 40, name="random_s7_y40")
    out["random_s8_y64_sparse"] = lambda: _random(8, 64, fill=0.3, name="random_s8_y64_sparse")
    out["random_s7_full_c800k"] = lambda: _random(7, 256, 800000, name="random_s7_full_c800k")
    out["random_s7_full_c8M"] = lambda: _random(7, 256, 8000000, name="random_s7_full_c8M")
    out["random_s9_y32_c20k"] = lambda: _random(9, 32, 20000, name="random_s9_y32_c20k")

    def checker():
        w = worlds.build_world("checker", -1, -1, 3, 3)
        return Case("checker", w, [(0, 0)], [(0, 0)])
    out["checker"] = checker

    def flat():
        w = worlds.build_world("flat", -1, -1, 3, 3)
        return Case("flat", w, [(0, 0)], [(0, 0)])
    out["flat"] = flat

    def hole():
        w = worlds.build_world("terrain", -1, -1, 3, 3, seed=31)
        del w.chunks[(1, 0)]
        del w.chunks[(-1, -1)]
        return Case("hole", w, [(0, 0)], [(0, 0), (0, 1)])
    out["hole"] = hole

    def nobacklight():
        c = _c1(3)
        c.name = "c1_s3_nobacklight"
        c.backlight = False
        return c
    out["c1_s3_nobacklight"] = nobacklight

    def nodense():
        c = _c1(4)
        c.name = "c1_s4_nodense"
        c.dense_render = False
        return c
    out["c1_s4_nodense"] = nodense

    def twice():
        c = _multi(23, 4)
        c.name = "multi_s23_twice"
        c.light_twice = True
        return c
    out["multi_s23_twice"] = twice

    def noexpand():
        c = _multi(24, 4)
        c.name = "multi_s24_noexpand"
        c.expand = False
        return c
    out["multi_s24_noexpand"] = noexpand

    for seed in (41, 42):
        out["edits_s%d" % seed] = (lambda s=seed: _edits(s))
    return out


FAST_CASES = ["c1_s1", "c1_s2", "flat", "hole", "random_s9_y32_c20k", "c1_s3_nobacklight",
              "c1_s4_nodense"]


def run_case(backend_factory, case, batch_light=False):
    """Runs a case on a backend; returns {label: numpy array}.

    backend_factory(table, ox, oz, w, d, settings) -> backend.  With
    batch_light the backend's light_build(keys) is used instead of one
    light_chunk call per key (the GPU path; same result by SURVEY §8a)."""
    w = case.world
    be = backend_factory(table(), w.ox, w.oz, w.w, w.d, case.settings())
    out = {}
    try:
        be.put_world(w)
        if case.prebuild:
            if batch_light:
                be.prebuild_batch(w.keys())
            else:
                for k in w.keys():
                    be.prebuild_sky(*k)
        rounds = 2 if case.light_twice else 1
        for _ in range(rounds):
            if batch_light:
                be.light_build(case.light_keys, case.expand)
            else:
                for k in case.light_keys:
                    be.light_chunk(k[0], k[1], case.expand)
        if case.edits:
            if batch_light:
                be.edit_batch(case.edits)
            else:
                for (x, y, z, bid, st) in case.edits:
                    be.set_block(x, y, z, bid, st)
        for k in w.keys():
            out["light_%d_%d" % k] = be.get_light(*k)
            info = be.get_info(*k)
            out["info_%d_%d" % k] = np.array(
                [info.bottom, info.top, info.highest_point, info.lighted], dtype=np.int32)
            out["modified_%d_%d" % k] = np.array([info.modified], dtype=np.int32)
        meshes = be.mesh_batch_read(case.mesh_keys, case.capacity) if batch_light else \
            [be.mesh_chunk(k[0], k[1], case.capacity) for k in case.mesh_keys]
        for k, m in zip(case.mesh_keys, meshes):
            out["mesh_v_%d_%d" % k] = m.vertices
            out["mesh_i_%d_%d" % k] = m.indices
            out["mesh_e_%d_%d" % k] = np.frombuffer(m.entries.tobytes(), dtype=np.uint32)
            out["mesh_f_%d_%d" % k] = m.entry_floats
    finally:
        be.close()
    return out


def digest(arrays):
    """{label: sha256 hex} plus sizes, for tests/golden/*.json."""
    return {k: {"sha256": hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest(),
                "n": int(v.size)} for k, v in sorted(arrays.items())}
