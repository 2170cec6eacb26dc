"""Terrain fill: WorldGenerator::generate without structure
placements — generate_pole, generateLand, generatePlants, struct_air -> air
(world/generator/WorldGenerator.cpp:87-116,362-461).  The goldens come from the
reference's own WorldGenerator driven through a native GeneratorScript
(tests/golden/make_golden_terrain.py).  CPU part: the oracle port against them
(and against the reference itself where it is built).  GPU part:
vx_terrain_fill against the oracle and the goldens, voxel for voxel, and the
generated world lit and meshed like the oracle's."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

import cases
import make_golden_terrain as mg
import pyoracle
import terrain_cases
from vxgpu import abi, terrain
from vxgpu import content as ct

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "terrain.json")))


@pytest.mark.parametrize("name", ["demo", "odd"])
@pytest.mark.parametrize("kind", ["port", "reference"])
def test_oracle_matches_reference_goldens(name, kind):
    if not pyoracle.available(kind):
        pytest.skip("library not built")
    tables, items, window = mg.all_cases()[name]
    got = mg.run(kind, tables, items, window)
    assert got == GOLDEN[name]


def test_demo_world_looks_like_the_demo_generator():
    """Sanity of the restated base:demo rules: bazalt floor, water up to the sea level, plants on grass."""
    tables, items, window = mg.all_cases()["demo"]
    o = pyoracle.OracleWorld(cases.table(), *window, "port")
    o.terrain_fill(tables, items)
    ids = np.concatenate([o.get_voxels(cx, cz) & 0xFFFF for (cx, cz, _, _) in items]).reshape(len(items), 256, 256)
    assert (ids[:, 0, :] == ct.BAZALT).all()
    present = set(np.unique(ids))
    assert {ct.STONE, ct.DIRT, ct.GRASS_BLOCK, ct.SAND, ct.WATER, ct.GRASS} <= present
    assert not (ids[:, 65:, :] == ct.WATER).any() and (ids[:, 64, :] == ct.WATER).any()
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["demo", "odd"])
def test_gpu_terrain_fill_matches_oracle_and_goldens(name):
    from vxgpu import gpu
    t = cases.table()
    tables, items, window = mg.all_cases()[name]
    g = gpu.GpuWorld(t, *window)
    o = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    g.terrain_fill(tables, items)
    o.terrain_fill(tables, items)
    for (cx, cz, _, _) in items:
        a, b = g.get_voxels(cx, cz), o.get_voxels(cx, cz)
        assert np.array_equal(a, b), (name, cx, cz, int((a != b).sum()), np.flatnonzero(a != b)[:5])
        assert hashlib.sha256(a.tobytes()).hexdigest() == GOLDEN[name]["%d,%d" % (cx, cz)]["sha256"]
        ia, ib = g.get_info(cx, cz), o.get_info(cx, cz)
        assert (ia.bottom, ia.top) == (ib.bottom, ib.top)
        assert not g.get_light(cx, cz).any()
    # the generated world behaves like an uploaded one: prebuild, light, mesh
    keys = [(cx, cz) for (cx, cz, _, _) in items]
    inner = [(x, z) for (x, z) in keys if -2 < x < 1 and -2 < z < 1]
    g.prebuild_batch(keys)
    for k in keys:
        o.prebuild_sky(*k)
    g.light_build(inner, True)
    o.light_batch(inner, True)
    for k in keys:
        assert np.array_equal(g.get_light(*k), o.get_light(*k)), (name, "light", k)
    for k in inner:
        m1, m2 = g.mesh_chunk(*k), o.mesh_chunk(*k)
        assert np.array_equal(m1.vertices, m2.vertices) and np.array_equal(m1.indices, m2.indices), (name, "mesh", k)
        assert np.array_equal(m1.entry_floats, m2.entry_floats)
    g.close()
    o.close()


@pytest.mark.gpu
def test_gpu_terrain_fill_rejects_bad_tables():
    from vxgpu import gpu
    g = gpu.GpuWorld(cases.table(), 0, 0, 2, 2)
    too_many = terrain.Tables([dict(ground=[(ct.STONE, 1, True)] * (abi.GEN_MAX_LAYERS + 1), sea=[])], 10)
    h, b = np.full(256, 0.2, np.float32), np.zeros(256, np.uint8)
    with pytest.raises(RuntimeError, match="layers"):
        g.terrain_fill(too_many, [(0, 0, h, b)])
    assert not g.get_info(0, 0).present
    ok = terrain.Tables([dict(ground=[(ct.STONE, -1, True)], sea=[])], 10)
    g.terrain_fill(ok, [(0, 0, h, b), (9, 9, h, b)][:1])
    assert g.get_info(0, 0).present and g.get_info(0, 0).top == int(np.float32(0.2) * 256) + 1
    g.close()
