"""Generates tests/golden/terrain.json from the reference's own WorldGenerator
(world/generator/*.cpp compiled into oracle/_ref/libvxref.so, driven through a
native GeneratorScript — see oracle/ref_build/ref_harness.cpp): SHA-256 of the
voxels WorldGenerator::generate produces for the inputs of tests/terrain_cases.py.
Run here, where /root/reference exists:  python tests/golden/make_golden_terrain.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "voxelengine-cpp_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
import pyoracle  # noqa: E402
import terrain_cases  # noqa: E402
from vxgpu import terrain  # noqa: E402


def run(kind, tables, items, window):
    o = pyoracle.OracleWorld(cases.table(), *window, kind)
    o.terrain_fill(tables, items)
    out = {}
    for (cx, cz, _, _) in items:
        v = o.get_voxels(cx, cz)
        info = o.get_info(cx, cz)
        out["%d,%d" % (cx, cz)] = {"sha256": hashlib.sha256(v.tobytes()).hexdigest(),
                                   "nonzero": int((v != 0).sum()), "bottom": info.bottom, "top": info.top}
    o.close()
    return out


def all_cases():
    keys = [(x, z) for z in range(-2, 2) for x in range(-2, 2)]
    return {
        "demo": (terrain.demo_tables(), terrain_cases.demo_inputs(keys), (-2, -2, 4, 4)),
        "odd": (terrain_cases.odd_tables(), terrain_cases.odd_inputs(keys, 31), (-2, -2, 4, 4)),
    }


if __name__ == "__main__":
    assert pyoracle.available("reference"), "build oracle/_ref first (make -C oracle reference)"
    out = {"_generator": "oracle/_ref/libvxref.so: WorldGenerator::generate (world/generator/WorldGenerator.cpp)"}
    for name, (tables, items, window) in all_cases().items():
        out[name] = run("reference", tables, items, window)
        print(name, len(items), "chunks", sum(c["nonzero"] for c in out[name].values()), "voxels")
    with open(os.path.join(HERE, "terrain.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
