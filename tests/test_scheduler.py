"""Scheduling layer: vx_window_update takes, for the whole
window at once, the decisions ChunksController::loadVisible / buildLights
(logic/ChunksController.cpp:63-128) and ChunksRenderer::retrieveChunk /
getOrRender (graphics/render/ChunksRenderer.cpp:128-183) take per frame.  The
policy is restated below on top of the oracle and every chunk flag, lightmap
and mesh is compared bit for bit, through chunk arrival, cached lights
(flags.loadedLights), block edits and a late chunk."""
import numpy as np
import pytest

import cases
import pyoracle
from vxgpu import abi
from vxgpu import content as ct
from vxgpu import worlds

pytestmark = pytest.mark.gpu

CAP = 800000


class RefPolicy:
    """The reference's per-frame policy run to completion on the oracle."""

    def __init__(self, oracle, window, padding):
        self.o = oracle
        self.ox, self.oz, self.w, self.d = window
        self.padding = padding
        self.present = {}   # key -> loaded_lights
        self.lighted = set()
        self.meshed = set()

    def put(self, k, voxels=None, encoded=None):
        if encoded is not None:
            self.o.put_chunk_encoded(k[0], k[1], encoded[0], encoded[1])
            self.present[k] = encoded[1] is not None
        else:
            self.o.put_chunk(k[0], k[1], voxels)
            self.present[k] = False
        if not self.present[k]:
            self.o.prebuild_sky(*k)      # ChunksController::createChunk :146-148
        self.lighted.discard(k)
        self.meshed.discard(k)

    def update(self):
        lit = []
        p = self.padding
        for z in range(self.oz + p, self.oz + self.d - p):          # loadVisible :72-83
            for x in range(self.ox + p, self.ox + self.w - p):
                k = (x, z)
                if k not in self.present or k in self.lighted:
                    continue
                if all((x + dx, z + dz) in self.present for dx in (-1, 0, 1) for dz in (-1, 0, 1)):
                    cached = self.present[k]                         # buildLights :116-123
                    self.o.light_chunk(x, z, abi.LIGHT_NO_SKY if cached else abi.LIGHT_EXPAND)
                    self.lighted.add(k)
                    lit.append(k)
        meshes = []
        for k in sorted(self.present, key=lambda k: (k[1], k[0])):   # retrieveChunk / getOrRender
            if k not in self.lighted:
                continue
            if k not in self.meshed or self.o.get_info(*k).modified:
                meshes.append((k, self.o.mesh_chunk(k[0], k[1], CAP)))
                self.meshed.add(k)
        # ChunksRenderer::render clears flags.modified of what it renders
        self._clear_modified([k for k, _ in meshes])
        return lit, meshes

    def _clear_modified(self, keys):
        # the oracle API clears all flags at once; keep the others
        keep = [k for k in self.present if k not in keys and self.o.get_info(*k).modified]
        self.o.clear_modified()
        for k in keep:
            self.o.lib.vxo_mark_modified(self.o.h, k[0], k[1])


def _same_mesh(a, b):
    return (a.cancelled == b.cancelled and np.array_equal(a.vertices, b.vertices)
            and np.array_equal(a.indices, b.indices) and a.entries.tobytes() == b.entries.tobytes()
            and np.array_equal(a.entry_floats, b.entry_floats))


def _compare(g, pol, keys, lit_g, meshes_g, lit_o, meshes_o, what):
    assert lit_g == len(lit_o), (what, lit_g, sorted(lit_o))
    assert [k for k, _ in meshes_g] == [k for k, _ in meshes_o], what
    for (k, a), (_, b) in zip(meshes_g, meshes_o):
        assert _same_mesh(a, b), (what, "mesh", k)
    for k in keys:
        if k not in pol.present:
            continue
        assert np.array_equal(g.get_light(*k), pol.o.get_light(*k)), (what, "light", k)
        a, b = g.get_info(*k), pol.o.get_info(*k)
        assert (a.lighted, a.modified, a.bottom, a.top) == (b.lighted, b.modified, b.bottom, b.top), (what, k)


def test_window_update_matches_reference_policy():
    from vxgpu import gpu
    t = cases.table()
    window = (-3, -3, 7, 7)
    w = worlds.build_world("terrain", *window, seed=91)
    rng = np.random.default_rng(91)
    for k in list(w.chunks):
        w.chunks[k] = worlds.sprinkle(w.chunks[k], rng, t, 10, 16)
    keys = w.keys()
    # lights cache: a fully lit copy of the world gives the cached sky nibbles
    pre = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    pre.put_world(w)
    for k in keys:
        pre.prebuild_sky(*k)
    for k in w.inner_keys():
        pre.light_chunk(k[0], k[1], True)
    cached = {k: pre.encode_chunk(*k) for k in [(0, 0), (1, 0), (-1, 1), (2, 2)]}
    pre.close()
    late = (1, -1)                       # arrives later
    never = (-3, 2)                      # a hole on the window edge

    o = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    if not hasattr(o.lib, "vxo_mark_modified"):
        pytest.skip("oracle library without vxo_mark_modified")
    pol = RefPolicy(o, window, padding=1)
    g = gpu.GpuWorld(t, *window)
    for k in keys:
        if k in (late, never):
            continue
        if k in cached:
            pol.put(k, encoded=cached[k])
            g.put_chunk_encoded(k[0], k[1], cached[k][0], cached[k][1])
        else:
            pol.put(k, voxels=w.chunks[k])
            g.put_chunk(k[0], k[1], w.chunks[k])
            g.prebuild_sky(*k)
    # frame 1: everything that can be lit is lit, then meshed
    lit_o, meshes_o = pol.update()
    lit_g, meshes_g = g.window_update(1, CAP)
    assert lit_g > 0 and meshes_g
    _compare(g, pol, keys, lit_g, meshes_g, lit_o, meshes_o, "frame 1")
    # frame 2: nothing left to do
    lit_o, meshes_o = pol.update()
    lit_g, meshes_g = g.window_update(1, CAP)
    assert lit_g == 0 and not meshes_g and not meshes_o
    # frame 3: block edits re-mesh exactly the chunks the light path marked
    edits = [(5, 90, 5, ct.LAMP, 0), (-17, 70, 18, ct.TORCH, 4), (16, 64, 0, ct.STONE, 0), (5, 90, 5, 0, 0)]
    for (x, y, z, bid, st) in edits:
        o.set_block(x, y, z, bid, st)
    g.edit_batch(edits)
    lit_o, meshes_o = pol.update()
    lit_g, meshes_g = g.window_update(1, CAP)
    assert meshes_g
    _compare(g, pol, keys, lit_g, meshes_g, lit_o, meshes_o, "frame 3")
    # frame 4: the late chunk arrives: it and the neighbours that waited for it are lit
    pol.put(late, voxels=w.chunks[late])
    g.put_chunk(late[0], late[1], w.chunks[late])
    g.prebuild_sky(*late)
    lit_o, meshes_o = pol.update()
    lit_g, meshes_g = g.window_update(1, CAP)
    assert lit_g > 0
    _compare(g, pol, keys, lit_g, meshes_g, lit_o, meshes_o, "frame 4")
    g.close()
    o.close()
