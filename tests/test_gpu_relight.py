"""Relighting sequences (run with -m gpu).  The seeding kernels take short cuts
when a chunk's lightmap is known to be exactly what Lighting::prebuildSkyLight
(lighting/Lighting.cpp:37-61) leaves behind; these sequences walk every way a
chunk enters and leaves that state — Lighting::clear (:21-35), prebuild of only
part of the window, block edits that move the sky columns, light arriving from
a cache — and compare lightmaps and chunk flags with the oracle after each."""
import numpy as np
import pytest

import cases
import pyoracle
from vxgpu import content as ct
from vxgpu import worlds

pytestmark = pytest.mark.gpu


def _check(g, o, keys, what):
    for k in keys:
        a, b = g.get_light(*k), o.get_light(*k)
        assert np.array_equal(a, b), (what, "light", k, int(np.count_nonzero(a != b)))
        ia, ib = g.get_info(*k), o.get_info(*k)
        assert (ia.lighted, ia.modified, ia.highest_point) == (ib.lighted, ib.modified, ib.highest_point), (what, k)


def _both(g, o, fn):
    fn(g)
    fn(o)


@pytest.mark.parametrize("kind,seed", [("terrain", 5), ("flat", 6), ("random", 7)])
def test_clear_prebuild_build_sequences(kind, seed):
    from vxgpu import gpu
    t = cases.table()
    window = (-3, -2, 6, 6)
    kw = {"ymax": 96, "fill": 0.35, "table": t} if kind == "random" else {}
    w = worlds.build_world(kind, *window, seed=seed, **kw)
    rng = np.random.default_rng(seed)
    if kind != "random":
        for k in list(w.chunks):
            w.chunks[k] = worlds.sprinkle(w.chunks[k], rng, t, 6, 8)
    keys, inner = w.keys(), w.inner_keys()
    g = gpu.GpuWorld(t, *window)
    o = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    g.put_chunks([(k[0], k[1], w.chunks[k], None) for k in keys])
    o.put_world(w)

    def prebuild(ks):
        g.prebuild_batch(ks)
        for k in ks:
            o.prebuild_sky(*k)

    def build(ks, expand=True):
        g.light_build(ks, expand)
        o.light_batch(ks, expand)

    # 1. the plain path: everything prebuilt, inner chunks lit
    prebuild(keys)
    build(inner)
    _check(g, o, keys, "first build")

    # 2. edits that move sky columns, add and remove emitters, on and off chunk borders
    edits = [(0, 200, 0, ct.STONE, 0), (15, 180, 16, ct.STONE, 0), (-1, 150, -1, ct.LAMP, 0),
             (4, 120, 4, ct.GLASS, 0), (0, 200, 0, 0, 0), (16, 130, 3, ct.LAMP, 0), (-17, 140, 8, ct.STONE, 0)]
    for e in edits:
        o.set_block(*e)
    g.edit_batch(edits)
    _check(g, o, keys, "edits")

    # 3. Lighting::clear, prebuild again (sky columns now differ), build again
    _both(g, o, lambda x: x.light_clear())
    _check(g, o, keys, "clear")
    prebuild(keys)
    build(inner)
    _check(g, o, keys, "second build")

    # 4. build a second time on top of the built light (nothing pristine any more)
    build(inner[: len(inner) // 2])
    _check(g, o, keys, "rebuild")

    # 5. clear, prebuild only every other chunk, build: zero and prebuilt chunks side by side
    _both(g, o, lambda x: x.light_clear())
    prebuild(keys[::2])
    build(inner)
    _check(g, o, keys, "partial prebuild")

    # 6. clear, prebuild, an edit in between, then build without expand
    _both(g, o, lambda x: x.light_clear())
    prebuild(keys)
    e = (3, 250, 3, ct.LAMP, 0)
    o.set_block(*e)
    g.edit_batch([e])
    build(inner, expand=False)
    _check(g, o, keys, "edit before build")

    # 7. prebuild twice, then build one chunk at a time (each build leaves its solve set non-pristine)
    _both(g, o, lambda x: x.light_clear())
    prebuild(keys)
    prebuild(keys)
    for k in inner[:5]:
        build([k])
    _check(g, o, keys, "one by one")
    g.close()
    o.close()


@pytest.mark.parametrize("env", ["VXGPU_NO_PRISTINE", "VXGPU_NO_VIRTUAL"])
def test_fallback_paths_still_match(env):
    """Seeding without the pristine-lightmap short cuts, and with lightmaps always written out (no lazy
    clear / prebuild), stays in the library behind environment switches (read when the context is
    created): the short cuts must not change a bit."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r]\n"
        "import test_gpu_relight as t\n"
        "t.test_clear_prebuild_build_sequences('terrain', 5)\n"
        "print('fallback ok')\n"
    ) % (here, os.path.join(here, "..", "oracle"), os.path.join(here, "..", "voxelengine-cpp_b200"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **{env: "1"}), capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0 and "fallback ok" in out.stdout, out.stderr[-2000:]


def test_edits_on_the_global_lightmaps_match():
    """An edit is first tried in a shared-memory box around it and runs on the lightmaps in global memory
    only when it reaches outside (a sky column falling more than 47 voxels, as edits[0] of the sequences
    above does).  VXGPU_EDIT_GLOBAL=1 sends every edit down that second path: same bits."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r]\n"
        "import test_gpu_relight as t, test_gpu_parity as p\n"
        "t.test_clear_prebuild_build_sequences('terrain', 5)\n"
        "for name in ('edits_s41', 'edits_s42'):\n"
        "    p.test_gpu_matches_oracle_and_golden(name)\n"
        "print('global ok')\n"
    ) % (here, os.path.join(here, "..", "oracle"), os.path.join(here, "..", "voxelengine-cpp_b200"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, VXGPU_EDIT_GLOBAL="1"), capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0 and "global ok" in out.stdout, (out.stdout[-500:], out.stderr[-2000:])


def test_edits_use_the_box_and_fall_back():
    """Of the edit sequence of test_clear_prebuild_build_sequences the stone placed at y = 200 under open
    sky darkens a column far longer than the box is tall: it must be counted as run on the global
    lightmaps, the small ones must not."""
    from vxgpu import gpu
    t = cases.table()
    window = (-2, -2, 5, 5)
    w = worlds.build_world("terrain", *window, seed=5)
    keys, inner = w.keys(), w.inner_keys()
    g = gpu.GpuWorld(t, *window)
    o = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    g.put_chunks([(k[0], k[1], w.chunks[k], None) for k in keys])
    o.put_world(w)
    g.prebuild_batch(keys)
    for k in keys:
        o.prebuild_sky(*k)
    g.light_build(inner, True)
    o.light_batch(inner, True)
    hm = int(worlds.heightmap(np.array([4]), np.array([4]), 5)[0])
    small = [(4, hm + 3, 4, ct.LAMP, 0), (4, hm + 3, 4, 0, 0), (20, hm + 2, 9, ct.TORCH, 0)]
    for e in small:
        o.set_block(*e)
    g.edit_batch(small)
    assert g.edit_global() == 0
    tall = [(7, 250, 7, ct.STONE, 0)]
    for e in tall:
        o.set_block(*e)
    g.edit_batch(tall)
    assert g.edit_global() == 1
    _check(g, o, keys, "box and fallback")
    g.close()
    o.close()


def test_one_call_with_more_edits_than_fit_a_launch():
    """450 edits in one vx_light_edit call: more than one cooperative launch can hold (one CTA per SM), the
    predecessor search goes through its cell hash (more than 256 edits), and edits wait for predecessors that
    ran in an earlier launch.  Lightmaps and chunk flags against the oracle, which takes them one by one."""
    from vxgpu import gpu
    t = cases.table()
    window = (-1, -1, 8, 8)
    w = worlds.build_world("terrain", *window, seed=77)
    keys, inner = w.keys(), w.inner_keys()
    rng = np.random.default_rng(78)
    hm = int(worlds.heightmap(np.array([40]), np.array([40]), 77)[0])
    pool = [ct.LAMP, ct.TORCH, ct.RED_LAMP, ct.GREEN_LAMP, ct.BLUE_LAMP, ct.LIGHTBULB, ct.STONE, ct.GLASS, ct.WATER,
            ct.LEAVES]
    edits = cases._random_edits(w, rng, 450, inner, max(1, hm - 20), min(250, hm + 30), pool)
    g = gpu.GpuWorld(t, *window)
    o = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    g.put_chunks([(k[0], k[1], w.chunks[k], None) for k in keys])
    o.put_world(w)
    g.prebuild_batch(keys)
    for k in keys:
        o.prebuild_sky(*k)
    g.light_build(inner, True)
    o.light_batch(inner, True)
    for e in edits:
        o.set_block(*e)
    g.edit_batch(edits)
    assert g.edit_stats()[2] >= 2, "expected more than one launch"
    _check(g, o, keys, "450 edits, one call")
    g.close()
    o.close()


@pytest.mark.parametrize("seed", [101, 102])
def test_dense_edits_in_a_small_area(seed):
    """1,500 edits crowded into 3x3 chunks, in batches of 64: nearly every edit overlaps several earlier ones of
    its batch (long predecessor chains, launches cut where an edit has more than eight), tall placements leave
    the shared-memory box, opaque / translucent / emissive blocks are placed and broken again.  Lightmaps and
    chunk flags against the oracle after every fifth batch."""
    from vxgpu import gpu
    t = cases.table()
    window = (-2, -2, 7, 7)
    w = worlds.build_world("terrain", *window, seed=seed)
    keys, inner = w.keys(), w.inner_keys()
    crowd = [(cx, cz) for cx in (0, 1, 2) for cz in (0, 1, 2)]
    rng = np.random.default_rng(seed)
    hm = int(worlds.heightmap(np.array([24]), np.array([24]), seed)[0])
    pool = [ct.LAMP, ct.TORCH, ct.RED_LAMP, ct.GREEN_LAMP, ct.BLUE_LAMP, ct.LIGHTBULB, ct.STONE, ct.STONE, ct.GLASS,
            ct.WATER, ct.LEAVES]
    edits = cases._random_edits(w, rng, 1400, crowd, max(1, hm - 12), min(250, hm + 20), pool)
    edits += cases._random_edits(w, rng, 100, crowd, 180, 250, [ct.STONE, ct.LAMP])  # sky columns that fall far
    g = gpu.GpuWorld(t, *window)
    o = pyoracle.OracleWorld(t, *window, pyoracle.best_kind())
    g.put_chunks([(k[0], k[1], w.chunks[k], None) for k in keys])
    o.put_world(w)
    g.prebuild_batch(keys)
    for k in keys:
        o.prebuild_sky(*k)
    g.light_build(inner, True)
    o.light_batch(inner, True)
    left_box = 0
    for b, at in enumerate(range(0, len(edits), 64)):
        batch = edits[at:at + 64]
        for e in batch:
            o.set_block(*e)
        g.edit_batch(batch)
        left_box += g.edit_global()
        if b % 5 == 4:
            _check(g, o, keys, "dense edits, batch %d" % b)
    _check(g, o, keys, "dense edits, end")
    assert left_box > 0, "no edit took the global path: the tall placements should have"
    g.close()
    o.close()
