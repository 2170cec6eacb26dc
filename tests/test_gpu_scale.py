"""GPU parity at BASELINE.json's full sizes (run with -m gpu on a B200).

configs[1]  1,024 terrain chunks: every lightmap and every mesh against the oracle.
configs[2]  4,096 worst-case random chunks: a 16x16 block (256 chunks) and a sample of
            chunks elsewhere against the oracle run on their neighbourhoods (exact by the
            locality of light, see vxgpu/shard.py), the capacity cut-off invariants on all
            4,096, and idempotence of the mesh build.
configs[3]  10,000 place/break edits over a 64x64-chunk lit window, lightmaps and
            re-meshed chunks compared with the oracle at checkpoints.
configs[4]  the 64x64 world sharded over 2 / 4 / 8 chunk-column ranges equals
            the single-world result bit for bit (ranks emulated as separate
            vx_ctx on one GPU).
"""
import hashlib
import multiprocessing as mp
import os

import numpy as np
import pytest

import cases
import pyoracle
from vxgpu import abi, content as ct, shard, worlds

pytestmark = pytest.mark.gpu

CAP = 800000


def _gen_terrain(k):
    return k, worlds.terrain_chunk(k[0], k[1], 2019)


def _gen_random(k):
    return k, worlds.random_chunk(k[0], k[1], 7, 0.5, None, cases.table())


def _generate(fn, keys, tag):
    cache = "/tmp/vxtest_%s_%d.npy" % (tag, len(keys))
    if os.path.exists(cache):
        arr = np.load(cache)
        if arr.shape == (len(keys), abi.CHUNK_VOL):
            return {k: arr[i] for i, k in enumerate(keys)}
    procs = max(1, min(len(os.sched_getaffinity(0)), 48))
    with mp.get_context("fork").Pool(procs) as pool:
        res = dict(pool.map(fn, keys, chunksize=max(1, len(keys) // (procs * 4))))
    try:
        np.save(cache, np.stack([res[k] for k in keys]))
    except OSError:
        pass
    return res


def _rect(x0, z0, w, d):
    return [(cx, cz) for cz in range(z0, z0 + d) for cx in range(x0, x0 + w)]


@pytest.fixture(scope="module")
def terrain66():
    keys = _rect(-1, -1, 66, 66)
    return _generate(_gen_terrain, keys, "terrain66")


def _gpu(table, window, settings=None):
    from vxgpu import gpu
    return gpu.GpuWorld(table, *window, settings)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _gpu_mesh_hashes(gw, keys, batch=1024):
    """{key: sha of (vertices, indices, entries, entry floats)} via the arena API."""
    out = {}
    for at in range(0, len(keys), batch):
        part = keys[at:at + batch]
        gw.mesh_build(part, CAP)
        sz = gw.mesh_arena_sizes()
        v = np.empty(max(1, sz[0]), np.uint32)
        i = np.empty(max(1, sz[1]), np.int32)
        e = np.zeros(max(1, sz[2]), dtype=pyoracle.ENTRY_DTYPE)
        f = np.empty(max(1, sz[3]), np.uint32)
        gw.mesh_read_arena(v.ctypes.data, i.ctypes.data, e.ctypes.data, f.ctypes.data)
        for j, k in enumerate(part):
            info, sl = gw.mesh_info(j), gw.mesh_slice(j)
            h = hashlib.sha256()
            h.update(v[sl.vertex_off:sl.vertex_off + info.n_vertex_floats].tobytes())
            h.update(i[sl.index_off:sl.index_off + info.n_indices].tobytes())
            h.update(e[sl.entry_off:sl.entry_off + info.n_entries].tobytes())
            h.update(f[sl.entry_float_off:sl.entry_float_off + info.n_entry_floats].tobytes())
            out[k] = (h.hexdigest(), info.n_vertex_floats, info.n_indices, info.n_entries)
    return out


def _oracle_mesh_hash(ow, k):
    m = ow.mesh_chunk(k[0], k[1], CAP)
    h = hashlib.sha256()
    for a in (m.vertices, m.indices, m.entries, m.entry_floats):
        h.update(np.ascontiguousarray(a).tobytes())
    return (h.hexdigest(), m.vertices.size, m.indices.size, m.entries.size)


def test_config1_1024_terrain_chunks_full_compare(terrain66):
    table = cases.table()
    sh = shard.plan(-1, -1, 34, 34, 1)[0]
    gw = _gpu(table, sh.window)
    gw.put_chunks((k[0], k[1], terrain66[k], None) for k in sh.resident)
    gw.prebuild_batch(sh.resident)
    gw.light_build(sh.lit, True)
    ow = pyoracle.OracleWorld(table, *sh.window, pyoracle.best_kind())
    for k in sh.resident:
        ow.put_chunk(k[0], k[1], terrain66[k])
    for k in sh.resident:
        ow.prebuild_sky(*k)
    ow.light_batch(sh.lit, True)
    lights = gw.get_light_batch(sh.resident).reshape(len(sh.resident), -1)
    bad = [k for j, k in enumerate(sh.resident) if not np.array_equal(lights[j], ow.get_light(*k))]
    assert not bad, "lightmaps differ in %d chunks, e.g. %s" % (len(bad), bad[:4])
    for k in sh.resident:
        a, b = gw.get_info(*k), ow.get_info(*k)
        assert (a.bottom, a.top, a.highest_point, a.lighted, a.modified) == \
               (b.bottom, b.top, b.highest_point, b.lighted, b.modified), k
    got = _gpu_mesh_hashes(gw, sh.owned)
    bad = [k for k in sh.owned if got[k] != _oracle_mesh_hash(ow, k)]
    assert not bad, "meshes differ in %d chunks, e.g. %s" % (len(bad), bad[:4])
    gw.close()
    ow.close()


def test_config2_4096_worst_case_chunks():
    table = cases.table()
    keys = _rect(-1, -1, 66, 66)
    world = _generate(_gen_random, keys, "random66")
    sh = shard.plan(-1, -1, 66, 66, 1)[0]
    gw = _gpu(table, sh.window)
    for at in range(0, len(keys), 512):
        gw.put_chunks((k[0], k[1], world[k], None) for k in keys[at:at + 512])
    gw.prebuild_batch(keys)
    gw.light_build(sh.lit, True)
    rng = np.random.default_rng(3)
    sample = [sh.owned[int(i)] for i in rng.choice(len(sh.owned), 10, replace=False)]
    sample += [(0, 0), (63, 63), (0, 63)]
    # every chunk: capacity cut-off invariants
    got = {}
    for at in range(0, len(sh.owned), 1024):
        part = sh.owned[at:at + 1024]
        got.update(_gpu_mesh_hashes(gw, part))
    for k, (_, nv, ni, ne) in got.items():
        assert nv <= CAP and nv % 6 == 0 and ni <= nv // 4 + 8, (k, nv, ni)
    again = _gpu_mesh_hashes(gw, sh.owned[:1024])
    assert all(again[k] == got[k] for k in sh.owned[:1024]), "mesh build is not idempotent"
    # a 16x16 block of the world (256 chunks) against the oracle: the block with a two-chunk apron is
    # rebuilt on the CPU, the 18x18 chunks around the block lit (the ring of lit chunks around the
    # block plays the part of a shard's halo, vxgpu/shard.py), the block meshed on the thread pool
    bx, bz, bn = 23, 31, 16
    sub = _rect(bx - 2, bz - 2, bn + 4, bn + 4)
    ow = pyoracle.OracleWorld(table, bx - 2, bz - 2, bn + 4, bn + 4, pyoracle.best_kind())
    for k in sub:
        ow.put_chunk(k[0], k[1], world[k])
    for k in sub:
        ow.prebuild_sky(*k)
    ow.light_batch(_rect(bx - 1, bz - 1, bn + 2, bn + 2), True)
    block = _rect(bx, bz, bn, bn)
    lights = gw.get_light_batch(block).reshape(len(block), -1)
    bad = [k for j, k in enumerate(block) if not np.array_equal(lights[j], ow.get_light(*k))]
    assert not bad, "lightmaps of the 16x16 block differ in %d chunks, e.g. %s" % (len(bad), bad[:4])
    ow.mesh_batch(block, CAP, min(16, len(os.sched_getaffinity(0))))   # warms nothing: just the thread-pool path
    gblock = _gpu_mesh_hashes(gw, block)
    bad = [k for k in block if gblock[k] != _oracle_mesh_hash(ow, k)]
    assert not bad, "meshes of the 16x16 block differ in %d chunks, e.g. %s" % (len(bad), bad[:4])
    ow.close()
    # sampled chunks elsewhere: oracle on the 5x5 neighbourhood, inner 3x3 lit
    for (cx, cz) in sample:
        ow = pyoracle.OracleWorld(table, cx - 2, cz - 2, 5, 5, pyoracle.best_kind())
        near = [k for k in _rect(cx - 2, cz - 2, 5, 5) if k in world]
        for k in near:
            ow.put_chunk(k[0], k[1], world[k])
        for k in near:
            ow.prebuild_sky(*k)
        for k in _rect(cx - 1, cz - 1, 3, 3):
            if k in set(sh.lit):
                ow.light_chunk(k[0], k[1], True)
        assert np.array_equal(gw.get_light(cx, cz), ow.get_light(cx, cz)), ("light", cx, cz)
        assert got[(cx, cz)] == _oracle_mesh_hash(ow, (cx, cz)), ("mesh", cx, cz)
        ow.close()
    gw.close()


def test_config3_10k_edits(terrain66):
    table = cases.table()
    sh = shard.plan(-1, -1, 66, 66, 1)[0]
    keys = sh.resident
    gw = _gpu(table, sh.window)
    for at in range(0, len(keys), 512):
        gw.put_chunks((k[0], k[1], terrain66[k], None) for k in keys[at:at + 512])
    gw.prebuild_batch(keys)
    gw.light_build(sh.lit, True)
    ow = pyoracle.OracleWorld(table, *sh.window, pyoracle.best_kind())
    for k in keys:
        ow.put_chunk(k[0], k[1], terrain66[k])
    for k in keys:
        ow.prebuild_sky(*k)
    ow.light_batch(sh.lit, True)
    gw.clear_modified()
    ow.clear_modified()

    edits = worlds.edit_script(10000, 11, table)
    for at in range(0, len(edits), 100):
        batch = edits[at:at + 100]
        gw.edit_batch(batch)
        for (x, y, z, bid, st) in batch:
            ow.set_block(x, y, z, bid, st)
        if (at // 100) % 10 != 9:
            continue
        # checkpoint every 1,000 edits: flags, lightmaps and meshes of touched chunks
        mod_o = [k for k in keys if ow.get_info(*k).modified]
        mod_g = [k for k in keys if gw.get_info(*k).modified]
        assert mod_g == mod_o, ("modified sets differ after %d edits" % (at + 100))
        lights = gw.get_light_batch(mod_o).reshape(len(mod_o), -1)
        bad = [k for j, k in enumerate(mod_o) if not np.array_equal(lights[j], ow.get_light(*k))]
        assert not bad, "after %d edits lightmaps differ in %s" % (at + 100, bad[:4])
        for k in mod_o:
            assert np.array_equal(gw.get_voxels(*k), ow.get_voxels(*k)), k
            a, b = gw.get_info(*k), ow.get_info(*k)
            assert (a.bottom, a.top) == (b.bottom, b.top), k
        meshed = [k for k in mod_o if k in set(sh.owned)]
        got = _gpu_mesh_hashes(gw, meshed)
        bad = [k for k in meshed if got[k] != _oracle_mesh_hash(ow, k)]
        assert not bad, "after %d edits meshes differ in %s" % (at + 100, bad[:4])
        gw.clear_modified()
        ow.clear_modified()
    # a build on top of the edited world reads the border-plane copies the edits had to keep in
    # step with the lightmaps (k_seed's halo): relight the chunks around the last edits
    again = sorted({(e[0] >> 4, e[2] >> 4) for e in edits[-400:]}, key=lambda k: (k[1], k[0]))
    gw.light_build(again, True)
    ow.light_batch(again, True)
    near = sorted({(cx + dx, cz + dz) for (cx, cz) in again for dx in (-1, 0, 1) for dz in (-1, 0, 1)} & set(keys),
                  key=lambda k: (k[1], k[0]))
    lights = gw.get_light_batch(near).reshape(len(near), -1)
    bad = [k for j, k in enumerate(near) if not np.array_equal(lights[j], ow.get_light(*k))]
    assert not bad, "a build after the edits differs in %s" % bad[:4]
    gw.close()
    ow.close()


def test_config4_sharded_equals_single(terrain66):
    table = cases.table()

    def run(sh):
        gw = _gpu(table, sh.window)
        for at in range(0, len(sh.resident), 512):
            gw.put_chunks((k[0], k[1], terrain66[k], None) for k in sh.resident[at:at + 512])
        gw.prebuild_batch(sh.resident)
        gw.light_build(sh.lit, True)
        lights = gw.get_light_batch(sh.owned).reshape(len(sh.owned), -1)
        meshes = _gpu_mesh_hashes(gw, sh.owned)
        out = {k: (_sha(lights[j]),) + meshes[k] for j, k in enumerate(sh.owned)}
        gw.close()
        return out

    single = run(shard.plan(-1, -1, 66, 66, 1)[0])
    assert len(single) == 64 * 64
    for n in (2, 4, 8):
        merged = {}
        for sh in shard.plan(-1, -1, 66, 66, n):
            merged.update(run(sh))
        assert merged.keys() == single.keys()
        bad = [k for k in single if merged[k] != single[k]]
        assert not bad, "%d shards: %d chunks differ, e.g. %s" % (n, len(bad), bad[:4])


def test_build_after_a_much_smaller_build_is_requeued():
    """The whole-voxel emission pass is launched over what the last build of the context listed (plus a
    margin), not over the whole unit list.  A build that lists far more such voxels than the one before it —
    one terrain chunk, then 25 chunks full of panes, pipes, grass and custom models — has to notice and run
    again: every mesh against the oracle, then the small build once more."""
    table = cases.table()
    window = (-1, -1, 7, 7)
    w = worlds.build_world("random", *window, seed=31, fill=0.4, table=table, ymax=48)
    tw = worlds.build_world("terrain", *window, seed=31)
    w.chunks[(0, 0)] = tw.chunks[(0, 0)]
    keys, inner = w.keys(), w.inner_keys()
    gw = _gpu(table, window)
    ow = pyoracle.OracleWorld(table, *window, pyoracle.best_kind())
    gw.put_chunks([(k[0], k[1], w.chunks[k], None) for k in keys])
    ow.put_world(w)
    gw.prebuild_batch(keys)
    for k in keys:
        ow.prebuild_sky(*k)
    gw.light_build(inner, True)
    ow.light_batch(inner, True)
    for part in ([(0, 0)], inner, [(0, 0)], inner[:3]):
        got = _gpu_mesh_hashes(gw, part)
        bad = [k for k in part if got[k] != _oracle_mesh_hash(ow, k)]
        assert not bad, "meshes differ in %s" % bad[:4]
    gw.close()
    ow.close()
