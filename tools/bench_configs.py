"""Secondary measurements on one B200 (the headline stays bench.py): the other
configs of BASELINE.json / SURVEY §8(d) and the §8(f) rows, each as one JSON
line on stdout.  GPU times are CUDA-event / wall times around the C-ABI calls;
the CPU figure beside each one is the oracle (reference TUs when built) on a
bounded sample of the same work, single-threaded unless stated.

  edits      configs[4]: 10,000 place/break edits on a lit 64x64-chunk window in
             batches of 100, re-meshing the chunks each batch marks modified
  worst      configs[3b]: 4,096 random chunks, full light + mesh at capacity 800,000
  flat       configs[2a]: 1,024 one-layer `core:obstacle` chunks
  codec      region-file blobs (extrle16 / extrle8) -> HBM and back, 1,024 chunks
  terrain    vx_terrain_fill of 1,024 base:demo-like chunks

Usage: python tools/bench_configs.py [edits worst flat codec terrain]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("voxelengine-cpp_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

import numpy as np  # noqa: E402

import cases  # noqa: E402
import pyoracle  # noqa: E402
import test_gpu_scale as ts  # noqa: E402  (world generators, edit script)
from vxgpu import abi, gpu, shard, terrain, worlds  # noqa: E402
from vxgpu import content as ct  # noqa: E402

CAP = 800000


def emit(**kw):
    print(json.dumps(kw), flush=True)


def load_world(gw, world, keys):
    for at in range(0, len(keys), 512):
        gw.put_chunks((k[0], k[1], world[k], None) for k in keys[at:at + 512])


def bench_edits():
    table = cases.table()
    keys = ts._rect(-1, -1, 66, 66)
    world = ts._generate(ts._gen_terrain, keys, "terrain66")
    sh = shard.plan(-1, -1, 66, 66, 1)[0]
    gw = gpu.GpuWorld(table, *sh.window)
    load_world(gw, world, keys)
    gw.prebuild_batch(keys)
    gw.light_build(sh.lit, True)
    import ctypes as C

    def frame():
        # one frame of the device-side scheduler: meshes stay in the HBM arenas (no read-back here)
        st = abi.UpdateStats()
        gw._check(gw.lib.vx_window_update(gw.h, 1, CAP, C.byref(st)))
        return st.n_meshed
    frame()              # first frame: every lighted chunk gets its first mesh
    gw.clear_modified()

    class H(dict):
        def __contains__(self, k):
            return True

        def __missing__(self, k):
            return worlds.heightmap(np.array([k[0]]), np.array([k[1]]), 2019)[0]
    edits = ts._edit_script(sh.owned, H(), 10000, 11)
    t_edit = t_mesh = 0.0
    remeshed = 0
    for at in range(0, len(edits), 100):
        t0 = time.perf_counter()
        gw.edit_batch(edits[at:at + 100])
        t1 = time.perf_counter()
        # ChunksRenderer: every lighted chunk whose flags.modified is set gets a new mesh
        n = frame()
        t2 = time.perf_counter()
        t_edit += t1 - t0
        t_mesh += t2 - t1
        remeshed += n
    gw.close()
    # CPU: the same first 1,000 edits through the oracle (blocks_agent::set + Lighting::onBlockSet)
    ow = pyoracle.OracleWorld(table, *sh.window, pyoracle.best_kind())
    near = [k for k in keys]
    for k in near:
        ow.put_chunk(k[0], k[1], world[k])
    for k in near:
        ow.prebuild_sky(*k)
    ow.light_batch(sh.lit, True)
    t0 = time.perf_counter()
    for e in edits[:1000]:
        ow.set_block(*e)
    t_cpu = time.perf_counter() - t0
    ow.close()
    emit(config="configs[4] edits", edits=len(edits), batch=100,
         gpu_edits_per_s=round(len(edits) / t_edit, 1), gpu_ms_per_batch=round(1e3 * t_edit / (len(edits) / 100), 3),
         remeshed_chunks=remeshed, gpu_remesh_chunks_per_s=round(remeshed / t_mesh, 1),
         cpu_edits_per_s=round(1000 / t_cpu, 1), cpu_kind=pyoracle.best_kind(), cpu_sample="first 1,000 edits, 1 thread")


def bench_world(tag, gen, n_side, what, sample_side=8):
    table = cases.table()
    keys = ts._rect(-1, -1, n_side + 2, n_side + 2)
    world = ts._generate(gen, keys, tag)
    sh = shard.plan(-1, -1, n_side + 2, n_side + 2, 1)[0]
    gw = gpu.GpuWorld(table, *sh.window)
    load_world(gw, world, keys)

    def step():
        gw.light_clear()
        gw.prebuild_batch(keys)
        gw.light_build(sh.lit, True)
        for at in range(0, len(sh.owned), 1024):
            gw.mesh_build(sh.owned[at:at + 1024], CAP)
    for _ in range(2):
        step()
    t0 = time.perf_counter()
    K = 3
    for _ in range(K):
        step()
    dt = (time.perf_counter() - t0) / K
    sz = gw.mesh_arena_sizes()
    gw.profile_enable(True)
    step()
    prof = gw.profile()
    gw.close()
    # CPU sample: sample_side^2 chunks with their ring, light on one thread + mesh on one thread
    skeys = ts._rect(-1, -1, sample_side + 2, sample_side + 2)
    ssh = shard.plan(-1, -1, sample_side + 2, sample_side + 2, 1)[0]
    ow = pyoracle.OracleWorld(table, *ssh.window, pyoracle.best_kind())
    for k in skeys:
        ow.put_chunk(k[0], k[1], world[k])
    t0 = time.perf_counter()
    for k in skeys:
        ow.prebuild_sky(*k)
    ow.light_batch(ssh.lit, True)
    ow.mesh_batch(ssh.owned, CAP, 1)
    t_cpu = time.perf_counter() - t0
    ow.close()
    emit(config=what, chunks=len(sh.owned), gpu_chunks_per_s=round(len(sh.owned) / dt, 1),
         gpu_ms_per_step=round(dt * 1e3, 3), last_batch_vertex_floats=int(sz[0]),
         kernel_ms={k: round(v[0], 4) for k, v in prof.items() if v[1]},
         cpu_chunks_per_s=round(len(ssh.owned) / t_cpu, 1), cpu_kind=pyoracle.best_kind(),
         cpu_sample="%d chunks, 1 thread" % len(ssh.owned))


def _gen_flat(k):
    v = np.zeros(abi.CHUNK_VOL, np.uint32)
    v[:256] = ct.OBSTACLE   # res/generators/default: one core:obstacle layer at y = 0
    return k, v


def bench_codec():
    table = cases.table()
    sh = shard.plan(-1, -1, 34, 34, 1)[0]
    keys = sh.resident
    world = ts._generate(ts._gen_terrain, ts._rect(-1, -1, 66, 66), "terrain66")
    gw = gpu.GpuWorld(table, *sh.window)
    load_world(gw, world, keys)
    gw.prebuild_batch(keys)
    gw.light_build(sh.lit, True)
    # HBM -> region-file blobs
    for _ in range(2):
        vb = gw.compress_chunks(keys, abi.LAYER_VOXELS)
        lb = gw.compress_chunks(keys, abi.LAYER_LIGHTS)
    t0 = time.perf_counter()
    vb = gw.compress_chunks(keys, abi.LAYER_VOXELS)
    lb = gw.compress_chunks(keys, abi.LAYER_LIGHTS)
    t_enc = time.perf_counter() - t0
    gw2 = gpu.GpuWorld(table, *sh.window)
    items = [(k[0], k[1], vb[i], lb[i]) for i, k in enumerate(keys)]
    gw2.put_chunks_compressed(items)
    t0 = time.perf_counter()
    gw2.put_chunks_compressed(items)
    t_dec = time.perf_counter() - t0
    ok = np.array_equal(gw2.get_voxels(*keys[40]), world[keys[40]])
    gw.close()
    gw2.close()
    # CPU: extrle encode + decode of the same blobs for 64 chunks
    ow = pyoracle.OracleWorld(table, *sh.window, pyoracle.best_kind())
    for k in keys[:64]:
        ow.put_chunk(k[0], k[1], world[k])
    t0 = time.perf_counter()
    for k in keys[:64]:
        v, l = ow.encode_chunk(*k)
        pyoracle.extrle_encode(v, 1)
        pyoracle.extrle_encode(l, 0)
    t_cpu_enc = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(64):
        pyoracle.extrle_decode(vb[i], 1, abi.CHUNK_DATA_LEN)
        pyoracle.extrle_decode(lb[i], 0, abi.LIGHTMAP_DATA_LEN)
    t_cpu_dec = time.perf_counter() - t0
    ow.close()
    emit(config="region-file blobs (extrle16 voxels + extrle8 lights)", chunks=len(keys), round_trip_ok=bool(ok),
         blob_bytes_per_chunk=round(sum(map(len, vb)) / len(keys) + sum(map(len, lb)) / len(keys), 1),
         raw_bytes_per_chunk=abi.CHUNK_DATA_LEN + abi.LIGHTMAP_DATA_LEN,
         gpu_compress_chunks_per_s=round(len(keys) / t_enc, 1), gpu_put_compressed_chunks_per_s=round(len(keys) / t_dec, 1),
         cpu_encode_chunks_per_s=round(64 / t_cpu_enc, 1), cpu_decode_chunks_per_s=round(64 / t_cpu_dec, 1),
         cpu_kind=pyoracle.best_kind(), cpu_sample="64 chunks, 1 thread (via ctypes)")


def bench_terrain():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import terrain_cases
    table = cases.table()
    keys = ts._rect(0, 0, 32, 32)
    tables = terrain.demo_tables()
    t0 = time.perf_counter()
    items = terrain_cases.demo_inputs(keys)
    t_inputs = time.perf_counter() - t0
    gw = gpu.GpuWorld(table, 0, 0, 32, 32)
    gw.terrain_fill(tables, items)
    t0 = time.perf_counter()
    K = 3
    for _ in range(K):
        gw.terrain_fill(tables, items)
    dt = (time.perf_counter() - t0) / K
    prof_on = gw.profile_enable(True)
    gw.terrain_fill(tables, items)
    prof = gw.profile()
    gw.close()
    ow = pyoracle.OracleWorld(table, 0, 0, 32, 32, pyoracle.best_kind())
    t0 = time.perf_counter()
    ow.terrain_fill(tables, items[:128])
    t_cpu = time.perf_counter() - t0
    ow.close()
    emit(config="terrain fill (base:demo land + plants)", chunks=len(keys),
         gpu_chunks_per_s=round(len(keys) / dt, 1), gpu_call_ms=round(dt * 1e3, 3),
         kernel_ms={k: round(v[0], 4) for k, v in prof.items()},
         cpu_chunks_per_s=round(128 / t_cpu, 1), cpu_kind=pyoracle.best_kind(), cpu_sample="128 chunks, 1 thread",
         note="heightmap / biome inputs are generated on the host (%.1f s in numpy here) and are not part of either time" % t_inputs)


if __name__ == "__main__":
    which = sys.argv[1:] or ["edits", "worst", "flat", "codec", "terrain"]
    if "edits" in which:
        bench_edits()
    if "worst" in which:
        bench_world("random66", ts._gen_random, 64, "configs[3b] 4,096 worst-case random chunks, capacity 800,000")
    if "flat" in which:
        bench_world("flat34", _gen_flat, 32, "configs[2a] 1,024 one-layer core:obstacle chunks")
    if "codec" in which:
        bench_codec()
    if "terrain" in which:
        bench_terrain()
