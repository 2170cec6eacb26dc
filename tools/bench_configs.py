"""Secondary measurements on one B200 (the headline stays bench.py): the other
configs of BASELINE.json / SURVEY §8(d) and the §8(f) rows, each as one JSON
line on stdout.  GPU times are CUDA-event / wall times around the C-ABI calls;
the CPU figure beside each one is the oracle (reference TUs when built) on a
bounded sample of the same work, single-threaded unless stated.

  (the light + mesh configs — terrain, flat, worst, edits, strong scaling — are `bench.py --config ...`)
  codec      region-file blobs (extrle16 / extrle8) -> HBM and back, 1,024 chunks
  terrain    vx_terrain_fill of 1,024 base:demo-like chunks

Usage: python tools/bench_configs.py [codec terrain]
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
    which = sys.argv[1:] or ["codec", "terrain"]
    if "codec" in which:
        bench_codec()
    if "terrain" in which:
        bench_terrain()
