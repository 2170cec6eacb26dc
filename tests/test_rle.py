"""Region-file compression of the voxels / lights layers:
extrle::encode16 / decode16 and extrle::encode / decode (coders/rle.cpp:93-205),
which WorldRegions applies to Chunk::encode / Lightmap::encode bytes
(world/files/WorldRegions.cpp:60-72).  CPU part: the oracle port against golden
digests made from the reference's own coders/rle.cpp, against that code itself
where it is built, and the round trip.  GPU part: vx_chunks_put_compressed /
vx_chunks_compress against the oracle, byte for byte, incl. malformed blobs."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import pyoracle
import rle_streams
from vxgpu import abi, worlds

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "extrle.json")))
STREAMS = rle_streams.streams()


@pytest.mark.parametrize("name", sorted(STREAMS))
def test_port_matches_golden_and_round_trips(name):
    raw, wide = STREAMS[name]
    enc = pyoracle.extrle_encode(raw, wide, "port")
    assert len(enc) == GOLDEN[name]["len"]
    assert hashlib.sha256(enc).hexdigest() == GOLDEN[name]["sha256"]
    assert pyoracle.extrle_decode(enc, wide, len(raw), "port") == raw


def test_port_matches_reference_on_random_streams():
    if not pyoracle.available("reference"):
        pytest.skip("oracle/_ref/libvxref.so not built")
    rng = np.random.default_rng(77)
    for trial in range(60):
        wide = trial & 1
        n = int(rng.integers(1, 40000))
        vals = rng.integers(0, 65536 if wide else 256, 12)
        a = vals[rng.integers(0, 12, n)]
        a = np.repeat(a, rng.integers(1, 1 + int(rng.integers(1, 300)), n))[:n]
        raw = a.astype(np.uint16 if wide else np.uint8).tobytes()
        ep, er = pyoracle.extrle_encode(raw, wide, "port"), pyoracle.extrle_encode(raw, wide, "reference")
        assert ep == er, trial
        assert pyoracle.extrle_decode(er, wide, len(raw), "port") == raw
        assert pyoracle.extrle_decode(ep, wide, len(raw), "reference") == raw


def _crafted_wires(seed, n):
    """n chunks in wire format whose run lengths sit on every header boundary."""
    rng = np.random.default_rng(seed)
    mx16, mx8 = 0x4000, 0x8000
    edge16 = [1, 2, 63, 64, 65, 66, 127, 128, 129, mx16 - 1, mx16, mx16 + 1, 2 * mx16, 2 * mx16 + 3]
    out = []
    for i in range(n):
        def half(values, edges, total, dtype):
            parts, left = [], total
            order = list(rng.permutation(len(edges))) if i else list(range(len(edges)))
            while left > 0:
                for e in order:
                    r = min(left, edges[e] if rng.random() < 0.7 else int(rng.integers(1, 40)))
                    parts.append(np.full(r, values[int(rng.integers(0, len(values)))], dtype=dtype))
                    left -= r
                    if left == 0:
                        break
            return np.concatenate(parts)
        if i == 1:      # an incompressible chunk: 4 bytes per element in the state half
            ids = rng.integers(0, 30, abi.CHUNK_VOL).astype(np.uint16)
            sts = rng.integers(256, 65536, abi.CHUNK_VOL).astype(np.uint16)
            lights = rng.integers(0, 256, abi.LIGHTMAP_DATA_LEN).astype(np.uint8)
        elif i == 2:    # a uniform chunk: ids and states merge into runs across the half boundary
            ids = np.zeros(abi.CHUNK_VOL, np.uint16)
            sts = np.zeros(abi.CHUNK_VOL, np.uint16)
            lights = np.full(abi.LIGHTMAP_DATA_LEN, 0xFF, np.uint8)   # exactly max_sequence + 1 bytes
        else:
            ids = half(list(range(30)), edge16, abi.CHUNK_VOL, np.uint16)
            sts = half([0, 1, 255, 256, 257, 0x1234, 0xFFFF], edge16, abi.CHUNK_VOL, np.uint16)
            lights = half([0, 0xFF, 0xF0, 0x0F, 0x5A], [1, 2, 127, 128, 129, 130, 4000, mx8 - 1], abi.LIGHTMAP_DATA_LEN,
                          np.uint8)
        wire_v = np.concatenate([ids, sts]).astype("<u2").view(np.uint8)
        out.append((wire_v, lights))
    return out


@pytest.mark.gpu
def test_gpu_compressed_put_and_compress_on_crafted_chunks():
    from vxgpu import gpu
    t = cases.table()
    wires = _crafted_wires(3, 6)
    keys = [(i % 3 - 1, i // 3 - 1) for i in range(len(wires))]
    blobs = [(pyoracle.extrle_encode(v, 1), pyoracle.extrle_encode(l, 0)) for v, l in wires]
    g = gpu.GpuWorld(t, -1, -1, 3, 3)
    g.put_chunks_compressed((k[0], k[1], b[0], b[1]) for k, b in zip(keys, blobs))
    gv, gl = g.encode_chunks(keys)
    for i, k in enumerate(keys):
        assert np.array_equal(gv[i], wires[i][0]), ("voxels", k)
        assert np.array_equal(gl[i], wires[i][1]), ("lights", k)
    cv = g.compress_chunks(keys, abi.LAYER_VOXELS)
    cl = g.compress_chunks(keys, abi.LAYER_LIGHTS)
    for i, k in enumerate(keys):
        assert cv[i] == blobs[i][0], ("voxels blob", k, len(cv[i]), len(blobs[i][0]))
        assert cl[i] == blobs[i][1], ("lights blob", k, len(cl[i]), len(blobs[i][1]))
    # a missing chunk compresses as zeros; a too small output buffer fails and says what it needs
    zeros = g.compress_chunks([(50, 50)], abi.LAYER_VOXELS)[0]
    assert zeros == pyoracle.extrle_encode(bytes(abi.CHUNK_DATA_LEN), 1)
    with pytest.raises(RuntimeError, match="needs"):
        g.compress_chunks(keys, abi.LAYER_VOXELS, capacity=16)
    g.close()


@pytest.mark.gpu
def test_gpu_compressed_world_lights_and_meshes_like_the_oracle():
    from vxgpu import gpu
    t = cases.table()
    w = worlds.build_world("terrain", -1, -1, 3, 3, seed=12)
    rng = np.random.default_rng(12)
    for k in list(w.chunks):
        w.chunks[k] = worlds.sprinkle(w.chunks[k], rng, t, 12, 24)
    o = pyoracle.OracleWorld(t, w.ox, w.oz, w.w, w.d, pyoracle.best_kind())
    o.put_world(w)
    for k in w.keys():
        o.prebuild_sky(*k)
    o.light_chunk(0, 0, True)
    # what WorldRegions::put would have written
    blobs = {}
    for k in w.keys():
        v, l = o.encode_chunk(*k)
        blobs[k] = (pyoracle.extrle_encode(v, 1), pyoracle.extrle_encode(l, 0))
    assert max(len(b[0]) for b in blobs.values()) < abi.CHUNK_DATA_LEN // 8   # terrain compresses well
    g = gpu.GpuWorld(t, w.ox, w.oz, w.w, w.d)
    o2 = pyoracle.OracleWorld(t, w.ox, w.oz, w.w, w.d, pyoracle.best_kind())
    g.put_chunks_compressed((k[0], k[1], blobs[k][0], blobs[k][1] if k != (1, 1) else None) for k in w.keys())
    for k in w.keys():
        v = np.frombuffer(pyoracle.extrle_decode(blobs[k][0], 1, abi.CHUNK_DATA_LEN), dtype=np.uint8)
        l = np.frombuffer(pyoracle.extrle_decode(blobs[k][1], 0, abi.LIGHTMAP_DATA_LEN), dtype=np.uint8)
        o2.put_chunk_encoded(k[0], k[1], v, l if k != (1, 1) else None)
    g.light_build([(0, 0)], False)      # flags.loadedLights: onChunkLoaded(expand = false)
    o2.light_chunk(0, 0, False)
    for k in w.keys():
        assert np.array_equal(g.get_voxels(*k), o2.get_voxels(*k)), k
        assert np.array_equal(g.get_light(*k), o2.get_light(*k)), k
    m1, m2 = g.mesh_chunk(0, 0), o2.mesh_chunk(0, 0)
    assert np.array_equal(m1.vertices, m2.vertices) and np.array_equal(m1.indices, m2.indices)
    # and back: the device's blobs are the reference's
    for layer, idx in ((abi.LAYER_VOXELS, 0), (abi.LAYER_LIGHTS, 1)):
        got = g.compress_chunks(w.keys(), layer)
        for k, b in zip(w.keys(), got):
            ev, el = o2.encode_chunk(*k)
            assert b == pyoracle.extrle_encode((ev, el)[idx], 1 - idx), (layer, k)
    for x in (g, o, o2):
        x.close()


@pytest.mark.gpu
def test_gpu_malformed_blobs_fail_and_put_nothing():
    from vxgpu import gpu
    t = cases.table()
    wire_v, wire_l = _crafted_wires(4, 1)[0]
    good_v, good_l = pyoracle.extrle_encode(wire_v, 1), pyoracle.extrle_encode(wire_l, 0)
    rng = np.random.default_rng(9)
    bad = {
        "truncated header": (good_v[:-1], good_l),
        "short": (good_v[: len(good_v) // 2], good_l),
        "long": (good_v + bytes([5, 1]), good_l),
        "empty": (b"", good_l),
        "noise": (rng.integers(0, 256, 3000, dtype=np.uint8).tobytes(), good_l),
        "huge runs": (bytes([0xFF, 0xFF, 0xFF, 0xFF]) * 64, good_l),
        "bad lights": (good_v, good_l[:-1]),
        "long lights": (good_v, good_l + bytes([0, 0])),
    }
    for what, (v, l) in bad.items():
        g = gpu.GpuWorld(t, -1, -1, 3, 3)
        with pytest.raises(RuntimeError, match="expected decompressed size"):
            g.put_chunks_compressed([(0, 0, good_v, good_l), (1, 0, v, l)])
        assert not g.get_info(1, 0).present and not g.get_info(0, 0).present, what
        # the context is still usable
        g.put_chunks_compressed([(0, 0, good_v, good_l)])
        assert np.array_equal(g.encode_chunk(0, 0)[0], wire_v), what
        g.close()
