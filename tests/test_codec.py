"""Chunk wire format: Chunk::encode / decode and
Lightmap::encode / decode.  CPU part: the oracle port against the reference's
own functions, and the round trip the reference tests in test/voxels/Chunk.cpp:
5-25 (random voxels survive encode -> decode).  GPU part: the device codec and
the lights-cache fast path (loadedLights => onChunkLoaded(expand = false),
logic/ChunksController.cpp:118-122) against the oracle, bit for bit."""
import numpy as np
import pytest

import cases
import pyoracle
from vxgpu import abi, worlds


def _random_world(seed, lit=True):
    t = cases.table()
    w = worlds.build_world("terrain", -1, -1, 3, 3, seed=seed)
    rng = np.random.default_rng(seed)
    for k in list(w.chunks):
        v = worlds.sprinkle(w.chunks[k], rng, t, 12, 24)
        # user bits / rotations in the state half, as the reference's Chunk test does
        v = v | ((rng.integers(0, 256, size=v.size).astype(np.uint32) * (v != 0)) << np.uint32(24))
        w.chunks[k] = v
    return t, w


def _oracle(t, w, kind):
    o = pyoracle.OracleWorld(t, w.ox, w.oz, w.w, w.d, kind)
    o.put_world(w)
    for k in w.keys():
        o.prebuild_sky(*k)
    o.light_chunk(0, 0, True)
    return o


def test_port_codec_matches_reference():
    if not pyoracle.available("reference"):
        pytest.skip("oracle/_ref/libvxref.so not built")
    t, w = _random_world(5)
    a, b = _oracle(t, w, "reference"), _oracle(t, w, "port")
    for k in w.keys():
        va, la = a.encode_chunk(*k)
        vb, lb = b.encode_chunk(*k)
        assert np.array_equal(va, vb) and np.array_equal(la, lb), k
    a.close()
    b.close()


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_encode_decode_round_trip(kind):
    """test/voxels/Chunk.cpp:5-25: decode(encode(chunk)) == chunk; lights keep the sky nibble only."""
    if not pyoracle.available(kind):
        pytest.skip("library not built")
    t, w = _random_world(6)
    o = _oracle(t, w, kind)
    o2 = pyoracle.OracleWorld(t, w.ox, w.oz, w.w, w.d, kind)
    for k in w.keys():
        v, l = o.encode_chunk(*k)
        assert v.size == abi.CHUNK_DATA_LEN and l.size == abi.LIGHTMAP_DATA_LEN
        # layout: ids then states, little endian
        ids = v[:abi.CHUNK_VOL * 2].view("<u2")
        sts = v[abi.CHUNK_VOL * 2:].view("<u2")
        assert np.array_equal(ids, (w.chunks[k] & 0xFFFF).astype(np.uint16))
        assert np.array_equal(sts, (w.chunks[k] >> 16).astype(np.uint16))
        o2.put_chunk_encoded(k[0], k[1], v, l)
        assert np.array_equal(o2.get_voxels(*k), w.chunks[k])
        assert np.array_equal(o2.get_light(*k), o.get_light(*k) & 0xF000)
        a, b = o.get_info(*k), o2.get_info(*k)
        assert (a.bottom, a.top) == (b.bottom, b.top)
    o.close()
    o2.close()


@pytest.mark.gpu
def test_gpu_codec_and_lights_cache_path():
    from vxgpu import gpu
    t, w = _random_world(7)
    o = _oracle(t, w, pyoracle.best_kind())
    g = gpu.GpuWorld(t, w.ox, w.oz, w.w, w.d)
    g.put_world(w)
    g.prebuild_batch(w.keys())
    g.light_build([(0, 0)], True)
    # encode on the device == the reference's encode
    gv, gl = g.encode_chunks(w.keys())
    enc = {}
    for i, k in enumerate(w.keys()):
        v, l = o.encode_chunk(*k)
        assert np.array_equal(gv[i], v) and np.array_equal(gl[i], l), k
        enc[k] = (v, l)
    miss_v, miss_l = g.encode_chunks([(99, 99)])
    assert not miss_v.any() and not miss_l.any()
    # reload from the wire format with cached lights, then light without expand
    # (ChunksController::buildLights with flags.loadedLights)
    g2 = gpu.GpuWorld(t, w.ox, w.oz, w.w, w.d)
    o2 = pyoracle.OracleWorld(t, w.ox, w.oz, w.w, w.d, pyoracle.best_kind())
    g2.put_chunks_encoded((k[0], k[1], enc[k][0], enc[k][1]) for k in w.keys())
    for k in w.keys():
        o2.put_chunk_encoded(k[0], k[1], enc[k][0], enc[k][1])
    for k in w.keys():
        assert np.array_equal(g2.get_voxels(*k), o2.get_voxels(*k)), k
        assert np.array_equal(g2.get_light(*k), o2.get_light(*k)), k
    g2.light_build([(0, 0)], False)
    o2.light_chunk(0, 0, False)
    for k in w.keys():
        assert np.array_equal(g2.get_light(*k), o2.get_light(*k)), k
        a, b = g2.get_info(*k), o2.get_info(*k)
        assert (a.bottom, a.top, a.lighted, a.modified) == (b.bottom, b.top, b.lighted, b.modified), k
    m1, m2 = g2.mesh_chunk(0, 0), o2.mesh_chunk(0, 0)
    assert np.array_equal(m1.vertices, m2.vertices) and np.array_equal(m1.indices, m2.indices)
    assert np.array_equal(m1.entry_floats, m2.entry_floats)
    # chunks without cached lights decode to a zero lightmap
    g3 = gpu.GpuWorld(t, w.ox, w.oz, w.w, w.d)
    g3.put_chunks_encoded((k[0], k[1], enc[k][0], None) for k in w.keys())
    assert not g3.get_light(0, 0).any()
    assert np.array_equal(g3.get_voxels(0, 0), w.chunks[(0, 0)])
    for x in (g, g2, g3, o, o2):
        x.close()
