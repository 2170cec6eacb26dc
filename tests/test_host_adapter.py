"""The compiled C++ host adapter (voxelengine-cpp_b200/host/GpuChunkBuild.{hpp,cpp}: GpuContext,
GpuLighting, GpuRendererBatcher) against the reference's own classes, in C++ (run with -m gpu).

oracle/_ref/libvxadapter_test.so holds the reference translation units, the adapter and the
self-test of tests/native/adapter_test.cpp (built by oracle/ref_build/Makefile where
/root/reference exists; it travels to the GPU box like the other built libraries).  The test makes
two worlds of the reference's Content / Chunks / Chunk objects, lights and meshes one with
Lighting + BlocksRenderer and the other through the adapter, and memcmp()s every lightmap, flag
and ChunkMeshData; it also walks the inwork / update / unload / clear contract of
ChunksRenderer.cpp:95-139 and the Block -> vx_block_def flattening.  This file only supplies the
block table, the voxels and the edit list.
"""
import ctypes as C
import os

import numpy as np
import pytest

import cases
from vxgpu import abi, content as ct, worlds

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libvxadapter_test.so")


def _lib():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libvxadapter_test.so is not built (needs /root/reference at build time)")
    lib = C.CDLL(LIB)
    lib.vxh_adapter_selftest.restype = C.c_int
    lib.vxh_adapter_selftest.argtypes = [C.POINTER(abi.Content), C.POINTER(abi.Settings), C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_char_p, C.c_int32]
    return lib


def _edits(rng, table, ox, oz, w, d, n):
    pool = [ct.LAMP, ct.TORCH, ct.RED_LAMP, ct.GLASS, ct.STONE, ct.LIGHTBULB, ct.WATER]
    variants = table.rot_variants()
    arr = (abi.Edit * n)()
    placed = []
    for i in range(n):
        if placed and rng.random() < 0.4:
            x, y, z = placed.pop(int(rng.integers(0, len(placed))))
            bid, st = 0, 0
        else:
            x = int(rng.integers((ox + 1) * 16, (ox + w - 1) * 16))
            z = int(rng.integers((oz + 1) * 16, (oz + d - 1) * 16))
            y = int(rng.integers(30, 110))
            bid = int(pool[int(rng.integers(0, len(pool)))])
            st = int(rng.integers(0, variants[bid]))
            placed.append((x, y, z))
        arr[i].x, arr[i].y, arr[i].z, arr[i].id, arr[i].state = x, y, z, bid, st
    return arr


@pytest.mark.parametrize("kind,seed,dense", [("terrain", 2019, 1), ("random", 5, 1), ("terrain", 77, 0)])
def test_cpp_adapter_equals_reference_classes(kind, seed, dense):
    lib = _lib()
    table = cases.table()
    ox, oz, w, d = -2, 3, 5, 4
    rng = np.random.default_rng(seed)
    vox = np.empty((d, w, abi.CHUNK_VOL), np.uint32)
    for cz in range(d):
        for cx in range(w):
            if kind == "terrain":
                v = worlds.sprinkle(worlds.terrain_chunk(ox + cx, oz + cz, seed), rng, table, 6, 10)
            else:
                v = worlds.random_chunk(ox + cx, oz + cz, seed, 0.35, None, table, ymax=96)
            vox[cz, cx] = v
    settings = ct.default_settings()
    settings.dense_render = dense
    edits = _edits(rng, table, ox, oz, w, d, 40)
    msg = C.create_string_buffer(512)
    rc = lib.vxh_adapter_selftest(C.byref(table.struct), C.byref(settings), ox, oz, w, d, vox.ctypes.data,
                                  C.cast(edits, C.c_void_p), len(edits), 0, msg, len(msg))
    assert rc == 0, "adapter self-test failed at step %d: %s" % (rc, msg.value.decode())
