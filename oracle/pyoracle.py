"""ctypes binding of the CPU checkers (oracle/vxoracle.h).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "voxelengine-cpp_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from vxgpu import abi  # noqa: E402

REF_LIB = os.path.join(_HERE, "_ref", "libvxref.so")
PORT_LIB = os.path.join(_HERE, "libvxoracle.so")


def _bind(path):
    lib = C.CDLL(path)
    P = C.POINTER
    lib.vxo_kind.restype = C.c_char_p
    lib.vxo_create.restype = C.c_void_p
    lib.vxo_create.argtypes = [P(abi.Content), P(abi.Settings), C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32]
    lib.vxo_destroy.argtypes = [C.c_void_p]
    lib.vxo_put_chunk.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.vxo_put_chunk_encoded.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.vxo_encode_chunk.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.vxo_prebuild_sky.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.vxo_light_chunk.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.vxo_light_clear.argtypes = [C.c_void_p]
    lib.vxo_set_block.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_uint16, C.c_uint16]
    lib.vxo_get_light.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.vxo_get_voxels.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.vxo_get_info.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(abi.ChunkInfo)]
    lib.vxo_clear_modified.argtypes = [C.c_void_p]
    lib.vxo_mark_modified.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    if hasattr(lib, "vxo_terrain_fill"):
        lib.vxo_terrain_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    if hasattr(lib, "vxo_extrle_encode"):
        for f in (lib.vxo_extrle_encode, lib.vxo_extrle_decode):
            f.restype = C.c_uint64
            f.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32]
    lib.vxo_mesh_chunk.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64,
                                   P(abi.MeshInfo)]
    lib.vxo_mesh_read.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    lib.vxo_mesh_batch.restype = C.c_double
    lib.vxo_mesh_batch.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_uint64,
                                   C.c_int32, P(C.c_uint64)]
    lib.vxo_light_batch.restype = C.c_double
    lib.vxo_light_batch.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_int32]
    return lib


_LIBS = {}


def load(kind):
    """kind: 'reference' (oracle/_ref/libvxref.so) or 'port' (oracle/libvxoracle.so)."""
    if kind not in _LIBS:
        path = REF_LIB if kind == "reference" else PORT_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _LIBS[kind] = _bind(path)
    return _LIBS[kind]


def available(kind):
    return os.path.exists(REF_LIB if kind == "reference" else PORT_LIB)


def best_kind():
    return "reference" if available("reference") else "port"


class Mesh:
    """One ChunkMeshData copied to numpy.  Vertex words are kept as uint32 (the
    6th word is a packed-light bit pattern, BlocksRenderer.cpp:50-60)."""

    def __init__(self, info, vertices, indices, entries, entry_floats):
        self.cancelled = bool(info.cancelled)
        self.vertices = vertices
        self.indices = indices
        self.entries = entries  # structured array
        self.entry_floats = entry_floats

    def same_as(self, other):
        return (self.cancelled == other.cancelled
                and np.array_equal(self.vertices, other.vertices)
                and np.array_equal(self.indices, other.indices)
                and self.entries.tobytes() == other.entries.tobytes()
                and np.array_equal(self.entry_floats, other.entry_floats))


ENTRY_DTYPE = np.dtype([("position", np.uint32, 3), ("first_float", np.uint32),
                        ("n_floats", np.uint32)])


def extrle_encode(data, wide, kind=None):
    """extrle::encode16 (wide) / extrle::encode of a bytes-like; returns bytes."""
    lib = load(kind or best_kind())
    src = np.frombuffer(bytes(data), dtype=np.uint8)
    dst = np.empty(2 * src.size + 8, dtype=np.uint8)
    n = lib.vxo_extrle_encode(src.ctypes.data, src.size, dst.ctypes.data, int(wide))
    return dst[:n].tobytes()


def extrle_decode(data, wide, out_len, kind=None):
    """extrle::decode16 / decode into out_len bytes (the caller vouches for the size)."""
    lib = load(kind or best_kind())
    src = np.frombuffer(bytes(data), dtype=np.uint8)
    dst = np.zeros(out_len + 8, dtype=np.uint8)
    n = lib.vxo_extrle_decode(src.ctypes.data, src.size, dst.ctypes.data, int(wide))
    assert n <= out_len, (n, out_len)
    return dst[:n].tobytes()


def keys_array(keys):
    arr = (abi.ChunkKey * len(keys))()
    for i, (cx, cz) in enumerate(keys):
        arr[i].cx, arr[i].cz = cx, cz
    return arr


class OracleWorld:
    def __init__(self, table, ox, oz, w, d, kind=None, settings=None):
        from vxgpu import content as ct
        self.kind = kind or best_kind()
        self.lib = load(self.kind)
        self.table = table
        self.settings = settings or ct.default_settings()
        self.h = self.lib.vxo_create(C.byref(table.struct), C.byref(self.settings),
                                     ox, oz, w, d)
        if not self.h:
            raise RuntimeError("vxo_create failed")

    def close(self):
        if self.h:
            self.lib.vxo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def put_chunk(self, cx, cz, voxels, light=None):
        voxels = np.ascontiguousarray(voxels, dtype=np.uint32)
        lp = None
        if light is not None:
            light = np.ascontiguousarray(light, dtype=np.uint16)
            lp = light.ctypes.data
        rc = self.lib.vxo_put_chunk(self.h, cx, cz, voxels.ctypes.data, lp)
        if rc:
            raise RuntimeError("put_chunk(%d,%d) outside window" % (cx, cz))

    def put_chunk_encoded(self, cx, cz, voxels, lights=None):
        voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
        lp = None
        if lights is not None:
            lights = np.ascontiguousarray(lights, dtype=np.uint8)
            lp = lights.ctypes.data
        if self.lib.vxo_put_chunk_encoded(self.h, cx, cz, voxels.ctypes.data, lp):
            raise RuntimeError("put_chunk_encoded(%d,%d) outside window" % (cx, cz))

    def encode_chunk(self, cx, cz):
        """(Chunk::encode bytes, Lightmap::encode bytes)"""
        v = np.empty(abi.CHUNK_VOL * 4, dtype=np.uint8)
        l = np.empty(abi.CHUNK_VOL // 2, dtype=np.uint8)
        if self.lib.vxo_encode_chunk(self.h, cx, cz, v.ctypes.data, l.ctypes.data):
            return None, None
        return v, l

    def put_world(self, world):
        for (cx, cz) in world.keys():
            self.put_chunk(cx, cz, world.chunks[(cx, cz)])

    def terrain_fill(self, tables, items):
        arr, keep = tables.chunks(items)
        if self.lib.vxo_terrain_fill(self.h, C.byref(tables.gen), arr, len(arr) if keep else 0):
            raise RuntimeError("vxo_terrain_fill failed")

    def prebuild_sky(self, cx, cz):
        return self.lib.vxo_prebuild_sky(self.h, cx, cz)

    def light_chunk(self, cx, cz, expand=True):
        return self.lib.vxo_light_chunk(self.h, cx, cz, int(expand))

    def light_clear(self):
        return self.lib.vxo_light_clear(self.h)

    def set_block(self, x, y, z, bid, state=0):
        return self.lib.vxo_set_block(self.h, x, y, z, bid, state)

    def get_light(self, cx, cz):
        out = np.empty(abi.CHUNK_VOL, dtype=np.uint16)
        if self.lib.vxo_get_light(self.h, cx, cz, out.ctypes.data):
            return None
        return out

    def get_voxels(self, cx, cz):
        out = np.empty(abi.CHUNK_VOL, dtype=np.uint32)
        if self.lib.vxo_get_voxels(self.h, cx, cz, out.ctypes.data):
            return None
        return out

    def get_info(self, cx, cz):
        info = abi.ChunkInfo()
        self.lib.vxo_get_info(self.h, cx, cz, C.byref(info))
        return info

    def clear_modified(self):
        self.lib.vxo_clear_modified(self.h)

    def mesh_chunk(self, cx, cz, capacity=800000):
        info = abi.MeshInfo()
        rc = self.lib.vxo_mesh_chunk(self.h, cx, cz, capacity, C.byref(info))
        if rc:
            raise RuntimeError("mesh_chunk: missing chunk")
        if info.cancelled:
            return Mesh(info, np.zeros(0, np.uint32), np.zeros(0, np.int32),
                        np.zeros(0, ENTRY_DTYPE), np.zeros(0, np.uint32))
        v = np.empty(info.n_vertex_floats, dtype=np.uint32)
        i = np.empty(info.n_indices, dtype=np.int32)
        e = np.zeros(info.n_entries, dtype=ENTRY_DTYPE)
        f = np.empty(info.n_entry_floats, dtype=np.uint32)
        self.lib.vxo_mesh_read(self.h, v.ctypes.data, i.ctypes.data, e.ctypes.data,
                               f.ctypes.data)
        return Mesh(info, v, i, e, f)

    def mesh_batch(self, keys, capacity=800000, threads=1):
        faces = C.c_uint64(0)
        arr = keys_array(keys)
        sec = self.lib.vxo_mesh_batch(self.h, arr, len(keys), capacity, threads,
                                      C.byref(faces))
        return sec, faces.value

    def light_batch(self, keys, expand=True):
        arr = keys_array(keys)
        return self.lib.vxo_light_batch(self.h, arr, len(keys), int(expand))
