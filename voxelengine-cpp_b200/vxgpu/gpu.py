"""ctypes binding of libvxgpu.so (include/vxgpu.h) and GpuWorld, the host-side
mirror of the reference's Chunks + Lighting + BlocksRenderer seam used by the
tests and the bench.  There is no CPU fallback: importing works anywhere, but
creating a GpuWorld without the built library or without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from . import content as ct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvxgpu.so")

# every symbol include/vxgpu.h declares
SYMBOLS = [
    "vx_ctx_create", "vx_ctx_destroy", "vx_last_error", "vx_window_configure",
    "vx_chunks_put", "vx_chunks_remove", "vx_chunk_get_info", "vx_chunk_get_voxels",
    "vx_chunks_clear_modified", "vx_light_prebuild", "vx_light_build", "vx_light_clear",
    "vx_light_edit", "vx_light_get", "vx_mesh_build", "vx_mesh_get_info", "vx_mesh_read",
    "vx_last_timing", "vx_last_edit_stats", "vx_last_edit_global", "vx_host_alloc", "vx_host_free", "vx_light_get_batch",
    "vx_mesh_get_slice", "vx_mesh_arena_sizes", "vx_mesh_read_arena", "vx_timer_begin",
    "vx_timer_end", "vx_profile_enable", "vx_profile_get", "vx_chunks_put_encoded",
    "vx_chunks_encode", "vx_window_update", "vx_chunks_put_compressed", "vx_chunks_compress", "vx_terrain_fill",
]

_lib = None


def load():
    """Loads libvxgpu.so; raises if it was not built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libvxgpu.so is not built: run `make -C voxelengine-cpp_b200/csrc` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.vx_ctx_create.restype = C.c_void_p
    lib.vx_ctx_create.argtypes = [P(abi.Content), P(abi.Settings), C.c_int]
    lib.vx_ctx_destroy.argtypes = [C.c_void_p]
    lib.vx_last_error.restype = C.c_char_p
    lib.vx_last_error.argtypes = [C.c_void_p]
    lib.vx_window_configure.argtypes = [C.c_void_p] + [C.c_int32] * 4
    lib.vx_chunks_put.argtypes = [C.c_void_p, P(abi.ChunkIn), C.c_int32]
    lib.vx_chunks_remove.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32]
    lib.vx_chunk_get_info.argtypes = [C.c_void_p, C.c_int32, C.c_int32, P(abi.ChunkInfo)]
    lib.vx_chunk_get_voxels.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.vx_chunks_clear_modified.argtypes = [C.c_void_p]
    lib.vx_light_prebuild.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32]
    lib.vx_light_build.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_int32]
    lib.vx_light_clear.argtypes = [C.c_void_p]
    lib.vx_light_edit.argtypes = [C.c_void_p, P(abi.Edit), C.c_int32]
    lib.vx_light_get.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.vx_mesh_build.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_uint64]
    lib.vx_mesh_get_info.argtypes = [C.c_void_p, C.c_int32, P(abi.MeshInfo)]
    lib.vx_mesh_read.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 4
    lib.vx_last_timing.argtypes = [C.c_void_p, P(C.c_float), P(C.c_float), P(C.c_int64)]
    lib.vx_last_edit_stats.argtypes = [C.c_void_p, P(C.c_float), P(C.c_int32), P(C.c_int32)]
    lib.vx_last_edit_global.argtypes = [C.c_void_p]
    lib.vx_last_edit_global.restype = C.c_int32
    lib.vx_light_get_batch.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_void_p]
    lib.vx_mesh_get_slice.argtypes = [C.c_void_p, C.c_int32, P(abi.MeshSlice)]
    lib.vx_mesh_arena_sizes.argtypes = [C.c_void_p, P(C.c_uint64)]
    lib.vx_mesh_read_arena.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.vx_timer_begin.argtypes = [C.c_void_p]
    lib.vx_timer_end.argtypes = [C.c_void_p, P(C.c_float)]
    lib.vx_profile_enable.argtypes = [C.c_void_p, C.c_int32]
    lib.vx_profile_get.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, P(C.c_float),
                                   P(C.c_int64)]
    lib.vx_chunks_put_encoded.argtypes = [C.c_void_p, P(abi.ChunkEncoded), C.c_int32]
    lib.vx_chunks_encode.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_void_p, C.c_void_p]
    lib.vx_terrain_fill.argtypes = [C.c_void_p, P(abi.GenDef), P(abi.GenChunk), C.c_int32]
    lib.vx_chunks_put_compressed.argtypes = [C.c_void_p, P(abi.ChunkCompressed), C.c_int32]
    lib.vx_chunks_compress.argtypes = [C.c_void_p, P(abi.ChunkKey), C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_uint64, P(C.c_uint64)]
    lib.vx_window_update.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, P(abi.UpdateStats)]
    lib.vx_host_alloc.restype = C.c_void_p
    lib.vx_host_alloc.argtypes = [C.c_size_t]
    lib.vx_host_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


ENTRY_DTYPE = np.dtype([("position", np.uint32, 3), ("first_float", np.uint32),
                        ("n_floats", np.uint32)])


def keys_array(keys):
    if isinstance(keys, C.Array):  # already built (hot loops build it once)
        return keys
    arr = (abi.ChunkKey * max(1, len(keys)))()
    for i, (cx, cz) in enumerate(keys):
        arr[i].cx, arr[i].cz = cx, cz
    return arr


class Mesh:
    """One ChunkMeshData copied to numpy; vertex words stay uint32 (the 6th word
    is a packed-light bit pattern, BlocksRenderer.cpp:50-60)."""

    def __init__(self, cancelled, vertices, indices, entries, entry_floats):
        self.cancelled = bool(cancelled)
        self.vertices = vertices
        self.indices = indices
        self.entries = entries
        self.entry_floats = entry_floats


class VxError(RuntimeError):
    pass


class GpuWorld:
    """Chunks window + Lighting + BlocksRenderer on one B200 (one vx_ctx)."""

    def __init__(self, table, ox, oz, w, d, settings=None, device=0):
        self.lib = load()
        self.table = table
        self.settings = settings or ct.default_settings()
        self.h = self.lib.vx_ctx_create(C.byref(table.struct), C.byref(self.settings), device)
        if not self.h:
            raise VxError(self.lib.vx_last_error(None).decode())
        self._check(self.lib.vx_window_configure(self.h, ox, oz, w, d))

    def _check(self, rc):
        if rc != 0:
            raise VxError(self.lib.vx_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.vx_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Chunks
    def configure(self, ox, oz, w, d):
        self._check(self.lib.vx_window_configure(self.h, ox, oz, w, d))

    def put_chunks(self, items):
        """items: iterable of (cx, cz, voxels u32[65536], light u16[65536] or None)."""
        items = list(items)
        arr = (abi.ChunkIn * max(1, len(items)))()
        keep = []
        for i, (cx, cz, vox, light) in enumerate(items):
            vox = np.ascontiguousarray(vox, dtype=np.uint32)
            keep.append(vox)
            arr[i].cx, arr[i].cz = cx, cz
            arr[i].voxels = vox.ctypes.data
            if light is not None:
                light = np.ascontiguousarray(light, dtype=np.uint16)
                keep.append(light)
                arr[i].light = light.ctypes.data
        self._check(self.lib.vx_chunks_put(self.h, arr, len(items)))

    def put_chunk(self, cx, cz, voxels, light=None):
        self.put_chunks([(cx, cz, voxels, light)])

    def put_world(self, world):
        self.put_chunks((cx, cz, world.chunks[(cx, cz)], None) for (cx, cz) in world.keys())

    def put_chunks_encoded(self, items):
        """items: (cx, cz, Chunk::encode bytes, Lightmap::encode bytes or None)."""
        items = list(items)
        arr = (abi.ChunkEncoded * max(1, len(items)))()
        keep = []
        for i, (cx, cz, vox, lights) in enumerate(items):
            vox = np.ascontiguousarray(vox, dtype=np.uint8)
            assert vox.size == abi.CHUNK_DATA_LEN
            keep.append(vox)
            arr[i].cx, arr[i].cz = cx, cz
            arr[i].voxels = vox.ctypes.data
            if lights is not None:
                lights = np.ascontiguousarray(lights, dtype=np.uint8)
                assert lights.size == abi.LIGHTMAP_DATA_LEN
                keep.append(lights)
                arr[i].lights = lights.ctypes.data
        self._check(self.lib.vx_chunks_put_encoded(self.h, arr, len(items)))

    def put_chunk_encoded(self, cx, cz, voxels, lights=None):
        self.put_chunks_encoded([(cx, cz, voxels, lights)])

    def encode_chunks(self, keys):
        """(n x CHUNK_DATA_LEN, n x LIGHTMAP_DATA_LEN) uint8 arrays."""
        v = np.empty((len(keys), abi.CHUNK_DATA_LEN), dtype=np.uint8)
        l = np.empty((len(keys), abi.LIGHTMAP_DATA_LEN), dtype=np.uint8)
        self._check(self.lib.vx_chunks_encode(self.h, keys_array(keys), len(keys), v.ctypes.data,
                                              l.ctypes.data))
        return v, l

    def encode_chunk(self, cx, cz):
        v, l = self.encode_chunks([(cx, cz)])
        return v[0], l[0]

    def terrain_fill(self, tables, items):
        """WorldGenerator::generate (land + plants) on the device for (cx, cz, heights, biomes) items;
        `tables` is a vxgpu.terrain.Tables."""
        arr, keep = tables.chunks(items)
        self._check(self.lib.vx_terrain_fill(self.h, C.byref(tables.gen), arr, len(arr) if keep else 0))

    def put_chunks_compressed(self, items):
        """items: (cx, cz, region-file voxels blob, lights blob or None) — extrle16 / extrle8 bytes."""
        items = list(items)
        arr = (abi.ChunkCompressed * max(1, len(items)))()
        keep = []
        for i, (cx, cz, vox, lights) in enumerate(items):
            vox = np.frombuffer(bytes(vox), dtype=np.uint8)
            keep.append(vox)
            arr[i].cx, arr[i].cz = cx, cz
            arr[i].voxels, arr[i].voxels_len = vox.ctypes.data, vox.size
            if lights is not None:
                lights = np.frombuffer(bytes(lights), dtype=np.uint8)
                keep.append(lights)
                arr[i].lights, arr[i].lights_len = lights.ctypes.data, lights.size
        self._check(self.lib.vx_chunks_put_compressed(self.h, arr, len(items)))

    def compress_chunks(self, keys, layer, capacity=None):
        """Region-file blobs (bytes) of one layer for a batch of chunks."""
        n = len(keys)
        offsets = (C.c_uint64 * (n + 1))()
        if capacity is None:
            capacity = n * (2 * abi.CHUNK_DATA_LEN if layer == abi.LAYER_VOXELS else 2 * abi.LIGHTMAP_DATA_LEN)
        out = np.empty(max(1, capacity), dtype=np.uint8)
        self._check(self.lib.vx_chunks_compress(self.h, keys_array(keys), n, layer, out.ctypes.data,
                                                capacity, offsets))
        return [out[offsets[i]:offsets[i + 1]].tobytes() for i in range(n)]

    def remove_chunks(self, keys):
        self._check(self.lib.vx_chunks_remove(self.h, keys_array(keys), len(keys)))

    def get_info(self, cx, cz):
        info = abi.ChunkInfo()
        self._check(self.lib.vx_chunk_get_info(self.h, cx, cz, C.byref(info)))
        return info

    def get_voxels(self, cx, cz):
        out = np.empty(abi.CHUNK_VOL, dtype=np.uint32)
        self._check(self.lib.vx_chunk_get_voxels(self.h, cx, cz, out.ctypes.data))
        return out

    def clear_modified(self):
        self._check(self.lib.vx_chunks_clear_modified(self.h))

    # ---- Lighting
    def prebuild_batch(self, keys):
        self._check(self.lib.vx_light_prebuild(self.h, keys_array(keys), len(keys)))

    def prebuild_sky(self, cx, cz):
        self.prebuild_batch([(cx, cz)])

    def light_build(self, keys, expand=True):
        self._check(self.lib.vx_light_build(self.h, keys_array(keys), len(keys), int(expand)))

    def light_chunk(self, cx, cz, expand=True):
        self.light_build([(cx, cz)], expand)

    def light_clear(self):
        self._check(self.lib.vx_light_clear(self.h))

    def edit_batch(self, edits):
        arr = (abi.Edit * max(1, len(edits)))()
        for i, (x, y, z, bid, st) in enumerate(edits):
            arr[i].x, arr[i].y, arr[i].z, arr[i].id, arr[i].state = x, y, z, bid, st
        self._check(self.lib.vx_light_edit(self.h, arr, len(edits)))

    def set_block(self, x, y, z, bid, state=0):
        self.edit_batch([(x, y, z, bid, state)])

    def get_light(self, cx, cz):
        out = np.empty(abi.CHUNK_VOL, dtype=np.uint16)
        self._check(self.lib.vx_light_get(self.h, cx, cz, out.ctypes.data))
        return out

    # ---- BlocksRenderer
    def mesh_build(self, keys, capacity=800000):
        self._check(self.lib.vx_mesh_build(self.h, keys_array(keys), len(keys), capacity))

    def mesh_info(self, i):
        info = abi.MeshInfo()
        self._check(self.lib.vx_mesh_get_info(self.h, i, C.byref(info)))
        return info

    def mesh_read(self, i):
        info = self.mesh_info(i)
        if info.cancelled:
            return Mesh(True, np.zeros(0, np.uint32), np.zeros(0, np.int32),
                        np.zeros(0, ENTRY_DTYPE), np.zeros(0, np.uint32))
        v = np.empty(info.n_vertex_floats, dtype=np.uint32)
        ix = np.empty(info.n_indices, dtype=np.int32)
        e = np.zeros(info.n_entries, dtype=ENTRY_DTYPE)
        f = np.empty(info.n_entry_floats, dtype=np.uint32)
        self._check(self.lib.vx_mesh_read(self.h, i, v.ctypes.data, ix.ctypes.data,
                                          e.ctypes.data, f.ctypes.data))
        return Mesh(False, v, ix, e, f)

    def mesh_batch_read(self, keys, capacity=800000):
        self.mesh_build(keys, capacity)
        return [self.mesh_read(i) for i in range(len(keys))]

    def mesh_chunk(self, cx, cz, capacity=800000):
        return self.mesh_batch_read([(cx, cz)], capacity)[0]

    def put_chunks_raw(self, chunk_in_array, n):
        """vx_chunks_put on a prebuilt abi.ChunkIn array (pinned buffers)."""
        self._check(self.lib.vx_chunks_put(self.h, chunk_in_array, n))

    def get_light_batch(self, keys, out=None):
        """n consecutive lightmaps into `out` (numpy u16 or a raw address)."""
        if out is None:
            out = np.empty(len(keys) * abi.CHUNK_VOL, dtype=np.uint16)
        addr = out.ctypes.data if isinstance(out, np.ndarray) else out
        self._check(self.lib.vx_light_get_batch(self.h, keys_array(keys), len(keys), addr))
        return out

    def mesh_arena_sizes(self):
        sz = (C.c_uint64 * 4)()
        self._check(self.lib.vx_mesh_arena_sizes(self.h, sz))
        return list(sz)

    def mesh_read_arena(self, v, i, e, f):
        """Raw addresses (or None) of host buffers sized by mesh_arena_sizes()."""
        self._check(self.lib.vx_mesh_read_arena(self.h, v, i, e, f))

    def mesh_slice(self, i):
        sl = abi.MeshSlice()
        self._check(self.lib.vx_mesh_get_slice(self.h, i, C.byref(sl)))
        return sl

    def window_update(self, padding=0, capacity=800000):
        """One frame of ChunksController::loadVisible + ChunksRenderer::getOrRender for
        the whole window; returns (n_lit, [(key, Mesh), ...])."""
        st = abi.UpdateStats()
        self._check(self.lib.vx_window_update(self.h, padding, capacity, C.byref(st)))
        meshes = []
        for i in range(st.n_meshed):
            info = self.mesh_info(i)
            meshes.append(((info.cx, info.cz), self.mesh_read(i)))
        return st.n_lit, meshes

    def timer_begin(self):
        self._check(self.lib.vx_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._check(self.lib.vx_timer_end(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._check(self.lib.vx_profile_enable(self.h, int(on)))

    def profile(self):
        """{kernel name: (total ms, launches)} since profile_enable(True)."""
        out = {}
        n = self.lib.vx_profile_get(self.h, -1, None, 0, None, None)
        for k in range(n):
            name = C.create_string_buffer(64)
            ms, cnt = C.c_float(0), C.c_int64(0)
            self.lib.vx_profile_get(self.h, k, name, 64, C.byref(ms), C.byref(cnt))
            if cnt.value:
                out[name.value.decode()] = (ms.value, cnt.value)
        return out

    def edit_stats(self):
        """(device ms, waves, launches) of the last edit_batch."""
        ms, w, n = C.c_float(0), C.c_int32(0), C.c_int32(0)
        self.lib.vx_last_edit_stats(self.h, C.byref(ms), C.byref(w), C.byref(n))
        return ms.value, w.value, n.value

    def edit_global(self):
        """Edits of the last edit_batch that left their shared-memory box and ran on the global lightmaps."""
        return int(self.lib.vx_last_edit_global(self.h))

    def timing(self):
        a, b, n = C.c_float(0), C.c_float(0), C.c_int64(0)
        self.lib.vx_last_timing(self.h, C.byref(a), C.byref(b), C.byref(n))
        return a.value, b.value, n.value
