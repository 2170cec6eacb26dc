"""ctypes mirror of include/vxgpu.h (struct layouts only, no behaviour).

Used by the ctypes bindings of the product library (vxgpu.gpu) and by the test
oracles (oracle/pyoracle.py); keeping one definition guarantees both sides see
the same bytes.
"""
import ctypes as C

CHUNK_W = 16
CHUNK_H = 256
CHUNK_D = 16
CHUNK_VOL = CHUNK_W * CHUNK_H * CHUNK_D
BLOCK_VOID = 0xFFFF
VERTEX_SIZE = 6

MODEL_NONE, MODEL_BLOCK, MODEL_XSPRITE, MODEL_AABB, MODEL_CUSTOM = range(5)
CULLING_DEFAULT, CULLING_OPTIONAL, CULLING_DISABLED = range(3)
ROT_NONE, ROT_PIPE, ROT_PANE = range(3)


class BlockDef(C.Structure):
    _fields_ = [
        ("emission", C.c_uint8 * 4),
        ("draw_group", C.c_uint8),
        ("model", C.c_uint8),
        ("culling", C.c_uint8),
        ("rotation_profile", C.c_uint8),
        ("light_passing", C.c_uint8),
        ("sky_light_passing", C.c_uint8),
        ("shadeless", C.c_uint8),
        ("ambient_occlusion", C.c_uint8),
        ("translucent", C.c_uint8),
        ("solid", C.c_uint8),
        ("has_hitbox", C.c_uint8),
        ("reserved0", C.c_uint8),
        ("hitbox_a", C.c_float * 3),
        ("hitbox_b", C.c_float * 3),
        ("model_mesh_begin", C.c_int32),
        ("model_mesh_count", C.c_int32),
    ]


class ModelVertex(C.Structure):
    _fields_ = [
        ("coord", C.c_float * 3),
        ("uv", C.c_float * 2),
        ("normal", C.c_float * 3),
    ]


class ModelMesh(C.Structure):
    _fields_ = [("vertex_begin", C.c_int32), ("vertex_count", C.c_int32)]


class Content(C.Structure):
    _fields_ = [
        ("defs", C.POINTER(BlockDef)),
        ("n_defs", C.c_int32),
        ("uv_regions", C.POINTER(C.c_float)),
        ("meshes", C.POINTER(ModelMesh)),
        ("n_meshes", C.c_int32),
        ("vertices", C.POINTER(ModelVertex)),
        ("n_vertices", C.c_int32),
    ]


class Settings(C.Structure):
    _fields_ = [("backlight", C.c_int32), ("dense_render", C.c_int32)]


class ChunkKey(C.Structure):
    _fields_ = [("cx", C.c_int32), ("cz", C.c_int32)]


class ChunkIn(C.Structure):
    _fields_ = [
        ("cx", C.c_int32),
        ("cz", C.c_int32),
        ("voxels", C.c_void_p),
        ("light", C.c_void_p),
    ]


class ChunkInfo(C.Structure):
    _fields_ = [
        ("present", C.c_int32),
        ("bottom", C.c_int32),
        ("top", C.c_int32),
        ("highest_point", C.c_int32),
        ("modified", C.c_int32),
        ("lighted", C.c_int32),
    ]


class Edit(C.Structure):
    _fields_ = [
        ("x", C.c_int32),
        ("y", C.c_int32),
        ("z", C.c_int32),
        ("id", C.c_uint16),
        ("state", C.c_uint16),
    ]


class SortEntry(C.Structure):
    _fields_ = [
        ("position", C.c_float * 3),
        ("first_float", C.c_uint32),
        ("n_floats", C.c_uint32),
    ]


class MeshInfo(C.Structure):
    _fields_ = [
        ("cx", C.c_int32),
        ("cz", C.c_int32),
        ("cancelled", C.c_int32),
        ("n_vertex_floats", C.c_uint32),
        ("n_indices", C.c_uint32),
        ("n_entries", C.c_uint32),
        ("n_entry_floats", C.c_uint32),
    ]


class MeshSlice(C.Structure):
    _fields_ = [
        ("vertex_off", C.c_uint64),
        ("index_off", C.c_uint64),
        ("entry_off", C.c_uint64),
        ("entry_float_off", C.c_uint64),
    ]


class ChunkEncoded(C.Structure):
    _fields_ = [
        ("cx", C.c_int32),
        ("cz", C.c_int32),
        ("voxels", C.c_void_p),
        ("lights", C.c_void_p),
    ]


class ChunkCompressed(C.Structure):
    _fields_ = [
        ("cx", C.c_int32),
        ("cz", C.c_int32),
        ("voxels", C.c_void_p),
        ("voxels_len", C.c_uint32),
        ("lights", C.c_void_p),
        ("lights_len", C.c_uint32),
    ]


LAYER_VOXELS = 0
LAYER_LIGHTS = 1


class GenLayer(C.Structure):
    _fields_ = [("block", C.c_uint16), ("height", C.c_int16), ("below_sea_level", C.c_uint8), ("pad", C.c_uint8 * 3)]


class GenPlant(C.Structure):
    _fields_ = [("block", C.c_uint16), ("pad", C.c_uint16), ("weight", C.c_float)]


class GenBiome(C.Structure):
    _fields_ = [
        ("ground_begin", C.c_int32), ("ground_count", C.c_int32),
        ("sea_begin", C.c_int32), ("sea_count", C.c_int32),
        ("ground_last_layers_height", C.c_uint32), ("sea_last_layers_height", C.c_uint32),
        ("plants_begin", C.c_int32), ("plants_count", C.c_int32),
        ("plants_chance", C.c_float),
    ]


class GenDef(C.Structure):
    _fields_ = [
        ("biomes", C.c_void_p), ("n_biomes", C.c_int32),
        ("layers", C.c_void_p), ("n_layers", C.c_int32),
        ("plants", C.c_void_p), ("n_plants", C.c_int32),
        ("sea_level", C.c_uint32),
    ]


class GenChunk(C.Structure):
    _fields_ = [("cx", C.c_int32), ("cz", C.c_int32), ("heights", C.c_void_p), ("biomes", C.c_void_p)]


GEN_MAX_LAYERS = 8


CHUNK_DATA_LEN = CHUNK_VOL * 4
LIGHTMAP_DATA_LEN = CHUNK_VOL // 2


class UpdateStats(C.Structure):
    _fields_ = [("n_lit", C.c_int32), ("n_meshed", C.c_int32)]


LIGHT_EXPAND = 1
LIGHT_NO_SKY = 2
