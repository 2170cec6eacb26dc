"""Block table of the reference's base content pack, flattened to vx_block_def.

Values restate res/content/base/blocks/*.json (ids in res/content/base/
content.json order after the three core blocks of src/core_defs.cpp:13-66),
with defaults from src/voxels/Block.hpp:150-210 and the runtime derivations of
src/content/ContentBuilder.cpp:29-37.  The UV table is a synthetic 16x16 atlas
grid (the real atlas needs libpng).  Block 29 is a synthetic
custom-model block (two boxes, built the way model::Mesh::addBox does,
src/graphics/commons/Model.cpp:42-50) so that BlockModel::custom is covered.
"""
import ctypes as C

import numpy as np

from . import abi

AIR, OBSTACLE, STRUCT_AIR, DIRT, GRASS_BLOCK, LAMP, GLASS, PLANKS, WOOD, LEAVES, \
    STONE, WATER, SAND, BAZALT, GRASS, FLOWER, BRICK, METAL, RUST, RED_LAMP, \
    GREEN_LAMP, BLUE_LAMP, PANE, PIPE, LIGHTBULB, TORCH, WOODEN_DOOR, COAL_ORE, \
    ICE, CUSTOM = range(30)

NAMES = [
    "core:air", "core:obstacle", "core:struct_air", "base:dirt",
    "base:grass_block", "base:lamp", "base:glass", "base:planks", "base:wood",
    "base:leaves", "base:stone", "base:water", "base:sand", "base:bazalt",
    "base:grass", "base:flower", "base:brick", "base:metal", "base:rust",
    "base:red_lamp", "base:green_lamp", "base:blue_lamp", "base:pane",
    "base:pipe", "base:lightbulb", "base:torch", "base:wooden_door",
    "base:coal_ore", "base:ice", "vx:custom",
]

# variantsCount of each BlockRotProfile (src/voxels/Block.cpp:57-92)
ROT_VARIANTS = {abi.ROT_NONE: 1, abi.ROT_PIPE: 6, abi.ROT_PANE: 4}


def _f32(x):
    return np.float32(x)


def _hitbox(x, y, z, w, h, d):
    # BlockLoader.cpp:150-160: a = vec3(x,y,z); b = vec3(w,h,d); b += a
    a = np.array([x, y, z], dtype=np.float32)
    b = np.array([w, h, d], dtype=np.float32) + a
    return a, b


def _defaults():
    return dict(
        emission=(0, 0, 0, 0), draw_group=0, model=abi.MODEL_BLOCK,
        culling=abi.CULLING_DEFAULT, rotation=abi.ROT_NONE, lp=0, slp=0,
        shadeless=0, ao=1, translucent=0,
        hitbox=_hitbox(0, 0, 0, 1, 1, 1), meshes=(0, 0),
    )


def base_pack_defs():
    """Python dicts for ids 0..29."""
    t = [_defaults() for _ in range(30)]
    t[AIR].update(draw_group=1, lp=1, slp=1, model=abi.MODEL_NONE)
    t[STRUCT_AIR].update(draw_group=255, lp=1, slp=1)
    t[LAMP].update(emission=(15, 14, 13, 0), shadeless=1)
    t[GLASS].update(draw_group=2, lp=1, slp=1, translucent=1)
    t[WOOD].update(rotation=abi.ROT_PIPE)
    t[LEAVES].update(draw_group=5, culling=abi.CULLING_OPTIONAL)
    t[WATER].update(draw_group=3, lp=1, slp=0, translucent=1)
    t[GRASS].update(draw_group=1, lp=1, model=abi.MODEL_XSPRITE,
                    hitbox=_hitbox(0.15, 0.0, 0.15, 0.7, 0.5, 0.7))
    t[FLOWER].update(draw_group=1, lp=1, model=abi.MODEL_XSPRITE,
                     hitbox=_hitbox(0.15, 0.0, 0.15, 0.7, 0.7, 0.7))
    t[RED_LAMP].update(emission=(15, 0, 0, 0), shadeless=1)
    t[GREEN_LAMP].update(emission=(0, 15, 0, 0), shadeless=1)
    t[BLUE_LAMP].update(emission=(0, 0, 15, 0), shadeless=1)
    t[PANE].update(model=abi.MODEL_AABB, lp=1, slp=1, rotation=abi.ROT_PANE,
                   hitbox=_hitbox(0.0, 0.0, 0.0, 1.0, 1.0, 0.2))
    t[PIPE].update(model=abi.MODEL_AABB, lp=1, rotation=abi.ROT_PIPE,
                   hitbox=_hitbox(0.25, 0.0, 0.25, 0.5, 1.0, 0.5))
    t[LIGHTBULB].update(emission=(15, 14, 13, 0), model=abi.MODEL_AABB, lp=1,
                        slp=1, shadeless=1, rotation=abi.ROT_PIPE,
                        hitbox=_hitbox(0.25, 0.0, 0.25, 0.5, 0.5, 0.5))
    t[TORCH].update(emission=(13, 13, 12, 0), model=abi.MODEL_AABB, lp=1,
                    shadeless=1, rotation=abi.ROT_PIPE,
                    hitbox=_hitbox(0.4375, 0.0, 0.4375, 0.125, 0.5, 0.125))
    t[WOODEN_DOOR].update(model=abi.MODEL_AABB, lp=1, slp=1, ao=0,
                          rotation=abi.ROT_PANE,
                          hitbox=_hitbox(0.0, 0.0, 0.8, 1.0, 2.0, 0.2))
    t[ICE].update(draw_group=4, lp=1, translucent=1)
    t[CUSTOM].update(model=abi.MODEL_CUSTOM, lp=1, rotation=abi.ROT_PIPE,
                     meshes=(0, 2))
    return t


def _add_plane(out, pos, right, up, norm, uv):
    # model::Mesh::addPlane with a UVRegion (Model.cpp:27-40)
    u1, v1, u2, v2 = uv
    p = lambda a: tuple(np.asarray(a, dtype=np.float32))
    out.append((p(pos - right - up), (u1, v1), p(norm)))
    out.append((p(pos + right - up), (u2, v1), p(norm)))
    out.append((p(pos + right + up), (u2, v2), p(norm)))
    out.append((p(pos - right - up), (u1, v1), p(norm)))
    out.append((p(pos + right + up), (u2, v2), p(norm)))
    out.append((p(pos - right + up), (u1, v2), p(norm)))


def _add_box(out, pos, size, uv):
    # model::Mesh::addBox (Model.cpp:52-63)
    pos = np.asarray(pos, dtype=np.float32)
    size = np.asarray(size, dtype=np.float32)
    X = np.array([1, 0, 0], dtype=np.float32)
    Y = np.array([0, 1, 0], dtype=np.float32)
    Z = np.array([0, 0, 1], dtype=np.float32)
    _add_plane(out, pos + Z * size, X * size, Y * size, Z, uv)
    _add_plane(out, pos - Z * size, -X * size, Y * size, -Z, uv)
    _add_plane(out, pos + Y * size, X * size, -Z * size, Y, uv)
    _add_plane(out, pos - Y * size, X * size, Z * size, -Y, uv)
    _add_plane(out, pos + X * size, -Z * size, Y * size, X, uv)
    _add_plane(out, pos - X * size, Z * size, Y * size, -X, uv)


def synthetic_uv(n_defs):
    k = np.arange(n_defs * 6)
    u1 = ((k % 16).astype(np.float32)) / np.float32(16)
    v1 = ((k // 16 % 16).astype(np.float32)) / np.float32(16)
    uv = np.stack([u1, v1, u1 + np.float32(1 / 16), v1 + np.float32(1 / 16)], axis=1)
    return np.ascontiguousarray(uv, dtype=np.float32)


class ContentTable:
    """Owns the ctypes arrays behind a vx_content."""

    def __init__(self, defs=None):
        defs = defs if defs is not None else base_pack_defs()
        n = len(defs)
        self.py_defs = defs
        self.uv = synthetic_uv(n)
        # custom model: two boxes in two meshes (a slab and a post on top)
        verts0, verts1 = [], []
        uvc = tuple(self.uv[CUSTOM * 6 + 3]) if n > CUSTOM else (0, 0, 1, 1)
        _add_box(verts0, (0.5, 0.25, 0.5), (0.5, 0.25, 0.5), uvc)
        _add_box(verts1, (0.5, 0.75, 0.5), (0.2, 0.25, 0.3), uvc)
        all_verts = verts0 + verts1
        self.mesh_arr = (abi.ModelMesh * 2)()
        self.mesh_arr[0].vertex_begin = 0
        self.mesh_arr[0].vertex_count = len(verts0)
        self.mesh_arr[1].vertex_begin = len(verts0)
        self.mesh_arr[1].vertex_count = len(verts1)
        self.vert_arr = (abi.ModelVertex * len(all_verts))()
        for i, (c, uv, nrm) in enumerate(all_verts):
            for k in range(3):
                self.vert_arr[i].coord[k] = float(c[k])
                self.vert_arr[i].normal[k] = float(nrm[k])
            self.vert_arr[i].uv[0] = float(uv[0])
            self.vert_arr[i].uv[1] = float(uv[1])
        self.def_arr = (abi.BlockDef * n)()
        for i, d in enumerate(defs):
            s = self.def_arr[i]
            for c in range(4):
                s.emission[c] = d["emission"][c]
            s.draw_group = d["draw_group"]
            s.model = d["model"]
            s.culling = d["culling"]
            s.rotation_profile = d["rotation"]
            s.light_passing = d["lp"]
            s.sky_light_passing = d["slp"]
            s.shadeless = d["shadeless"]
            s.ambient_occlusion = d["ao"]
            s.translucent = d["translucent"]
            s.solid = 1 if d["model"] == abi.MODEL_BLOCK else 0
            s.has_hitbox = 1
            a, b = d["hitbox"]
            for k in range(3):
                s.hitbox_a[k] = float(a[k])
                s.hitbox_b[k] = float(b[k])
            s.model_mesh_begin, s.model_mesh_count = d["meshes"]
        self.struct = abi.Content()
        self.struct.defs = C.cast(self.def_arr, C.POINTER(abi.BlockDef))
        self.struct.n_defs = n
        self.struct.uv_regions = self.uv.ctypes.data_as(C.POINTER(C.c_float))
        self.struct.meshes = C.cast(self.mesh_arr, C.POINTER(abi.ModelMesh))
        self.struct.n_meshes = 2
        self.struct.vertices = C.cast(self.vert_arr, C.POINTER(abi.ModelVertex))
        self.struct.n_vertices = len(all_verts)

    @property
    def n_defs(self):
        return len(self.py_defs)

    def flag_array(self, key):
        return np.array([d[key] for d in self.py_defs], dtype=np.uint8)

    def rot_variants(self):
        return np.array([ROT_VARIANTS[d["rotation"]] for d in self.py_defs],
                        dtype=np.int32)


def default_settings(backlight=True, dense_render=True):
    s = abi.Settings()
    s.backlight = int(backlight)
    s.dense_render = int(dense_render)
    return s
