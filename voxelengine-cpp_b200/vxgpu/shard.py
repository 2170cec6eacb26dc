"""Sharding of a world across GPUs by chunk-column ranges.

A world is the rectangle of chunks [x0, x0+w) x [z0, z0+d); the chunks the
reference would light are the ones with all 8 neighbours present, i.e. the
rectangle shrunk by one (logic/ChunksController.cpp:106-128).  Rank r owns a
contiguous range of chunk columns (x) of that lit rectangle and keeps a
one-chunk halo ring around it.  Light travels at most 14 voxels from a seed
(max level 15, -1 per step, lighting/LightSolver.cpp:116-121), i.e. never more
than one chunk, and a chunk's seeds depend only on its own voxels and on the
y* rows of its side neighbours' border columns; so a shard that lights its
halo chunks whenever the global run lights them reproduces the neighbour
shard's influence on every owned chunk bit for bit, with no exchange at all.
The halo chunks themselves are NOT final on this rank (their far side is
missing) and are never reported.  Meshing needs final light of the halo only
within 2 voxels of the border, which the same argument covers.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

Key = Tuple[int, int]


@dataclass
class Shard:
    rank: int
    window: Tuple[int, int, int, int]       # ox, oz, w, d of the rank's Chunks window
    resident: List[Key] = field(default_factory=list)  # chunks to upload (owned + halo)
    owned: List[Key] = field(default_factory=list)     # chunks whose results this rank reports
    lit: List[Key] = field(default_factory=list)       # chunks to pass to vx_light_build


def column_ranges(n_columns: int, n_ranks: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) column ranges, sizes differing by at most one."""
    base, extra = divmod(n_columns, n_ranks)
    out, at = [], 0
    for r in range(n_ranks):
        size = base + (1 if r < extra else 0)
        out.append((at, at + size))
        at += size
    return out


def plan(x0: int, z0: int, w: int, d: int, n_ranks: int) -> List[Shard]:
    """Shards of the world rectangle [x0,x0+w) x [z0,z0+d) (all chunks present)."""
    if w < 3 or d < 3:
        raise ValueError("a world needs at least 3x3 chunks to have a lit chunk")
    lx0, lz0, lw, ld = x0 + 1, z0 + 1, w - 2, d - 2      # the lit rectangle
    if n_ranks > lw:
        raise ValueError("more ranks than lit chunk columns")
    shards = []
    for r, (a, b) in enumerate(column_ranges(lw, n_ranks)):
        ox, oz = lx0 + a - 1, z0
        sw, sd = (b - a) + 2, d
        sh = Shard(r, (ox, oz, sw, sd))
        for cz in range(oz, oz + sd):
            for cx in range(ox, ox + sw):
                sh.resident.append((cx, cz))
                in_lit = lx0 <= cx < lx0 + lw and lz0 <= cz < lz0 + ld
                if in_lit:
                    sh.lit.append((cx, cz))
                    if lx0 + a <= cx < lx0 + b:
                        sh.owned.append((cx, cz))
        shards.append(sh)
    return shards
