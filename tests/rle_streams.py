"""Seeded byte streams for the extrle tests (tests/test_rle.py) and their golden
digests (tests/golden/make_golden_rle.py).  Each entry: name -> (raw bytes, wide)."""
import numpy as np


def _runs(rng, n_elems, max_run, values):
    out = []
    total = 0
    while total < n_elems:
        r = int(rng.integers(1, max_run + 1))
        out.append(np.full(r, values[int(rng.integers(0, len(values)))], dtype=np.uint16))
        total += r
    return np.concatenate(out)[:n_elems]


def streams():
    rng = np.random.default_rng(2024)
    s = {}
    for wide in (0, 1):
        dt = np.uint16 if wide else np.uint8
        top = 65536 if wide else 256
        tag = "w16" if wide else "w8"
        mx = 0x4000 if wide else 0x8000   # max_sequence + 1
        s[tag + "_single"] = (np.array([7], dtype=dt), wide)
        s[tag + "_zeros"] = (np.zeros(131072 if wide else 32768, dtype=dt), wide)
        s[tag + "_noise"] = (rng.integers(0, top, 50000).astype(dt), wide)
        s[tag + "_short_runs"] = (_runs(rng, 70000, 6, list(range(0, 600, 7))).astype(dt), wide)
        s[tag + "_mixed_runs"] = (_runs(rng, 131072, 900, [0, 1, 3, 255, 256, 257, 4000, 65535]).astype(dt), wide)
        # run lengths around every header boundary: 63/64/65 (16), 127/128/129 (8), max_sequence +- 1, 2x
        edges = []
        v = 1
        for r in (1, 2, 63, 64, 65, 127, 128, 129, mx - 1, mx, mx + 1, 2 * mx, 2 * mx + 1, 3):
            edges.append(np.full(r, v if not wide else (v * 100) % 65536, dtype=np.uint16))
            v += 1
        s[tag + "_edges"] = (np.concatenate(edges).astype(dt), wide)
        # the same value continuing after a max-length split, then a change on the split boundary
        s[tag + "_split"] = (np.concatenate([np.full(mx, 9), np.full(mx, 9), np.full(5, 10), np.full(mx, 10),
                                             np.full(1, 300)]).astype(dt), wide)
    return {k: (np.ascontiguousarray(a).tobytes(), w) for k, (a, w) in s.items()}
