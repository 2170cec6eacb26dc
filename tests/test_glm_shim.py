"""Pins the GLM stand-in under oracle/ref_build/glm (the only arithmetic beneath the reference
TUs of oracle/_ref that the reference itself does not supply; GLM is un-vendored and un-pinned
there, vcpkg.json:9).  SURVEY §8c lists the probe values: the six sun factors
`0.7 + dot(normalize(axis), SUN) * 0.3` of BlocksRenderer.cpp:131-132 as bit patterns, and
normalize() of scaled unit axes (exactly 1 for 0.2 / 0.125 / 0.5 / 1 / 2, 0.99999994 for 0.85).
cross / dot / normalize of a general vector are checked against numpy float32 following GLM
0.9.9's scalar definitions operation by operation.
"""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def _bits(x):
    return "%08x" % np.array([x], np.float32).view(np.uint32)[0]


def _probe(tmp_path):
    exe = str(tmp_path / "glm_probe")
    subprocess.check_call(["g++", "-std=c++17", "-O3", "-DNDEBUG", "-ffp-contract=off",
                           "-I" + os.path.join(ROOT, "oracle", "ref_build"),
                           os.path.join(ROOT, "tests", "native", "glm_probe.cpp"), "-o", exe])
    out = {}
    for line in subprocess.check_output([exe], text=True).splitlines():
        k, *v = line.split()
        out.setdefault(k, []).append(v)
    return out


def test_glm_shim_probe_values(tmp_path):
    got = _probe(tmp_path)
    # +X, -X, +Y, -Y, +Z, -Z
    assert [v[0] for v in got["sun"]] == ["3f44ac08", "3f21ba5e", "3f7b4cc2", "3ed63348", "3f2b7b4a", "3f3aeb1c"]
    # normalize(axis * s) for s = 0.2, 0.125, 0.5, 1, 2 is exactly 1; 0.85 gives 0.99999994
    assert [v[0] for v in got["norm"]] == ["3f800000"] * 5 + ["3f7fffff"]
    assert f32(np.array([0x3f7fffff], np.uint32).view(np.float32)[0]) == f32(0.99999994)
    assert got["trunc"] == [["1", "-1", "0"]]  # implicit vec3 -> ivec3 truncates towards zero


def test_glm_shim_matches_scalar_definitions(tmp_path):
    got = _probe(tmp_path)
    a, b, c = f32(0.3), f32(0.7), f32(-0.2)
    x, y = (a, b, c), (c, a, b)
    # cross(x, y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
    cr = (f32(f32(x[1] * y[2]) - f32(y[1] * x[2])), f32(f32(x[2] * y[0]) - f32(y[2] * x[0])),
          f32(f32(x[0] * y[1]) - f32(y[0] * x[1])))
    assert got["cross"] == [[_bits(v) for v in cr]]
    # dot(a, b) = a.x*b.x + a.y*b.y + a.z*b.z, left to right
    dot = f32(f32(f32(x[0] * y[0]) + f32(x[1] * y[1])) + f32(x[2] * y[2]))
    assert got["dot"] == [[_bits(dot)]]
    # normalize(v) = v * (1 / sqrt(dot(v, v)))
    d = f32(f32(f32(x[0] * x[0]) + f32(x[1] * x[1])) + f32(x[2] * x[2]))
    inv = f32(f32(1.0) / f32(np.sqrt(d)))
    assert got["normv"] == [[_bits(f32(v * inv)) for v in x]]
