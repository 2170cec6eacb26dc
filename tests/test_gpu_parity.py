"""GPU parity (run with -m gpu on a B200): every case of tests/cases.py goes
through the C ABI (libvxgpu.so via vxgpu.gpu.GpuWorld, batched light_build /
mesh_build) and must equal the CPU oracle bit for bit — lightmaps, chunk info,
Chunk::flags.modified, vertex words, indices, translucent entries — and the
committed golden digests produced by the reference's own translation units."""
import json
import os

import numpy as np
import pytest

import cases
import pyoracle

pytestmark = pytest.mark.gpu

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
NAMES = sorted(cases.named_cases().keys())


def _gpu_factory(t, ox, oz, w, d, s):
    from vxgpu import gpu
    return gpu.GpuWorld(t, ox, oz, w, d, s)


def _oracle_factory(t, ox, oz, w, d, s):
    return pyoracle.OracleWorld(t, ox, oz, w, d, pyoracle.best_kind(), s)


def _describe(label, a, b):
    if a.shape != b.shape:
        return "%s: size %d vs %d" % (label, a.size, b.size)
    bad = np.flatnonzero(a != b)
    head = ", ".join("%d: got %x want %x" % (i, int(a[i]), int(b[i])) for i in bad[:6])
    return "%s: %d of %d differ (%s)" % (label, bad.size, a.size, head)


@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_oracle_and_golden(name):
    case = cases.named_cases()[name]()
    got = cases.run_case(_gpu_factory, case, batch_light=True)
    want = cases.run_case(_oracle_factory, case)
    assert sorted(got.keys()) == sorted(want.keys())
    problems = []
    for k in sorted(want.keys()):
        a = np.ascontiguousarray(got[k]).view(np.uint8).reshape(-1)
        b = np.ascontiguousarray(want[k]).view(np.uint8).reshape(-1)
        if a.size != b.size or not np.array_equal(a, b):
            problems.append(_describe(k, np.asarray(got[k]).reshape(-1), np.asarray(want[k]).reshape(-1)))
    assert not problems, "%s: %s" % (name, "; ".join(problems[:10]))
    dig = cases.digest(got)
    bad = [k for k in GOLDEN[name] if dig[k] != GOLDEN[name][k]]
    assert not bad, "%s differs from the reference goldens in %s" % (name, bad[:8])


def test_gpu_library_is_the_one_loaded():
    """The product path is libvxgpu.so, not an oracle or a fallback."""
    from vxgpu import gpu
    lib = gpu.load()
    assert os.path.basename(gpu.LIB_PATH) == "libvxgpu.so"
    with open("/proc/self/maps") as f:
        assert "libvxgpu.so" in f.read()
    assert lib.vx_ctx_create is not None


def test_split_translucent_emission_matches_goldens():
    """vx_mesh_build lists translucent cube faces apart (own list, own staged emission pass) when they
    were a large share of the previous build's output, and with the slow kinds otherwise; the policy
    must not change a bit.  VXGPU_SPLIT_TRANSL=1 forces the split pass from the first build on
    (the switch is read in the library, so this runs in a process of its own); every case is then
    checked against the goldens of the reference's own translation units."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys; sys.path[:0] = [%r, %r, %r]\n"
        "import test_gpu_parity as t\n"
        "for name in t.NAMES:\n"
        "    t.test_gpu_matches_oracle_and_golden(name)\n"
        "print('split ok', len(t.NAMES))\n"
    ) % (here, os.path.join(here, "..", "oracle"), os.path.join(here, "..", "voxelengine-cpp_b200"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, VXGPU_SPLIT_TRANSL="1"),
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "split ok" in out.stdout, (out.stdout[-500:], out.stderr[-2000:])
