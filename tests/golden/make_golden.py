"""Generates tests/golden/golden.json from oracle/_ref/libvxref.so, i.e. from
the reference's own translation units compiled in this container
(oracle/ref_build/Makefile).  Run here, where /root/reference exists:

    python tests/golden/make_golden.py

The committed JSON holds SHA-256 digests + element counts of every output of
every case in tests/cases.py (lightmaps, chunk info, vertex/index buffers,
translucent entries), so the oracle port and the CUDA path can be pinned on
the GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "voxelengine-cpp_b200"), os.path.join(ROOT, "oracle"),
          os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
import pyoracle  # noqa: E402


def main():
    assert pyoracle.available("reference"), "build oracle/_ref first (make -C oracle reference)"
    out = {"_generator": "oracle/_ref/libvxref.so (reference TUs, g++ -O3 -DNDEBUG)"}
    for name, factory in cases.named_cases().items():
        case = factory()
        res = cases.run_case(
            lambda t, ox, oz, w, d, s: pyoracle.OracleWorld(t, ox, oz, w, d, "reference", s),
            case)
        out[name] = cases.digest(res)
        print(name, len(res), "outputs")
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
