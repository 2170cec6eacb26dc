"""Generates tests/golden/extrle.json from the reference's own coders/rle.cpp
(compiled into oracle/_ref/libvxref.so): SHA-256 + length of extrle::encode16 /
extrle::encode of every stream in tests/rle_streams.py.  Run here, where
/root/reference exists:  python tests/golden/make_golden_rle.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "voxelengine-cpp_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import pyoracle  # noqa: E402
import rle_streams  # noqa: E402

assert pyoracle.available("reference"), "build oracle/_ref first (make -C oracle reference)"
out = {"_generator": "oracle/_ref/libvxref.so: extrle::encode16 / extrle::encode (coders/rle.cpp)"}
for name, (raw, wide) in sorted(rle_streams.streams().items()):
    enc = pyoracle.extrle_encode(raw, wide, "reference")
    assert pyoracle.extrle_decode(enc, wide, len(raw), "reference") == raw, name
    out[name] = {"len": len(enc), "sha256": hashlib.sha256(enc).hexdigest()}
    print(name, len(raw), "->", len(enc))
with open(os.path.join(HERE, "extrle.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
