"""CPU-only: pins the oracle.  The golden digests (tests/golden/golden.json)
were produced by oracle/_ref/libvxref.so — the reference's own translation
units compiled here — because the reference has no tests or golden vectors on
the light/mesh path.  The port restatement must reproduce every
digest bit for bit; where oracle/_ref is present it is re-checked too."""
import json
import os

import numpy as np
import pytest

import cases
import pyoracle

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))
NAMES = sorted(cases.named_cases().keys())


def _factory(kind):
    return lambda t, ox, oz, w, d, s: pyoracle.OracleWorld(t, ox, oz, w, d, kind, s)


def _check(name, kind):
    case = cases.named_cases()[name]()
    got = cases.digest(cases.run_case(_factory(kind), case))
    want = GOLDEN[name]
    assert sorted(got.keys()) == sorted(want.keys())
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, "%s/%s differs from golden in %s" % (kind, name, bad[:8])


@pytest.mark.parametrize("name", NAMES)
def test_port_matches_golden(name):
    _check(name, "port")


@pytest.mark.parametrize("name", cases.FAST_CASES)
def test_reference_matches_golden(name):
    if not pyoracle.available("reference"):
        pytest.skip("oracle/_ref/libvxref.so not built (no /root/reference here)")
    _check(name, "reference")


def test_golden_covers_every_case():
    assert sorted(k for k in GOLDEN if not k.startswith("_")) == NAMES


def test_relight_is_idempotent():
    """Property of the oracle itself: lighting a lit chunk again
    changes nothing."""
    case = cases.named_cases()["c1_s5"]()
    w = case.world
    be = pyoracle.OracleWorld(cases.table(), w.ox, w.oz, w.w, w.d, "port")
    be.put_world(w)
    for k in w.keys():
        be.prebuild_sky(*k)
    be.light_chunk(0, 0, True)
    first = {k: be.get_light(*k).copy() for k in w.keys()}
    be.light_chunk(0, 0, True)
    for k in w.keys():
        assert np.array_equal(first[k], be.get_light(*k))
    be.close()


def test_light_order_independent():
    """Full relight is an order-independent fixpoint: the batch
    GPU solver relies on it."""
    case = cases.named_cases()["multi_s21"]()
    w = case.world
    results = []
    for order in (case.light_keys, list(reversed(case.light_keys))):
        be = pyoracle.OracleWorld(cases.table(), w.ox, w.oz, w.w, w.d, "port")
        be.put_world(w)
        for k in w.keys():
            be.prebuild_sky(*k)
        for k in order:
            be.light_chunk(k[0], k[1], True)
        results.append({k: be.get_light(*k) for k in w.keys()})
        be.close()
    for k in w.keys():
        assert np.array_equal(results[0][k], results[1][k])
