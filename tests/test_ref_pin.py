"""Pins the restatement to the reference's OWN object code where that can be built without third-party libraries
(oracle/ref_build.mk compiles /root/reference/src/datastructure/Interval.cpp and src/util/Parse2dCsvFile.cpp verbatim into
oracle/_ref/libhrm_ref.so; Eigen is only stubbed for the typedefs of DataType.h).  The interval algebra is what turns sweep
intervals into free segments (K3: FreeSpace-inl.h:101-129 -> Interval::unions / intersects / complements), the CSV reader is
what every scene file goes through."""
import os

import numpy as np
import pytest

from oracle import hrmo

pytestmark = pytest.mark.skipif(not os.path.exists(hrmo.ref_path()), reason="oracle/_ref not built (needs /root/reference at build time)")


def random_intervals(rng, n, kind):
    if kind == "grid":  # many ties, touching and nested intervals, zero-length ones
        s = rng.integers(-6, 6, size=n).astype(float)
        e = s + rng.integers(0, 5, size=n)
    elif kind == "float":
        s = rng.uniform(-30, 30, size=n)
        e = s + rng.exponential(4.0, size=n)
    else:  # unordered end points too (the reference never checks s <= e)
        s = rng.normal(size=n) * 10
        e = rng.normal(size=n) * 10
    return np.stack([s, e], axis=1)


@pytest.mark.parametrize("kind", ["grid", "float", "wild"])
def test_interval_algebra_equals_reference_object_code(kind):
    rng = np.random.default_rng({"grid": 1, "float": 2, "wild": 3}[kind])
    for trial in range(400):
        a = random_intervals(rng, int(rng.integers(0, 40)), kind)
        b = random_intervals(rng, int(rng.integers(0, 40)), kind)
        for op in ("unions", "intersects"):
            assert np.array_equal(hrmo.interval_op(op, a), hrmo.interval_op(op, a, impl="reference")), (op, trial)
        outer = a[:1] if kind != "wild" else a[: int(rng.integers(0, 3))]
        assert np.array_equal(hrmo.interval_op("complements", outer, b), hrmo.interval_op("complements", outer, b, impl="reference")), trial


def test_free_segment_chain_equals_reference_object_code():
    """computeSweepLineFreeSegment (FreeSpace-inl.h:101-129) = complements(intersects(arena), unions(obstacles)), formed from the
    reference's object code, against the oracle's free_segments() before enhancement (one line, so nothing is enhanced)."""
    rng = np.random.default_rng(7)
    for trial in range(200):
        So = int(rng.integers(0, 30))
        obs = random_intervals(rng, So, "float")
        nan = rng.random(So) < 0.3
        obs[nan] = np.nan
        arena = np.array([[-25.0, 25.0]])
        ref = hrmo.interval_op("complements", hrmo.interval_op("intersects", arena, impl="reference"),
                               hrmo.interval_op("unions", obs[~nan], impl="reference"), impl="reference")
        off, xL, xU, xM = hrmo.free_segments(arena[:, :1].T.reshape(1, 1), arena[:, 1:].T.reshape(1, 1), obs[:, 0].reshape(1, So), obs[:, 1].reshape(1, So))
        assert off[-1] == len(ref) and np.array_equal(xL, ref[:, 0]) and np.array_equal(xU, ref[:, 1])
        assert np.array_equal(xM, (ref[:, 0] + ref[:, 1]) / 2.0)


def test_csv_reader_equals_reference_object_code(tmp_path):
    """The product's host reader (hrm_b200/host/hrm_io.hpp: parse2DCsvFile) against the reference's, on files with comments, empty
    lines, scientific notation, more digits than a float holds, trailing commas and a non-numeric cell."""
    from hrm_b200 import hostlib

    rng = np.random.default_rng(11)
    lines = ["# a comment line", "1,2,3", "", "0.1,0.2000000000001,3e-3,-4.5E+2", "7.123456789012345,-0.000001,1e10,", "abc,5,6", "8.5"]
    for _ in range(50):
        lines.append(",".join(repr(float(v)) for v in rng.normal(size=int(rng.integers(1, 12))) * 10.0 ** int(rng.integers(-3, 4))))
    path = tmp_path / "t.csv"
    path.write_text("\n".join(lines) + "\n")
    ref = hrmo.ref_parse_csv(str(path))
    got = hostlib.parse_csv(str(path), stride=64)
    assert len(ref) == len(got) and len(ref) >= 55
    for a, b in zip(ref, got):
        assert np.array_equal(np.asarray(a), np.asarray(b))
