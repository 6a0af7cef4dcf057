"""The lattice snap of the grid build (csrc/gridbuild.cu: snap_point), restated in numpy f64 and checked on the CPU for the
claim the kernel relies on: lrint((p - min) / pitch) -- what the oracle computes (oracle/amcl3d_oracle.c: snap) -- equals
rint of the product with the rounded reciprocal unless that product sits within 1e-15 * |q| of a tie, in which case the
kernel takes the real divide.  Random coordinates, coordinates at ties and a few float steps either side of them, for the
lattice pitches the configurations use.  The GPU test test_snap_to_the_lattice_at_and_around_ties checks the kernel itself."""
import numpy as np
import pytest


def snap_reference(p, mn, pitch):
    return np.rint((p.astype(np.float64) - mn) / pitch)


def snap_kernel(p, mn, pitch):
    """-> (node, took_the_divide) with the kernel's rule."""
    t = p.astype(np.float64) - mn
    rinv = 1.0 / pitch
    q = t * rinv
    r = np.rint(q)
    slow = (0.5 - np.abs(q - r)) <= np.abs(q) * 1e-15
    return np.where(slow, np.rint(t / pitch), r), slow


@pytest.mark.parametrize("resol,sub", [(0.1, 1), (0.1, 2), (0.05, 2), (0.05, 1), (0.3, 2), (0.0371, 2)])
def test_reciprocal_snap_equals_the_quotient_snap(resol, sub):
    rng = np.random.RandomState(7)
    pitch = resol / sub
    for mn in (0.0, -12.3400001525878906, 3.0e3):
        k = rng.randint(0, 8000, 400_000).astype(np.float64)
        cases = [
            (mn + rng.uniform(0, 400.0, 400_000)).astype(np.float32),  # anywhere
            (mn + k * pitch).astype(np.float32),                         # on the nodes
            (mn + (k + 0.5) * pitch).astype(np.float32),                 # at the ties (up to the float rounding)
        ]
        tie = cases[2]
        for _ in range(3):  # and up to three float steps either side
            cases.append(np.nextafter(cases[-1], np.float32(np.inf)))
        cases.append(tie)
        for _ in range(3):
            cases.append(np.nextafter(cases[-1], np.float32(-np.inf)))
        for p in cases:
            got, slow = snap_kernel(p, mn, pitch)
            assert np.array_equal(got, snap_reference(p, mn, pitch))
        # the fast path is the rule, not the exception: of random float coordinates only the genuine ties take the divide
        # (e.g. 294.75 / 0.1 = 2947.5; around 3000 a float has 12 fraction bits, so about one coordinate in a thousand is one)
        _, slow = snap_kernel(cases[0], mn, pitch)
        assert slow.mean() < 5e-3
        q = (cases[0].astype(np.float64) - mn) * (1.0 / pitch)
        assert np.all(np.abs(np.abs(q[slow] - np.rint(q[slow])) - 0.5) < 1e-9)
