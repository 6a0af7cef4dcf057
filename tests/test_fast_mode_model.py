"""The placement rule of the fast (non-exact) sharded resample (csrc/shard.cu: shard_push_kernel), restated in numpy f32
and checked on the CPU for the property the kernel relies on: with every rank placing only the thresholds that fall into
its own stretch (base_r, base_{r+1}] of the global prefix -- rank 0 also those at or below 0, the last rank everything above
the total -- each output slot m is written by exactly one rank, whatever the weights (zero-weight ranks, one rank holding
everything, u01 = 0), and the result is a valid systematic resample (non-decreasing source indices).  The GPU tests
(tools/sharded_check.py) check the kernel itself against this rule's observable consequences."""
import numpy as np
import pytest

f32 = np.float32


def shard_bounds(n_total, world, rank):
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def seq_sum(v):
    acc = f32(0)
    for x in v:
        acc = f32(acc + x)
    return acc


def seq_prefix(v):
    out = np.empty(len(v), f32)
    acc = f32(0)
    for k, x in enumerate(v):
        acc = f32(acc + x)
        out[k] = acc
    return out


def threshold(r, factor, m):
    return f32(r + f32(factor * f32(m)))  # ParticleFilter.cpp:241


def first_above(r, factor, n, c):
    lo, hi = 0, n
    while lo < hi:
        mid = (lo + hi) >> 1
        if threshold(r, factor, mid) > c:
            hi = mid
        else:
            lo = mid + 1
    return lo


def fast_mode(weights, world, u01):
    """Returns (source index of every output, how many ranks wrote every output)."""
    n = len(weights)
    blocks = [shard_bounds(n, world, r) for r in range(world)]
    sums = [seq_sum(np.where(weights[lo:hi] > 0, weights[lo:hi], f32(0))) for lo, hi in blocks]
    total = seq_sum(sums)  # rank order
    norm = [(weights[lo:hi] / total if total > 0 else np.zeros(hi - lo, f32)).astype(f32) for lo, hi in blocks]
    prefix = [seq_prefix(w) for w in norm]
    locals_ = [p[-1] if len(p) else f32(0) for p in prefix]
    factor = f32(f32(1) / f32(n))
    r = f32(factor * f32(u01))
    src = np.full(n, -1, np.int64)
    writes = np.zeros(n, np.int64)
    base = f32(0)
    for rank, (lo, hi) in enumerate(blocks):
        n_local = hi - lo
        end = f32(base + locals_[rank]) if n_local else base
        m_a = 0 if rank == 0 else first_above(r, factor, n, base)
        m_b = m_a if n_local == 0 else (n if rank + 1 == world else first_above(r, factor, n, end))
        for m in range(m_a, m_b):
            u = threshold(r, factor, m)
            a, b = 0, n_local
            while a < b:
                mid = (a + b) >> 1
                if u > f32(base + prefix[rank][mid]):
                    a = mid + 1
                else:
                    b = mid
            src[m] = lo + (n_local - 1 if a >= n_local else a)
            writes[m] += 1
        base = f32(base + locals_[rank])  # the same sequential f32 sum every rank forms (combine_scalar_kernel)
    return src, writes


def exact_mode(weights, u01):
    n = len(weights)
    total = seq_sum(np.where(weights > 0, weights, f32(0)))
    norm = (weights / total).astype(f32)
    prefix = seq_prefix(norm)
    factor = f32(f32(1) / f32(n))
    r = f32(factor * f32(u01))
    idx = np.empty(n, np.int64)
    for m in range(n):
        i = int(np.searchsorted(prefix, threshold(r, factor, m), side="left"))  # first i with !(u > c_i)
        idx[m] = min(i, n - 1)
    return idx


CASES = {
    "random": lambda rng, n: rng.uniform(0, 1, n),
    "tracking": lambda rng, n: np.exp(-rng.uniform(0, 12, n)),
    "zero_first_rank": lambda rng, n: np.concatenate([np.zeros(n // 3), rng.uniform(0, 1, n - n // 3)]),
    "zero_last_rank": lambda rng, n: np.concatenate([rng.uniform(0, 1, n - n // 3), np.zeros(n // 3)]),
    "one_heavy": lambda rng, n: np.where(np.arange(n) == n // 2, 1.0, 1e-9),
    "sparse": lambda rng, n: np.where(rng.uniform(0, 1, n) < 0.05, rng.uniform(0, 1, n), 0.0),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("u01", [0.0, 0.37, 0.999])
def test_every_output_is_placed_exactly_once(case, world, u01):
    rng = np.random.RandomState(hash((case, world)) % 2**31)
    n = 997
    w = CASES[case](rng, n).astype(f32)
    if not np.any(w > 0):
        w[n // 2] = f32(1)
    src, writes = fast_mode(w, world, u01)
    assert np.all(writes == 1)
    assert np.all(np.diff(src) >= 0) and src.min() >= 0 and src.max() < n
    if world == 1:
        assert np.array_equal(src, exact_mode(w, u01))  # one rank: the rank-order combination IS the sequential chain
    else:
        # a differing pick is still next to the exact one in prefix space: never farther than a few positions here
        ex = exact_mode(w, u01)
        heavy = w[src] > 0
        assert np.all(heavy | (w[ex] == 0) | (src == ex) | (np.abs(src - ex) <= 64))
