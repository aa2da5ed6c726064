"""Synthetic R1CS workloads for bench.py and the full-size tests.

* synthetic_r1cs : non-degenerate circuit -- constraint i is (z_p + k_i) * z_q = z_new with p, q uniform over earlier
                   variables; two uniformly random seed witnesses; dense queries, witness values uniform-looking in
                   [0, r).  Sized so that num_constraints + num_instance_variables == 2^log_n exactly (the sizing
                   trick of benches/bench.rs:19-20).
* dummy_r1cs     : the reference's own DummyCircuit (benches/bench.rs:41-64): constant witness, A/B rows touching
                   two variables only (degenerate: almost every a/b query element is the identity).
Both return (ConstraintMatrices, full_assignment as Montgomery limbs, public inputs as ints).
"""
from __future__ import annotations

import numpy as np

from .api import ConstraintMatrices
from .codec import CurveCodec
from .params import get_curve


def _splitmix(seed: int):
    s = seed & 0xFFFFFFFFFFFFFFFF
    M = 0xFFFFFFFFFFFFFFFF
    while True:
        s = (s + 0x9E3779B97F4A7C15) & M
        z = s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        yield z ^ (z >> 31)


def _fr(gen, r):
    bits = r.bit_length()
    while True:
        v = next(gen) | (next(gen) << 64) | (next(gen) << 128) | (next(gen) << 192)
        v &= (1 << bits) - 1
        if v < r:
            return v


def synthetic_r1cs(curve, log_n: int, seed: int = 0, num_inputs: int = 1):
    c = get_curve(curve)
    cd = CurveCodec(c)
    r = c.r
    ninst = 1 + num_inputs
    nc = (1 << log_n) - ninst
    if nc < num_inputs + 1:
        raise ValueError("domain too small")
    gen = _splitmix(seed)
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    nwit = 2 + nc - num_inputs
    # variable creation order: two seeds, then one product per constraint
    cols = np.empty(2 + nc, dtype=np.uint32)
    cols[0], cols[1] = ninst, ninst + 1
    vals = [0] * (2 + nc)
    vals[0], vals[1] = _fr(gen, r), _fr(gen, r)
    # p, q uniform over the variables that exist when constraint i is written
    u1 = rs.random_sample(nc)
    u2 = rs.random_sample(nc)
    avail = np.arange(2, 2 + nc, dtype=np.float64)
    ps = np.minimum((u1 * avail).astype(np.int64), (avail - 1).astype(np.int64))
    qs = np.minimum((u2 * avail).astype(np.int64), (avail - 1).astype(np.int64))
    ks_lo = rs.randint(0, 1 << 62, size=nc, dtype=np.int64)
    ks_hi = rs.randint(0, 1 << 62, size=nc, dtype=np.int64)
    ks = [0] * nc
    n_w = 2
    n_i = 0
    first_input = nc - num_inputs
    for i in range(nc):
        k = (int(ks_hi[i]) << 62) | int(ks_lo[i])   # 124-bit coefficient; its size is irrelevant to the prover
        ks[i] = k
        v = (vals[ps[i]] + k) * vals[qs[i]] % r
        vals[2 + i] = v
        if i >= first_input:
            cols[2 + i] = 1 + n_i
            n_i += 1
        else:
            cols[2 + i] = ninst + n_w
            n_w += 1
    assert n_w == nwit and n_i == num_inputs
    # CSR: A = [(1, col_p), (k, One)], B = [(1, col_q)], C = [(1, col_new)]
    one = cd.fr.enc1(1)
    a_rp = np.arange(0, 2 * nc + 1, 2, dtype=np.uint32)
    a_col = np.empty(2 * nc, dtype=np.uint32)
    a_col[0::2] = cols[ps]
    a_col[1::2] = 0
    a_val = np.empty((2 * nc, 4), dtype=np.uint64)
    a_val[0::2] = one
    a_val[1::2] = cd.fr.enc(ks)
    b_rp = np.arange(0, nc + 1, dtype=np.uint32)
    b_col = cols[qs].astype(np.uint32)
    b_val = np.tile(one, (nc, 1))
    c_rp = np.arange(0, nc + 1, dtype=np.uint32)
    c_col = cols[2:].astype(np.uint32)
    c_val = np.tile(one, (nc, 1))
    m = ConstraintMatrices(ninst, nwit, nc, (a_rp, a_col, np.ascontiguousarray(a_val)),
                           (b_rp, np.ascontiguousarray(b_col), np.ascontiguousarray(b_val)),
                           (c_rp, np.ascontiguousarray(c_col), np.ascontiguousarray(c_val)))
    # full assignment: One, inputs, witnesses (in column order)
    full = [0] * (ninst + nwit)
    full[0] = 1
    for j in range(2 + nc):
        full[cols[j]] = vals[j]
    z = np.ascontiguousarray(cd.fr.enc(full))
    return m, z, full[1:ninst]


def dummy_r1cs(curve, num_variables: int, num_constraints: int, seed: int = 0):
    c = get_curve(curve)
    cd = CurveCodec(c)
    r = c.r
    gen = _splitmix(seed)
    a, b = _fr(gen, r), _fr(gen, r)
    nwit = num_variables - 1
    one = cd.fr.enc1(1)
    nc = num_constraints
    rp = np.concatenate([np.arange(0, nc, dtype=np.uint32), np.array([nc - 1], dtype=np.uint32)])  # last row empty
    mk = lambda col: (rp.copy(), np.full(nc - 1, col, dtype=np.uint32), np.ascontiguousarray(np.tile(one, (nc - 1, 1))))
    m = ConstraintMatrices(2, nwit, nc, mk(2), mk(3), mk(1))
    full = [1, a * b % r, a, b] + [a] * (num_variables - 3)
    z = np.ascontiguousarray(cd.fr.enc(full))
    return m, z, full[1:2]
