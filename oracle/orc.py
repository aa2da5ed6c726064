"""oracle/orc.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (oracle.cpp).

Takes and returns exactly the arrays of the product's C ABI (include/g16b200.h: Montgomery u64 limbs, affine x||y,
CSR matrices), so one set of inputs feeds both the CUDA path and this CPU restatement."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


class Csr(C.Structure):
    _fields_ = [("row_ptr", _u32p), ("col", _u32p), ("val", _u64p)]


class PkDesc(C.Structure):
    _fields_ = [("a_query", _u64p), ("a_len", C.c_uint64), ("b_g1_query", _u64p), ("b_g1_len", C.c_uint64),
                ("b_g2_query", _u64p), ("b_g2_len", C.c_uint64), ("h_query", _u64p), ("h_len", C.c_uint64),
                ("l_query", _u64p), ("l_len", C.c_uint64), ("alpha_g1", _u64p), ("beta_g1", _u64p),
                ("delta_g1", _u64p), ("beta_g2", _u64p), ("delta_g2", _u64p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/liboracle.so missing: run `make -C oracle`")
        _lib = C.CDLL(LIB)
        _lib.orc_hw_threads.restype = C.c_int
    return _lib


def hw_threads() -> int:
    return lib().orc_hw_threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c64(a, width=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def ntt(cid, log_n, vals, inverse=False, coset=False, threads=1):
    v = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4).copy()
    rc = lib().orc_ntt(cid, log_n, int(inverse), int(coset), _p(v), threads)
    assert rc == 0
    return v


def _csr(t, keep):
    rp, col, val = t
    rp = np.ascontiguousarray(rp, dtype=np.uint32)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val = np.ascontiguousarray(val, dtype=np.uint64)
    keep.extend([rp, col, val])
    s = Csr()
    s.row_ptr = rp.ctypes.data_as(_u32p)
    s.col = col.ctypes.data_as(_u32p) if col.size else None
    s.val = val.ctypes.data_as(_u64p) if val.size else None
    return s


def witness_map(cid, m, z, threads=1):
    """m: groth16_b200.ConstraintMatrices (or anything with the same fields); z: Montgomery limbs."""
    keep = []
    a, b, c = (_csr(t, keep) for t in (m.a, m.b, m.c))
    need = m.num_constraints + m.num_instance_variables
    log_n = max(need - 1, 0).bit_length()
    h = np.zeros((1 << log_n, 4), dtype=np.uint64)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    rc = lib().orc_witness_map(cid, m.num_instance_variables, m.num_constraints, m.num_witness_variables,
                               C.byref(a), C.byref(b), C.byref(c), _p(z), _p(h), threads)
    if rc == 1:
        raise ValueError("PolynomialDegreeTooLarge")
    assert rc == 0
    return h


def msm_g1(cid, nq, bases, scalars, threads=1):
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * nq)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(3 * nq, dtype=np.uint64)
    rc = lib().orc_msm_g1(cid, _p(bases), _p(scalars), C.c_uint64(n), _p(out), threads)
    assert rc == 0
    return out


def msm_g2(cid, nq, bases, scalars, threads=1):
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 4 * nq)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(6 * nq, dtype=np.uint64)
    rc = lib().orc_msm_g2(cid, _p(bases), _p(scalars), C.c_uint64(n), _p(out), threads)
    assert rc == 0
    return out


def prove(cid, nq, pk, m, z, r, s, threads=1):
    """pk: groth16_b200.ProvingKey-like (ABI arrays); r, s: Montgomery limbs.  Returns (proof limbs, [wm_ms, msm_ms])."""
    keep = []
    d = PkDesc()
    for name, width in (("a_query", 2 * nq), ("b_g1_query", 2 * nq), ("b_g2_query", 4 * nq), ("h_query", 2 * nq),
                        ("l_query", 2 * nq)):
        arr = np.ascontiguousarray(getattr(pk, name), dtype=np.uint64).reshape(-1, width)
        keep.append(arr)
        setattr(d, name, arr.ctypes.data_as(_u64p) if arr.size else None)
        setattr(d, name.replace("_query", "_len"), arr.shape[0])
    for k, v in (("alpha_g1", pk.vk.alpha_g1), ("beta_g1", pk.beta_g1), ("delta_g1", pk.delta_g1),
                 ("beta_g2", pk.vk.beta_g2), ("delta_g2", pk.vk.delta_g2)):
        arr = np.ascontiguousarray(v, dtype=np.uint64)
        keep.append(arr)
        setattr(d, k, arr.ctypes.data_as(_u64p))
    a, b, c = (_csr(t, keep) for t in (m.a, m.b, m.c))
    z = np.ascontiguousarray(z, dtype=np.uint64)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    proof = np.zeros(8 * nq, dtype=np.uint64)
    tms = (C.c_double * 2)()
    rc = lib().orc_prove(cid, C.byref(d), m.num_instance_variables, m.num_constraints, m.num_witness_variables,
                         C.byref(a), C.byref(b), C.byref(c), _p(z), _p(r), _p(s), _p(proof), threads, tms)
    if rc == 1:
        raise ValueError("PolynomialDegreeTooLarge")
    assert rc == 0
    return proof, [tms[0], tms[1]]


def batch_mul_g1(cid, nq, g, scalars, threads=1):
    """out[i] = scalars[i] * g (Montgomery Fr scalars) as affine ABI points."""
    g = np.ascontiguousarray(g, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((scalars.shape[0], 2 * nq), dtype=np.uint64)
    rc = lib().orc_batch_mul_g1(cid, _p(g), _p(scalars), C.c_uint64(scalars.shape[0]), _p(out), threads)
    assert rc == 0
    return out


def batch_mul_g2(cid, nq, g, scalars, threads=1):
    g = np.ascontiguousarray(g, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((scalars.shape[0], 4 * nq), dtype=np.uint64)
    rc = lib().orc_batch_mul_g2(cid, _p(g), _p(scalars), C.c_uint64(scalars.shape[0]), _p(out), threads)
    assert rc == 0
    return out


def set_msm_mode(mode: int) -> int:
    """0 = chunk-parallel Pippenger (default, the stronger CPU baseline), 1 = ark-ec's window-parallel scheme.  Returns the
    previous mode."""
    return lib().orc_set_msm_mode(int(mode))


def setup_scalars(cid, m, toxic, threads=1):
    """Exponents of every proving-key element (generator.rs:47-127, r1cs_to_qap.rs:128-170,237-247).
    m: ConstraintMatrices-like; toxic: (5, 4) Montgomery limbs alpha, beta, gamma, delta, tau.
    Returns dict(a, b, l, h, gamma_abc) of Montgomery Fr limb arrays."""
    keep = []
    a, b, c = (_csr(t, keep) for t in (m.a, m.b, m.c))
    ni, nc, nw = m.num_instance_variables, m.num_constraints, m.num_witness_variables
    n = 1 << max(nc + ni - 1, 0).bit_length()
    tx = np.ascontiguousarray(toxic, dtype=np.uint64).reshape(5, 4)
    z = lambda k: np.zeros((k, 4), dtype=np.uint64)
    out = dict(a=z(ni + nw), b=z(ni + nw), l=z(nw), h=z(n - 1), gamma_abc=z(ni))
    rc = lib().orc_setup_scalars(cid, ni, nc, nw, C.byref(a), C.byref(b), C.byref(c), _p(tx), _p(out["a"]), _p(out["b"]),
                                 _p(out["l"]), _p(out["h"]), _p(out["gamma_abc"]), threads)
    if rc == 1:
        raise ValueError("PolynomialDegreeTooLarge")
    assert rc == 0, f"orc_setup_scalars rc={rc}"
    return out


def generate_parameters(cid, nq, m, toxic, g1, g2, threads=1):
    """Complete CPU trusted setup with explicit toxic waste (generator.rs:47-208): exponents from setup_scalars, every
    group element by the fixed-base routine (generator.rs:129-183).  toxic: (5,4) Montgomery limbs; g1 / g2: ABI affine
    generators.  Returns a dict with the ABI arrays of ProvingKey / VerifyingKey (the caller wraps them; this module
    must not import the product package)."""
    ex = setup_scalars(cid, m, toxic, threads)
    tx = np.ascontiguousarray(toxic, dtype=np.uint64).reshape(5, 4)
    single1 = batch_mul_g1(cid, nq, g1, tx[[0, 1, 3]], threads)          # alpha_g1, beta_g1, delta_g1
    single2 = batch_mul_g2(cid, nq, g2, tx[[1, 2, 3]], threads)          # beta_g2, gamma_g2, delta_g2
    return dict(a_query=batch_mul_g1(cid, nq, g1, ex["a"], threads), b_g1_query=batch_mul_g1(cid, nq, g1, ex["b"], threads),
                b_g2_query=batch_mul_g2(cid, nq, g2, ex["b"], threads), h_query=batch_mul_g1(cid, nq, g1, ex["h"], threads),
                l_query=batch_mul_g1(cid, nq, g1, ex["l"], threads),
                gamma_abc_g1=batch_mul_g1(cid, nq, g1, ex["gamma_abc"], threads),
                alpha_g1=single1[0], beta_g1=single1[1], delta_g1=single1[2],
                beta_g2=single2[0], gamma_g2=single2[1], delta_g2=single2[2], exponents=ex)
