"""Host-side mirror of ark-groth16's proving interface over the C ABI of libg16b200.so.

Names, argument meaning and error behaviour follow /root/reference:
  Groth16.create_proof_with_reduction_and_matrices  <- prover.rs:26-51
  Groth16.generate_parameters_with_qap              <- generator.rs:47-208 (explicit toxic waste)
  ProvingKey / VerifyingKey / Proof                 <- data_structures.rs:9-16,32-47,126-143
  ConstraintMatrices                                <- ark-relations `ConstraintMatrices` as consumed at r1cs_to_qap.rs:172-218
  SynthesisError variants                           <- r1cs_to_qap.rs:134,179 ; verifier.rs:30
Field elements cross this layer as numpy uint64 limb arrays in Montgomery form (codec.py converts Python ints).
The Rust shim a maintainer would write against the same symbols is in INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .codec import CurveCodec
from .params import CurveParams, get_curve


class SynthesisError(Exception):
    """ark_relations::r1cs::SynthesisError"""


class PolynomialDegreeTooLarge(SynthesisError):
    pass


class MalformedKey(SynthesisError):
    pass


class CudaError(RuntimeError):
    pass


def _check(rc: int):
    if rc == _lib.G16_OK:
        return
    msg = _lib.last_error()
    if rc == _lib.ERR_POLYNOMIAL_DEGREE_TOO_LARGE:
        raise PolynomialDegreeTooLarge(msg)
    if rc == _lib.ERR_MALFORMED_KEY:
        raise MalformedKey(msg)
    if rc == _lib.ERR_CUDA:
        raise CudaError(msg)
    raise ValueError(msg)


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.uint64
    return a.ctypes.data_as(C.c_void_p)


def _u64p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.uint64
    return a.ctypes.data_as(_lib.u64p)


def _u32p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"] and a.dtype == np.uint32
    return a.ctypes.data_as(_lib.u32p)


@dataclass
class ConstraintMatrices:
    """CSR form of ark-relations' ConstraintMatrices (rows of (coeff, column); column < num_instance_variables is an
    instance variable, otherwise witness index + num_instance_variables)."""
    num_instance_variables: int
    num_witness_variables: int
    num_constraints: int
    a: Tuple[np.ndarray, np.ndarray, np.ndarray]  # row_ptr u32 [nc+1], col u32 [nnz], val u64 [nnz,4] Montgomery
    b: Tuple[np.ndarray, np.ndarray, np.ndarray]
    c: Tuple[np.ndarray, np.ndarray, np.ndarray]

    @staticmethod
    def from_rows(curve, num_instance: int, num_witness: int, a_rows, b_rows, c_rows) -> "ConstraintMatrices":
        cd = CurveCodec(get_curve(curve))

        def csr(rows):
            rp = np.zeros(len(rows) + 1, dtype=np.uint32)
            cols, vals = [], []
            for i, row in enumerate(rows):
                for cf, idx in row:
                    cols.append(idx)
                    vals.append(cf)
                rp[i + 1] = len(cols)
            col = np.asarray(cols, dtype=np.uint32)
            val = cd.fr.enc(vals) if vals else np.zeros((0, 4), dtype=np.uint64)
            return rp, col, np.ascontiguousarray(val)

        return ConstraintMatrices(num_instance, num_witness, len(a_rows), csr(a_rows), csr(b_rows), csr(c_rows))


@dataclass
class VerifyingKey:
    alpha_g1: np.ndarray
    beta_g2: np.ndarray
    gamma_g2: Optional[np.ndarray]
    delta_g2: np.ndarray
    gamma_abc_g1: Optional[np.ndarray]


@dataclass
class ProvingKey:
    vk: VerifyingKey
    beta_g1: np.ndarray
    delta_g1: np.ndarray
    a_query: np.ndarray
    b_g1_query: np.ndarray
    b_g2_query: np.ndarray
    h_query: np.ndarray
    l_query: np.ndarray


@dataclass
class Proof:
    a: np.ndarray  # G1 affine limbs
    b: np.ndarray  # G2 affine limbs
    c: np.ndarray  # G1 affine limbs


class Groth16:
    """One instance = one curve on one GPU (a g16_ctx).  The circuit (matrices) and the proving key are made
    resident once and reused by every proof, like a long-lived prover process would."""

    def __init__(self, curve, device: int = 0):
        self.curve: CurveParams = get_curve(curve)
        self.codec = CurveCodec(self.curve)
        self._lib = _lib.load()
        h = C.c_void_p()
        _check(self._lib.g16_ctx_create(self.curve.cid, device, C.byref(h)))
        self._ctx = h
        self.nq = self._lib.g16_fq_limbs(self._ctx)
        self._matrices: Optional[ConstraintMatrices] = None
        self._pk_resident = False
        self._pk_obj: Optional[ProvingKey] = None   # identity of the resident key (None: minted by g16_setup and not exported)
        self.world = 1

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.g16_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the two dependency-level operations (ark-poly / ark-ec) ----
    def ntt(self, values: np.ndarray, inverse: bool = False, coset: bool = False) -> np.ndarray:
        """Radix2EvaluationDomain::{fft,ifft}_in_place / coset variants on 2^k Montgomery Fr elements."""
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4).copy()
        n = v.shape[0]
        log_n = max(n - 1, 0).bit_length()
        if (1 << log_n) != n:
            raise ValueError("length must be a power of two (ark resizes to domain.size(); do that in the caller)")
        _check(self._lib.g16_ntt(self._ctx, log_n, int(inverse), int(coset), _ptr(v)))
        return v

    def ntt_log(self, log_n: int, values: np.ndarray, inverse=False, coset=False) -> np.ndarray:
        """Same transform with the domain size given explicitly (error-path tests: log_n above the two-adicity).  The C side
        copies 32 << log_n bytes in and out of `values`, so the length is checked here."""
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4).copy()
        if log_n < 0 or (log_n <= self.curve.two_adicity and v.shape[0] != (1 << log_n)):
            raise ValueError(f"values must hold exactly 2^{log_n} field elements")
        # log_n above the two-adicity: the C side returns PolynomialDegreeTooLarge before it touches the buffer
        _check(self._lib.g16_ntt(self._ctx, log_n, int(inverse), int(coset), _ptr(v)))
        return v

    def witness_map_from_evals(self, a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, 4)
        n = a.shape[0]
        log_n = max(n - 1, 0).bit_length()
        if (1 << log_n) != n or b.shape != a.shape or c.shape != a.shape:
            raise ValueError("a, b, c must have the same power-of-two length")
        h = np.empty_like(a)
        _check(self._lib.g16_witness_map_evals(self._ctx, log_n, _ptr(a), _ptr(b), _ptr(c), _ptr(h)))
        return h

    def msm_g1(self, bases: np.ndarray, scalars: np.ndarray) -> np.ndarray:
        """VariableBaseMSM::msm_bigint on G1: truncates to the shorter operand like ark (prover.rs:66 relies on it)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * self.nq)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = min(bases.shape[0], scalars.shape[0])
        out = np.zeros(3 * self.nq, dtype=np.uint64)
        _check(self._lib.g16_msm_g1(self._ctx, _ptr(bases), _ptr(scalars), n, _ptr(out)))
        return out

    def msm_g2(self, bases: np.ndarray, scalars: np.ndarray) -> np.ndarray:
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 4 * self.nq)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = min(bases.shape[0], scalars.shape[0])
        out = np.zeros(6 * self.nq, dtype=np.uint64)
        _check(self._lib.g16_msm_g2(self._ctx, _ptr(bases), _ptr(scalars), n, _ptr(out)))
        return out

    def prepare_inputs(self, vk: VerifyingKey, public_inputs) -> np.ndarray:
        """Groth16::prepare_inputs (verifier.rs:25-39): gamma_abc_g1[0] + sum_i x_i * gamma_abc_g1[i + 1], computed as ONE G1
        MSM with scalars (1, x_0, x_1, ...).  `public_inputs`: Python ints, or an (l, 4) array of Montgomery Fr limbs (ark's
        memory image).  Returns the projective point in msm_g1's encoding.  A length mismatch is
        SynthesisError::MalformedVerifyingKey (verifier.rs:30)."""
        if vk.gamma_abc_g1 is None:
            raise MalformedKey("verifying key has no gamma_abc_g1")
        abc = np.ascontiguousarray(vk.gamma_abc_g1, dtype=np.uint64).reshape(-1, 2 * self.nq)
        if isinstance(public_inputs, np.ndarray):
            xs = self.codec.fr.dec(np.ascontiguousarray(public_inputs, dtype=np.uint64).reshape(-1, 4))
        else:
            xs = [int(x) % self.curve.r for x in public_inputs]
        if len(xs) + 1 != abc.shape[0]:
            raise MalformedKey("public input count does not match the verifying key")
        return self.msm_g1(abc, self.codec.fr.bigint([1] + xs))

    # ---- resident state ----
    def load_matrices(self, m: ConstraintMatrices):
        def csr(t):
            rp, col, val = t
            s = _lib.Csr()
            rpc = np.ascontiguousarray(rp, dtype=np.uint32)
            colc = np.ascontiguousarray(col, dtype=np.uint32)
            valc = np.ascontiguousarray(val, dtype=np.uint64)
            s.row_ptr = _u32p(rpc)
            s.col = _u32p(colc) if colc.size else None
            s.val = _u64p(valc) if valc.size else None
            return s, (rpc, colc, valc)

        keep = []
        structs = []
        for t in (m.a, m.b, m.c):
            s, k = csr(t)
            structs.append(s)
            keep.append(k)
        _check(self._lib.g16_circuit_load(self._ctx, m.num_instance_variables, m.num_constraints,
                                          m.num_witness_variables, C.byref(structs[0]), C.byref(structs[1]),
                                          C.byref(structs[2])))
        self._matrices = m
        self._pk_resident = False
        self._pk_obj = None

    def load_proving_key(self, pk: ProvingKey, rank: int = 0, world: int = 1):
        d = _lib.PkDesc()
        arrs = {}
        for name in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
            arr = np.ascontiguousarray(getattr(pk, name), dtype=np.uint64)
            arrs[name] = arr
            width = 4 * self.nq if name == "b_g2_query" else 2 * self.nq
            arr2 = arr.reshape(-1, width)
            setattr(d, name, _u64p(arr) if arr.size else None)
            setattr(d, name.replace("_query", "_len"), arr2.shape[0])
        singles = dict(alpha_g1=pk.vk.alpha_g1, beta_g1=pk.beta_g1, delta_g1=pk.delta_g1, beta_g2=pk.vk.beta_g2,
                       delta_g2=pk.vk.delta_g2)
        for k, v in singles.items():
            arrs[k] = np.ascontiguousarray(v, dtype=np.uint64)
            setattr(d, k, _u64p(arrs[k]))
        _check(self._lib.g16_pk_load(self._ctx, C.byref(d), rank, world))
        self._pk_resident = True
        self._pk_obj = pk
        self.world = world

    # ---- generator.rs:47-208 with explicit toxic waste and generators ----
    def generate_parameters_with_qap(self, matrices: ConstraintMatrices, alpha, beta, gamma, delta, tau, g1_generator,
                                     g2_generator, export: bool = True) -> Optional[ProvingKey]:
        """alpha..tau: Python ints (canonical); g1/g2 generators: affine int tuples.  The key becomes resident."""
        self.load_matrices(matrices)
        cd = self.codec
        sc = [np.ascontiguousarray(cd.fr.enc1(x)) for x in (alpha, beta, gamma, delta, tau)]
        g1 = np.ascontiguousarray(cd.enc_g1([g1_generator])[0])
        g2 = np.ascontiguousarray(cd.enc_g2([g2_generator])[0])
        _check(self._lib.g16_setup(self._ctx, *[_ptr(x) for x in sc], _ptr(g1), _ptr(g2)))
        self._pk_resident = True
        self.world = 1
        self._pk_obj = self.export_proving_key() if export else None
        return self._pk_obj

    def export_proving_key(self) -> ProvingKey:
        m = self._matrices
        nq = self.nq
        nv = m.num_instance_variables + m.num_witness_variables
        n = 1 << self._lib.g16_domain_log(self._ctx)
        z = lambda rows, w: np.zeros((rows, w), dtype=np.uint64)
        out = dict(a_query=z(nv, 2 * nq), b_g1_query=z(nv, 2 * nq), b_g2_query=z(nv, 4 * nq), h_query=z(n - 1, 2 * nq),
                   l_query=z(m.num_witness_variables, 2 * nq), alpha_g1=z(1, 2 * nq), beta_g1=z(1, 2 * nq),
                   delta_g1=z(1, 2 * nq), beta_g2=z(1, 4 * nq), gamma_g2=z(1, 4 * nq), delta_g2=z(1, 4 * nq),
                   gamma_abc_g1=z(m.num_instance_variables, 2 * nq))
        d = _lib.PkExportDesc()
        for k, v in out.items():
            setattr(d, k, _u64p(v) if v.size else None)
        _check(self._lib.g16_pk_export(self._ctx, C.byref(d)))
        vk = VerifyingKey(out["alpha_g1"][0], out["beta_g2"][0], out["gamma_g2"][0], out["delta_g2"][0], out["gamma_abc_g1"])
        return ProvingKey(vk, out["beta_g1"][0], out["delta_g1"][0], out["a_query"], out["b_g1_query"], out["b_g2_query"],
                          out["h_query"], out["l_query"])

    # ---- prover.rs:26-51 ----
    def create_proof_with_reduction_and_matrices(self, pk: Optional[ProvingKey], r, s,
                                                 matrices: Optional[ConstraintMatrices], num_inputs: int,
                                                 num_constraints: int, full_assignment: np.ndarray,
                                                 flags: int = 0) -> Proof:
        """r, s, full_assignment: Montgomery Fr limbs (r, s may also be Python ints).  `pk` / `matrices` may be None
        to reuse what is already resident on the GPU."""
        if matrices is not None and matrices is not self._matrices:
            self.load_matrices(matrices)
        # the reference always proves under the `pk` argument (prover.rs:26): a key other than the resident one is loaded
        if pk is not None and (not self._pk_resident or pk is not self._pk_obj):
            self.load_proving_key(pk)
        m = self._matrices
        if m is None or not self._pk_resident:
            raise ValueError("matrices and proving key must be loaded")
        if num_inputs != m.num_instance_variables or num_constraints != m.num_constraints:
            raise ValueError("num_inputs / num_constraints do not match the matrices")
        rr = self._fr_arg(r)
        ss = self._fr_arg(s)
        z = np.ascontiguousarray(full_assignment, dtype=np.uint64).reshape(-1, 4)
        if z.shape[0] != m.num_instance_variables + m.num_witness_variables:
            raise ValueError("full_assignment has the wrong length")
        nq = self.nq
        out = np.zeros(8 * nq, dtype=np.uint64)
        _check(self._lib.g16_prove(self._ctx, _ptr(rr), _ptr(ss), _ptr(z), flags, _ptr(out)))
        return Proof(out[:2 * nq].copy(), out[2 * nq:6 * nq].copy(), out[6 * nq:].copy())

    def prove_raw(self, r_limbs: np.ndarray, s_limbs: np.ndarray, z_ptr, flags: int, out: np.ndarray):
        """Thin call used by bench.py: everything already in ABI form; z_ptr is a host or device address."""
        _check(self._lib.g16_prove(self._ctx, _ptr(r_limbs), _ptr(s_limbs), C.c_void_p(z_ptr), flags, _ptr(out)))

    # ---- pipelined proving: two slots per context (g16_prove_submit / g16_prove_wait) ----
    def prove_submit_raw(self, slot: int, r_limbs: np.ndarray, s_limbs: np.ndarray, z_ptr, flags: int):
        """Enqueue a whole proof on `slot` and return; r/s/z buffers must stay alive until prove_wait_raw(slot)."""
        _check(self._lib.g16_prove_submit(self._ctx, slot, _ptr(r_limbs), _ptr(s_limbs), C.c_void_p(z_ptr), flags))

    def prove_wait_raw(self, slot: int, out: np.ndarray):
        _check(self._lib.g16_prove_wait(self._ctx, slot, _ptr(out)))

    def prove_partial_submit_raw(self, slot: int, r_limbs: np.ndarray, z_ptr, flags: int):
        _check(self._lib.g16_prove_partial_submit(self._ctx, slot, _ptr(r_limbs), C.c_void_p(z_ptr), flags))

    def prove_partial_wait_raw(self, slot: int, out: np.ndarray):
        _check(self._lib.g16_prove_partial_wait(self._ctx, slot, _ptr(out)))

    def prove_partial_raw(self, r_limbs: np.ndarray, z_ptr, flags: int, out: np.ndarray):
        _check(self._lib.g16_prove_partial(self._ctx, _ptr(r_limbs), C.c_void_p(z_ptr), flags, _ptr(out)))

    # ---- sharded proving with the NCCL exchange inside the library ----
    def comm_unique_id(self) -> np.ndarray:
        out = np.zeros(256, dtype=np.uint8)
        _check(self._lib.g16_comm_unique_id(out.ctypes.data_as(C.c_void_p)))
        return out

    def comm_init(self, unique_id: np.ndarray, rank: int, world: int):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 256
        _check(self._lib.g16_comm_init(self._ctx, uid.ctypes.data_as(C.c_void_p), rank, world))

    def prove_sharded_raw(self, r_limbs, s_limbs, z_ptr, flags: int, out: np.ndarray):
        _check(self._lib.g16_prove_sharded(self._ctx, _ptr(r_limbs), _ptr(s_limbs), C.c_void_p(z_ptr), flags, _ptr(out)))

    def prove_sharded_submit_raw(self, slot: int, r_limbs, s_limbs, z_ptr, flags: int):
        _check(self._lib.g16_prove_sharded_submit(self._ctx, slot, _ptr(r_limbs), _ptr(s_limbs), C.c_void_p(z_ptr), flags))

    def prove_sharded_wait_raw(self, slot: int, out: np.ndarray):
        _check(self._lib.g16_prove_sharded_wait(self._ctx, slot, _ptr(out)))

    def prove_assemble_prepare(self, r, s):
        """start the (r, s)-only scalar multiplications on a helper thread (overlaps GPU work and the gather)"""
        self._asm_keep = (self._fr_arg(r), self._fr_arg(s))
        _check(self._lib.g16_prove_assemble_prepare(self._ctx, _ptr(self._asm_keep[0]), _ptr(self._asm_keep[1])))

    def prove_assemble(self, r, s, partials: np.ndarray) -> Proof:
        rr, ss = self._fr_arg(r), self._fr_arg(s)
        pl = self._lib.g16_partial_limbs(self._ctx)
        p = np.ascontiguousarray(partials, dtype=np.uint64).reshape(-1, pl)
        nq = self.nq
        out = np.zeros(8 * nq, dtype=np.uint64)
        _check(self._lib.g16_prove_assemble(self._ctx, _ptr(rr), _ptr(ss), _ptr(p), p.shape[0], _ptr(out)))
        return Proof(out[:2 * nq].copy(), out[2 * nq:6 * nq].copy(), out[6 * nq:].copy())

    def partial_limbs(self) -> int:
        return self._lib.g16_partial_limbs(self._ctx)

    def witness_map_from_matrices(self, matrices: Optional[ConstraintMatrices], num_inputs: int, num_constraints: int,
                                  full_assignment: np.ndarray) -> np.ndarray:
        """R1CSToQAP::witness_map_from_matrices (r1cs_to_qap.rs:172-235) -> domain_size Montgomery Fr coefficients."""
        if matrices is not None and matrices is not self._matrices:
            self.load_matrices(matrices)
        m = self._matrices
        z = np.ascontiguousarray(full_assignment, dtype=np.uint64).reshape(-1, 4)
        if z.shape[0] != m.num_instance_variables + m.num_witness_variables:
            raise ValueError("full_assignment has the wrong length")
        n = 1 << self._lib.g16_domain_log(self._ctx)
        h = np.zeros((n, 4), dtype=np.uint64)
        _check(self._lib.g16_witness_map(self._ctx, _ptr(z), 0, _ptr(h)))
        return h

    def timings(self) -> dict:
        t = _lib.Timings()
        _check(self._lib.g16_get_timings(self._ctx, C.byref(t)))
        names = ["h", "l", "a", "b_g1", "b_g2"]
        return dict(total_ms=t.total_ms, h2d_ms=t.h2d_ms, witness_map_ms=t.witness_map_ms,
                    msm_ms={n: t.msm_ms[i] for i, n in enumerate(names)},
                    msm_accum_ms={n: t.msm_accum_ms[i] for i, n in enumerate(names)},
                    msm_pairs={n: int(t.msm_pairs[i]) for i, n in enumerate(names)},
                    msm_entries={n: int(t.msm_entries[i]) for i, n in enumerate(names)},
                    msm_begin_ms={n: t.msm_begin_ms[i] for i, n in enumerate(names)},
                    msm_end_ms={n: t.msm_end_ms[i] for i, n in enumerate(names)},
                    host_finish_ms=t.host_finish_ms, launches=int(t.launches), h2d_bytes=int(t.h2d_bytes),
                    d2h_bytes=int(t.d2h_bytes))

    def set_option(self, key: str, value: int):
        """MSM launch-geometry knobs (g16_set_option; results never depend on them)."""
        _check(self._lib.g16_set_option(self._ctx, key.encode(), int(value)))

    def config(self) -> dict:
        """Launch geometry of the resident key's MSMs plus two derived figures bench.py reports: the field products per
        bucket entry of the G1 accumulation stage (XYZZ mixed addition = 10; batched-affine addition = 6 + the combine's
        share) and the stage's name."""
        c = _lib.Config()
        _check(self._lib.g16_get_config(self._ctx, C.byref(c)))
        d = {n: int(getattr(c, n)) for n, _ in _lib.Config._fields_ if n != "reserved"}
        R = d["ba_rounds_g1"]
        ba_mul = 6.0 + 3.0 / max(1, d["ba_m"])            # forward 1 + backward 5 + combine 3 per thread product
        frac = 1.0 - 0.5 ** R
        d["imad_per_g1_entry_mul"] = frac * ba_mul + (1.0 - frac) * 10.0 if R > 0 else 10.0
        d["g1_accum_stage"] = (f"G1 bucket accumulation: {R} batched-affine rounds (ba_forward / ba_combine / ba_backward) + "
                               "msm_accum_l0<Fq> on the last list; one stage per G1 MSM" if R > 0 else
                               "msm_accum_l0<Fq> (G1 bucket accumulation, XYZZ mixed additions); one launch per G1 MSM")
        return d

    def _fr_arg(self, x) -> np.ndarray:
        if isinstance(x, (int, np.integer)):
            return np.ascontiguousarray(self.codec.fr.enc1(int(x)))
        return np.ascontiguousarray(x, dtype=np.uint64).reshape(4)
