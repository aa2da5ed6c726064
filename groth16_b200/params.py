"""Field moduli and limb counts of the three supported pairing curves.

Only what the host-side codecs need; the CUDA side has its own generated table (csrc/g16_constants.h)."""
from dataclasses import dataclass


@dataclass(frozen=True)
class CurveParams:
    name: str
    cid: int            # curve id at the C ABI (include/g16b200.h)
    r: int              # scalar field modulus
    q: int              # base field modulus
    fr_generator: int   # Fr::GENERATOR (coset offset, r1cs_to_qap.rs:204)
    two_adicity: int

    @property
    def fq_limbs(self) -> int:
        return (self.q.bit_length() + 63) // 64


BLS12_381 = CurveParams(
    "bls12_381", 0,
    0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    7, 32)
BN254 = CurveParams(
    "bn254", 1,
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    5, 28)
BLS12_377 = CurveParams(
    "bls12_377", 2,
    8444461749428370424248824938781546531375899335154063827935233455917409239041,
    258664426012969094010652733694893533536393512754914660539884262666720468348340822774968888139573360124440321458177,
    22, 47)

CURVES = {c.name: c for c in (BLS12_381, BN254, BLS12_377)}


def get_curve(curve) -> CurveParams:
    if isinstance(curve, CurveParams):
        return curve
    if isinstance(curve, str):
        return CURVES[curve]
    if hasattr(curve, "name"):
        return CURVES[curve.name]
    raise KeyError(curve)
