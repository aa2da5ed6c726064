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


# Fixed points of prime order r in G1 / G2, used as the "random group generators" of the trusted setup
# (generator.rs:26-32 samples them; any r-torsion points are valid).  Derived once by hashing to an x coordinate and
# clearing the cofactor; tests/test_constants.py re-derives them with the oracle and checks order and curve equation.
GENERATORS = {
    "bls12_381": dict(
        g1=(0x7225faa9ea1508c4c911e95beee000273bcbcbbd1d41ce0f18e8659251c0df081f9327b47e7a275f5a8031e433aa871,
            0x1949d82fc886648068a620dbb0a53b4c66213f267efa964cc41d733acdea86ce2c00e22e3202c3c594f339f30c2e3051),
        g2=((0x14e99e3b657acbd60979fa3525ae77af164a311785d46a08e67c54463326a88c60f04d6ded6397b2731df815d95892ca,
             0xa4601a9a48444765251e2a65f0b5619c3ea7290b0d3f6da7a8f3363808b58bb9e978796daba529742b6458d9e939b23),
            (0x1254a4d6508c091d0c5099f847013293e8d60d995850c5e88b1f18a7ef33f05d80bd1fedd8b272c6d123bdbfc18cb40d,
             0x4f838240ae2b657dcbd587add1ee8bde44e8802453d401454eec303e0e0bdfdf2fc11f9c642b8e2857205c8e32f01be))),
    "bn254": dict(
        g1=(0x2fda4996c18c5417c7ea845438e13c5d8c7e190a961ff9abbf8263e27d10ae7c,
            0x11911b5eaf5b93b8f1c774ba78cf95255929d38f32bfaad3167451e7c220715d),
        g2=((0x1df968558dbed90f366d524d5060557cc2f5b61c513ff08d463bc2843848b870,
             0x15ea7aac5c64a54af33865da65cf9ae6bbd4a0dd0b08317fc4aca44cd4560c50),
            (0xff98893b3d0be07b26b9485d5237dfcb32473d07a50d810132babe65d792aeb,
             0x2ea59f614a18e9f222cde53f20176036f654442bbd915ee616b7c422c0f8d69f))),
    "bls12_377": dict(
        g1=(0xe09f3eb12c7f3e7b77a61139ee7e62bb97ff9d88b8df1791fcf66b0dec04e4f24be4e0989fae248218a13843ed03d9,
            0x18144cece60d3a8fb0b6d28d6f8fe3a915e2be042f16b230110fb6b4a133d233236e7a189192b3dc4e31109664765d1),
        g2=((0x12cd874591f305b74d3cda047a5f555e752911ced088131536de5023ee8c1aed7aadb3ff7c15f6357f832f9698b963d,
             0x895a8006e7536bdf93f71f27bd5b280fa531cc66f63449312a91ec01c684a6ccff1ebd00e2931c19b0ac9bf61d4f68),
            (0x17663bd2b96d697799583fe676e7df81723dc2c223265cc2685c69e2b7d4c8464c342be5846f0eeeeec44de888db212,
             0x1b49e02eade86f46ff617db109925f68fc7bd69f1dbcbae76ff26e3388801324d585e56fbfb1cc438029a7a8b7f6b3))),
}
