"""Point types at the native API surface (ark G1Affine / G2Affine / G1Projective / G2Projective as
used by /root/reference/src/miller_loop_native.rs:1 and src/miller_loop_native_optimized.rs:1).
Marshalling only; generator coordinates as in src/fields_as_trees/g1_curve.rs:53-74, g2_curve.rs:63-117."""
from dataclasses import dataclass

from ..fields.types import Fq, Fq2


@dataclass(frozen=True)
class G1Affine:
    x: Fq
    y: Fq
    infinity: bool = False

    @staticmethod
    def identity():
        return G1Affine(Fq(0), Fq(0), True)

    @staticmethod
    def generator():
        return G1Affine(
            Fq(0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB),
            Fq(0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1))

    def limbs(self):
        return self.x.limbs() + self.y.limbs()


@dataclass(frozen=True)
class G2Affine:
    x: Fq2
    y: Fq2
    infinity: bool = False

    @staticmethod
    def identity():
        return G2Affine(Fq2.zero(), Fq2.zero(), True)

    @staticmethod
    def generator():
        return G2Affine(
            Fq2(Fq(0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8),
                Fq(0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E)),
            Fq2(Fq(0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801),
                Fq(0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)))

    def limbs(self):
        return self.x.limbs() + self.y.limbs()


@dataclass(frozen=True)
class G1Projective:
    """Jacobian (x, y, z) as ark-ec 0.4 short_weierstrass::Projective."""
    x: Fq
    y: Fq
    z: Fq

    @staticmethod
    def from_affine(p: G1Affine):
        # ark-ec 0.4 Projective::zero() and From<Affine> give (1, 1, 0) for the identity
        return G1Projective(Fq(1), Fq(1), Fq(0)) if p.infinity else G1Projective(p.x, p.y, Fq(1))

    @staticmethod
    def generator():
        return G1Projective.from_affine(G1Affine.generator())

    def limbs(self):
        return self.x.limbs() + self.y.limbs() + self.z.limbs()


@dataclass(frozen=True)
class G2Projective:
    x: Fq2
    y: Fq2
    z: Fq2

    @staticmethod
    def from_affine(q: G2Affine):
        return G2Projective(Fq2.one(), Fq2.one(), Fq2.zero()) if q.infinity else G2Projective(q.x, q.y, Fq2.one())

    @staticmethod
    def generator():
        return G2Projective.from_affine(G2Affine.generator())

    def limbs(self):
        return self.x.limbs() + self.y.limbs() + self.z.limbs()
