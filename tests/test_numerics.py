"""CPU: arithmetic identities the CUDA kernels rely on to stay bit-identical to numpy while being cheaper."""
import math
import random
from fractions import Fraction


def _rn(fr: Fraction) -> float:
    return float(fr)  # Fraction -> float is correctly rounded (round-half-even)


def _fma(a: float, b: float, c: float) -> float:
    return _rn(Fraction(a) * Fraction(b) + Fraction(c))


def test_markstein_division_by_small_integers_is_correctly_rounded():
    """limb_score.cuh computes np.linspace's step = delta / (m - 1) as q0 = RN(delta*y), r = RN(delta - q0*d),
    q = RN(q0 + r*y) with y = RN(1/d).  It must equal the correctly rounded quotient for every d it can see."""
    rnd = random.Random(7)
    for d in range(1, 64):
        y = _rn(Fraction(1, d))
        for _ in range(300):
            kind = rnd.random()
            if kind < 0.5:
                x = rnd.uniform(-2100.0, 2100.0)                      # coordinate differences on maps up to 2048 px
            elif kind < 0.8:
                x = (rnd.getrandbits(53) / 2 ** 53) * rnd.choice([1e-3, 1.0, 37.0, 511.0]) * rnd.choice([-1, 1])
            else:
                x = rnd.choice([-1, 1]) * math.ldexp(rnd.getrandbits(53) | 1, rnd.randint(-80, 10))
            q0 = x * y
            q = _fma(_fma(-q0, float(d), x), y, q0)
            assert q == _rn(Fraction(x) / d), (x, d)


def test_prior_sign_is_decided_by_norm_vs_half_extent():
    """limb_score.cuh skips `0.5*extent/norm - 1` unless norm > 0.5*extent: the reference's min(prior, 0) is 0 otherwise."""
    rnd = random.Random(11)
    for _ in range(20000):
        half = 0.5 * rnd.choice([46, 64, 128, 368, 512, 1333])
        norm = half * (1.0 + rnd.choice([-1, 1]) * rnd.choice([0.0, 2.0 ** -52, 2.0 ** -51, 1e-12, 1e-6, 0.3]))
        if norm <= 0:
            continue
        prior = half / norm - 1.0
        assert (prior < 0) == (norm > half)


def test_mid_num_clamp_needs_no_sqrt():
    """n2 >= mid_num^2 implies min(round(sqrt(n2) + 1), mid_num) == mid_num."""
    rnd = random.Random(13)
    for M in (1, 5, 10, 20, 63):
        for _ in range(2000):
            n2 = M * M * (1.0 + rnd.choice([0.0, 2.0 ** -52, 1e-9, 0.5, 30.0]))
            assert min(round(math.sqrt(n2) + 1), M) == M
