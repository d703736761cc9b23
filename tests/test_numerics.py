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


def test_magic_number_rounding_equals_rint():
    """limb_score rounds to nearest-even without F2I: the low bits of float32(x + 1.5*2^23) hold rint(x) for
    0 <= x < 2^22, and the low word of float64(x + 1.5*2^52) holds rint(x) for |x| < 2^31."""
    import numpy as np
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.random(200000, dtype=np.float32) * np.float32(4.0e6 / 1.0),
                        (np.arange(0, 4096, dtype=np.float32) + np.float32(0.5)),      # exact ties
                        np.float32([0.0, 0.49999997, 0.5, 1.5, 2.5, 4194302.5, 4194303.0])])
    x = x[x < 4194304.0]
    bits = (x + np.float32(12582912.0)).view(np.int32) - np.int32(0x4B400000)
    assert np.array_equal(bits, np.rint(x).astype(np.int32))
    y = np.concatenate([(rng.random(200000) - 0.5) * 4.0e9, np.arange(-2048, 2048) + 0.5, [0.0, -0.0, -1e-14, 2147483646.5]])
    y = y[np.abs(y) < 2147483647.0]
    lo = (y + 6755399441055744.0).view(np.int64).astype(np.int32)  # truncation keeps the low word, like __double2loint
    assert np.array_equal(lo, np.rint(y).astype(np.int64).astype(np.int32))


def test_screen_positions_and_sample_counts_are_conservative():
    """The f32 screen of limb_score (limb_score_persist.cuh: screen_pair) may only report CERTAIN failures.  Modelled here
    in numpy float32, step for step: whenever it trusts its sample count m it equals the reference's
    min(round(norm + 1), mid_num) (evaluate.py:226), and every sample whose position is more than 2/64 px (in the
    kernel's biased units: u & 63 > 2) away from a rounding boundary lands on the pixel the reference's f64
    np.linspace + round() picks (:232-235) -- on maps up to 2048 px."""
    import numpy as np
    f32 = np.float32
    rng = np.random.default_rng(12)
    mid = 20
    checked = trusted = 0
    for size in (128, 512, 2048):
        n = 4000
        ax, ay = rng.random(n) * (size - 1), rng.random(n) * (size - 1)
        # a mix of short and long limbs, some nearly axis-aligned, some landing close to pixel boundaries
        length = np.where(rng.random(n) < 0.5, rng.random(n) * 25.0, rng.random(n) * size)
        ang = rng.random(n) * 2 * np.pi
        bx = np.clip(ax + length * np.cos(ang), 0, size - 1)
        by = np.clip(ay + length * np.sin(ang), 0, size - 1)
        snap = rng.random(n) < 0.2
        ax = np.where(snap, np.rint(ax) + 0.5 - 1e-7 * rng.integers(-3, 4, n), ax).clip(0, size - 1)
        for i in range(n):
            fax, fay, fbx, fby = f32(ax[i] * 64.0), f32(ay[i] * 64.0), f32(bx[i] * 64.0), f32(by[i] * 64.0)
            dx, dy = f32(fbx - fax), f32(fby - fay)
            n2 = f32(f32(f32(dx * dx) + f32(dy * dy)) * f32(1.0 / 4096.0))
            if not n2 > f32(1e-6):
                continue
            norm = math.sqrt((bx[i] - ax[i]) ** 2 + (by[i] - ay[i]) ** 2)
            m_ref = min(int(round(norm + 1)), mid)
            m = -1
            rs0 = f32(1.0) / np.sqrt(n2, dtype=f32)
            for rs in (rs0, np.nextafter(np.nextafter(rs0, f32(0)), f32(0)), np.nextafter(np.nextafter(rs0, f32(9)), f32(9))):
                # MUFU.RSQ is good to ~2 ulp: whenever the screen trusts its m under any such value, m is the reference's
                qf = f32(f32(n2 * rs) + f32(1.0))
                r = np.rint(qf)
                longp = qf >= f32(mid) + f32(0.51)
                if not longp and not abs(f32(qf - r)) < f32(0.49):
                    continue  # tie guard: the pair survives unscreened
                mm = mid if longp else min(int(r), mid)
                assert mm == m_ref, (size, i, mm, m_ref, norm)
                if rs is rs0:
                    m = mm
            if m < 0:
                continue
            trusted += 1
            if m < 2:
                continue
            inv = f32(1.0) / f32(m - 1)
            sx, sy = f32(dx * inv), f32(dy * inv)
            ax_o, ay_o = f32(fax + f32(33.0)), f32(fay + f32(33.0))
            xs = np.linspace(ax[i], bx[i], m)
            ys = np.linspace(ay[i], by[i], m)
            for t in range(m):
                tf = f32(t)
                ux = int(np.rint(f32(np.float64(tf) * np.float64(sx) + np.float64(ax_o))))  # FFMA, then RN to an integer
                uy = int(np.rint(f32(np.float64(tf) * np.float64(sy) + np.float64(ay_o))))
                if (ux & 63) > 2 and (uy & 63) > 2:
                    assert (ux >> 6, uy >> 6) == (int(round(xs[t])), int(round(ys[t]))), (size, i, t)
                    checked += 1
    assert trusted > 8000 and checked > 100000
