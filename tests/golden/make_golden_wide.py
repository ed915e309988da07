"""Generates tests/golden/golden_n16_l3_wide.npz from the CPU oracle: the operations either side of the
multiply/rotate core on the same tiny parameter set as golden_n16_l3.npz.

    python tests/golden/make_golden_wide.py

Deterministic: fixed numpy seed; psi = documented default root per prime (also for the primes of the custom basis)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fhe_oracle as O  # noqa: E402


def main():
    degree, t, nmod = 16, 1153, 3
    rng = np.random.default_rng(20260925)
    par = O.BfvParameters(degree, t, moduli_sizes=[62] * nmod)
    ctx = par.context_at_level(0)
    sk = O.SecretKey(par, rng)
    count = 4
    cta = [sk.encrypt(O.simd_encode(par, rng.integers(0, t, degree)), 0, rng) for _ in range(count)]
    ctb = [sk.encrypt(O.simd_encode(par, rng.integers(0, t, degree)), 0, rng) for _ in range(count)]
    out = dict(degree=degree, t=t, moduli=np.array(par.moduli, dtype=np.uint64), sk=sk.coeffs,
               a=np.stack([c.to_array() for c in cta]), b=np.stack([c.to_array() for c in ctb]))
    # sub / neg / switch_down / switch_to_level
    out["sub"] = np.stack([x.sub(y).to_array() for x, y in zip(cta, ctb)])
    out["neg"] = np.stack([x.neg().to_array() for x in cta])
    out["switch_down"] = np.stack([x.copy().switch_down().to_array() for x in cta])
    out["switch_to_2"] = np.stack([x.copy().switch_to_level(2).to_array() for x in cta])
    # ct (+,-,*) pt
    pv = rng.integers(0, t, degree)
    dp = O.plaintext_to_poly(par, O.simd_encode(par, pv), 0)             # Plaintext::to_poly
    pn = O.Poly.from_u64(ctx, O.simd_encode(par, pv), O.NTT)             # Plaintext::poly_ntt
    out["pt_to_poly"], out["pt_poly_ntt"] = dp.c, pn.c
    plus, minus, times = [], [], []
    for x in cta:
        p = x.copy(); p.c[0] = p.c[0].copy().iadd(dp); plus.append(p.to_array())
        m = x.copy(); m.c[0] = m.c[0].copy().isub(dp); minus.append(m.to_array())
        times.append(np.stack([q.mul(pn).c for q in x.c]))
    out["add_plain"], out["sub_plain"], out["mul_plain"] = np.stack(plus), np.stack(minus), np.stack(times)
    # dot_product_scalar: 2 groups of 2 terms
    pts = [O.Poly.random(ctx, O.NTT, rng) for _ in range(count)]
    out["dot_pts"] = np.stack([p.c for p in pts])
    out["dot"] = np.stack([O.dot_product_scalar(cta[g * 2:g * 2 + 2], pts[g * 2:g * 2 + 2]).to_array() for g in range(2)])
    # 3 x 2 part product
    c3 = [x.mul(y) for x, y in zip(cta, ctb)]
    out["mul_3x2"] = np.stack([x.mul(y).to_array() for x, y in zip(c3, ctb)])
    # second multiplication strategy (mul.rs:369-418), with and without relinearization
    basis = list(par.moduli)
    for _ in range(3):
        basis.append(O.generate_prime(62, 2 * degree, basis[-1]))
    P = 1
    for q in basis[3:]:
        P *= q
    Q = ctx.modulus()
    rk = O.RelinearizationKey(sk, rng)
    out["rk_c0"], out["rk_c1"] = rk.ksk.arrays()
    out["basis"] = np.array(basis, dtype=np.uint64)
    m2 = O.Multiplicator(par, O.ScalingFactor.one(), O.ScalingFactor(P, Q), basis, O.ScalingFactor(t, P))
    out["strategy2"] = np.stack([m2.multiply(x, y).to_array() for x, y in zip(cta, ctb)])
    m2.enable_relinearization(rk)
    out["strategy2_relin"] = np.stack([m2.multiply(x, y).to_array() for x, y in zip(cta, ctb)])
    # key switch with a single-modulus key (level 2): base-2^31 decomposition
    ctx2 = par.context_at_level(2)
    frm = O.Poly.random(ctx2, O.POWER_BASIS, rng)
    k2 = O.KeySwitchingKey(sk, frm, 2, 2, rng)
    out["k2_c0"], out["k2_c1"] = k2.arrays()
    xin = np.stack([O.Poly.random(ctx2, O.POWER_BASIS, rng).c for _ in range(count)])
    out["k2_in"] = xin
    res = [k2.key_switch(O.Poly(ctx2, O.POWER_BASIS, x.copy())) for x in xin]
    out["k2_out"] = np.stack([np.stack([c0.c, c1.c]) for c0, c1 in res])
    # wire format of the first ciphertext's polynomials (Rq.coefficients, power basis, 62-bit packed)
    out["packed"] = np.stack([np.frombuffer(O.poly_to_rq_coefficients(p), dtype=np.uint8) for p in cta[0].c])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_n16_l3_wide.npz"), **out)


if __name__ == "__main__":
    main()
