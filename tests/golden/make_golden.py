"""Generates tests/golden/golden_n16_l3.npz from the CPU oracle (the Rust reference cannot run in
this image -- no cargo/rustc -- so fixtures come from the oracle that tests/test_oracle_pinning.py
pins against the reference's own KATs and property tests).

    python tests/golden/make_golden.py

Deterministic: fixed numpy seed; psi = documented default root per prime."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fhe_oracle as O  # noqa: E402


def main():
    degree, t, nmod = 16, 1153, 3
    rng = np.random.default_rng(20260924)
    par = O.BfvParameters(degree, t, moduli_sizes=[62] * nmod)
    sk = O.SecretKey(par, rng)
    count = 4
    cta = [sk.encrypt(rng.integers(0, t, degree), 0, rng) for _ in range(count)]
    ctb = [sk.encrypt(rng.integers(0, t, degree), 0, rng) for _ in range(count)]
    rk = O.RelinearizationKey(sk, rng)
    gk = O.GaloisKey(sk, 3, rng)
    m = O.Multiplicator.default(rk)
    out = dict(degree=degree, t=t, moduli=np.array(par.moduli, dtype=np.uint64),
               psi=np.array([O.default_psi(q, degree) for q in par.moduli + par.extended_basis], dtype=np.uint64),
               sk=sk.coeffs,
               a=np.stack([c.to_array() for c in cta]), b=np.stack([c.to_array() for c in ctb]))
    out["rk_c0"], out["rk_c1"] = rk.ksk.arrays()
    out["gk_c0"], out["gk_c1"] = gk.ksk.arrays()
    out["add"] = np.stack([x.add(y).to_array() for x, y in zip(cta, ctb)])
    out["mul3"] = np.stack([x.mul(y).to_array() for x, y in zip(cta, ctb)])
    out["mul_relin"] = np.stack([m.multiply(x, y).to_array() for x, y in zip(cta, ctb)])
    m.enable_mod_switching()
    out["mul_relin_ms"] = np.stack([m.multiply(x, y).to_array() for x, y in zip(cta, ctb)])
    out["galois3"] = np.stack([gk.relinearize(x).to_array() for x in cta])
    out["a_pb"] = np.stack([np.stack([p.copy().into_power_basis().c for p in x.c]) for x in cta])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_n16_l3.npz"), **out)


if __name__ == "__main__":
    main()
