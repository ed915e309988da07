"""Small mul+relin + rotate at N = 2^13 through the TMA kernels, for compute-sanitizer runs (memcheck / racecheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import fhe_oracle as O
import fhe_rs_b200 as F

degree, L, count = 1 << 13, 3, 4
opar = O.BfvParameters(degree, 786433, moduli_sizes=[62] * L)
gpar = F.BfvParameters(degree, 786433, moduli=opar.moduli, device=0)
ctx = opar.context_at_level(0)
rng = np.random.default_rng(3)


def rnd(*prefix):
    a = np.zeros(tuple(prefix) + (L, degree), np.uint64)
    for i, q in enumerate(ctx.moduli):
        a[..., i, :] = rng.integers(0, q, size=tuple(prefix) + (degree,), dtype=np.uint64)
    return a


kc, a, b = rnd(2, L), rnd(count, 2), rnd(count, 2)
ork = O.RelinearizationKey.from_ksk(O.KeySwitchingKey.from_arrays(opar, kc[0], kc[1]))
grk = F.RelinearizationKey.from_arrays(gpar, kc[0], kc[1])
A, B = F.Ciphertext.from_host(gpar, a), F.Ciphertext.from_host(gpar, b)
P = F.Multiplicator.default(grk).multiply(A, B).to_host()
exp = O.Multiplicator.default(ork).multiply(O.Ciphertext.from_array(opar, a[0], 0), O.Ciphertext.from_array(opar, b[0], 0))
assert (P[0] == exp.to_array()).all()
R = F.GaloisKey.from_arrays(gpar, 3, kc[1], kc[0]).relinearize(A).to_host()
X = F.Ciphertext.from_host(gpar, a)
assert (X.into_power_basis().into_ntt().to_host() == a).all()
# power-basis substitute, message pack / unpack (device halves of to_bytes / from_bytes)
Y = F.Ciphertext.from_host(gpar, a).into_power_basis().substitute(3).into_ntt()
assert (Y.to_host() == F.Ciphertext.from_host(gpar, a).substitute(3).to_host()).all()
assert (F.Ciphertext.from_bytes(gpar, A.to_bytes()).to_host() == a).all()
print("sanitize probe ok", R.shape)
