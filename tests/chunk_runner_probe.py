"""Run by tests/test_gpu_parity.py::test_chunk_runner_entry_points in a subprocess with a tiny FHE_B200_CHUNK: every chunked
entry point of the C ABI on a batch that spans several chunks (dealt over the side streams) must give, ciphertext by
ciphertext, what the same call gives on one-ciphertext batches (which are never split)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_rs_b200 as F  # noqa: E402

degree, nmod, t, count = 64, 3, 1153, 7
par = F.BfvParameters(degree, t, moduli_sizes=[62] * nmod, device=0)
moduli = par.moduli()
rng = np.random.default_rng(int(os.environ.get("FHE_B200_CHUNK", "0")) + 100)


def rnd(*prefix, limbs=nmod):
    a = np.zeros(tuple(prefix) + (limbs, degree), np.uint64)
    for i in range(limbs):
        a[..., i, :] = rng.integers(0, moduli[i], size=tuple(prefix) + (degree,), dtype=np.uint64)
    return a


a, b, kc, gc = rnd(count, 2), rnd(count, 2), rnd(2, nmod), rnd(2, nmod)
rk = F.RelinearizationKey.from_arrays(par, kc[0], kc[1])
gk = F.GaloisKey.from_arrays(par, 3, gc[0], gc[1])
ms = F.Multiplicator.default(rk)
msw = F.Multiplicator.default(rk)
msw.enable_mod_switching()
Q = 1
for q in moduli:
    Q *= q
custom = F.Multiplicator.new(F.ScalingFactor.one(), F.ScalingFactor.one(), par.mul_basis(0), F.ScalingFactor(t, Q), par)
custom_rk = F.Multiplicator.new(F.ScalingFactor.one(), F.ScalingFactor.one(), par.mul_basis(0), F.ScalingFactor(t, Q), par)
custom_rk.enable_relinearization(rk)

ops = {
    "mul_relin": lambda A, B: ms.multiply(A, B),
    "mul_relin + mod_switch": lambda A, B: msw.multiply(A, B),
    "ct * ct": lambda A, B: A * B,
    "relinearizes": lambda A, B: rk.relinearizes(A * B),
    "multiplicator (3 parts)": lambda A, B: custom.multiply(A, B),
    "multiplicator + key": lambda A, B: custom_rk.multiply(A, B),
    "galois": lambda A, B: gk.relinearize(A),
    "key_switch": lambda A, B: rk.ksk.key_switch(A.clone().into_power_basis(), part=1),
}
for name, op in ops.items():
    whole = op(F.Ciphertext.from_host(par, a), F.Ciphertext.from_host(par, b)).to_host()
    for i in range(count):
        one = op(F.Ciphertext.from_host(par, a[i:i + 1]), F.Ciphertext.from_host(par, b[i:i + 1])).to_host()
        assert (whole[i] == one[0]).all(), (name, i)
print("chunk runner probe ok", len(ops), "entry points,", count, "ciphertexts, chunk", os.environ.get("FHE_B200_CHUNK"),
      "streams", os.environ.get("FHE_B200_STREAMS", "2"))
