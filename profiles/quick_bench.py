"""Device-resident mul+relin (and rotate) rate at set C for the current environment switches: a fast A/B tool.
    [FHE_B200_...=...] python profiles/quick_bench.py [batch] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fhe_rs_b200 as F
from fhe_rs_b200._capi import check
from bench import fill_uniform, DEGREE, N_MODULI, PLAINTEXT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
L = F._capi.lib()
par = F.BfvParameters(DEGREE, PLAINTEXT, moduli_sizes=[62] * N_MODULI, device=0)
moduli = par.moduli()
A, Bt, out = F.Ciphertext(par, B, 2), F.Ciphertext(par, B, 2), F.Ciphertext(par, B, 2)
fill_uniform(torch, A, moduli, 1)
fill_uniform(torch, Bt, moduli, 2)
rng = np.random.default_rng(7)
kc = np.zeros((2, N_MODULI, N_MODULI, DEGREE), np.uint64)
for i, q in enumerate(moduli):
    kc[:, :, i, :] = rng.integers(0, q, size=(2, N_MODULI, DEGREE), dtype=np.uint64)
rk = F.RelinearizationKey.from_arrays(par, kc[0], kc[1])
gk = F.GaloisKey.from_arrays(par, 3, kc[1], kc[0])
res = {}
for name, fn in (("mul_relin", lambda: check(L.fhe_b200_mul_relin(A._h, Bt._h, rk.ksk._h, 0, out._h, None))),
                 ("rotate", lambda: check(L.fhe_b200_galois(A._h, 3, gk.ksk._h, out._h, None)))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    res[name] = round(B * steps / (e0.elapsed_time(e1) * 1e-3), 1)
cs = int(torch.as_tensor(__import__("bench").DevArray(out.device_ptr(), 2 * N_MODULI * DEGREE), device="cuda").sum().item())
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("FHE_B200_")}, "batch": B, **res, "checksum": cs}))
