"""Small driver for ncu captures: runs the hot path once (or a selected primitive) at the north-star size.
    python profiles/probe.py ntt|mulrelin [count]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import fhe_rs_b200 as F
from bench import fill_uniform, DEGREE, N_MODULI, PLAINTEXT

what = sys.argv[1] if len(sys.argv) > 1 else "mulrelin"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if what == "ntt14":  # BASELINE config 2: [256][8][2^14] forward + inverse (the bench.py roofline shape)
    par = F.BfvParameters(1 << 14, PLAINTEXT, moduli_sizes=[62] * 8, device=0)
    X = F.Ciphertext(par, 256, 1, repr=F.POWER_BASIS)
    fill_uniform(torch, X, par.moduli(), 3)
    for _ in range(3):
        X.into_ntt()
        X.into_power_basis()
    torch.cuda.synchronize()
    print("probe done ntt14")
    sys.exit(0)
par = F.BfvParameters(DEGREE, PLAINTEXT, moduli_sizes=[62] * N_MODULI, device=0)
moduli = par.moduli()
A = F.Ciphertext(par, count, 2)
fill_uniform(torch, A, moduli, 1)
if what == "rotate":  # BASELINE config 4: GaloisKey rotate (exponent 3)
    rng = np.random.default_rng(9)
    gc = np.zeros((2, N_MODULI, N_MODULI, DEGREE), np.uint64)
    for i, q in enumerate(moduli):
        gc[:, :, i, :] = rng.integers(0, q, size=(2, N_MODULI, DEGREE), dtype=np.uint64)
    gk = F.GaloisKey.from_arrays(par, 3, gc[0], gc[1])
    for _ in range(2):
        out = gk.relinearize(A)
    out.sync()
elif what == "ntt":
    for _ in range(2):
        A.into_power_basis()
        A.into_ntt()
else:
    B = F.Ciphertext(par, count, 2)
    fill_uniform(torch, B, moduli, 2)
    rng = np.random.default_rng(7)
    kc = np.zeros((2, N_MODULI, N_MODULI, DEGREE), np.uint64)
    for i, q in enumerate(moduli):
        kc[:, :, i, :] = rng.integers(0, q, size=(2, N_MODULI, DEGREE), dtype=np.uint64)
    rk = F.RelinearizationKey.from_arrays(par, kc[0], kc[1])
    m = F.Multiplicator.default(rk)
    for _ in range(2):
        out = m.multiply(A, B)
    out.sync()
torch.cuda.synchronize()
print("probe done", what, count)
