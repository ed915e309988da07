"""Times the batched NTT (forward + inverse through the C ABI) for one kernel family.

    FHE_B200_NTT=tma|fast python profiles/ntt_bench.py [--shape B|C] [--reps 20]

Shape B = BASELINE configs[1] ([256][8][2^14]); shape C = 64 ciphertexts x 2 x 14 rows of N = 2^15.
Prints one JSON line: microseconds per limb-NTT (forward, inverse) and algorithmic GB/s (16*N bytes per limb-NTT)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="B")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import fhe_rs_b200 as F
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
    from bench import fill_uniform
    degree, L, count, parts = (1 << 14, 8, 256, 1) if a.shape == "B" else (1 << 15, 14, 64, 2)
    par = F.BfvParameters(degree, 786433, moduli_sizes=[62] * L, device=0)
    X = F.Ciphertext(par, count, parts, repr=F.POWER_BASIS)
    fill_uniform(torch, X, par.moduli(), 3)
    ref = X.clone()
    for _ in range(3):
        X.into_ntt(); X.into_power_basis()
    torch.cuda.synchronize()
    ok = bool((torch.as_tensor(_dev(X), device="cuda") == torch.as_tensor(_dev(ref), device="cuda")).all())
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = ti = 0.0
    for _ in range(a.reps):
        ev[0].record(); X.into_ntt(); ev[1].record(); X.into_power_basis(); ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); ti += ev[1].elapsed_time(ev[2])
    rows = count * parts * L
    tf, ti = tf / a.reps, ti / a.reps
    gbs = lambda ms: 16.0 * degree * rows / (ms * 1e-3) / 1e9
    print(json.dumps({"family": os.environ.get("FHE_B200_NTT", "auto"), "shape": a.shape, "rows": rows, "N": degree,
                      "roundtrip_identity": ok, "fwd_us_per_ntt": 1e3 * tf / rows, "inv_us_per_ntt": 1e3 * ti / rows,
                      "fwd_gbs": gbs(tf), "inv_gbs": gbs(ti), "both_gbs": gbs((tf + ti) / 2)}))


def _dev(ct):
    from bench import DevArray
    c, p, l, n = ct.shape()
    return DevArray(ct.device_ptr(), c * p * l * n)


if __name__ == "__main__":
    main()
