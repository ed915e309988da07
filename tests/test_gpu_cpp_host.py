"""Compiles tests/cpp/host_api_test.cpp against include/fhe_b200.hpp + libfhe_b200.so and runs it on the GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _records(msgs):
    return b"".join(struct.pack("<I", len(m)) + m for m in msgs)


def _read_records(path):
    data, pos, out = open(path, "rb").read(), 0, []
    while pos < len(data):
        (n,) = struct.unpack_from("<I", data, pos)
        out.append(data[pos + 4: pos + 4 + n])
        pos += 4 + n
    return out


@pytest.mark.parametrize("degree,nmod", [(64, 3), (8192, 2)])
def test_cpp_host_api(tmp_path, oracle, degree, nmod):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the reference's protobuf messages, produced by the CPU oracle (oracle/fhe_wire.py), for the C++ host to consume
    import fhe_wire as ow
    rng = np.random.default_rng(70 + nmod)
    opar = oracle.BfvParameters(degree, 1153, moduli_sizes=[62] * nmod)
    sk = oracle.SecretKey(opar, rng)
    cts = [sk.encrypt(rng.integers(0, 1153, degree), 0, rng) for _ in range(3)]
    ork, ogk = oracle.RelinearizationKey(sk, rng), oracle.GaloisKey(sk, 3, rng)
    msgs = [ow.ciphertext_to_bytes(c) for c in cts]
    (tmp_path / "cts.bin").write_bytes(_records(msgs))
    (tmp_path / "cts_seeded.bin").write_bytes(_records([ow.ciphertext_to_bytes(c, seed=b"s" * 32) for c in cts]))
    (tmp_path / "halves.bin").write_bytes(_records([np.stack([c.c[1].c for c in cts]).tobytes()]))
    (tmp_path / "relin.bin").write_bytes(_records([ow.relin_key_to_bytes(ork)]))
    (tmp_path / "galois.bin").write_bytes(_records([ow.galois_key_to_bytes(ogk)]))
    exe = str(tmp_path / "host_api_test")
    lib_dir = os.path.join(ROOT, "fhe_rs_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp"), "-o", exe,
                           "-L", lib_dir, "-lfhe_b200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe, str(degree), str(nmod), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
    words = np.fromfile(str(tmp_path / "out_words.bin"), dtype=np.uint64)
    assert (words == np.stack([c.to_array() for c in cts]).ravel()).all()
    assert _read_records(str(tmp_path / "out_msgs.bin")) == msgs
    om = oracle.Multiplicator.default(ork)
    assert _read_records(str(tmp_path / "out_mul.bin")) == [ow.ciphertext_to_bytes(om.multiply(c, c)) for c in cts]
    assert _read_records(str(tmp_path / "out_rot.bin")) == [ow.ciphertext_to_bytes(ogk.relinearize(c)) for c in cts]
