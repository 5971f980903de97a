"""The C++ host mirror of the reference's entry points (block_b200/host: SpinBlock::multiplyH / diagonalH / RenormaliseFrom /
transform_operators, operatorfunctions::TensorMultiply, Linear::block_davidson) driven by tests/cpp/host_mirror_test.cpp on
records of the real reference."""
import os
import subprocess

import numpy as np
import pytest

from oracle import dmrg_oracle as O
from oracle import dumpio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "block_b200", "lib", "host_mirror_test")


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_cpp_mirror_builds_and_fails_loudly_without_a_device(tmp_path, golden):
    """-m 'not gpu': the driver links against libb2dhost.so / libblockb200.so and, on a box without CUDA, aborts with the
    library's message instead of silently computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the gpu test")
    assert os.path.exists(EXE), "run __graft_entry__.build()"
    rec, _ = golden
    inp = tmp_path / "in.bin"
    dumpio.write_records(inp, rec)
    res = subprocess.run([EXE, str(inp), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert res.returncode != 0
    assert "b2d_create failed" in res.stderr and "not available" in res.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_reference(tmp_path, golden):
    rec, big = golden
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    dumpio.write_records(inp, rec)
    res = subprocess.run([EXE, str(inp), str(outp), "0"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    out = dumpio.read_records(outp)
    nroots = int(rec["meta"][4])
    assert int(out["mask_mismatch"][0]) == 0                                  # SparseMatrix::allocate rule, bit-exact
    assert rel(out["sigma"], rec["rsigma"]) < 1e-10
    assert rel(out["diag"], rec["diag"]) < 1e-13
    c = big.unflatten(rec["rpsi"])
    v = big.zeros()
    O.tensor_multiply(big, O.View(big.left.get(O.HAM, (), 0)), O.View(big.right.get(O.OVERLAP, (), 0)), c, v, 0, 1.0)
    assert rel(out["tm_ham_left"], big.flatten(v)) < 1e-12
    if "tm_ccd_cre_t" in out:
        cre = next(o for o in big.right.ops if o.optype == O.CRE)
        ccd = big.left.get(O.CRE_CRE_DESCOMP, cre.orbs, 0)
        v = big.zeros()
        O.tensor_multiply(big, O.View(ccd), O.View(cre, True), c, v, 0, 1.0)
        assert rel(out["tm_ccd_cre_t"], big.flatten(v)) < 1e-12
    assert np.abs(out["dav_evals"] - rec["dav_evals"][:nroots]).max() < 1e-8
    assert np.abs(out["energies"] - rec["energies"][:nroots]).max() < 1e-8
    ref_rot = dumpio.rotation_from(rec)
    assert list(out["kept"]) == [r.shape[1] for r in ref_rot]
    assert abs(out["error"][0] - rec["error"][0]) < 1e-9
    N = dumpio.block_from(rec, "N.")
    assert list(out["N.dims"]) == list(N.dims)
