"""The C-ABI library loads here (no GPU) and exports every symbol include/zkp_b200.h
declares; the product has no CPU fallback."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    out = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            out |= set(re.findall(r"\b(zkp_[a-z0-9_]+)\s*\(", txt))
    return out


def test_library_exports_every_declared_symbol():
    import dusk_plonk_b200 as z
    lib = z.load_library()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libzkp_b200.so does not export %s" % s
    # and the Python binding covers the header exactly
    assert syms == set(z.SIGNATURES), syms ^ set(z.SIGNATURES)


def test_host_side_constants_match_oracle():
    import dusk_plonk_b200 as z
    from oracle.fields import R_MOD, domain_generator, fr_from_mont_limbs
    for k in (0, 1, 9, 12, 19, 26, 32):
        w = domain_generator(k)
        assert fr_from_mont_limbs(z.fft_constant(k, 0)) == [w]
        assert fr_from_mont_limbs(z.fft_constant(k, 1)) == [pow(w, -1, R_MOD)]
        assert fr_from_mont_limbs(z.fft_constant(k, 2)) == [pow(1 << k, -1, R_MOD)]
    assert fr_from_mont_limbs(z.fft_constant(5, 3)) == [7]
    assert fr_from_mont_limbs(z.fft_constant(5, 4)) == [pow(7, -1, R_MOD)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked tests")
    import dusk_plonk_b200 as z
    with pytest.raises(z.ZkpError):
        z.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dusk-plonk_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "zkp_oracle" not in txt, f


def test_round_driver_entry_points_reject_bad_arguments_without_a_gpu():
    """Argument checks of the native round driver happen before any CUDA call: ZKP_ERR_INVALID, not a crash."""
    import ctypes
    import dusk_plonk_b200 as z
    lib = z.load_library()
    out = ctypes.c_void_p()
    assert lib.zkp_prover_create(None, None, None, ctypes.byref(out)) == z.ZKP_ERR_INVALID
    assert lib.zkp_prover_prove(None, None, None, None, None, None, None, None, None, None, None) == z.ZKP_ERR_INVALID
    assert lib.zkp_prover_prove_witness(None, None, None, 0, None, None, None, None, None, None) == z.ZKP_ERR_INVALID
    assert lib.zkp_prover_set_wiring(None, None, 0, None, 0) == z.ZKP_ERR_INVALID
    assert lib.zkp_prover_destroy(None) == z.ZKP_OK
    assert lib.zkp_poly_eval2_dev(None, None, None, 0, None, None) == z.ZKP_ERR_INVALID
    assert lib.zkp_transcript_append(None, b"x", None, 0) == z.ZKP_ERR_INVALID
    assert lib.zkp_linearization_scalars(5, None, None, None) == z.ZKP_ERR_INVALID
    assert lib.zkp_g1_compress(None, None) == z.ZKP_ERR_INVALID


def test_multi_gpu_and_verifier_entry_points_reject_bad_arguments_without_a_gpu():
    """The round-2 entry points (zkp_comm, sharded commit / prover, coset transforms, verifier glue) check their
    arguments before any CUDA or NCCL call: ZKP_ERR_INVALID, not a crash -- and the library loads and answers
    on a box with neither a GPU nor an NCCL communicator."""
    import ctypes
    import numpy as np
    import dusk_plonk_b200 as z
    lib = z.load_library()
    out = ctypes.c_void_p()
    assert lib.zkp_comm_create(None, None, 0, 1, ctypes.byref(out)) == z.ZKP_ERR_INVALID
    assert lib.zkp_comm_destroy(None) == z.ZKP_OK
    assert lib.zkp_comm_rank(None) == 0 and lib.zkp_comm_size(None) == 1
    assert lib.zkp_comm_all_to_all_dev(None, None, 0, None, 0, 0) == z.ZKP_ERR_INVALID
    assert lib.zkp_commit_batch_sharded_dev(None, None, None, None, 0, None, None) == z.ZKP_ERR_INVALID
    assert lib.zkp_coset8_ntt_dev(None, None, 0, 0, None, 0, 3, 0, 8) == z.ZKP_ERR_INVALID
    assert lib.zkp_prover_create_sharded(None, None, None, None, ctypes.byref(out)) == z.ZKP_ERR_INVALID
    assert lib.zkp_twiddle_transpose_dev(None, None, 0, None, 0, 1, 1, 0, 4, 0) == z.ZKP_ERR_INVALID
    assert lib.zkp_srs_trim(None, None, 0, ctypes.byref(out)) == z.ZKP_ERR_INVALID
    assert lib.zkp_buf_upload_2d(None, None, 0, None, 0, 0, 0) == z.ZKP_ERR_INVALID
    assert lib.zkp_verify(None, None, None, None, None, None, None, 0) == z.ZKP_ERR_INVALID
    assert lib.zkp_kzg_batch_check(None, None, None, None, None, 0, None) == z.ZKP_ERR_INVALID
    assert lib.zkp_g2_generator_mul(None, None) == z.ZKP_ERR_INVALID
    # a G2 element that is not on the twist is an invalid argument, not a rejected proof
    bad = np.zeros(24, dtype=np.uint64); bad[3] = 9
    g1 = np.zeros(12, dtype=np.uint64)
    assert lib.zkp_pairing_check(ctypes.c_void_p(g1.ctypes.data), ctypes.c_void_p(bad.ctypes.data), 1) == z.ZKP_ERR_INVALID
    assert lib.zkp_pairing_check(None, None, 0) == z.ZKP_OK          # the empty product is one
    assert lib.zkp_strerror(z.ZKP_ERR_VERIFY) == b"proof rejected"
