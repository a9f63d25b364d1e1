"""Host-side pieces of the native round driver (csrc/create_proof.cu), checked on the CPU against the
Python mirror they replace: Merlin transcript operations on a serialised state, the dusk encodings
(compressed G1, 64-byte wide reduction), and the scalar side of the linearisation
(src/prover/linearization_poly.rs:75-105,136-225).  No kernel is launched."""
import ctypes
import random

import numpy as np
import pytest

from dusk_plonk_b200 import ffi, field, transcript, widgets

R = field.R_MOD
P = field.P_MOD


@pytest.fixture(scope="module")
def lib():
    return ffi.load_library()


def _state(tr):
    s = tr.strobe
    return np.frombuffer(bytes(s.state) + bytes([s.pos, s.pos_begin, s.cur_flags]), dtype=np.uint8).copy()


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def test_transcript_ops_match_python_merlin(lib):
    rng = random.Random(5)
    tr = transcript.Transcript(b"plonk")
    st = _state(tr)
    for step in range(60):
        label = bytes(rng.choice(b"abcdefgh_") for _ in range(rng.randrange(1, 30)))
        if rng.random() < 0.6:
            msg = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 32, 48, 165, 166, 167, 400])))
            tr.append_message(label, msg)
            m = np.frombuffer(msg, dtype=np.uint8).copy() if msg else np.zeros(1, dtype=np.uint8)
            assert lib.zkp_transcript_append(_ptr(st), label, _ptr(m), len(msg)) == 0
        else:
            n = rng.choice([1, 32, 64, 200])
            want = tr.challenge_bytes(label, n)
            out = np.zeros(n, dtype=np.uint8)
            assert lib.zkp_transcript_challenge(_ptr(st), label, _ptr(out), n) == 0
            assert bytes(out) == want
        assert bytes(st) == bytes(_state(tr)), step


def test_merlin_known_answer_through_native_ops(lib):
    """Merlin's published "test protocol" vector, driven through the native operations."""
    tr = transcript.MerlinTranscript(b"test protocol")
    st = _state(tr)
    msg = np.frombuffer(b"some data", dtype=np.uint8).copy()
    assert lib.zkp_transcript_append(_ptr(st), b"some label", _ptr(msg), len(msg)) == 0
    out = np.zeros(32, dtype=np.uint8)
    assert lib.zkp_transcript_challenge(_ptr(st), b"challenge", _ptr(out), 32) == 0
    assert bytes(out).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_transcript_init_is_merlin_new(lib):
    """zkp_transcript_init == merlin's Transcript::new(label): a host that keeps Merlin inside the library
    never needs the crate's (private) STROBE state.  The KAT again, with no Python Merlin in the loop."""
    st = np.zeros(203, dtype=np.uint8)
    lab = np.frombuffer(b"test protocol", dtype=np.uint8).copy()
    assert lib.zkp_transcript_init(_ptr(st), _ptr(lab), len(lab)) == 0
    assert bytes(st) == bytes(_state(transcript.MerlinTranscript(b"test protocol")))
    msg = np.frombuffer(b"some data", dtype=np.uint8).copy()
    assert lib.zkp_transcript_append(_ptr(st), b"some label", _ptr(msg), len(msg)) == 0
    out = np.zeros(32, dtype=np.uint8)
    assert lib.zkp_transcript_challenge(_ptr(st), b"challenge", _ptr(out), 32) == 0
    assert bytes(out).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    assert lib.zkp_transcript_init(None, _ptr(lab), 3) == ffi.ZKP_ERR_INVALID


def test_wide_reduction(lib):
    rng = random.Random(6)
    cases = [bytes(64), bytes([255]) * 64, R.to_bytes(32, "little") + bytes(32), bytes(32) + R.to_bytes(32, "little")]
    cases += [bytes(rng.randrange(256) for _ in range(64)) for _ in range(50)]
    for b in cases:
        out = np.zeros(4, dtype=np.uint64)
        assert lib.zkp_fr_from_wide(_ptr(np.frombuffer(b, dtype=np.uint8).copy()), _ptr(out)) == 0
        assert field.fr_from_mont(out.reshape(1, 4))[0] == int.from_bytes(b, "little") % R


def _g1_mont(pt):
    rm = (1 << 384) % P
    x, y = pt
    raw = (x * rm % P).to_bytes(48, "little") + (y * rm % P).to_bytes(48, "little")
    return np.frombuffer(raw, dtype=np.uint64).copy()


def test_g1_compress(lib):
    gx = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
    gy = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
    pts = [(gx, gy), (gx, P - gy)]
    for pt in pts:
        out = np.zeros(48, dtype=np.uint8)
        assert lib.zkp_g1_compress(_ptr(_g1_mont(pt)), _ptr(out)) == 0
        assert bytes(out) == transcript.g1_compress(pt)
        assert field.g1_decompress(bytes(out)) == pt
    out = np.zeros(48, dtype=np.uint8)
    assert lib.zkp_g1_compress(_ptr(np.zeros(12, dtype=np.uint64)), _ptr(out)) == 0
    assert bytes(out) == transcript.g1_compress(None)
    # the generator's compressed form is a published constant (zkcrypto bls12_381 G1 generator)
    out = np.zeros(48, dtype=np.uint8)
    lib.zkp_g1_compress(_ptr(_g1_mont((gx, gy))), _ptr(out))
    assert bytes(out).hex() == ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
                                "6c55e83ff97a1aeffb3af00adb22c6bb")


EVAL_ORDER = ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
              "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval",
              "q_l_eval", "q_r_eval", "perm_eval")


@pytest.mark.parametrize("k", [5, 9, 16, 20])
def test_linearization_scalars_match_widgets(lib, k):
    rng = random.Random(100 + k)
    for trial in range(8):
        if trial == 0:      # degenerate values exercise the subtractions that wrap
            ch = [1, 0, R - 1, 2, 3, R - 2, 5, 7]
            ev = {nm: (i * 17) % 5 for i, nm in enumerate(EVAL_ORDER)}
        else:
            ch = [rng.randrange(R) for _ in range(8)]
            ev = {nm: rng.randrange(R) for nm in EVAL_ORDER}
        want = widgets.linearization_scalars(1 << k, tuple(ch), ev)
        chm = field.fr_to_mont(ch)
        evm = field.fr_to_mont([ev[nm] for nm in EVAL_ORDER])
        out = np.zeros((12, 4), dtype=np.uint64)
        assert lib.zkp_linearization_scalars(k, _ptr(chm), _ptr(evm), _ptr(out)) == 0
        got = field.fr_from_mont(out)
        assert [nm for nm, _ in want] == ["q_m", "q_l", "q_r", "q_o", "q_d", "q_c", "q_range", "q_logic",
                                          "q_fixed_group_add", "q_variable_group_add", "z", "s_sigma_4"]
        assert got == [s for _, s in want]
