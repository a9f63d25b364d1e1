"""CPU: the oracle's own Merlin (oracle/merlin.py) against independent pins -- Python's hashlib SHA-3
(same Keccak-f), Merlin's published "test protocol" vector -- and against the product's transcript
(pure-Python body, no .so needed), so the two sides of every proof-parity test hash with unrelated code."""
import hashlib

from oracle import merlin
from oracle.rng import SplitMix64


def test_keccak_f_reproduces_hashlib_sha3():
    for n in (0, 1, 135, 136, 137, 500):
        data = bytes((7 * i + n) & 255 for i in range(n))
        assert merlin.sponge(136, 0x06, data, 32) == hashlib.sha3_256(data).digest()
        assert merlin.sponge(168, 0x1F, data, 300) == hashlib.shake_128(data).digest(300)


def test_round_constants_and_rotations_are_the_standard_ones():
    assert merlin._IOTA[0] == 1 and merlin._IOTA[1] == 0x8082 and merlin._IOTA[23] == 0x8000000080008008
    rot = {dst: r for _, dst, r in merlin._WALK}
    assert len(rot) == 24 and sorted(r for r in rot.values())[:3] == [1, 2, 3]


def test_merlin_published_known_answer():
    # merlin crate, transcript.rs `equivalence_simple` conformance vector
    t = merlin.Merlin(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == \
        "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_agrees_with_the_product_transcript():
    from dusk_plonk_b200 import transcript as ptr
    rng = SplitMix64(5)
    a, b = merlin.Transcript(b"plonk"), ptr.Transcript(b"plonk")
    pts = [None, (rng.fr(), rng.fr() | 1 << 300), (3, merlin.FQ - 2)]
    for i in range(40):
        if i % 3 == 0:
            s = rng.fr()
            a.append_scalar(b"s%d" % i, s); b.append_scalar(b"s%d" % i, s)
        elif i % 3 == 1:
            a.append_commitment(b"c", pts[i % len(pts)]); b.append_commitment(b"c", pts[i % len(pts)])
        else:
            assert a.challenge_scalar(b"ch") == b.challenge_scalar(b"ch")
    c = a.clone()
    assert c.challenge_bytes(b"x", 200) == a.challenge_bytes(b"x", 200)
    vk = [(b"q_m", pts[1]), (b"q_l", None)]
    assert merlin.Transcript.base(b"demo", vk, 287).challenge_scalar(b"z") == \
        ptr.Transcript.base(b"demo", vk, 287).challenge_scalar(b"z")
