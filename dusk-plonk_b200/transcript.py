"""Fiat-Shamir transcript of the prover (host side; never on the GPU).

The reference drives a ``zksnarks::plonk::Transcript`` through ``TranscriptProtocol``
(``append_scalar`` / ``append_commitment`` / ``challenge_scalar``; call sites and labels
``src/prover.rs:99-105,139-199,203-226,268-295,321-405,435-450``, ``Transcript::base`` at
``src/prover.rs:54-55``).  That crate is absent from the reference tree; upstream dusk-plonk
0.13 builds it on Merlin (STROBE-128 over Keccak-f[1600]).  This module restates Merlin
from its published specification -- pinned by Merlin's own "test protocol" known-answer
vector in ``tests/test_oracle_plonk.py`` -- and the dusk encodings on top of it
([EXT-RECALL]: scalars as 32-byte little-endian canonical, commitments as 48-byte
compressed G1, challenges as 64 squeezed bytes reduced mod r).  The prover takes the
transcript as an object, so a Rust host keeps its own and nothing on the device depends
on these encodings.
"""

import ctypes

_R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
_P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_M64 = (1 << 64) - 1

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & _M64 if n else v


def keccak_f1600(state: bytearray) -> None:
    """In place on a 200-byte state.  Uses the C routine of libzkp_b200.so (host code) when the
    library is built; the pure-Python body below is the specification it is tested against."""
    fn = _native_keccak()
    if fn is not None:
        buf = (ctypes.c_char * 200).from_buffer(state)
        fn(ctypes.addressof(buf))
        return
    keccak_f1600_py(state)


_NATIVE = [False]


def _native_keccak():
    if _NATIVE[0] is False:
        try:
            from .ffi import load_library
            _NATIVE[0] = load_library().zkp_keccak_f1600
        except Exception:
            _NATIVE[0] = None
    return _NATIVE[0]


def keccak_f1600_py(state: bytearray) -> None:
    a = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = a[x][y].to_bytes(8, "little")


class Strobe128:
    R = 166
    FLAG_I, FLAG_A, FLAG_C, FLAG_T, FLAG_M, FLAG_K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state = st
        self.pos = 0
        self.pos_begin = 0
        self.cur_flags = 0
        self.meta_ad(protocol_label, False)

    def clone(self):
        c = object.__new__(Strobe128)
        c.state = bytearray(self.state)
        c.pos, c.pos_begin, c.cur_flags = self.pos, self.pos_begin, self.cur_flags
        return c

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for byte in data:
            self.state[self.pos] ^= byte
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray(n)
        for i in range(n):
            out[i] = self.state[self.pos]
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & self.FLAG_T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.FLAG_C | self.FLAG_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(self.FLAG_M | self.FLAG_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(self.FLAG_A, more)
        self._absorb(data)

    def prf(self, n, more):
        self._begin_op(self.FLAG_I | self.FLAG_A | self.FLAG_C, more)
        return self._squeeze(n)


class MerlinTranscript:
    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def clone(self):
        c = object.__new__(type(self))
        c.__dict__.update(self.__dict__)
        c.strobe = self.strobe.clone()
        return c

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(len(message).to_bytes(4, "little"), True)
        self.strobe.ad(message, False)

    def append_u64(self, label: bytes, x: int):
        self.append_message(label, x.to_bytes(8, "little"))

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(n.to_bytes(4, "little"), True)
        return self.strobe.prf(n, False)


def g1_compress(pt) -> bytes:
    """48-byte compressed G1 (zcash/dusk ``G1Affine::to_bytes``): big-endian x, bit 7 =
    compressed, bit 6 = infinity, bit 5 = y is the lexicographically larger root."""
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > _P_MOD - y:
        b[0] |= 0x20
    return bytes(b)


class Transcript(MerlinTranscript):
    """``TranscriptProtocol`` over canonical ints / affine points ((x, y) or None)."""

    def append_scalar(self, label: bytes, s: int):
        self.append_message(label, (s % _R_MOD).to_bytes(32, "little"))

    def append_commitment(self, label: bytes, pt):
        self.append_message(label, g1_compress(pt))

    def challenge_scalar(self, label: bytes) -> int:
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % _R_MOD

    def circuit_domain_sep(self, n: int):
        self.append_message(b"dom-sep", b"circuit_size")
        self.append_u64(b"n", n)

    @classmethod
    def base(cls, label: bytes, vk_commitments, constraints: int):
        """``Transcript::base(label, &verifier_key, constraints)`` (src/prover.rs:54-55,
        src/verifier.rs:33-34).  ``vk_commitments`` is the ordered (label, point) list the
        verification key seeds the transcript with."""
        t = cls(label)
        t.circuit_domain_sep(constraints)
        for lab, pt in vk_commitments:
            t.append_commitment(lab, pt)
        t.circuit_domain_sep(constraints)
        return t
