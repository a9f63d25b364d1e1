"""TEST INFRASTRUCTURE -- the oracle's OWN Fiat-Shamir transcript (never imported by the product).

Independent restatement of Merlin (STROBE-128 over Keccak-f[1600]) and of the dusk
``TranscriptProtocol`` encodings, so that the oracle side of every proof-parity test derives its
challenges with code that shares nothing with ``dusk-plonk_b200/transcript.py`` or the C++ Merlin
of ``csrc/create_proof.cu``.  Reference call sites: ``Transcript::base`` at ``src/prover.rs:54-55`` /
``src/verifier.rs:33-34``; labels and order ``src/prover.rs:99-105,139-199,203-226,268-295,321-405,
435-450``.  The ``zksnarks`` crate that implements it is absent from the reference tree ([EXT-RECALL]:
upstream dusk-plonk 0.13 = merlin 3 + 32-byte LE scalars, 48-byte compressed G1, 64-byte wide
challenges).

Pins (tests/test_oracle_merlin.py): the permutation against ``hashlib.sha3_256`` / ``shake_128`` (a
sponge built on this module's Keccak-f must reproduce Python's own SHA-3), Merlin's published
"test protocol" known-answer vector, and agreement with the product's two transcripts.

Structure differs from the product's on purpose: the state is 25 lanes (ints), rho/pi come from the
(t+1)(t+2)/2 walk of the Keccak specification rather than from a table, and STROBE is a single
``_operate`` routine driven by flag bits.
"""

MASK = (1 << 64) - 1
FR = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FQ = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def _round_constants():
    """iota constants from the degree-8 LFSR of the Keccak specification."""
    out, reg = [], 1
    for _ in range(24):
        rc = 0
        for j in range(7):
            if reg & 1:
                rc |= 1 << ((1 << j) - 1)
            reg <<= 1
            if reg & 0x100:
                reg ^= 0x171
        out.append(rc)
    return out


def _rho_pi():
    """(source lane, destination lane, rotation) triples: the walk (x, y) -> (y, 2x + 3y) starting at
    (1, 0) with rotation offsets (t + 1)(t + 2) / 2."""
    steps, x, y = [], 1, 0
    for t in range(24):
        nx, ny = y, (2 * x + 3 * y) % 5
        steps.append((x + 5 * y, nx + 5 * ny, ((t + 1) * (t + 2) // 2) % 64))
        x, y = nx, ny
    return steps


_IOTA = _round_constants()
_WALK = _rho_pi()


def keccak_f(lanes):
    """Keccak-f[1600] on a list of 25 little-endian 64-bit lanes (index x + 5 y); returns a new list."""
    s = list(lanes)
    for rc in _IOTA:
        col = [s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20] for x in range(5)]
        for x in range(5):
            r = col[(x + 1) % 5]
            d = col[(x + 4) % 5] ^ (((r << 1) | (r >> 63)) & MASK)
            for y in range(0, 25, 5):
                s[x + y] ^= d
        t = list(s)
        for src, dst, rot in _WALK:
            v = s[src]
            t[dst] = ((v << rot) | (v >> (64 - rot))) & MASK if rot else v
        for y in range(0, 25, 5):
            row = t[y:y + 5]
            for x in range(5):
                s[y + x] = row[x] ^ (~row[(x + 1) % 5] & MASK & row[(x + 2) % 5])
        s[0] ^= rc
    return s


def _permute_bytes(buf):
    lanes = [int.from_bytes(buf[8 * i:8 * i + 8], "little") for i in range(25)]
    out = keccak_f(lanes)
    for i, v in enumerate(out):
        buf[8 * i:8 * i + 8] = v.to_bytes(8, "little")


def sponge(rate, suffix, data, outlen):
    """Plain Keccak sponge (used only to pin ``keccak_f`` against hashlib)."""
    st = bytearray(200)
    data = bytes(data) + bytes([suffix])
    data += bytes(-len(data) % rate)
    data = bytearray(data)
    data[-1] |= 0x80
    for off in range(0, len(data), rate):
        for i in range(rate):
            st[i] ^= data[off + i]
        _permute_bytes(st)
    out = b""
    while len(out) < outlen:
        out += bytes(st[:rate])
        if len(out) < outlen:
            _permute_bytes(st)
    return out[:outlen]


# STROBE flag bits
F_I, F_A, F_C, F_T, F_M, F_K = 1, 2, 4, 8, 16, 32
RATE = 166   # STROBE-128: 200 - 128 / 4 - 2


class Strobe:
    def __init__(self, protocol):
        self.st = bytearray(200)
        self.st[:18] = bytes([1, RATE + 2, 1, 0, 1, 12 * 8]) + b"STROBEv1.0.2"
        _permute_bytes(self.st)
        self.pos = self.begin = self.flags = 0
        self._operate(F_M | F_A, protocol, None, False)

    def copy(self):
        c = Strobe.__new__(Strobe)
        c.st, c.pos, c.begin, c.flags = bytearray(self.st), self.pos, self.begin, self.flags
        return c

    def _f(self):
        self.st[self.pos] ^= self.begin
        self.st[self.pos + 1] ^= 4
        self.st[RATE + 1] ^= 0x80
        _permute_bytes(self.st)
        self.pos = self.begin = 0

    def _duplex(self, data, squeeze):
        """absorb ``data`` (bytes) or squeeze ``squeeze`` bytes (overwriting the rate with zeros)."""
        out = bytearray()
        for i in range(len(data) if data is not None else squeeze):
            if data is not None:
                self.st[self.pos] ^= data[i]
            else:
                out.append(self.st[self.pos])
                self.st[self.pos] = 0
            self.pos += 1
            if self.pos == RATE:
                self._f()
        return bytes(out)

    def _operate(self, flags, data, squeeze, more):
        if more:
            if flags != self.flags:
                raise ValueError("continued operation with different flags")
        else:
            if flags & F_T:
                raise ValueError("transport operations are not used by Merlin")
            prev = self.begin
            self.begin = self.pos + 1
            self.flags = flags
            self._duplex(bytes([prev, flags]), None)
            if flags & (F_C | F_K) and self.pos:
                self._f()
        return self._duplex(data, squeeze)

    def meta_ad(self, data, more=False):
        self._operate(F_M | F_A, data, None, more)

    def ad(self, data, more=False):
        self._operate(F_A, data, None, more)

    def prf(self, n):
        return self._operate(F_I | F_A | F_C, None, n, False)


class Merlin:
    def __init__(self, label):
        self.s = Strobe(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def clone(self):
        c = type(self).__new__(type(self))
        c.s = self.s.copy()
        return c

    def append_message(self, label, msg):
        self.s.meta_ad(label)
        self.s.meta_ad(len(msg).to_bytes(4, "little"), more=True)
        self.s.ad(msg)

    def challenge_bytes(self, label, n):
        self.s.meta_ad(label)
        self.s.meta_ad(n.to_bytes(4, "little"), more=True)
        return self.s.prf(n)


def compress_g1(pt):
    """48-byte big-endian x with the three flag bits of the zcash encoding (compressed, infinity,
    y > -y)."""
    if pt is None:
        return b"\xc0" + bytes(47)
    x, y = pt
    top = 0x80 | (0x20 if 2 * y > FQ else 0)
    raw = x.to_bytes(48, "big")
    return bytes([raw[0] | top]) + raw[1:]


class Transcript(Merlin):
    """``TranscriptProtocol`` over canonical ints / affine points ((x, y) or None)."""

    def append_scalar(self, label, s):
        self.append_message(label, (s % FR).to_bytes(32, "little"))

    def append_commitment(self, label, pt):
        self.append_message(label, compress_g1(pt))

    def challenge_scalar(self, label):
        return int.from_bytes(self.challenge_bytes(label, 64), "little") % FR

    def _domain_sep(self, constraints):
        self.append_message(b"dom-sep", b"circuit_size")
        self.append_message(b"n", constraints.to_bytes(8, "little"))

    @classmethod
    def base(cls, label, vk_commitments, constraints):
        t = cls(label)
        t._domain_sep(constraints)
        for lab, pt in vk_commitments:
            t.append_commitment(lab, pt)
        t._domain_sep(constraints)
        return t
