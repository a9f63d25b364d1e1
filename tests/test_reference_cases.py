"""CPU: the case tables of the reference's integration tests (tests/boolean.rs, tests/logic.rs,
tests/decomposition.rs, tests/range.rs), run through the host composer.  The reference proves every positive
case and expects create_proof to fail on every negative one; here each case is decided by the gate identities
(every assert_* of the reference is a gate, so satisfiability is exactly "no row violated") and one negative
per file is also taken through the restated prover, which has to refuse it the way the reference does."""
import pytest

from host_mirror.composer import Plonk, SynthesizedCircuit, JUBJUB_GENERATOR, jubjub_mul
from oracle import plonk
from oracle.fields import R_MOD
from oracle.rng import SplitMix64

from test_oracle_plonk import rows_violated, setup


def logic_circuit(a, b, c, bits, xor):
    """tests/logic.rs:57-70 / :214-227: c == a (AND | XOR) b over `bits` bits."""
    cs = Plonk.initialize()
    wa, wb, wc = cs.append_witness(a), cs.append_witness(b), cs.append_witness(c)
    wx = (cs.append_logic_xor if xor else cs.append_logic_and)(wa, wb, bits)
    cs.assert_equal(wc, wx)
    return cs


def logic_expected(a, b, bits, xor):
    # the gadget reads the low `bits` bits most-significant first and consumes whole quads (src/lib.rs:291-304):
    # with an odd count the last bit is left over, so the result is the window shifted down by one
    window = ((a ^ b) if xor else (a & b)) & ((1 << bits) - 1)
    return window >> (bits & 1)


@pytest.mark.parametrize("xor", [False, True])
@pytest.mark.parametrize("bits", [256, 30, 0, 55])
def test_logic_cases(bits, xor):
    """tests/logic.rs:78-176, :235-333: default (256), small (30), zero and odd (55) bit counts."""
    rng = SplitMix64(1000 + bits + xor)
    a, b = rng.fr(), rng.fr()
    good = logic_circuit(a, b, logic_expected(a, b, bits, xor), bits, xor)
    assert rows_violated(good) == []
    assert good.m() == 6 + (min(bits, 256) >> 1) + 1 + 1           # one gate per quad + the closing row + assert
    if bits >= 2:
        # tests/logic.rs:92-112: c computed from another operand does not satisfy the circuit
        m = rng.fr()
        wrong = logic_expected(a, m, bits, xor)
        assert wrong != logic_expected(a, b, bits, xor)
        assert rows_violated(logic_circuit(a, b, wrong, bits, xor)) != []


def boolean_circuit(a):
    """tests/boolean.rs:45-54."""
    cs = Plonk.initialize()
    cs.component_boolean(cs.append_witness(a))
    return cs


def test_boolean_cases():
    """tests/boolean.rs:61-92: 0 and 1 pass, 2 fails."""
    assert rows_violated(boolean_circuit(0)) == []
    assert rows_violated(boolean_circuit(1)) == []
    assert rows_violated(boolean_circuit(2)) != []
    assert rows_violated(boolean_circuit(R_MOD - 1)) != []


def select_circuit(bit, a, b, res, zero_bit, zero_a, zero_res, one_bit, one_a, one_res,
                   point_bit, point_a, point_b, point_res, identity_bit, identity_a, identity_res):
    """tests/boolean.rs:230-275: the five select gadgets, each asserted against a witness."""
    cs = Plonk.initialize()
    w = cs.append_witness
    x = cs.component_select(w(bit), w(a), w(b))
    cs.assert_equal(x, w(res))
    x = cs.component_select_zero(w(zero_bit), w(zero_a))
    cs.assert_equal(x, w(zero_res))
    x = cs.component_select_one(w(one_bit), w(one_a))
    cs.assert_equal(x, w(one_res))
    x = cs.component_select_point(w(point_bit), cs.append_point(point_a), cs.append_point(point_b))
    cs.assert_equal_point(x, cs.append_point(point_res))
    x = cs.component_select_identity(w(identity_bit), cs.append_point(identity_a))
    cs.assert_equal_point(x, cs.append_point(identity_res))
    return cs


def select_honest(bit, rng):
    """DummyCircuit::new of tests/boolean.rs:128-188 with one bit for all five gadgets."""
    a, b, za, oa = rng.fr(), rng.fr(), rng.fr(), rng.fr()
    pa, pb, ia = (jubjub_mul(JUBJUB_GENERATOR, rng.fr()) for _ in range(3))
    return dict(bit=bit, a=a, b=b, res=a if bit else b, zero_bit=bit, zero_a=za, zero_res=za if bit else 0,
                one_bit=bit, one_a=oa, one_res=oa if bit else 1, point_bit=bit, point_a=pa, point_b=pb,
                point_res=pa if bit else pb, identity_bit=bit, identity_a=ia, identity_res=ia if bit else (0, 1))


def test_select_cases():
    """tests/boolean.rs:282-460: both bits pass; each gadget's result, when wrong, fails on its own."""
    rng = SplitMix64(77)
    for bit in (1, 0):
        assert rows_violated(select_circuit(**select_honest(bit, rng))) == []
    base = select_honest(1, rng)
    other = jubjub_mul(JUBJUB_GENERATOR, 123456789)
    for key, val in (("res", base["b"]), ("zero_res", 0), ("one_res", 1),
                     ("point_res", base["point_b"]), ("identity_res", other)):
        bad = dict(base, **{key: val})
        assert rows_violated(select_circuit(**bad)) != [], key


def decomposition_circuit(a, bits):
    """tests/decomposition.rs:52-71 with N = len(bits)."""
    cs = Plonk.initialize()
    wa = cs.append_witness(a)
    wbits = [cs.append_witness(v) for v in bits]
    for w, x in zip(wbits, cs.component_decomposition(wa, len(bits))):
        cs.assert_equal(w, x)
    return cs


def test_decomposition_cases():
    """tests/decomposition.rs:80-105: 256 bits of a random scalar; bit 10 flipped fails."""
    a = SplitMix64(5).fr()
    bits = [(a >> i) & 1 for i in range(256)]
    good = decomposition_circuit(a, bits)
    assert rows_violated(good) == [] and good.m() == 6 + (2 * 256 + 1) + 256
    bits[10] ^= 1
    assert rows_violated(decomposition_circuit(a, bits)) != []


def range_circuit(a, bits):
    cs = Plonk.initialize()
    cs.component_range(cs.append_witness(a), bits)
    return cs


def test_range_cases():
    """tests/range.rs:66-97: default (76 bits) passes, -(2^77) fails, an odd bit count (77) still composes
    (the reference only requires that compilation does not panic)."""
    a = SplitMix64(6).fr() & ((1 << 76) - 1)
    assert rows_violated(range_circuit(a, 76)) == []
    assert rows_violated(range_circuit((-(1 << 77)) % R_MOD, 76)) != []
    odd = range_circuit(a, 77)
    SynthesizedCircuit.from_composer(odd)
    assert odd.m() == range_circuit(a, 78).m()


@pytest.mark.parametrize("which", ["logic", "boolean", "decomposition"])
def test_restated_prover_refuses_the_negative_case(which):
    """create_proof(...).expect_err(...) of the three files: keys from the honest circuit, witness of the bad one."""
    rng = SplitMix64(9)
    a, b = rng.fr() & ((1 << 30) - 1), rng.fr() & ((1 << 30) - 1)
    bits = [(a >> i) & 1 for i in range(32)]
    flipped = list(bits)
    flipped[10] ^= 1
    good, bad = {
        "logic": (logic_circuit(a, b, a & b, 30, False), logic_circuit(a, b, (a & b) ^ 4, 30, False)),
        "boolean": (boolean_circuit(1), boolean_circuit(2)),
        "decomposition": (decomposition_circuit(a, bits), decomposition_circuit(a, flipped)),
    }[which]
    circ, tau, commit, pk, vk, tr, bl = setup(good)
    proof, pi = plonk.create_proof(pk, circ, commit, tr, bl)
    assert plonk.verify(vk, pk.n, proof, circ.pi_indexes, pi, tr, plonk.trapdoor_kzg_check(tau))
    with pytest.raises(plonk.ProverError):
        plonk.create_proof(pk, SynthesizedCircuit.from_composer(bad), commit, tr, bl)
