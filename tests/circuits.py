"""Circuits shared by the CPU and GPU tests: the reference's own test circuits
(tests/range.rs, tests/logic.rs, tests/ecc.rs, README.md TestCircuit) restated on the
Python composer."""
from host_mirror.composer import (Plonk, Constraint, JUBJUB_GENERATOR, jubjub_mul, R_MOD)


def range_circuit(a, bits=76):
    """tests/range.rs:47-58 DummyCircuit."""
    cs = Plonk.initialize()
    w = cs.append_witness(a)
    cs.component_range(w, bits)
    return cs


def readme_circuit(a=20, b=5, e=2, c=None, d=None, f=None):
    """README.md:24-71 TestCircuit: a + b = c (PI), range checks, a * b = d (PI), e * G = f (PI)."""
    cs = Plonk.initialize()
    c = (a + b) % R_MOD if c is None else c
    d = a * b % R_MOD if d is None else d
    f = jubjub_mul(JUBJUB_GENERATOR, e) if f is None else f
    wa = cs.append_witness(a)
    wb = cs.append_witness(b)
    cs.append_gate(Constraint().left(1).right(1).public(-c).a(wa).b(wb))
    cs.component_range(wa, 1 << 6)
    cs.component_range(wb, 1 << 5)
    cs.append_gate(Constraint().mult(1).public(-d).a(wa).b(wb))
    we = cs.append_witness(e)
    res = cs.component_mul_generator(we, JUBJUB_GENERATOR)
    cs.assert_equal_public_point(res, f)
    return cs


def logic_curve_circuit(a=0x1234567890ABCDEF, b=0xFEDCBA0987654321, s1=77, s2=1234567, bad=False):
    """tests/logic.rs + tests/ecc.rs (component_add_point) in one circuit."""
    cs = Plonk.initialize()
    wa = cs.append_witness(a)
    wb = cs.append_witness(b)
    r1 = cs.append_logic_and(wa, wb, 64)
    r2 = cs.append_logic_xor(wa, wb, 64)
    cs.assert_equal_constant(r1, (a & b) + (1 if bad else 0), None)
    cs.assert_equal_constant(r2, a ^ b, None)
    p1 = cs.append_point(jubjub_mul(JUBJUB_GENERATOR, s1))
    p2 = cs.append_point(jubjub_mul(JUBJUB_GENERATOR, s2))
    p3 = cs.component_add_point(p1, p2)
    cs.assert_equal_public_point(p3, jubjub_mul(JUBJUB_GENERATOR, s1 + s2))
    return cs


def arithmetic_chain(m_target, seed=3):
    """Synthetic add / mul gate chain: m_target gates in total."""
    cs = Plonk.initialize()
    x = cs.append_witness(seed)
    y = cs.append_witness(seed + 4)
    i = 0
    while cs.m() < m_target - 1:
        if i & 1:
            x = cs.gate_mul(Constraint().mult(1).a(x).b(y))
        else:
            y = cs.gate_add(Constraint().left(1).right(1).constant(i).a(x).b(y))
        i += 1
    pub = cs.append_public(cs[x]) if cs.m() < m_target else None
    return cs


def boolean_select_circuit(bit=1, a=11, b=22):
    """tests/boolean.rs: component_boolean + the select family."""
    cs = Plonk.initialize()
    wbit = cs.append_witness(bit)
    cs.component_boolean(wbit)
    wa, wb = cs.append_witness(a), cs.append_witness(b)
    x = cs.component_select(wbit, wa, wb)
    cs.assert_equal_constant(x, a if bit else b, None)
    z0 = cs.component_select_zero(wbit, wa)
    cs.assert_equal_constant(z0, a if bit else 0, None)
    o1 = cs.component_select_one(wbit, wa)
    cs.assert_equal_constant(o1, a if bit else 1, None)
    p1 = cs.append_point(jubjub_mul(JUBJUB_GENERATOR, 5))
    p2 = cs.append_point(jubjub_mul(JUBJUB_GENERATOR, 9))
    sp = cs.component_select_point(wbit, p1, p2)
    cs.assert_equal_public_point(sp, jubjub_mul(JUBJUB_GENERATOR, 5 if bit else 9))
    si = cs.component_select_identity(wbit, p1)
    cs.assert_equal_public_point(si, jubjub_mul(JUBJUB_GENERATOR, 5) if bit else (0, 1))
    return cs


def decomposition_circuit(a=23, n_bits=64):
    """tests/decomposition.rs: bits of a, asserted against witnesses."""
    cs = Plonk.initialize()
    wa = cs.append_witness(a)
    bits = cs.component_decomposition(wa, n_bits)
    for i, w in enumerate(bits):
        cs.assert_equal_constant(w, (a >> i) & 1, None)
    return cs


def mul_point_circuit(scalar=0x1234567890ABCDEF1234567, base_mult=7):
    """tests/ecc.rs mul_point: variable-base scalar multiplication (decomposition + curve-add gates)."""
    cs = Plonk.initialize()
    ws = cs.append_witness(scalar)
    pt = cs.append_point(jubjub_mul(JUBJUB_GENERATOR, base_mult))
    res = cs.component_mul_point(ws, pt)
    cs.assert_equal_public_point(res, jubjub_mul(JUBJUB_GENERATOR, base_mult * scalar))
    return cs


# ---- the three DummyCircuits of tests/ecc.rs, with free (a, b, c) so that the reference's negative cases
# (an honest prover key, a witness that violates the curve relation) can be replayed
def ecc_mul_generator_circuit(a=7, b=None):
    """tests/ecc.rs:19-62 (mul_generator_works): b == a * G; negative case :81-97 passes b = 8 G for a = 7."""
    cs = Plonk.initialize()
    wa = cs.append_witness(a)
    wb = cs.append_point(jubjub_mul(JUBJUB_GENERATOR, a) if b is None else b)
    wx = cs.component_mul_generator(wa, JUBJUB_GENERATOR)
    cs.assert_equal_point(wb, wx)
    return cs


def ecc_add_point_circuit(a=None, b=None, c=None, sa=7, sb=8):
    """tests/ecc.rs:109-160 (add_point_works): c == a + b; negative case :216-232 passes c = 9 G."""
    from host_mirror.composer import jubjub_add
    a = jubjub_mul(JUBJUB_GENERATOR, sa) if a is None else a
    b = jubjub_mul(JUBJUB_GENERATOR, sb) if b is None else b
    c = jubjub_add(a, b) if c is None else c
    cs = Plonk.initialize()
    wa, wb, wc = cs.append_point(a), cs.append_point(b), cs.append_point(c)
    wx = cs.component_add_point(wa, wb)
    cs.assert_equal_point(wc, wx)
    return cs


def ecc_mul_point_circuit(a=7, b=None, c=None, sb=8):
    """tests/ecc.rs:236-290 (mul_point_works): c == a * b; negative case :303-318 passes an unrelated c."""
    b = jubjub_mul(JUBJUB_GENERATOR, sb) if b is None else b
    c = jubjub_mul(b, a) if c is None else c
    cs = Plonk.initialize()
    wa = cs.append_witness(a)
    wb, wc = cs.append_point(b), cs.append_point(c)
    wx = cs.component_mul_point(wa, wb)
    cs.assert_equal_point(wc, wx)
    return cs
