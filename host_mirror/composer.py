"""Host-side constraint system: the workload generator in front of the hot path.

Mirror of the reference's ``Plonk<C>`` composer (``src/lib.rs:100-1198``) and of
``Permutation`` bookkeeping (``src/permutation.rs:22-200``) with the same method names and
gate layouts, in Python ints (canonical Fr values).  It is serial host work in the
reference too; nothing here runs on the GPU.  Its output -- selector
columns, wire -> witness indices, witness values, the sigma mapping -- is what
``PlonkKey.compile`` / ``Prover.create_proof`` upload.

The ``Constraint`` builder lives in the absent ``zksnarks`` crate; its selector semantics
are [EXT-RECALL] from upstream dusk-plonk 0.13 and corroborated by the in-tree gate
equation ``q_m a b + q_l a + q_r b + q_o o + q_4 d + q_c + PI = 0`` (``src/lib.rs:544-545``).
"""
import numpy as np

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

# JubJub: -x^2 + y^2 = 1 + D x^2 y^2 over Fr (absent `jub-jub` crate; public parameters)
EDWARDS_D = (-(10240 * pow(10241, -1, R_MOD))) % R_MOD
JUBJUB_GENERATOR = (0x3FD2814C43AC65A6F1FBF02D0FD6CCE62E3EBB21FD6C54ED4DF7B7FFEC7BEACA, 0x12)
JUBJUB_IDENTITY = (0, 1)

SELECTORS = ("q_m", "q_l", "q_r", "q_o", "q_c", "q_d", "q_arith", "q_range", "q_logic",
             "q_fixed_group_add", "q_variable_group_add")


def jubjub_add(p, q):
    x1, y1 = p
    x2, y2 = q
    t = EDWARDS_D * x1 * x2 % R_MOD * y1 % R_MOD * y2 % R_MOD
    x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, R_MOD) % R_MOD
    y3 = (y1 * y2 + x1 * x2) * pow(1 - t, -1, R_MOD) % R_MOD
    return (x3, y3)


def jubjub_neg(p):
    return ((-p[0]) % R_MOD, p[1])


def jubjub_mul(p, k):
    acc = JUBJUB_IDENTITY
    while k:
        if k & 1:
            acc = jubjub_add(acc, p)
        p = jubjub_add(p, p)
        k >>= 1
    return acc


def jubjub_on_curve(p):
    x, y = p
    return (-x * x + y * y - 1 - EDWARDS_D * x * x % R_MOD * y * y) % R_MOD == 0


def compute_windowed_naf(k, width=2):
    """256 entries, least-significant first, each in {-1, 0, 1} for width 2."""
    res = [0] * 256
    i = 0
    while k >= 1:
        if k & 1:
            ki = k % (1 << width)
            if ki >= (1 << (width - 1)):
                ki -= 1 << width
            res[i] = ki
            k -= ki
        k >>= 1
        i += 1
    return res


class Constraint:
    """``zksnarks::plonk::Constraint``: selectors + four wires + optional public input."""
    __slots__ = SELECTORS + ("w_a", "w_b", "w_o", "w_d", "public_input")

    def __init__(self):
        for s in SELECTORS:
            setattr(self, s, 0)
        self.w_a = self.w_b = self.w_o = self.w_d = 0  # Plonk::ZERO
        self.public_input = None

    def copy(self):
        c = Constraint.__new__(Constraint)
        for s in self.__slots__:
            setattr(c, s, getattr(self, s))
        return c

    def _set(self, name, v):
        c = self.copy()
        setattr(c, name, v % R_MOD if isinstance(v, int) and name.startswith("q_") else v)
        return c

    # external selectors
    def mult(self, v): return self._set("q_m", v)
    def left(self, v): return self._set("q_l", v)
    def right(self, v): return self._set("q_r", v)
    def output(self, v): return self._set("q_o", v)
    def fourth(self, v): return self._set("q_d", v)
    def constant(self, v): return self._set("q_c", v)
    def public(self, v): return self._set("public_input", v % R_MOD)
    # wires
    def a(self, w): return self._set("w_a", w)
    def b(self, w): return self._set("w_b", w)
    def o(self, w): return self._set("w_o", w)
    def d(self, w): return self._set("w_d", w)

    def _from_external(self):
        c = self.copy()
        for s in ("q_arith", "q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add"):
            setattr(c, s, 0)
        return c

    @staticmethod
    def arithmetic(s): return s._from_external()._set("q_arith", 1)
    @staticmethod
    def range(s): return s._from_external()._set("q_range", 1)
    @staticmethod
    def logic(s): return s._from_external()._set("q_c", 1)._set("q_logic", 1)
    @staticmethod
    def logic_xor(s): return s._from_external()._set("q_c", -1)._set("q_logic", -1)
    @staticmethod
    def group_add_curve_scalar(s): return s._from_external()._set("q_fixed_group_add", 1)
    @staticmethod
    def group_add_curve_addtion(s): return s._from_external()._set("q_variable_group_add", 1)


def _bits_msb_first(v):
    """``BitIterator8::new(v.to_raw_bytes())``: 256 bits of the canonical value, MSB first."""
    return [(v >> (255 - i)) & 1 for i in range(256)]


class Plonk:
    """``Plonk<JubjubAffine>`` (src/lib.rs:100-112)."""
    ZERO = 0
    ONE = 1

    def __init__(self):
        self.constraints = []
        self.instance = {}
        self.witness = []
        self.witness_map = []  # Permutation.witness_map: witness -> [(wire 0..3, gate)]

    # ConstraintSystem::initialize (src/lib.rs:121-134)
    @classmethod
    def initialize(cls):
        s = cls()
        zero = s.append_witness(0)
        one = s.append_witness(1)
        s.assert_equal_constant(zero, 0, None)
        s.assert_equal_constant(one, 1, None)
        s.append_dummy_gates()
        s.append_dummy_gates()
        return s

    def m(self):
        return len(self.constraints)

    def __getitem__(self, w):
        return self.witness[w]

    def public_input_indexes(self):
        return sorted(self.instance.keys())

    def instance_values(self):
        return [self.instance[i] for i in self.public_input_indexes()]

    @staticmethod
    def dense_public_inputs(indexes, values, size):
        out = [0] * size
        for i, v in zip(indexes, values):
            out[i] = v
        return out

    def append_witness(self, v):
        self.witness.append(v % R_MOD)
        self.witness_map.append([])
        return len(self.witness) - 1

    def append_custom_gate(self, c):
        n = len(self.constraints)
        self.constraints.append(c)
        if c.public_input is not None:
            self.instance[n] = c.public_input
        for wire, w in enumerate((c.w_a, c.w_b, c.w_o, c.w_d)):
            self.witness_map[w].append((wire, n))

    def append_gate(self, c):
        self.append_custom_gate(Constraint.arithmetic(c))

    def append_evaluated_output(self, s):
        a, b, d = self[s.w_a], self[s.w_b], self[s.w_d]
        pi = s.public_input or 0
        x = (s.q_m * a * b + s.q_l * a + s.q_r * b + s.q_d * d + s.q_c + pi) % R_MOD
        y = s.q_o
        if y == 0:
            return None
        o = (-x * pow(y, -1, R_MOD)) % R_MOD
        return self.append_witness(o)

    def append_dummy_gates(self):  # src/lib.rs:601-641
        six = self.append_witness(6)
        one = self.append_witness(1)
        seven = self.append_witness(7)
        min_twenty = self.append_witness(-20)
        c = Constraint().mult(1).left(2).right(3).fourth(1).constant(4).output(4) \
            .a(six).b(seven).d(one).o(min_twenty)
        self.append_gate(c)
        c = Constraint().mult(1).left(1).right(1).constant(127).output(1) \
            .a(min_twenty).b(six).o(seven)
        self.append_gate(c)

    def append_constant(self, v):
        w = self.append_witness(v)
        self.assert_equal_constant(w, v, None)
        return w

    def append_public(self, v):
        w = self.append_witness(v)
        self.assert_equal_constant(w, 0, (-v) % R_MOD)
        return w

    def assert_equal(self, a, b):
        self.append_gate(Constraint().left(1).right(-1).a(a).b(b))

    def assert_equal_constant(self, a, constant, public):
        c = Constraint().left(1).constant(-constant).a(a)
        if public is not None:
            c = c.public(public)
        self.append_gate(c)

    def assert_equal_public_point(self, point, public):
        self.assert_equal_constant(point[0], 0, (-public[0]) % R_MOD)
        self.assert_equal_constant(point[1], 0, (-public[1]) % R_MOD)

    def gate_add(self, s):
        s = Constraint.arithmetic(s).output(-1)
        o = self.append_evaluated_output(s)
        self.append_gate(s.o(o))
        return o

    gate_mul = gate_add  # identical bodies in the reference (src/lib.rs:1168-1197)

    def component_boolean(self, a):
        self.append_gate(Constraint().mult(1).output(-1).a(a).b(a).o(a).d(self.ZERO))

    def component_select_zero(self, bit, value):
        return self.gate_mul(Constraint().mult(1).a(bit).b(value))

    def component_select_one(self, bit, value):
        b, v = self[bit], self[value]
        f = self.append_witness(1 - b + b * v)
        self.append_gate(Constraint().mult(1).left(-1).output(-1).constant(1).a(bit).b(value).o(f))
        return f

    def component_select(self, bit, a, b):
        bit_times_a = self.gate_mul(Constraint().mult(1).a(bit).b(a))
        one_min_bit = self.gate_add(Constraint().left(-1).constant(1).a(bit))
        one_min_bit_b = self.gate_mul(Constraint().mult(1).a(one_min_bit).b(b))
        return self.gate_add(Constraint().left(1).right(1).a(one_min_bit_b).b(bit_times_a))

    def assert_equal_point(self, a, b):
        self.assert_equal(a[0], b[0])
        self.assert_equal(a[1], b[1])

    def component_select_point(self, bit, a, b):
        return (self.component_select(bit, a[0], b[0]), self.component_select(bit, a[1], b[1]))

    def component_select_identity(self, bit, a):
        return (self.component_select_zero(bit, a[0]), self.component_select_one(bit, a[1]))

    # src/lib.rs:881-917: N bits, little-endian, 2 N + 1 gates
    def component_decomposition(self, scalar, n_bits):
        assert 0 < n_bits <= 256
        v = self[scalar]
        acc = self.ZERO
        decomposition = []
        for i in range(n_bits):
            d = self.append_witness((v >> i) & 1)
            self.component_boolean(d)
            acc = self.gate_add(Constraint().left(pow(2, i, R_MOD)).right(1).a(d).b(acc))
            decomposition.append(d)
        self.assert_equal(acc, scalar)
        return decomposition

    # src/lib.rs:937-957: double-and-add over the 252 bits of the scalar
    def component_mul_point(self, jubjub, point):
        bits = self.component_decomposition(jubjub, 252)
        result = (self.ZERO, self.ONE)
        for bit in reversed(bits):
            result = self.component_add_point(result, result)
            to_add = self.component_select_identity(bit, point)
            result = self.component_add_point(result, to_add)
        return result

    # src/lib.rs:1041-1163
    def component_range(self, witness, num_bits):
        bits = _bits_msb_first(self[witness])
        bits.reverse()
        num_gates = num_bits >> 3
        if num_bits % 8 != 0:
            num_gates += 1
        num_quads = num_gates * 4
        pad = 1 + (((num_quads << 1) - num_bits) >> 1)
        used_gates = num_gates + 1
        base = Constraint.range(Constraint())
        constraints = [base.copy() for _ in range(used_gates)]
        accumulators = []
        accumulator = 0
        for i in range(pad, num_quads + 1):
            bit_index = (num_quads - i) << 1
            quad = bits[bit_index] + 2 * bits[bit_index + 1]
            accumulator = (4 * accumulator + quad) % R_MOD
            var = self.append_witness(accumulator)
            accumulators.append(var)
            idx = i // 4
            name = ("w_d", "w_o", "w_b", "w_a")[i % 4]
            setattr(constraints[idx], name, var)
        constraints[-1] = Constraint()
        if accumulators:
            constraints[-1].w_d = accumulators[-1]
        for c in constraints:
            self.append_custom_gate(c)
        if accumulators:
            self.assert_equal(accumulators[-1], witness)

    # src/lib.rs:283-390
    def _append_logic_component(self, a, b, num_bits, is_xor):
        num_bits = min(num_bits, 256)
        num_quads = num_bits >> 1
        left_acc = right_acc = out_acc = 0
        a_bits = _bits_msb_first(self[a])[256 - num_bits:]
        b_bits = _bits_msb_first(self[b])[256 - num_bits:]
        c = Constraint.logic_xor(Constraint()) if is_xor else Constraint.logic(Constraint())
        for i in range(num_quads):
            idx = i * 2
            lq = (a_bits[idx] << 1) + a_bits[idx + 1]
            rq = (b_bits[idx] << 1) + b_bits[idx + 1]
            oq = (lq ^ rq) if is_xor else (lq & rq)
            pq = lq * rq
            left_acc = (left_acc * 4 + lq) % R_MOD
            right_acc = (right_acc * 4 + rq) % R_MOD
            out_acc = (out_acc * 4 + oq) % R_MOD
            wa = self.append_witness(left_acc)
            wb = self.append_witness(right_acc)
            wc = self.append_witness(pq)
            wd = self.append_witness(out_acc)
            c = c.o(wc)
            self.append_custom_gate(c)
            c = c.a(wa).b(wb).d(wd)
        self.append_custom_gate(Constraint().a(c.w_a).b(c.w_b).d(c.w_d))
        return c.w_d

    def append_logic_and(self, a, b, num_bits):
        return self._append_logic_component(a, b, num_bits, False)

    def append_logic_xor(self, a, b, num_bits):
        return self._append_logic_component(a, b, num_bits, True)

    def append_point(self, p):
        return (self.append_witness(p[0]), self.append_witness(p[1]))

    # src/lib.rs:808-854
    def component_add_point(self, a, b):
        x_1, y_1 = a
        x_2, y_2 = b
        p3 = jubjub_add((self[x_1], self[y_1]), (self[x_2], self[y_2]))
        x1_y2 = self[x_1] * self[y_2] % R_MOD
        w_x1y2 = self.append_witness(x1_y2)
        x_3 = self.append_witness(p3[0])
        y_3 = self.append_witness(p3[1])
        c = Constraint.group_add_curve_addtion(Constraint().a(x_1).b(y_1).o(x_2).d(y_2))
        self.append_custom_gate(c)
        self.append_custom_gate(Constraint().a(x_3).b(y_3).d(w_x1y2))
        return (x_3, y_3)

    # src/lib.rs:399-537
    def component_mul_generator(self, jubjub, generator):
        bits = 256
        multiples = [generator]
        for _ in range(1, bits):
            multiples.append(jubjub_add(multiples[-1], multiples[-1]))
        multiples.reverse()
        scalar = self[jubjub]
        wnaf = compute_windowed_naf(scalar, 2)
        scalar_acc = [0]
        point_acc = [JUBJUB_IDENTITY]
        xy_alphas = []
        for i, entry in enumerate(reversed(wnaf)):
            if entry == 0:
                s_add, p_add = 0, JUBJUB_IDENTITY
            elif entry == -1:
                s_add, p_add = R_MOD - 1, jubjub_neg(multiples[i])
            elif entry == 1:
                s_add, p_add = 1, multiples[i]
            else:
                raise ValueError("UnsupportedWNAF2k")
            scalar_acc.append((2 * scalar_acc[i] + s_add) % R_MOD)
            point_acc.append(jubjub_add(point_acc[i], p_add))
            xy_alphas.append(p_add[0] * p_add[1] % R_MOD)
        for i in range(bits):
            acc_x = self.append_witness(point_acc[i][0])
            acc_y = self.append_witness(point_acc[i][1])
            accumulated_bit = self.append_witness(scalar_acc[i])
            if i == 0:
                self.assert_equal_constant(acc_x, 0, None)
                self.assert_equal_constant(acc_y, 1, None)
                self.assert_equal_constant(accumulated_bit, 0, None)
            x_beta, y_beta = multiples[i]
            xy_alpha = self.append_witness(xy_alphas[i])
            xy_beta = x_beta * y_beta % R_MOD
            c = Constraint.group_add_curve_scalar(Constraint()).left(x_beta).right(y_beta) \
                .constant(xy_beta).a(acc_x).b(acc_y).o(xy_alpha).d(accumulated_bit)
            self.append_custom_gate(c)
        acc_x = self.append_witness(point_acc[bits][0])
        acc_y = self.append_witness(point_acc[bits][1])
        last_bit = self.append_witness(scalar_acc[bits])
        self.append_gate(Constraint().a(acc_x).b(acc_y).d(last_bit))
        self.assert_equal(last_bit, jubjub)
        return (acc_x, acc_y)

    # ---------------------------------------------------------------- export
    def selector_columns(self):
        """{name: list of canonical ints of length m} (src/key.rs:103-119)."""
        return {s: [getattr(c, s) for c in self.constraints] for s in SELECTORS}

    def wire_indices(self):
        """(4, m) int64 witness index per wire (src/prover.rs:114-119)."""
        return np.array([[c.w_a for c in self.constraints], [c.w_b for c in self.constraints],
                         [c.w_o for c in self.constraints], [c.w_d for c in self.constraints]],
                        dtype=np.int64)

    def compute_sigma_permutations(self, n):
        """``Permutation::compute_sigma_permutations`` (src/permutation.rs:108-145):
        sigma[wire][gate] = (next wire, next gate) in the cycle of the shared witness."""
        sig_w = np.tile(np.arange(4, dtype=np.int64)[:, None], (1, n))
        sig_g = np.tile(np.arange(n, dtype=np.int64)[None, :], (4, 1))
        for wire_data in self.witness_map:
            k = len(wire_data)
            for j, (w, g) in enumerate(wire_data):
                nw, ng = wire_data[(j + 1) % k]
                sig_w[w, g] = nw
                sig_g[w, g] = ng
        return sig_w, sig_g


class SynthesizedCircuit:
    """Array form of a synthesized constraint system: what compile / prove consume.

    selectors: {name: (m,) canonical-int list or small-int numpy array}; wires: (4, m)
    witness indices; witness: list of canonical ints; sigma: (wire, gate) int arrays of
    shape (4, n); public inputs as (indexes, values)."""

    def __init__(self, m, selectors, wires, witness, sigma_w, sigma_g, pi_indexes, pi_values):
        self.m = m
        self.n = 1 << max(m - 1, 0).bit_length() if m > 1 else 1
        self.selectors = selectors
        self.wires = wires
        self.witness = witness
        self.sigma_w, self.sigma_g = sigma_w, sigma_g
        self.pi_indexes, self.pi_values = pi_indexes, pi_values

    @classmethod
    def from_composer(cls, cs: Plonk):
        m = cs.m()
        n = 1 << (m - 1).bit_length()
        sw, sg = sigma_from_wires(cs.wire_indices(), n)
        return cls(m, cs.selector_columns(), cs.wire_indices(), list(cs.witness), sw, sg,
                   cs.public_input_indexes(), cs.instance_values())


def sigma_from_wires(wires, n):
    """Vectorised ``compute_sigma_permutations`` (src/permutation.rs:108-145) from the (4, m)
    wire -> witness table: positions are visited gate by gate in a, b, o, d order, exactly
    the insertion order of ``add_witnesses_to_map``; each maps to the next position holding
    the same witness (cyclically)."""
    wires = np.asarray(wires, dtype=np.int64)
    m = wires.shape[1]
    wid = wires.T.reshape(-1)                      # position p = gate * 4 + wire
    order = np.argsort(wid, kind="stable")         # grouped by witness, insertion order inside
    sw = wid[order]
    first = np.ones(len(order), dtype=bool)
    first[1:] = sw[1:] != sw[:-1]
    starts = np.nonzero(first)[0]
    group_start = np.repeat(starts, np.diff(np.append(starts, len(order))))
    nxt = np.empty(len(order), dtype=np.int64)
    nxt[:-1] = order[1:]
    nxt[-1] = order[0]
    last = np.ones(len(order), dtype=bool)
    last[:-1] = sw[1:] != sw[:-1]
    nxt[last] = order[group_start[last]]
    sigma_pos = np.empty(len(order), dtype=np.int64)
    sigma_pos[order] = nxt
    sig_w = np.tile(np.arange(4, dtype=np.int64)[:, None], (1, n))
    sig_g = np.tile(np.arange(n, dtype=np.int64)[None, :], (4, 1))
    sp = sigma_pos.reshape(m, 4).T                 # (4, m)
    sig_w[:, :m] = sp % 4
    sig_g[:, :m] = sp // 4
    return sig_w, sig_g


def synthetic_circuit(k, seed=8349, slack=8):
    """Synthetic 2^k-gate workload: ``Plonk::initialize`` followed by an alternating
    ``gate_add`` / ``gate_mul`` chain and one public input, m = 2^k - slack gates so that
    ``additional_n`` stays 2^k (src/key.rs:81).  Built with numpy (the per-gate Python composer
    is too slow for 2^20 gates); returns a ``SynthesizedCircuit``."""
    cs = Plonk.initialize()
    x = cs.append_witness(seed)
    y = cs.append_witness(seed + 4)
    m0 = cs.m()
    m = (1 << k) - slack
    nchain = m - m0 - 1
    assert nchain > 0
    base = len(cs.witness)
    wit = list(cs.witness)
    xv, yv = wit[x], wit[y]
    wa = np.empty(nchain, dtype=np.int64)
    wb = np.empty(nchain, dtype=np.int64)
    xi, yi = x, y
    for i in range(nchain):
        wa[i], wb[i] = xi, yi
        if i & 1:
            xv = xv * yv % R_MOD
            wit.append(xv)
            xi = base + i
        else:
            yv = (xv + yv + i) % R_MOD
            wit.append(yv)
            yi = base + i
    wo = base + np.arange(nchain, dtype=np.int64)
    odd = (np.arange(nchain) & 1).astype(np.int64)
    sel0 = cs.selector_columns()
    W0 = cs.wire_indices()
    # final gate: append_public(x)
    pub = len(wit)
    wit.append(xv)
    sel = {}
    zeros = np.zeros(nchain, dtype=np.int64)
    chain = {"q_m": odd, "q_l": 1 - odd, "q_r": 1 - odd, "q_o": zeros - 1, "q_c": (1 - odd) * np.arange(nchain),
             "q_d": zeros, "q_arith": zeros + 1, "q_range": zeros, "q_logic": zeros, "q_fixed_group_add": zeros,
             "q_variable_group_add": zeros}
    last = {s: 0 for s in SELECTORS}
    last.update({"q_l": 1, "q_arith": 1})
    for s in SELECTORS:
        head = np.array([v if v < (1 << 62) else v - R_MOD for v in sel0[s]], dtype=np.int64)
        sel[s] = np.concatenate([head, chain[s], np.array([last[s]], dtype=np.int64)])
    wires = np.concatenate([W0, np.stack([wa, wb, wo, zeros]), np.array([[pub], [0], [0], [0]], dtype=np.int64)], axis=1)
    n = 1 << k
    sw, sg = sigma_from_wires(wires, n)
    return SynthesizedCircuit(m, sel, wires, wit, sw, sg, [m - 1], [(-xv) % R_MOD])
