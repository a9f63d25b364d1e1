"""``PlonkKey::compile_with_circuit`` on the GPU (host mirror of ``src/key.rs:63-327``).

Same sequence as the reference -- 11 selector iNTTs + 4 sigma iNTTs over n, 15 commits,
16 coset NTTs over 8n, the vanishing polynomial on the coset -- but the proving key is
born in HBM and stays there: ``Prover`` holds device buffers, never host vectors, and
``create_proof`` does not deep-clone them (the reference clones 17 x 8n Fr per proof,
``src/prover.rs:80-86``).  The sigma evaluations over the n-domain are kept too, so round 2
does not re-run four forward NTTs per proof (``src/permutation.rs:229-235``): dft(idft(x))
is x exactly.
"""
import numpy as np

from host_mirror.composer import SELECTORS, SynthesizedCircuit, Plonk
from .field import R_MOD, fr_column_to_mont, fr_from_mont, fr_to_mont, g1_from_mont
from .plonk_params import Error
from .transcript import Transcript

SIGMAS = ("s_sigma_1", "s_sigma_2", "s_sigma_3", "s_sigma_4")
# order in which the verification key seeds the transcript ([EXT-RECALL] dusk-plonk 0.13)
VK_TRANSCRIPT_ORDER = (("q_m", b"q_m"), ("q_l", b"q_l"), ("q_r", b"q_r"), ("q_o", b"q_o"),
                       ("q_c", b"q_c"), ("q_d", b"q_4"), ("q_arith", b"q_arith"),
                       ("q_range", b"q_range"), ("q_logic", b"q_logic"),
                       ("q_variable_group_add", b"q_variable_group_add"),
                       ("q_fixed_group_add", b"q_fixed_group_add"),
                       ("s_sigma_1", b"s_sigma_1"), ("s_sigma_2", b"s_sigma_2"),
                       ("s_sigma_3", b"s_sigma_3"), ("s_sigma_4", b"s_sigma_4"))


class ProvingKey:
    """``zksnarks::plonk::ProvingKey`` resident on the device (fields src/key.rs:247-302)."""

    def __init__(self):
        self.poly = {}    # name -> DeviceBuffer (n coefficients)
        self.eval8 = {}   # name -> DeviceBuffer (8n coset evaluations); plus "linear"
        self.sigma_evals = []  # 4 x DeviceBuffer (n)
        self.roots = None      # Fft::elements
        self.zh_inv = None     # (8, 4) Montgomery: 1 / Z_H on the coset (period 8)
        self.widget_mask = 0


class VerificationKey(dict):
    """name -> affine commitment ((x, y) or None) plus n (= m), n_inv, generator(_inv)
    (src/key.rs:203-214)."""

    def transcript_list(self):
        return [(lab, self[name]) for name, lab in VK_TRANSCRIPT_ORDER]


class PlonkKey:
    @staticmethod
    def compile_with_circuit(pp, label, circuit):
        """-> Prover.  ``circuit`` is a synthesized ``Plonk`` composer or a
        ``SynthesizedCircuit``; ``pp`` a ``PlonkParams``.  (The Verifier half of the pair is
        constant-size host work and lives with the acceptance oracle.)"""
        from .prover import Prover
        circ = SynthesizedCircuit.from_composer(circuit) if isinstance(circuit, Plonk) else circuit
        ctx = pp.ctx
        m, n = circ.m, circ.n
        k = n.bit_length() - 1
        additional_n = 1 << (m + 6 - 1).bit_length()
        keypair = pp.trim(additional_n)                                  # src/key.rs:82
        pk = ProvingKey()
        pk.n, pk.m, pk.k = n, m, k
        n8 = 8 * n
        pk.roots = ctx.fft_elements(k)
        vk = VerificationKey()
        gen = fr_from_mont([_const(ctx, k, 0), _const(ctx, k, 1), _const(ctx, k, 2)])
        vk["n"], vk["generator"], vk["generator_inv"], vk["n_inv"] = m, gen[0], gen[1], gen[2]
        # selectors: pad, iNTT, commit, coset NTT over 8n (src/key.rs:89-131,138-154,226-245)
        for s in SELECTORS:
            col = circ.selectors[s]
            buf = ctx.upload(fr_column_to_mont(col, n))
            ctx.ntt_dev(buf, n, buf, k, True, False)
            pk.poly[s] = buf
        # the 11 selector commitments as two batched launch sets; a selector polynomial has n
        # coefficients and the trimmed SRS at least n + 7 powers, so `.unwrap_or_default()`
        # (src/key.rs:138-154) can only ever see Ok here
        names = list(SELECTORS)
        for lo in range(0, len(names), 8):
            grp = names[lo:lo + 8]
            try:
                comms = keypair.commit_batch([pk.poly[s] for s in grp])
            except Error:   # SRS shorter than the circuit: keep the reference's per-selector default
                comms = [keypair.commit_or_default(pk.poly[s]) for s in grp]
            for s, c in zip(grp, comms):
                vk[s] = c.affine()
        for s in ("q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add"):
            if vk[s] is not None:
                pk.widget_mask |= 1 << ("q_range", "q_logic", "q_fixed_group_add", "q_variable_group_add").index(s)
        # sigma polynomials (src/permutation.rs:172-200, src/key.rs:134-159)
        for i, nm in enumerate(SIGMAS):
            enc = (np.asarray(circ.sigma_w[i], dtype=np.uint32) << np.uint32(30)) | \
                np.asarray(circ.sigma_g[i], dtype=np.uint32)
            ev = ctx.alloc(n)
            ctx.perm_lagrange(k, enc, pk.roots, ev)
            pk.sigma_evals.append(ev)
            buf = ctx.alloc(n)
            ctx.ntt_dev(ev, n, buf, k, True, False)
            pk.poly[nm] = buf
        for nm, c in zip(SIGMAS, keypair.commit_batch([pk.poly[nm] for nm in SIGMAS])):   # `?` in the reference
            vk[nm] = c.affine()
        # 8n-coset evaluations (src/key.rs:226-245).  A rank of a sharded proof keeps only its cosets
        # (g w_8n^u) H_n, u in [u0, u0 + nloc): n-point transforms, 8 / G of the memory
        comm = getattr(keypair, "native_comm", None)
        lin = ctx.upload(fr_to_mont([0, 1]))
        if comm is not None:
            nloc = 8 // comm.world
            u0 = comm.rank * nloc
            pk.cosets = (u0, nloc)
            for nm, p in list(pk.poly.items()) + [("linear", lin)]:
                e8 = ctx.alloc(nloc * n)
                ctx.coset8_ntt(p, 0, p.n, e8, 0, k, u0, nloc)
                pk.eval8[nm] = e8
        else:
            for nm, p in pk.poly.items():
                e8 = ctx.alloc(n8)
                ctx.ntt_dev(p, n, e8, k + 3, False, True)
                pk.eval8[nm] = e8
            e8 = ctx.alloc(n8)
            ctx.ntt_dev(lin, 2, e8, k + 3, False, True)
            pk.eval8["linear"] = e8
        # Z_H(g w8^i) = g^n (w8^n)^i - 1: eight distinct values (src/key.rs:291)
        g = 7
        w8 = fr_from_mont([_const(ctx, k + 3, 0)])[0]
        gn, wn = pow(g, n, R_MOD), pow(w8, n, R_MOD)
        pk.zh_inv = fr_to_mont([pow((gn * pow(wn, i, R_MOD) - 1) % R_MOD, -1, R_MOD) for i in range(8)])
        transcript = Transcript.base(label, vk.transcript_list(), m)      # src/prover.rs:54-55
        prover = Prover(ctx, keypair, pk, vk, transcript, circ.pi_indexes)
        prover.label = label
        return prover

    @staticmethod
    def compile_pair(pp, circuit, label=b"plonk"):
        """``PlonkKey::compile`` as the reference returns it (src/key.rs:46-50,304-327): (Prover, Verifier)."""
        prover = PlonkKey.compile_with_circuit(pp, label, circuit)
        return prover, prover.verifier()

    @staticmethod
    def compile(pp, circuit, label=b"plonk"):
        """``PlonkKey::compile`` (src/key.rs:46-50): label b"plonk"."""
        return PlonkKey.compile_with_circuit(pp, label, circuit)


def _const(ctx, k, kind):
    from .ffi import fft_constant
    return fft_constant(k, kind)
