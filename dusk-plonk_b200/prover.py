"""``Prover::create_proof`` with device-resident rounds (host mirror of ``src/prover.rs:67-474``).

Round structure, transcript labels and commit order follow the reference line by line;
everything between the transcript appends runs on the GPU through the C ABI and the
polynomials never leave HBM.  Host <-> device traffic per proof: the witness upload
(4n Fr), 11 affine commitments and 17 evaluations down, 8 challenges + 11 blinders up.
"""
import numpy as np

from .field import R_MOD, K1, K2, K3, fr_from_mont, fr_to_mont, fr_to_mont1, g1_from_mont
from .ffi import QuotientArgs, BufferView as _View
from host_mirror.composer import SELECTORS, SynthesizedCircuit, Plonk
from .widgets import linearization_scalars
from .plonk_params import PlonkParams, ShardedNativeParams

_r = R_MOD
EVAL_NAMES = ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
              "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval",
              "q_l_eval", "q_r_eval", "perm_eval", "r_poly_eval")
COMM_NAMES = ("a_comm", "b_comm", "c_comm", "d_comm", "z_comm", "t_low_comm", "t_mid_comm",
              "t_high_comm", "t_4_comm", "w_z_chall_comm", "w_z_chall_w_comm")


class Proof:
    """``Proof<P>`` (src/prover/proof.rs:38-66): 11 commitments + 16 evaluations, as affine
    points ((x, y) canonical ints or None) and canonical ints."""

    def __init__(self):
        self.evaluations = {}

    @classmethod
    def from_limbs(cls, comms, evals, wire_bytes=None):
        """Proof over the native driver's output (11 x 12 / 16 x 4 uint64 Montgomery limbs); the
        conversion to canonical integers happens on first access, not on the proving path."""
        p = object.__new__(cls)
        p.__dict__["_limbs"] = (np.array(comms, dtype=np.uint64), np.array(evals, dtype=np.uint64))
        if wire_bytes is not None:
            p.__dict__["wire_bytes"] = bytes(wire_bytes)
        return p

    def __getattr__(self, name):
        limbs = self.__dict__.pop("_limbs", None)
        if limbs is None:
            raise AttributeError(name)
        for i, c in enumerate(COMM_NAMES):
            self.__dict__[c] = g1_from_mont(limbs[0][i])
        self.__dict__["evaluations"] = dict(zip(EVAL_NAMES, fr_from_mont(limbs[1])))
        return getattr(self, name)

    def __eq__(self, o):
        return all(getattr(self, c) == getattr(o, c) for c in COMM_NAMES) and \
            dict(self.evaluations) == dict(o.evaluations)

    # Wire format (src/prover/proof.rs:36 derives SCALE Encode/Decode; fixed-size fields encode as their
    # bytes in declaration order): 11 compressed G1 points (48 B) then the 16 evaluations (32 B LE) in
    # the field order of ``Evaluations`` (src/prover/linearization_poly.rs:113-130) = 1040 bytes.
    # [EXT-RECALL] for the per-type encodings (zcash-style compressed G1, canonical little-endian Fr).
    WIRE_EVAL_ORDER = ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
                       "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval", "s_sigma_1_eval", "s_sigma_2_eval",
                       "s_sigma_3_eval", "r_poly_eval", "perm_eval")

    def to_bytes(self):
        from .transcript import g1_compress
        out = b"".join(g1_compress(getattr(self, c)) for c in COMM_NAMES)
        return out + b"".join((self.evaluations[k] % _r).to_bytes(32, "little") for k in self.WIRE_EVAL_ORDER)

    @classmethod
    def from_bytes(cls, data):
        """Decode and VALIDATE untrusted proof bytes with the library's host decoder (``zkp_proof_decode``): every
        commitment must be a canonical compressed point of the prime-order subgroup, every scalar canonical
        (src/prover/proof.rs:77: "subgroup checks are done when the proof is deserialized").  ``ValueError`` otherwise."""
        import ctypes
        from .ffi import load_library
        if len(data) != 48 * 11 + 32 * 16:
            raise ValueError("a proof is 1040 bytes")
        raw = np.frombuffer(bytes(data), dtype=np.uint8).copy()
        comms = np.zeros((11, 12), dtype=np.uint64)
        evals = np.zeros((16, 4), dtype=np.uint64)
        rc = load_library().zkp_proof_decode(ctypes.c_void_p(raw.ctypes.data), ctypes.c_void_p(comms.ctypes.data),
                                             ctypes.c_void_p(evals.ctypes.data))
        if rc:
            raise ValueError("malformed proof: non-canonical encoding, point off the curve or outside the subgroup")
        p = cls.from_limbs(comms, evals, bytes(data))
        p.evaluations          # decoded proofs are materialised at once (callers may edit fields)
        return p


class Prover:
    def __init__(self, ctx, keypair, prover_key, verifier_key, transcript, pi_indexes):
        self.ctx = ctx
        self.keypair = keypair
        self.prover_key = prover_key
        self.verifier_key = verifier_key
        self.transcript = transcript
        self.size = prover_key.n
        self.pi_indexes = list(pi_indexes)
        self._ws = None
        self._native = None
        self._wiring_of = None
        # the sharded proof (NCCL collectives between rounds) and traced proofs run the same rounds
        # from Python; everything else goes through the native driver
        self.native = True

    # ---------------------------------------------------------------- workspace
    def _workspace(self):
        if self._ws is None:
            n, c = self.size, self.ctx
            # the seven polynomials that go to the 8n coset live side by side (stride S) so their
            # transforms run as batched launches: a, b, c, d, PI (known after round 1) | z, L1
            S = n + 8
            P7 = c.alloc(7 * S)
            P7.zero()
            comm = self._comm()
            if comm is not None:
                # sharded proof: the buffers NCCL gathers into are torch tensors the kernels address too;
                # E7 is padded to a whole number of rounds of `world` polynomials
                # sharded proof: each GPU transforms one of the seven polynomials to the 8n coset and the
                # ranks exchange SLICES (all-to-all): rank s only ever needs the evaluations its part of the
                # quotient reads, [s per, (s + 1) per + 8) of each polynomial.  E7 holds those slices.
                torch = comm.torch
                G = comm.world
                slots = -(-7 // G) * G
                self._hl = 8 * n // G + 8
                dev = torch.device("cuda", c.device)
                self._t_E7 = torch.empty(slots * self._hl * 4, dtype=torch.int64, device=dev)
                self._t_T = torch.empty(8 * n * 4, dtype=torch.int64, device=dev)
                self._t_stage = torch.empty(8 * n * 4, dtype=torch.int64, device=dev)
                self._t_send = torch.empty(G * self._hl * 4, dtype=torch.int64, device=dev)
                self._full8 = c.wrap(self._t_stage.data_ptr(), 8 * n)
                E7 = c.wrap(self._t_E7.data_ptr(), slots * self._hl)
                Tbuf = c.wrap(self._t_T.data_ptr(), 8 * n)
            else:
                E7, Tbuf = c.alloc(7 * 8 * n), c.alloc(8 * n)
            e8n = 8 * n if comm is None else self._hl      # elements each coset vector occupies in E7
            ws = {"W": c.alloc(4 * n), "Z": c.alloc(n), "P7": P7, "E7": E7, "S": S,
                  "wp": [_View(P7, j * S, n + 2) for j in range(4)], "PI": _View(P7, 4 * S, n),
                  "zp": _View(P7, 5 * S, n + 3), "L1": _View(P7, 6 * S, n),
                  "e8": [_View(E7, j * e8n, e8n) for j in range(4)], "pi8": _View(E7, 4 * e8n, e8n),
                  "z8": _View(E7, 5 * e8n, e8n), "l18": _View(E7, 6 * e8n, e8n),
                  "T": Tbuf, "R": c.alloc(n + 3),
                  "AGG": c.alloc(5 * n), "WZ": c.alloc(5 * n), "SAGG": c.alloc(n + 3), "WZW": c.alloc(n + 3)}
            self._ws = ws
        return self._ws

    def _comm(self):
        """The communicator when the commit key is sharded over several GPUs, else None."""
        comm = getattr(self.keypair, "comm", None)
        return comm if comm is not None and getattr(comm, "world", 1) > 1 and hasattr(comm, "dist") else None

    def _side_context(self):
        """Second stream on the same device for work that does not depend on the transcript."""
        if getattr(self, "_side", None) is None:
            from .ffi import Context
            self._side = Context(self.ctx.device)
        return self._side

    def clone(self):
        """``Prover: Clone`` (src/prover.rs:28): a second prover over the SAME device-resident key and SRS with its own
        context (stream, MSM / NTT scratch) and workspace, so that several proofs can be in flight on one GPU --
        one proof's latency-bound bucket reductions and transcript synchronisations run under another's
        accumulation.  The library is re-entrant per context; clones are driven from separate host threads."""
        from .ffi import Context
        if getattr(self.keypair, "native_comm", None) is not None:
            raise ValueError("a prover sharded over GPUs is not cloned: its ranks already fill the GPUs")
        ctx2 = Context(self.ctx.device)
        kp2 = type(self.keypair)(ctx2, self.keypair.srs, getattr(self.keypair, "opening_key", None))
        c = Prover(ctx2, kp2, self.prover_key, self.verifier_key, self.transcript, self.pi_indexes)
        c.label = getattr(self, "label", b"plonk")
        c._owns_ctx = True
        return c

    def verifier(self, opening_key=None):
        """The ``Verifier`` half of ``PlonkKey::compile``'s pair (src/key.rs:316-325): same verification key, the
        commit key's ``verification_key()`` as opening key."""
        from .verifier import Verifier
        ok = opening_key if opening_key is not None else self.keypair.verification_key()
        return Verifier(getattr(self, "label", b"plonk"), self.verifier_key, ok, self.pi_indexes, self.size,
                        self.verifier_key["n"])

    def close(self):
        """Release the workspace and the side stream (the proving key stays with its buffers)."""
        if getattr(self, "_side", None) is not None:
            self._side.close()
            self._side = None
        if self._native is not None:
            self._native.close()
            self._native = None
        self._ws = None
        if getattr(self, "_owns_ctx", False):
            self.ctx.close()
            self._owns_ctx = False

    def _commit(self, buf, off=0, n=None):
        return self.keypair.commit(_View(buf, off, n)).affine()

    # ---------------------------------------------------------------- native round driver
    def _native_prover(self):
        """``zkp_prover`` over this key (csrc/create_proof.cu): rounds, transcript and the
        linearisation scalars in native code, one call per proof."""
        if self._native is None:
            from .ffi import NativeProver, ProvingKeyDesc
            from .key import SIGMAS
            pk, ref = self.prover_key, self.ctx.ref
            d = ProvingKeyDesc()
            d.k = pk.k
            comm = getattr(self.keypair, "native_comm", None)
            n = pk.n
            n8 = 8 * n if comm is None else pk.cosets[1] * n     # a sharded key holds this rank's cosets only
            for i, nm in enumerate(SELECTORS + SIGMAS):
                d.poly[i] = ref(pk.poly[nm], 0, n)
                d.eval8[i] = ref(pk.eval8[nm], 0, n8)
            d.linear8 = ref(pk.eval8["linear"], 0, n8)
            for j in range(4):
                d.sigma_evals[j] = ref(pk.sigma_evals[j], 0, n)
            d.roots = pk.roots.h
            gen = fr_to_mont1(self.verifier_key["generator"])
            for j in range(8):
                for l in range(4):
                    d.zh_inv[j][l] = int(pk.zh_inv[j, l])
            for l in range(4):
                d.generator[l] = int(gen[l])
            d.widget_mask = pk.widget_mask
            self._native = NativeProver(self.ctx, self.keypair.srs, d, (pk, self.keypair), comm, comm is not None)
        return self._native

    def _create_proof_native(self, tr, wa, bl):
        from .ffi import ZKP_ERR_DEGREE, ZkpError
        from .plonk_params import Error
        st = bytes(tr.strobe.state) + bytes([tr.strobe.pos, tr.strobe.pos_begin, tr.strobe.cur_flags])
        n = self.size
        np_ = self._native_prover()
        if isinstance(wa, WitnessValues):
            key = (id(wa.wires), id(wa.pi_indexes))
            if self._wiring_of != key:                    # same circuit shape as the last proof: already set
                np_.set_wiring(wa.wires, wa.pi_indexes)
                self._wiring_of = key
                self._wiring_keep = (wa.wires, wa.pi_indexes)   # keeps the ids alive
            rc, comms, evals, raw = np_.prove_witness(st, wa.witness_mont, wa.pi_values_mont, bl)
        else:
            wires_host = None if wa.wires_dev is not None else np.ascontiguousarray(wa.wires_mont).reshape(4 * n, 4)
            pi_host = None if wa.pi_dev is not None else np.ascontiguousarray(wa.dense_pi_mont)
            rc, comms, evals, raw = np_.prove(st, wires_host, wa.wires_dev, pi_host, wa.pi_dev, bl)
        if rc == ZKP_ERR_DEGREE:
            raise Error("polynomial degree exceeds the SRS")
        if rc:
            self.ctx.check(rc)
            raise ZkpError(rc)
        # wire_bytes: as serialised by the native driver (== to_bytes())
        return Proof.from_limbs(comms, evals, raw), list(wa.pi_values)

    # ---------------------------------------------------------------- create_proof
    def create_proof(self, blinders, circuit, trace=None):
        """``blinders``: the 11 scalars ``blind`` draws from the RNG, in draw order
        (2 each for a, b, o, d, then 3 for z; src/prover.rs:126-129,193).  ``circuit`` is a
        synthesized composer / ``SynthesizedCircuit`` / ``WitnessAssignment``.  Raises
        ``plonk_params.Error`` where the reference returns ``Err``.  Returns
        (Proof, public_inputs)."""
        n = self.size
        T = trace if trace is not None else None
        if isinstance(circuit, Plonk):
            circuit = SynthesizedCircuit.from_composer(circuit)
        use_native = self.native and trace is None and type(self.keypair) in (PlonkParams, ShardedNativeParams)
        if type(self.keypair) is ShardedNativeParams and not use_native:
            raise ValueError("a key sharded over GPUs proves through the native driver only")
        if isinstance(circuit, (WitnessAssignment, WitnessValues)):
            wa = circuit
        elif use_native:
            wa = WitnessValues.from_circuit(circuit)      # the wire gather happens on the device
        else:
            wa = WitnessAssignment.from_circuit(circuit, n)
        if isinstance(wa, WitnessValues) and not use_native:
            wa = wa.gathered(n)
        tr = self.transcript.clone()
        for pi in wa.pi_values:
            tr.append_scalar(b"pi", pi)
        bl = fr_to_mont(blinders)
        if use_native:
            return self._create_proof_native(tr, wa, bl)
        # only the Python-driven rounds (sharded / traced proofs) need the Python-side workspace: the native
        # driver owns an identical one (zkp_prover_create)
        ctx, pk, k = self.ctx, self.prover_key, self.prover_key.k
        ref = ctx.ref
        ws = self._workspace()
        proof = Proof()

        # round 1: wires -> iNTT -> blind -> commit (src/prover.rs:107-158)
        if wa.wires_dev is not None:      # witness already resident in HBM
            W = wa.wires_dev
        else:
            W = ws["W"]
            W.upload(wa.wires_mont.reshape(4 * n, 4))
        ctx.ntt_dev_batch(W, n, n, ws["P7"], ws["S"], k, True, False, 4)     # 4 wire iNTTs, one launch set
        for j in range(4):
            ctx.poly_blind(ws["wp"][j], 0, n, bl[2 * j:2 * j + 2])
        # The 8n-coset evaluations of a, b, c, d and PI depend on nothing the transcript still has to
        # produce: run them on a second stream while the wire commitments (whose bucket-reduction tail
        # leaves most of the GPU idle) are computed.  (reference order: src/prover.rs:229,
        # quotient_poly.rs:54-58,145 -- same values, earlier.)
        PI = ws["PI"]
        if wa.pi_dev is not None:
            ctx.ntt_dev(wa.pi_dev, n, PI, k, True, False)
        else:
            PI.upload(wa.dense_pi_mont)
            ctx.ntt_dev(PI, n, PI, k, True, False)
        k8, n8 = k + 3, 8 * n
        comm = self._comm()
        side = None
        if comm is None:
            side = self._side_context()
            ctx.sync()
            side.ntt_dev_batch(ws["P7"], ws["S"], n + 3, ws["E7"], n8, k8, False, True, 5)
        comms = [c.affine() for c in self.keypair.commit_batch([ws["wp"][j] for j in range(4)])]
        proof.a_comm, proof.b_comm, proof.c_comm, proof.d_comm = comms
        for lab, c in zip((b"a_w", b"b_w", b"c_w", b"d_w"), comms):
            tr.append_commitment(lab, c)

        # round 2: permutation accumulator (src/prover.rs:160-199)
        beta = tr.challenge_scalar(b"beta")
        tr.append_scalar(b"beta", beta)
        gamma = tr.challenge_scalar(b"gamma")
        ctx.perm_z(n, [ref(W, j * n, n) for j in range(4)], [ref(s, 0, n) for s in pk.sigma_evals], pk.roots,
                   fr_to_mont1(beta), fr_to_mont1(gamma), ws["Z"])
        ctx.ntt_dev(ws["Z"], n, ws["zp"], k, True, False)
        ctx.poly_blind(ws["zp"], 0, n, bl[8:11])
        proof.z_comm = self._commit(ws["zp"])
        tr.append_commitment(b"z", proof.z_comm)

        # round 3: quotient on the 8n coset (src/prover.rs:201-287, quotient_poly.rs)
        alpha = tr.challenge_scalar(b"alpha")
        rs = tr.challenge_scalar(b"range separation challenge")
        ls = tr.challenge_scalar(b"logic separation challenge")
        fs = tr.challenge_scalar(b"fixed base separation challenge")
        vs = tr.challenge_scalar(b"variable base separation challenge")
        ch7 = (alpha, beta, gamma, rs, ls, fs, vs)
        # L1 * alpha^2: idft of (alpha^2, 0, ..) has every coefficient alpha^2 / n (quotient_poly.rs:264-272)
        ctx.fill(ws["L1"], 0, n, fr_to_mont1(alpha * alpha % _r * pow(n, -1, _r) % _r))
        if comm is None:
            # z and L1 -> 8n coset (a, b, c, d, PI were transformed on the side stream during round 1)
            ctx.ntt_dev(ws["zp"], n + 3, ws["z8"], k8, False, True)
            ctx.ntt_dev(ws["L1"], n, ws["l18"], k8, False, True)
            side.sync()
        else:
            # sharded proof: the seven coset transforms are dealt out one per GPU; every rank then needs only
            # its slice of each (plus the 8 "next gate" evaluations), so the exchange is an all-to-all of
            # slices over NVLink -- 1/G of the bytes an all-gather of whole vectors would move
            G, rank, hl = comm.world, comm.rank, self._hl
            for r0 in range(0, 7, G):
                j = r0 + rank
                if j < 7:
                    ctx.ntt_dev(_View(ws["P7"], j * ws["S"], n + 3), n + 3, self._full8, k8, False, True)
                ctx.sync()
                comm.exchange_slices(self._t_E7[r0 * hl * 4:(r0 + G) * hl * 4], self._t_stage, self._t_send, n8)
        qa = QuotientArgs()
        wl = n8 if comm is None else self._hl       # whole coset vectors, or this rank's slices + halo
        for j in range(4):
            qa.wires[j] = ref(ws["e8"][j], 0, wl)
            qa.sigma[j] = ref(pk.eval8["s_sigma_%d" % (j + 1)], 0, n8)
        qa.z, qa.pi, qa.l1 = ref(ws["z8"], 0, wl), ref(ws["pi8"], 0, wl), ref(ws["l18"], 0, wl)
        qa.sliced = 0 if comm is None else 1
        for j, s in enumerate(SELECTORS):
            qa.sel[j] = ref(pk.eval8[s], 0, n8)
        qa.linear = ref(pk.eval8["linear"], 0, n8)
        chm = fr_to_mont(ch7)
        for j in range(7):
            for l in range(4):
                qa.challenges[j][l] = int(chm[j, l])
        for j in range(8):
            for l in range(4):
                qa.zh_inv[j][l] = int(pk.zh_inv[j, l])
        qa.widget_mask = pk.widget_mask
        Tb = ws["T"]
        if comm is None:
            ctx.quotient(k8, qa, Tb)
        else:
            # each GPU evaluates its slice of the 8n points; slices are all-gathered
            per = n8 // comm.world
            ctx.quotient(k8, qa, Tb, 0, comm.rank * per, per)
            ctx.sync()
            comm.all_gather_device(self._t_T, self._t_T[comm.rank * per * 4:(comm.rank + 1) * per * 4], self._t_stage)
        ctx.ntt_dev(Tb, n8, Tb, k8, True, True)   # coset_idft -> t coefficients
        tc = [c.affine() for c in self.keypair.commit_batch(
            [_View(Tb, 0, n), _View(Tb, n, n), _View(Tb, 2 * n, n), _View(Tb, 3 * n, 5 * n)])]
        proof.t_low_comm, proof.t_mid_comm, proof.t_high_comm, proof.t_4_comm = tc
        for lab, c in zip((b"t_low", b"t_mid", b"t_high", b"t_4"), tc):
            tr.append_commitment(lab, c)

        # round 4/5: evaluations, linearisation, openings (src/prover.rs:289-452)
        zc = tr.challenge_scalar(b"z_challenge")
        gen = self.verifier_key["generator"]
        zw = zc * gen % _r
        P = pk.poly
        at_z = [ref(Tb, 0, n8)] + [ref(ws["wp"][j], 0, n + 2) for j in range(4)] + \
            [ref(P[s], 0, n) for s in ("s_sigma_1", "s_sigma_2", "s_sigma_3", "q_arith", "q_c", "q_l", "q_r")]
        e1 = fr_from_mont(ctx.poly_eval(at_z, fr_to_mont1(zc)))
        e2 = fr_from_mont(ctx.poly_eval([ref(ws["wp"][0], 0, n + 2), ref(ws["wp"][1], 0, n + 2),
                                         ref(ws["wp"][3], 0, n + 2), ref(ws["zp"], 0, n + 3)], fr_to_mont1(zw)))
        ev = dict(zip(("a_eval", "b_eval", "c_eval", "d_eval", "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval",
                       "q_arith_eval", "q_c_eval", "q_l_eval", "q_r_eval"), e1[1:]))
        t_eval = e1[0]
        ev["a_next_eval"], ev["b_next_eval"], ev["d_next_eval"], ev["perm_eval"] = e2
        scal = linearization_scalars(n, ch7 + (zc,), ev)
        lin_refs = [ref(ws["zp"], 0, n + 3) if nm == "z" else ref(P[nm], 0, n) for nm, _ in scal]
        ctx.poly_lincomb(lin_refs, fr_to_mont([s for _, s in scal]), ws["R"], 0, n + 3)
        ev["r_poly_eval"] = fr_from_mont(ctx.poly_eval([ref(ws["R"], 0, n + 3)], fr_to_mont1(zc)))[0]
        for nm in ("a_eval", "b_eval", "c_eval", "d_eval", "a_next_eval", "b_next_eval", "d_next_eval",
                   "s_sigma_1_eval", "s_sigma_2_eval", "s_sigma_3_eval", "q_arith_eval", "q_c_eval", "q_l_eval",
                   "q_r_eval", "perm_eval"):
            tr.append_scalar(nm.encode(), ev[nm])
        tr.append_scalar(b"t_eval", t_eval)
        tr.append_scalar(b"r_eval", ev["r_poly_eval"])
        # W_z: (t_low + z^n t_mid + z^2n t_high + z^3n t_4) + v r + v^2 a + .. , divided by (X - z)
        z_n = pow(zc, n, _r)
        v1 = tr.challenge_scalar(b"v_challenge")
        vp = [pow(v1, i, _r) for i in range(9)]
        refs = [ref(Tb, 0, n), ref(Tb, n, n), ref(Tb, 2 * n, n), ref(Tb, 3 * n, 5 * n), ref(ws["R"], 0, n + 3)] + \
            [ref(ws["wp"][j], 0, n + 2) for j in range(4)] + [ref(P[s], 0, n) for s in ("s_sigma_1", "s_sigma_2", "s_sigma_3")]
        sc = [1, z_n, z_n * z_n % _r, pow(z_n, 3, _r)] + vp[1:]
        ctx.poly_lincomb(refs, fr_to_mont(sc), ws["AGG"], 0, 5 * n)
        ctx.poly_div_linear(ref(ws["AGG"], 0, 5 * n), fr_to_mont1(zc), ws["WZ"])
        # The second v_challenge follows the first with no transcript append in between
        # (src/prover.rs:435-450: the opening commitments are never appended by the prover), so both
        # witnesses can be built first and committed as one batch of two.
        v2 = tr.challenge_scalar(b"v_challenge")
        refs = [ref(ws["zp"], 0, n + 3), ref(ws["wp"][0], 0, n + 2), ref(ws["wp"][1], 0, n + 2),
                ref(ws["wp"][3], 0, n + 2)]
        ctx.poly_lincomb(refs, fr_to_mont([pow(v2, i, _r) for i in range(4)]), ws["SAGG"], 0, n + 3)
        ctx.poly_div_linear(ref(ws["SAGG"], 0, n + 3), fr_to_mont1(zw), ws["WZW"])
        wc = self.keypair.commit_batch([_View(ws["WZ"], 0, 5 * n - 1), _View(ws["WZW"], 0, n + 2)])
        proof.w_z_chall_comm, proof.w_z_chall_w_comm = wc[0].affine(), wc[1].affine()
        proof.evaluations = {nm: ev[nm] for nm in EVAL_NAMES}
        if T is not None:
            T.update({"challenges": ch7, "z_challenge": zc, "t_eval": t_eval, "workspace": ws})
        return proof, list(wa.pi_values)


class WitnessAssignment:
    """What a proof needs from the synthesized circuit, already in limb form: the four wire
    columns over the n-domain (src/prover.rs:109-119) and the dense public inputs
    (src/lib.rs:206-219).  Building it is the host-side gather; ``create_proof`` accepts it
    directly so repeated proofs do not redo the conversion."""

    def __init__(self, wires_mont, dense_pi_mont, pi_values):
        self.wires_mont = wires_mont          # (4, n, 4) uint64
        self.dense_pi_mont = dense_pi_mont    # (n, 4) uint64
        self.pi_values = list(pi_values)
        self.wires_dev = None
        self.pi_dev = None

    def to_device(self, ctx):
        """Upload once; later proofs read the witness from HBM."""
        n = self.wires_mont.shape[1]
        self.wires_dev = ctx.upload(self.wires_mont.reshape(4 * n, 4))
        self.pi_dev = ctx.upload(self.dense_pi_mont)
        return self

    @classmethod
    def from_circuit(cls, circ, n):
        wit = fr_to_mont(circ.witness)                      # (num_witness, 4)
        wires = np.zeros((4, n, 4), dtype=np.uint64)
        idx = np.asarray(circ.wires, dtype=np.int64)
        for j in range(4):
            wires[j, :idx.shape[1]] = wit[idx[j]]
        dense = np.zeros((n, 4), dtype=np.uint64)
        if len(circ.pi_indexes):
            dense[np.asarray(circ.pi_indexes, dtype=np.int64)] = fr_to_mont(circ.pi_values)
        return cls(wires, dense, circ.pi_values)


class WitnessValues:
    """A proof's inputs before the wire gather: the witness values, the (4, m) wire -> witness table
    of the synthesized circuit and the public inputs.  ``create_proof`` ships the values and gathers
    the four wire columns on the device (src/prover.rs:109-119), so the host never builds them."""

    def __init__(self, witness_mont, wires, pi_indexes, pi_values):
        self.witness_mont = witness_mont                     # (num_witness, 4) uint64 Montgomery
        self.wires = wires                                   # (4, m) uint32
        self.pi_indexes = np.asarray(pi_indexes, dtype=np.uint32)
        self.pi_values = list(pi_values)
        self.pi_values_mont = fr_to_mont(self.pi_values) if len(self.pi_values) else np.zeros((0, 4), dtype=np.uint64)

    @classmethod
    def from_circuit(cls, circ):
        return cls(fr_to_mont(circ.witness), np.ascontiguousarray(circ.wires, dtype=np.uint32), circ.pi_indexes,
                   circ.pi_values)

    def gathered(self, n):
        """The host-side gather (what ``WitnessAssignment.from_circuit`` builds)."""
        wires = np.zeros((4, n, 4), dtype=np.uint64)
        idx = self.wires.astype(np.int64)
        for j in range(4):
            wires[j, :idx.shape[1]] = self.witness_mont[idx[j]]
        dense = np.zeros((n, 4), dtype=np.uint64)
        if len(self.pi_values):
            dense[self.pi_indexes.astype(np.int64)] = self.pi_values_mont
        return WitnessAssignment(wires, dense, self.pi_values)
