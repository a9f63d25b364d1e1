"""``Verifier`` (host mirror of ``src/verifier.rs:15-81``) over the native verifier glue of libzkp_b200.so
(csrc/verifier.cu: ``Proof::verify`` src/prover/proof.rs:70-383, ``batch_check``
src/commitment_scheme.rs:24-66, BLS12-381 ate pairing).  Host code only: constant-size work, no GPU."""
import ctypes

import numpy as np

from .ffi import VerifierKeyDesc, ZKP_ERR_VERIFY, ZkpError, load_library
from .field import fr_to_mont, fr_to_mont1, g1_to_mont
from .key import SIGMAS
from host_mirror.composer import SELECTORS
from .plonk_params import Error
from .prover import COMM_NAMES, EVAL_NAMES
from .transcript import Transcript


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class EvaluationKey:
    """``poly_commit::EvaluationKey`` (src/commitment_scheme.rs:51-58): the G2 side of the SRS.  ``beta_h`` =
    [tau]_2 as 24 uint64 (x0 x1 y0 y1, Montgomery); g = [1]_1 and h = [1]_2 are the fixed generators."""

    def __init__(self, beta_h):
        self.beta_h = np.ascontiguousarray(beta_h, dtype=np.uint64).reshape(24)

    @classmethod
    def from_tau(cls, tau_mont):
        """For an SRS generated from a known tau (``PlonkParams.setup_synthetic``)."""
        out = np.zeros(24, dtype=np.uint64)
        t = np.ascontiguousarray(tau_mont, dtype=np.uint64).reshape(4)
        rc = load_library().zkp_g2_generator_mul(_ptr(t), _ptr(out))
        if rc:
            raise ZkpError(rc)
        return cls(out)


class Verifier:
    def __init__(self, label, verifier_key, opening_key, public_input_indexes, size, constraints):
        self.verifier_key = verifier_key
        self.opening_key = opening_key
        self.public_input_indexes = [int(i) for i in public_input_indexes]
        self.size = size
        self.transcript = Transcript.base(label, verifier_key.transcript_list(), constraints)   # src/verifier.rs:33-34
        d = VerifierKeyDesc()
        d.k = size.bit_length() - 1
        d.constraints = constraints
        for i, nm in enumerate(SELECTORS + SIGMAS):
            limbs = g1_to_mont(verifier_key[nm])
            for l in range(12):
                d.commitments[i][l] = int(limbs[l])
        self._desc = d

    def verify(self, proof, public_inputs):
        """Ok(()) -> None; Err -> ``plonk_params.Error`` (InconsistentPublicInputsLen / ProofVerificationError)."""
        if len(public_inputs) != len(self.public_input_indexes):
            raise Error("InconsistentPublicInputsLen: expected %d, provided %d" %
                        (len(self.public_input_indexes), len(public_inputs)))
        tr = self.transcript
        st = np.frombuffer(bytes(tr.strobe.state) + bytes([tr.strobe.pos, tr.strobe.pos_begin, tr.strobe.cur_flags]),
                           dtype=np.uint8).copy()
        comms = np.zeros((11, 12), dtype=np.uint64)
        for i, c in enumerate(COMM_NAMES):
            comms[i] = g1_to_mont(getattr(proof, c))
        evals = np.ascontiguousarray(fr_to_mont([proof.evaluations[e] for e in EVAL_NAMES]))
        idx = np.asarray(self.public_input_indexes, dtype=np.uint32)
        vals = np.ascontiguousarray(fr_to_mont(list(public_inputs))) if len(public_inputs) else np.zeros((0, 4), dtype=np.uint64)
        rc = load_library().zkp_verify(ctypes.byref(self._desc), _ptr(self.opening_key.beta_h), _ptr(st), _ptr(comms),
                                       _ptr(evals), _ptr(idx) if idx.size else None, _ptr(vals) if idx.size else None,
                                       idx.size)
        if rc == ZKP_ERR_VERIFY:
            raise Error("ProofVerificationError")
        if rc:
            raise ZkpError(rc)
