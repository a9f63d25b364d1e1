"""Importable alias for the ``dusk-plonk_b200/`` package directory (a hyphen is not a
valid Python identifier).  All code lives in ``dusk-plonk_b200/``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dusk-plonk_b200")
__path__.insert(0, _real)

from .ffi import *  # noqa: E402,F401,F403
from .poly_commit import Fft, Coefficients, PointsValue, Commitment  # noqa: E402,F401
from .plonk_params import PlonkParams, ShardedNativeParams, Error  # noqa: E402,F401
from host_mirror.composer import Plonk, Constraint, SynthesizedCircuit  # noqa: E402,F401
from .key import PlonkKey  # noqa: E402,F401
from .prover import Prover, Proof, WitnessAssignment, WitnessValues  # noqa: E402,F401
from .transcript import Transcript  # noqa: E402,F401
from .verifier import Verifier, EvaluationKey  # noqa: E402,F401
from . import sharding  # noqa: E402,F401
