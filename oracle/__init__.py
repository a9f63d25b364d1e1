"""CPU oracle for the zkplonk hot path (TEST INFRASTRUCTURE -- not product code).

This package restates, on the CPU, the algorithms the reference prover reaches on
its hot path: Fr / Fq arithmetic, the radix-2 NTT family of ``poly_commit::Fft``,
the KZG10 commit (= G1 multi-scalar multiplication) of ``PlonkParams::commit``,
and the prover-round formulas of ``src/prover.rs`` / ``src/prover/quotient_poly.rs``
/ ``src/permutation.rs`` / ``src/prover/linearization_poly.rs``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker.  The
product (``dusk-plonk_b200``) never imports ``oracle`` and has no CPU fallback.

PARITY UNPINNED.  The reference (``/root/reference``) cannot be compiled here (no
Rust toolchain) and the crates that hold its arithmetic (``poly-commit``,
``bls-12-381``, ``zksnarks``, ``zkstd``, ``ec-pairing``, ``jub-jub`` -- path
dependencies with NO pinned version, ``Cargo.toml:28-33``, no lockfile) are absent
from the tree.  The reference's own tests hold no golden vector for a commitment, an
NTT output or proof bytes.  What *is* pinned in-tree and checked by
``tests/test_oracle_constants.py``:

* ``MINUS_ONE`` limbs (``src/lib.rs:583-588``) => Fr is 4 x u64 little-endian
  Montgomery form with R = 2^256;
* ``K1, K2, K3 = 7, 13, 17`` (``src/permutation.rs:28-30``);
* the domain ordering ``elements[i] = w^i`` (``src/permutation.rs:148-166, 764-772``).

Everything else rests on the public BLS12-381 parameters and on the fact that an
MSM result and a DFT over Fr are unique functions of their inputs.
"""
