"""CPU oracle for the RAFT-family inference hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker (or as the timed CPU
baseline) -- never as the thing shipped.  The product path (``ptlflow_b200``) must
fail loudly when its CUDA library is missing; it never routes through this package.

Contents
--------
raft_oracle.py   plain torch-fp32 restatement of the reference algorithm, one function
                 per row of SURVEY.md section 8(a), each citing the reference file:line.
synth.py         platform-stable (numpy Philox) synthetic weights / images so that the
                 reference (in the build container) and the product (on the GPU box)
                 see bit-identical parameters without shipping multi-MB fixtures.
ref_shim.py      loads the *real* reference modules from /root/reference (build
                 container only; that path does not exist on the GPU box).
make_golden.py   runs the real reference through ref_shim and writes tests/golden/*.npz.

Parity status: PINNED.  The reference's own tests hold no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
generated here by make_golden.py and committed under tests/golden/.
"""
