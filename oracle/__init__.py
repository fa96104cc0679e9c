"""CPU oracle for the SynthSR training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be imported by
``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs, and only as the checker / the CPU baseline.  The product path (``synthsr_b200``) never imports it and
fails loudly when the CUDA library is missing.

PARITY UNPINNED: the reference (BBillot/SynthSR) has no tests, golden vectors or fixtures for this path, and
its arithmetic lives in un-vendored third-party modules (tensorflow-gpu==2.0.0, Keras==2.3.1) that cannot be
installed in this image (Python 3.12, no network).  The oracle is therefore a line-by-line restatement of the
reference's Python graph code (every function cites the reference file:line it follows) plus the published
definitions of the TF/Keras ops it calls; it is pinned only against the reference's docstring worked examples
and analytic known-answer tests (tests/test_oracle_*.py).
"""
