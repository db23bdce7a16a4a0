"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the structured-layer hot path.

This package restates, on the CPU (torch CPU ops / numpy), the forward pass of
every layer in ``structurednets.layers`` of the reference (MatthiasKi/structurednets)
so that torch autograd supplies the matching backward.  Each function cites the
reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
``structurednets_b200`` never imports anything from here and has no CPU fallback.

Parity pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the *unmodified reference modules* imported in
the build container (``oracle/ref_import.py`` + ``tests/golden/make_golden.py``);
the resulting fixtures are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.

Second, torch-independent pin for the SSS forward: the reference's own C program
``speed_comparison/run.c`` is compiled unmodified from where it lies into
``oracle/_ref/run`` (``oracle/run_c.py``, ``oracle/Makefile``; git-ignored) and must accept
the oracle's output -- and the CUDA kernels' -- as its "checksum" vector
(``tests/test_run_c_reference.py``).
"""
