"""TEST INFRASTRUCTURE ONLY -- CPU/torch restatement of the F Lite denoise hot path.

Nothing under ``oracle/`` is product code.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs, and there only as the checker / the reported CPU baseline -- never as the thing
that is measured or shipped.  The product path (``f-lite_b200``) calls hand-written
sm_100a CUDA kernels through ``libflite_b200.so`` and raises when that library is
missing; it never routes through this package.

Parity status: the reference (sippycoder/f-lite) ships no tests, golden vectors or
fixtures so the oracle is pinned *differentially*: the
unmodified ``/root/reference/f_lite/model.py`` is imported in the build container through
stub modules (``oracle/ref_shim.py``), run on seeded inputs, and its outputs are committed
as fixtures under ``tests/golden/`` by ``oracle/make_golden.py``.  ``tests/test_oracle.py``
checks this restatement against those fixtures on every CPU run.
"""
