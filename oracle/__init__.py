"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the audiogan hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline -- never as the thing measured or shipped.

Parity status: **parity unpinned by the reference** -- BarclayII/audiogan ships
no tests, golden vectors or known-answer values.  The
oracle is pinned instead against *outputs of the reference itself run here*:
``oracle/ref_loader.py`` executes the reference's own class definitions
(``/root/reference/audiogan.py:1-552``) on stock PyTorch fp32 CPU, and
``oracle/make_golden.py`` (committed) dumps their outputs into
``tests/golden/``.  ``oracle/restated.py`` is the stand-alone restatement that
travels to the GPU box (``/root/reference`` does not exist there); the CPU test
suite checks it against those golden vectors and, when ``/root/reference`` is
present, against the exec'd reference classes directly.
"""
