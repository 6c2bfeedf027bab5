"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package restates, on the CPU (torch-CPU / numpy, float64 accumulation where
noted), the algorithms of the reference's attention stack, Detect decode and NMS
(SURVEY.md section 8a).  Every function cites the reference ``file:line`` it follows.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- there only as the
checker or the timed CPU baseline.  The product package
(``small-object-detection-transformers_b200``) never imports it and has no CPU
fallback: it raises when the CUDA library is missing.

Parity pinning: the reference has no golden vectors or tests of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the UNMODIFIED
reference run in the build container (``tests/golden/make_golden.py`` imports it from
``/root/reference`` through three stub modules) and committed as fixtures under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every oracle function
against them.  The greedy suppression inside NMS lives in a third-party dependency
(``torchvision.ops.nms``, requirement ``torchvision>=0.8.1``, 0.26.0 installed); its
published algorithm is restated in ``oracle/nms_ref.py`` and pinned against the
installed torchvision on the same fixtures.
"""
