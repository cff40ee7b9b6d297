"""CPU oracle for the batched VI-ESKF hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / CPU baseline.  The product path (``dvi_ekf_b200``) never imports
this package and fails loudly when its CUDA library is missing.

Parity status: PINNED against the reference's only numeric artefacts
(``data/trajs/kf_best_mandala0_mono.txt`` and ``imu_ref_mandala0_mono*.txt``,
packed into ``tests/golden/reference_golden.npz`` by
``tests/golden/make_golden.py``) for the legacy preset; HEAD-only deltas
(Q7 frozen-DOF zeroing, Q11 xyz Euler gradient) and the notch!=0 /
interframe>1 filter arithmetic are unpinned by reference artefacts and are
anchored on the literal sympy transcription in ``oracle/symbolic_check.py``.
"""
