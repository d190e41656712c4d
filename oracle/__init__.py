"""CPU oracle for the WCT inference hot path (TEST INFRASTRUCTURE ONLY).

Everything under ``oracle/`` restates, on the CPU, the algorithm of the
reference (eridgd/WCT-TF) for the path
``stylize.py -> WCT.predict -> [VGG19 encoder -> WCT|AdaIN -> decoder] x levels``.

It is a checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``wct_tf_b200`` never imports ``oracle``.

Pinning status (see DESIGN.md "Oracle"):
  * ``wct_np``  -- PINNED against the reference's own ``ops.wct_np`` executed
    in the build container (``tests/golden/make_golden.py`` imports
    ``/root/reference/ops.py`` with TensorFlow/Keras stubbed and commits the
    outputs as ``tests/golden/wct_np_*.npz``).
  * ``wct_tf``, ``adain``, encoder, decoder, level wiring -- PARITY UNPINNED:
    TensorFlow/Keras are not installable here and the reference ships no
    tests / golden vectors / weights, so these are restatements of the
    reference source only (file:line cited on every function).
"""
