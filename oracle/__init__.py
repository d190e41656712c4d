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
  * ``wct_tf``, ``adain``, ``wct_style_swap``/``style_swap``, encoder, decoder, level wiring -- PINNED AT SOURCE LEVEL:
    ``tests/golden/make_pipeline_golden.py`` imports the reference's own model.py /
    ops.py / vgg_normalised.py / torchfile.py unmodified and evaluates them over
    ``tests/golden/np_tf1.py`` (an eager NumPy stand-in for the TensorFlow/Keras calls
    they make); the outputs are ``tests/golden/pipeline_*.npz`` and the oracle matches
    them to 1e-9 in float64.  TensorFlow's own kernels (conv, svd) are NOT exercised --
    TensorFlow is not installable offline -- and no trained weights exist on disk, so
    real-TF / real-weight parity stays unverified (file:line cited on every function).
"""
