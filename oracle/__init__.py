"""CPU oracle for the AMPConv hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker (or as the thing
timed for the CPU baseline) -- never as a fallback for the CUDA path.

Parity status: PINNED.  ``oracle/gen_golden.py`` executes the reference's own
``src/ampnet/conv/amp_conv.py`` (loaded by path, unmodified, behind the PyG
stand-in of ``oracle/pyg_stub.py``) in this container and commits its inputs
and outputs under ``tests/golden/``; ``tests/test_oracle.py`` holds both
restatements in this package to those vectors and to the one known-answer test
the reference ships (``synthetic_benchmark/testing_message_passing_pyg.py:37-40``).
"""
