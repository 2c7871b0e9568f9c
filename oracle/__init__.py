"""TEST INFRASTRUCTURE ONLY.

CPU restatements ("oracles") of the reference's Light-Head R-CNN hot path.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package; the product (``x-detector_b200/``) never does.
"""
