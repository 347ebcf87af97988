"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
``panopticdiffusionmodels_b200`` never imports it and has no CPU fallback.
"""
