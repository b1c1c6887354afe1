"""CPU oracle for the DreamScene rasterizer hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product
(``dreamscene_b200`` / ``diff_gaussian_rasterization``) never does.
See ``oracle/splat_ref.py`` for the parity-pinning statement.
"""
