"""TEST INFRASTRUCTURE — not product code.

CPU restatements of the reference's algorithm for the hot path (``oracle.port``), a
loader that executes the reference's own torch-only modules in place from
``/root/reference`` (``oracle.ref_loader``; build container only) and the script that
generates the golden fixtures under ``tests/golden/`` (``oracle.make_golden``).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package.  Nothing under ``dance_b200/`` does.
"""
