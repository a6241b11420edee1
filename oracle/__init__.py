"""Test infrastructure only: CPU oracles for the parse hot path.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker.  See ``oracle/README.md``.
"""
