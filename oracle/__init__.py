"""CPU oracle for the speaker-embedding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker or as
the timed CPU baseline -- never as a fallback of the CUDA path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the reference's own
``Modules.GE2E`` / ``Modules.GE2E_Loss`` (``/root/reference/Modules.py``) in the
authoring container and stores their outputs on seeded inputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures on every CPU run.
"""
