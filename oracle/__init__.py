"""CPU oracle for the VIP-ANT InfoNCE / retrieval-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vipant_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import or execute it, and there
only as the checker / the CPU baseline, never as the thing shipped.

Parity pinning: the reference ships no tests and no golden vectors (SURVEY.md
section 4), so the restatements in this package are pinned against the
reference's OWN code executed unmodified in the build container
(``oracle/reference_loader.py`` + ``oracle/make_golden.py``); the vectors it
produced are committed under ``tests/golden/`` and ``tests/test_oracle.py``
re-checks the restatement against them on every run (CPU-only, no access to
``/root/reference`` needed).
"""
