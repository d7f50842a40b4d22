"""CPU oracle for the ActiveZero stereo hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product package ``activezero_b200`` never imports this package and raises if
its CUDA library is missing.

Parity status
-------------
The reference ships no golden vectors.  The restatements in
``oracle/stereo_oracle.py`` are pinned against the reference's *own Python*,
imported from ``/root/reference`` and executed on seeded inputs by
``oracle/make_golden.py``; the resulting input/output pairs are committed under
``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.

One exception: the group-wise-correlation volume has no reference
implementation at all -- **parity unpinned** for that op;
its oracle is this package's own definition.
"""
