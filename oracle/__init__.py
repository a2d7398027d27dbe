"""CPU oracle for the DiffSplitting sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffsplitting_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  It is a from-scratch functional
restatement (torch-CPU fp32/fp64 tensor ops and numpy integer arithmetic) of the
reference algorithms, each function citing the reference file:line it follows.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
modules from ``/root/reference`` (available only in the build container),
checks every oracle function against them and writes the small fixtures under
``tests/golden/`` that the CPU test-suite re-checks the oracle against.  The
tiling oracle is additionally pinned by the reference's own
``tests/test_tiling_setup.py`` construction (arange frames -> tiles -> stitch ->
exact equality).
"""
