"""CPU oracle for the pulser-diff hot path.  TEST INFRASTRUCTURE ONLY.

This package is a slow, faithful restatement (torch CPU, sparse COO, tape
autograd) of what the reference does on the path
``TorchEmulator.run -> pyqtorch.sesolve/mesolve`` fed by
``Hamiltonian.build_ham_tensor``.  It exists to CHECK the CUDA path and to be
timed as the CPU baseline.  Nothing under ``pulser_diff_b200/`` imports it;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may.

PARITY STATUS: **unpinned by the reference at 1e-10**.  The reference ships no
golden vectors and its arithmetic lives in ``pyqtorch`` (unpinned dependency,
not installable here).  The restatement is pinned
only by the numbers printed in the reference notebooks (4-6 significant
digits; ``tests/golden/notebook_kats.json``), which fix every convention:
basis order, C6, interpolation quirk, solver semantics.
"""
