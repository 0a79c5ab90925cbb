"""CPU oracle for the BranchedGP hot path (AssignGP variational bound + gradient).

TEST INFRASTRUCTURE ONLY.  Nothing under ``branchedgp_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

PARITY UNPINNED: the reference (TensorFlow 2.4.0 / GPflow 2.1.2 / TFP 0.11.1, none installable here and
none vendored under /root/reference) stores no numeric golden values for the bound or its gradient
(SURVEY.md §4, §8c).  The oracle is therefore a *restatement* of the reference formulas:

* ``oracle.bgp_torch``  -- op-for-op float64 restatement on torch-CPU, gradients by autograd
                           (mirrors the reference's "TF autodiff" cost structure);
* ``oracle.bgp_numpy``  -- independent numpy/SciPy-LAPACK forward + hand-derived adjoint.

The two are checked against each other (tests/test_oracle_*.py) and against the reference's own
relational tests (testing/test_KL.py:102-103, testing/MBGP/test_equivalence_multi_to_single.py:54-56,
testing/test_pZ_construct.py:86-97).  Every parity claim in this repo carries the note
"oracle = restatement of TF 2.4.0 / GPflow 2.1.2 semantics".

The one set of outputs of the REAL reference that ships with it -- notebooks/syntheticdata/syntheticDataRun.p, written by
notebooks/runsynteticData.py:27-52 (FitModel, M = 10, maxiter = 20, 40 genes x 6 branching points) -- is committed as
tests/golden/synthetic20_reference_run.npz; ``oracle.fit_ref`` reproduces its argmax branching points and log Bayes factors
(tests/test_reference_run_cpu.py).  That pins the drop-in outputs loosely (the pickle comes from an older TF / GPflow build
and unconverged fits); the 1e-8 parity of the bound itself stays unpinned.
"""
