"""oracle/ — TEST INFRASTRUCTURE ONLY.

A CPU restatement (PyTorch-CPU fp32 functional ops + numpy integer arithmetic) of the reference
algorithm for the U-Net training-step hot path of LorenzoFramba/Continual-Learning.  Every function
cites the reference file:line it follows.  Nothing here is imported by the product package
`continual_learning_b200`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may use it, and only as the checker or the timed CPU baseline.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the reference modules themselves (`models/unet.py`, `metrics.py`) imported from
/root/reference in the build container: `tests/golden/make_golden.py` generated the committed
fixtures, and `tests/test_oracle_pinned.py` re-checks the oracle against the live reference whenever
/root/reference is present.  The distillation term (continual_ref.py) is NOT in the reference:
"parity unpinned" — its specification is BASELINE.json:north_star / SURVEY.md §8(c).
"""
