"""CPU oracle for the 3dgan training-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`3dgan_b200/`) may
import this; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do, and only as the checker / the
reported CPU baseline.

PARITY UNPINNED: the arithmetic of the reference
path lives in TensorFlow ~1.2 (unpinned, "a few commits ahead of 1.2",
/root/reference/doc/guide.tex:50), which is neither vendored in the reference
nor installable here (no network, Python 3.12).  The only golden vectors the
reference's own tests hold for this path are the `hem.rmse` known answers
(/root/reference/hem/ops/test_losses.py:7-27), which `tests/test_oracle.py`
checks.  Everything else is a restatement of published TF-1.x semantics
(SURVEY.md Appendix A), anchored on the reference's call sites and cross-checked
by independent derivations (autograd adjoints, float64 finite differences).
"""
