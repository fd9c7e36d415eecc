"""`from lamb_amp_opt.fused_lamb import FusedLAMBAMP` (run_pretraining.py:43): the 3-launch multi-tensor LAMB of libdle_b200.so behind
the reference optimizer's interface (lamb_amp_opt/fused_lamb/fused_lamb.py:10-307)."""
from deeplearningexamples_b200.lamb import FusedLAMBAMP  # noqa: F401
