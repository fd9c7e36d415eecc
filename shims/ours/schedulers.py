"""`import schedulers` for the unmodified reference scripts (PolyWarmUpScheduler / LinearWarmUpScheduler, schedulers.py:90-136)."""
from deeplearningexamples_b200.schedulers import LinearWarmUpScheduler, PolyWarmUpScheduler  # noqa: F401
