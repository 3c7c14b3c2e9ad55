"""Drop-in for the reference's models/FastRF.py: `from models.FastRF import FastRF` (main_protein.py:18) resolves
to the B200-native implementation in fastegnn_b200."""
from fastegnn_b200.FastEGNN import unsorted_segment_mean, unsorted_segment_sum  # noqa: F401
from fastegnn_b200.FastRF import E_GCL_vel, FastRF  # noqa: F401
