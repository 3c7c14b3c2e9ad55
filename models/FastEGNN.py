"""Drop-in for the reference's models/FastEGNN.py: `from models.FastEGNN import FastEGNN`
(main_nbody.py:17, main_protein.py:20, main_simulation.py:17, equivariant_test.py:9) resolves
to the B200-native implementation in fastegnn_b200."""
from fastegnn_b200.FastEGNN import E_GCL_vel, FastEGNN, unsorted_segment_mean, unsorted_segment_sum  # noqa: F401
