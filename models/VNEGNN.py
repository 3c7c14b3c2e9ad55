"""Import-path shim: `from models.VNEGNN import VNEGNN` (main_nbody.py:19) resolves to the B200-native sibling."""
from fastegnn_b200.VNEGNN import EGCL_A2A, EGCL_A2V, EGCL_V2A, VNEGNN  # noqa: F401
