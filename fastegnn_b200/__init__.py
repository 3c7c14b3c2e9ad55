"""fastegnn_b200 -- B200-native FastEGNN layer path (hand-written sm_100a CUDA behind a C ABI).

Public surface (mirrors models/FastEGNN.py of GLAD-RUC/FastEGNN):
    FastEGNN, E_GCL_vel            drop-in module classes
    FastRF                         the radial-field sibling (models/FastRF.py) on the same kernels
    VNEGNN                         the virtual-node EGNN sibling (models/VNEGNN.py): A2A / A2V / V2A stages on the same kernels
    mmd_loss                       the MMD regulariser of utils/train.py:111-165 as one op
    mse_mmd_loss                   the step's whole loss (MSE + weight * MMD, utils/train.py:104-163) as one op
    CsrGraph                       the once-per-batch CSR graph prep; CsrGraph.from_radius builds the graph on the device
    FusedAdam                      torch.optim.Adam's step (utils/train.py:168-170) as one launch over flat buffers
    PipelinedStep                  a training step as a CUDA graph with double-buffered inputs (H2D of batch k+1 under step k)
Importing the package loads fastegnn_b200/_C/libfegnn.so and raises if it is absent.
"""
from . import _lib  # noqa: F401  (fails loudly when the shared library has not been built)
from .FastEGNN import E_GCL_vel, FastEGNN, unsorted_segment_mean, unsorted_segment_sum  # noqa: F401
from .FastRF import FastRF  # noqa: F401
from .VNEGNN import VNEGNN  # noqa: F401
from .ops import CsrGraph, mmd_loss, mse_mmd_loss  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .runtime import PipelinedStep  # noqa: F401

__all__ = ["FastEGNN", "FastRF", "VNEGNN", "E_GCL_vel", "mmd_loss", "mse_mmd_loss", "CsrGraph", "FusedAdam", "PipelinedStep", "unsorted_segment_sum", "unsorted_segment_mean"]
