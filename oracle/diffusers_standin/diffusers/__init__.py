"""ORACLE-ONLY stand-in for the handful of `diffusers` symbols that the reference's models/unet_3d_condition.py and
models/unet_3d_blocks.py import (diffusers itself is not installed and not vendored in /root/reference).
It lets those two reference files be imported UNMODIFIED so that the oracle's wiring restatement
(oracle/unet3d_ref.py) can be pinned against the reference's own code.  The arithmetic lives in oracle/leaves.py.
Never imported by the product package."""
__version__ = "0.0-standin"
