"""Drop-in for /root/reference/tf_ops/3d_interpolation/tf_interpolate.py (same function names and argument order)."""
from learn_region_grow_b200.tfops import three_nn, three_interpolate, three_interpolate_grad  # noqa: F401
