"""Drop-in for /root/reference/tf_ops/sampling/tf_sampling.py (same function names and argument order)."""
from learn_region_grow_b200.tfops import farthest_point_sample, gather_point, gather_point_grad, prob_sample  # noqa: F401
