"""Drop-in for /root/reference/tf_ops/grouping/tf_grouping.py (same function names and argument order)."""
from learn_region_grow_b200.tfops import query_ball_point, select_top_k, group_point, group_point_grad, knn_point  # noqa: F401
