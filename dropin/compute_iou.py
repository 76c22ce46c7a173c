"""Drop-in for the hot-path functions of the reference's compute_iou.py (fast_hist, per_class_iu, label_mapping)."""
from adaptsegnet_b200.compute_iou import fast_hist, label_mapping, per_class_iu  # noqa: F401
