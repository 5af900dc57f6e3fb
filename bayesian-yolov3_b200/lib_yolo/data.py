"""`Prior` holder of the reference (lib_yolo/data.py:80-86); the GT-encoding helpers of that file are training-only."""
from byolo.priors import Prior  # noqa: F401
