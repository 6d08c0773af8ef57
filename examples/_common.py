import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def str2bool(v):
    """The reference declares its boolean flags with type=bool, so any non-empty string is True
    (main_NonLinElliptic2d.py:44-45).  Here "false"/"0"/"no" mean False."""
    return str(v).lower() not in ("false", "0", "no", "")
