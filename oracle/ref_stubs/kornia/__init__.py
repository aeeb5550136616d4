"""Import-only stub so /root/reference's modules load in the build container (oracle/make_golden.py).
The hot path never calls kornia; geometry_utils.py:1 and generic_utils.py:85-92 only need the name."""
from . import filters  # noqa: F401
