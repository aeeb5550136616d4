"""Import-only stub (networks.py:3 imports timm; the hot path never calls it)."""
