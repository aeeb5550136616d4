"""Import-only stub (networks.py:1 imports antialiased_cnns; the hot path never calls it)."""
