"""Import shim (reference nflows/nn/nde/made.py:6 imports matplotlib.pyplot at module top)."""
