"""Synthetic weights come from the product-side generator (same dict feeds oracle and engine in parity tests)."""
from groma_b200.synth import make_state_dict  # noqa: F401
