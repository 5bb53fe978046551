"""Host helpers the reference's eval scripts import from `groma.utils` (reference `groma/utils.py`)."""
import torch


def disable_torch_init():
    """Skip the default nn.Linear / nn.LayerNorm initialisers (weights are always loaded afterwards); same effect as the
    reference helper (`groma/utils.py` disable_torch_init)."""
    setattr(torch.nn.Linear, "reset_parameters", lambda self: None)
    setattr(torch.nn.LayerNorm, "reset_parameters", lambda self: None)
