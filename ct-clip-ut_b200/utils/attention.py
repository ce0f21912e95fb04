"""Drop-in for the reference's `src/utils/attention.py`: the same class names, as parameter containers
with identical state-dict keys (the fused CUDA path never executes them module by module)."""
from ctclip_b200.modules import (Attention, ContinuousPositionBias, FeedForward, LayerNorm, PEG,  # noqa: F401
                                 Transformer)
