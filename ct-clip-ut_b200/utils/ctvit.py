"""Drop-in for the reference's `src/utils/ctvit.py` (`from utils.ctvit import CTViT`)."""
from ctclip_b200.modules import CTViT  # noqa: F401
