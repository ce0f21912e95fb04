"""Drop-in for the reference's `src/models/ctclip.py`: with `ct-clip-ut_b200/` on sys.path in place of
the reference's `src/`, `from models.ctclip import CTCLIP` resolves to the sm_100a-backed module."""
from ctclip_b200.modules import CTCLIP  # noqa: F401
