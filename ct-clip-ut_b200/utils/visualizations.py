"""Drop-in for the reference's `src/utils/visualizations.py` (`from utils.visualizations import Visualizations`)."""
from ctclip_b200.attribution import PATHOLOGIES, Visualizations  # noqa: F401
