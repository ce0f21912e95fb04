"""Drop-in for the reference's `src/utils/InferenceDataset.py` (`from utils.InferenceDataset import InferenceDataset`)."""
from ctclip_b200.dataset import DeviceLoader, InferenceDataset  # noqa: F401
