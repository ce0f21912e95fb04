"""Drop-in for the reference's `src/utils/preprocess.py` (`from utils.preprocess import process_file`)."""
from ctclip_b200.preprocess import process_file, process_volume, read_nii_data, read_nii_raw  # noqa: F401
