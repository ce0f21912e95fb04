"""Fixture: the state-dict keys and shapes of the UNMODIFIED reference CTCLIP(CTViT) at the benchmark configuration
(src/inference_ctclip.py:21-39), i.e. what `ctclip_v2.pt` must contain.  Built on the meta device (no
memory), text tower excluded.      python tests/golden/make_golden_keys.py [--ref /root/reference]
Output: tests/golden/state_dict_keys.json {key: [shape...]}"""
import argparse
import json
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import FakeText, import_reference  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    _, ref_ctvit, ref_ctclip, _ = import_reference(args.ref)
    with torch.device("meta"):
        vit = ref_ctvit.CTViT(dim=512, codebook_size=8192, image_size=480, patch_size=20, temporal_patch_size=10,
                              spatial_depth=4, temporal_depth=4, dim_head=32, heads=8)
        clip = ref_ctclip.CTCLIP(text_encoder=FakeText(), image_encoder=vit, dim_text=768, dim_image=294912,
                                 dim_latent=512)
    keys = {k: list(v.shape) for k, v in clip.state_dict().items() if not k.startswith("text_transformer.")}
    (HERE / "state_dict_keys.json").write_text(json.dumps(keys, indent=0, sort_keys=True))
    print(len(keys), "keys")


if __name__ == "__main__":
    main()
