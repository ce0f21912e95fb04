"""Golden fixture for the zero-shot scoring loop: runs the UNMODIFIED reference
`CTClipInference.zeroshot` (/root/reference/src/utils/CTClipInference.py:147-201) on CPU.

    python tests/golden/make_golden_zeroshot.py [--ref /root/reference]

The class cannot be constructed here (Accelerate, the dataset classes and the metric / plotting helpers are
absent), so the method is executed on a bare instance whose collaborators are stand-ins that carry no
arithmetic: a data loader of 3 dummy samples, a tokenizer that passes the two prompt strings through, an
accelerator whose gather is the identity, and a `model` that returns seeded unit latents for (sample, prompt)
as a 6-tuple — the committed loop unpacks six values from CTCLIP.forward's five (:169), so it cannot run
against the real CTCLIP; everything after that line (validate_prompts, diag, softmax, float64 store, stack,
gather) is the reference's own code.  The predictions are captured at the `calculate_metrics` call (:194).
Output: zero_shot.npz {image_latents [3,16], pair_latents [18,2,16], temp, predictions [3,18] f64}.
"""
import argparse
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import _Anything, _stub_module  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.path.insert(0, str(Path(args.ref) / "src"))
    captured = {}

    def calculate_metrics(pred, targets, names):
        captured["pred"], captured["targets"] = np.array(pred), np.array(targets)
        return {}

    import transformers  # noqa: F401  (before the accelerate stub: it probes for the real package)
    from transformers import BertTokenizer  # noqa: F401
    _stub_module("accelerate", Accelerator=_Anything("Accelerator"))
    _stub_module("accelerate.utils", InitProcessGroupKwargs=_Anything(), DistributedDataParallelKwargs=_Anything())
    _stub_module("utils.InferenceDataset", InferenceDataset=_Anything())
    m = _stub_module("utils.metrics", calculate_metrics=calculate_metrics, save_metrics=lambda *a, **k: None,
                     plot_precision_recall_curve=lambda *a, **k: None, plot_roc_curve=lambda *a, **k: None,
                     plot_per_class_f1=lambda *a, **k: None)
    m.__all__ = ["calculate_metrics", "save_metrics", "plot_precision_recall_curve", "plot_roc_curve",
                 "plot_per_class_f1"]
    _stub_module("utils.visualizations", Visualizations=_Anything())
    _stub_module("models.ctclip", CTCLIP=_Anything())
    _stub_module("scipy.ndimage", zoom=_Anything())
    import utils.CTClipInference as R

    P, d, n = len(R.PATHOLOGIES), 16, 3
    g = torch.Generator().manual_seed(11)
    unit = lambda t: t / t.norm(dim=-1, keepdim=True)
    il = unit(torch.randn(n, d, generator=g))
    # a shared direction keeps the present/absent gap small enough that the probabilities are not saturated
    pair = unit(torch.randn(P, 1, d, generator=g) + 0.35 * torch.randn(P, 2, d, generator=g))
    temp = torch.tensor(1.0).exp() * 4.0

    class Tok:
        def __call__(self, prompts, **kw):
            self_ = types.SimpleNamespace(prompts=prompts)
            self_.to = lambda dev: self_
            return self_

    state = {"sample": -1}

    class Loader:
        dataset = list(range(n))

        def __iter__(self):
            for i in range(n):
                state["sample"] = i
                labels = torch.zeros(1, P)
                yield (torch.zeros(1, 1, 2, 2, 2), "report", labels, "name", "path")

    def model(text_tokens, images):
        name = text_tokens.prompts[0][len("There is "):-1]
        assert text_tokens.prompts[1] == f"There is no {name}."
        j = R.PATHOLOGIES.index(name)
        i = state["sample"]
        return None, il[i:i + 1], pair[j], temp, None, None

    inst = object.__new__(R.CTClipInference)
    torch.nn.Module.__init__(inst)
    inst.dl, inst.tokenizer, inst.model = Loader(), Tok(), model
    inst.accelerator = types.SimpleNamespace(device=torch.device("cpu"), process_index=0, is_main_process=True,
                                             gather_for_metrics=lambda x: x)
    inst.metrics, inst.results_folder = [], Path("/tmp")
    inst.zeroshot()
    pred = captured["pred"]
    assert pred.shape == (n, P) and pred.dtype == np.float64
    assert 0.02 < pred.min() and pred.max() < 0.98
    np.savez(HERE / "zero_shot.npz", image_latents=il.numpy(), pair_latents=pair.numpy(), temp=temp.numpy(),
             predictions=pred)
    print("zero_shot.npz: predictions", pred.shape, "range", float(pred.min()), float(pred.max()))


if __name__ == "__main__":
    main()
