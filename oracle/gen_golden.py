#!/usr/bin/env python3
"""Generate ``tests/golden/*.npz`` by running the REFERENCE's own functions.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container, where
``/root/reference`` exists; the GPU box never executes this (the fixtures it writes are
committed).  The reference modules import plotting / launcher packages that are not
installed here (matplotlib, umap, plotly, accelerate); none of them is on the hot
path, so they are replaced by inert stubs before the import.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

What is pinned (all seeds fixed, all inputs synthetic):
  a1  calculate_color_embedding_correlation   token_embedding_analysis.py:183-260
  a3  average_embeddings_for_tokens           token_embedding_analysis_imagenet.py:261-286
  a4  MLLM.forward -> loss / logits           mllm.py:90-121 (tiny random-init Gemma3 + ViT)
  a5  evaluate_model                          multimodal_training.py:250-316
  a7  vision CrossEntropyLoss + torch.max     vision_training.py:81-83,116,132 (nn modules)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub_missing():
    # import the real third-party packages first so their own availability probes
    # (transformers looks for accelerate via importlib) never see a stub
    import sklearn.metrics.pairwise  # noqa: F401
    import transformers  # noqa: F401
    from transformers import Gemma3ForCausalLM, ViTModel  # noqa: F401
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.lines",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "umap", "plotly",
                 "plotly.graph_objects", "plotly.express", "accelerate", "wandb"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__getattr__ = lambda attr, _n=name: mock.MagicMock(name=f"{_n}.{attr}")  # type: ignore
            m.__path__ = []  # behave like a package
            sys.modules[name] = m


sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.tiny_models import ToyTokenizer, tiny_lm_config, tiny_vit_config  # noqa: E402


def color_fixture(seed=0, V=512, D=64):
    g = torch.Generator().manual_seed(seed)
    mapping_path = os.path.join(REF, "experiments/multimodal/color/12colors_4ood_labels_mapping.json")
    if not os.path.exists(mapping_path):
        cands = sorted(p for p in os.listdir(os.path.join(REF, "experiments/multimodal/color"))
                       if p.endswith("labels_mapping.json"))
        mapping_path = os.path.join(REF, "experiments/multimodal/color", cands[0])
    with open(mapping_path) as f:
        labels_mapping = json.load(f)
    ood = [v for v in labels_mapping.values() if v.startswith("<ood")]
    reg = [v for v in labels_mapping.values() if not v.startswith("<ood")]
    table0 = torch.randn(V, D, generator=g).to(torch.bfloat16)
    table1 = (torch.randn(V, D, generator=g) * 0.02 + torch.randn(1, D, generator=g)).to(torch.bfloat16)
    # the reference initialises OOD rows as copies of rows 0..n-1 (mllm.py:73)
    table0[-len(ood):] = table0[:len(ood)].clone()
    perm = torch.randperm(V - len(ood), generator=g)
    reg_ids = perm[:len(reg)].tolist()
    ood_ids = list(range(V - len(ood), V))
    emb = {"initial": table0, "epoch_0": table0.clone(), "epoch_3": table1}
    return labels_mapping, ood, reg, ood_ids, reg_ids, emb, os.path.basename(mapping_path)


def gen_a1():
    import src.multimodal.token_embedding_analysis as tea  # the reference module
    labels_mapping, ood, reg, ood_ids, reg_ids, emb, name = color_fixture()
    out = {}
    for tag, e in [("last", emb), ("initial_only", {"initial": emb["initial"]})]:
        with contextlib.redirect_stdout(io.StringIO()):
            r = tea.calculate_color_embedding_correlation(e, ood, reg, ood_ids, reg_ids, labels_mapping)
        out[f"r_{tag}"] = np.float64(r)
    with open(os.path.join(OUT, name), "w") as f:   # the label map is part of the fixture
        json.dump(labels_mapping, f, indent=1)
    np.savez(os.path.join(OUT, "a1_color_correlation.npz"),
             mapping_name=name, ood_ids=np.array(ood_ids), reg_ids=np.array(reg_ids),
             table_initial=emb["initial"].float().numpy(), table_epoch3=emb["epoch_3"].float().numpy(),
             **out)
    print("a1", out)


def gen_a3():
    import src.multimodal.token_embedding_analysis_imagenet as teai
    g = torch.Generator().manual_seed(1)
    V, D = 997, 48
    tok = ToyTokenizer(V)
    names = ["red", "dark olive green", "light-goldenrod yellow", "", "a b c d e f g h", "blue"]
    emb = {"initial": torch.randn(V, D, generator=g).to(torch.bfloat16),
           "epoch_0": torch.randn(V, D, generator=g)}  # one bf16, one fp32 table
    res = teai.average_embeddings_for_tokens(tok, emb, names)
    np.savez(os.path.join(OUT, "a3_average_embeddings.npz"), V=V, D=D, names=np.array(names),
             table_initial=emb["initial"].float().numpy(), table_epoch0=emb["epoch_0"].numpy(),
             out_initial=res["initial"].float().numpy(), out_epoch0=res["epoch_0"].numpy(),
             out_initial_is_bf16=np.bool_(res["initial"].dtype == torch.bfloat16))
    print("a3", {k: tuple(v.shape) for k, v in res.items()})


def tiny_mllm(seed=2):
    from transformers import Gemma3ForCausalLM, ViTModel
    from src.multimodal.mllm import MLLM
    torch.manual_seed(seed)
    m = MLLM.__new__(MLLM)                 # bypass from_pretrained (no weights offline)
    torch.nn.Module.__init__(m)
    m.vision_model_name = "tiny-vit"
    m.language_model_name = "tiny-gemma3"
    m.num_vision_tokens = 5
    m.vision_model = ViTModel(tiny_vit_config())
    m.language_model = Gemma3ForCausalLM(tiny_lm_config()).to(torch.bfloat16)
    m.projector = torch.nn.Linear(32, 64)
    m.tokenizer = ToyTokenizer(640)
    m.labels_mapping = None
    return m


def _state_arrays(module, prefix):
    """state_dict as npz-able arrays; bf16 tensors are stored as fp32 (exact) + a dtype marker."""
    out = {}
    for k, v in module.state_dict().items():
        out[f"{prefix}{k}"] = v.detach().float().numpy() if v.dtype == torch.bfloat16 else v.detach().numpy()
        out[f"{prefix}{k}::bf16"] = np.bool_(v.dtype == torch.bfloat16)
    return out


def gen_a4_a5():
    import src.multimodal.multimodal_training as mt   # needs the accelerate stub
    m = tiny_mllm().eval()
    V = m.language_model.config.vocab_size
    g = torch.Generator().manual_seed(3)
    B, T = 3, 12
    images = torch.randn(B, 3, 32, 32, generator=g)
    input_ids = torch.randint(0, V, (B, T), generator=g)
    attention_mask = torch.ones(B, T, dtype=torch.long)
    labels = torch.full((B, T), -100, dtype=torch.long)
    labels[0, 9:11] = torch.tensor([7, 1])      # answer-only supervision (imagenet_dataset.py:171-175)
    labels[1, 10] = 11
    # sample 2 keeps all labels ignored (exercises the `continue` at :283)
    hidden = {}

    def grab(mod, args, kwargs, out):
        hidden["h"] = out.last_hidden_state.detach()
    h = m.language_model.model.register_forward_hook(grab, with_kwargs=True)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        out = m(images=images, input_ids=input_ids, attention_mask=attention_mask, labels=labels)
    h.remove()
    table = m.language_model.get_input_embeddings().weight.detach()
    assert m.language_model.lm_head.weight.data_ptr() == table.data_ptr(), "lm_head must be tied"

    # a5: the reference's evaluate_model over a one-batch loader
    class Acc:  # the only Accelerator surface evaluate_model touches
        is_main_process = False
        def unwrap_model(self, model): return model
        def autocast(self): return torch.autocast("cpu", dtype=torch.bfloat16)
    cfg = types.SimpleNamespace(disable_tqdm=True)
    batch = {"images": images, "input_ids": input_ids, "attention_mask": attention_mask, "labels": labels}
    mt.tqdm = lambda it, **kw: it
    with contextlib.redirect_stdout(io.StringIO()):
        ev = mt.evaluate_model(m, [batch], cfg, Acc())
    # the whole tiny model + the batch, so that the GPU test can run the DROP-IN forward /
    # evaluate_model (shims.mllm.fused_forward bound as MLLM.forward) end to end
    np.savez(os.path.join(OUT, "a4_a5_mllm_model.npz"), images=images.numpy(), input_ids=input_ids.numpy(),
             attention_mask=attention_mask.numpy(), **_state_arrays(m, "sd::"))
    np.savez(os.path.join(OUT, "a4_a5_mllm_head.npz"),
             hidden=hidden["h"].float().numpy(), hidden_is_bf16=np.bool_(hidden["h"].dtype == torch.bfloat16),
             table=table.float().numpy(), labels=labels.numpy(),
             logits=out.logits.float().numpy(), logits_is_bf16=np.bool_(out.logits.dtype == torch.bfloat16),
             loss=np.float64(out.loss.item()), softcap=np.float64(m.language_model.config.final_logit_softcapping or 0.0),
             test_loss=np.float64(ev["test_loss"]), test_acc=np.float64(ev["test_acc"]),
             yes_id=7, no_id=11)
    print("a4", float(out.loss), tuple(out.logits.shape), out.logits.dtype, "a5", ev)


def gen_a7():
    g = torch.Generator().manual_seed(4)
    B, Dv, C = 16, 96, 100
    feats = torch.randn(B, Dv, generator=g)
    lin = torch.nn.Linear(Dv, C)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(C, Dv, generator=g) * 0.1)
        lin.bias.zero_()
        lin.weight[C - 1] = lin.weight[0]          # an exact tie: first index must win
    labels = torch.randint(0, C, (B,), generator=g)
    out = {}
    for eps in (0.0, 0.1):
        crit = torch.nn.CrossEntropyLoss(label_smoothing=eps) if eps > 0 else torch.nn.CrossEntropyLoss()
        logits = lin(feats)
        out[f"loss_{eps}"] = np.float64(crit(logits, labels).item())
    _, predicted = torch.max(lin(feats).data, 1)
    np.savez(os.path.join(OUT, "a7_vision_head.npz"), feats=feats.numpy(), weight=lin.weight.detach().numpy(),
             labels=labels.numpy(), predicted=predicted.numpy(), **out)
    print("a7", out)


def gen_a7_model():
    """vision_training.py:81-83,115-116,132 on a tiny ViTForImageClassification (the class the
    reference trains from scratch, vision_training.py:259-275): criterion(outputs.logits, labels)
    and torch.max(outputs.logits.data, 1), under the fp32 the CPU run uses."""
    from transformers import ViTForImageClassification
    torch.manual_seed(5)
    C = 37
    model = ViTForImageClassification(tiny_vit_config(num_labels=C)).eval()
    with torch.no_grad():                  # a head that discriminates (HF initialises it near zero)
        model.classifier.weight.mul_(40.0)
        model.classifier.bias.normal_(generator=torch.Generator().manual_seed(8))
    g = torch.Generator().manual_seed(6)
    images = torch.randn(9, 3, 32, 32, generator=g)
    labels = torch.randint(0, C, (9,), generator=g)
    out = {}
    with torch.no_grad():
        outputs = model(images)
        for eps in (0.0, 0.1):
            criterion = torch.nn.CrossEntropyLoss(label_smoothing=eps) if eps > 0 else torch.nn.CrossEntropyLoss()
            out[f"loss_{eps}"] = np.float64(criterion(outputs.logits, labels).item())
        _, predicted = torch.max(outputs.logits.data, 1)
    np.savez(os.path.join(OUT, "a7_vit_model.npz"), images=images.numpy(), labels=labels.numpy(),
             logits=outputs.logits.numpy(), predicted=predicted.numpy(), num_labels=C,
             **out, **_state_arrays(model, "sd::"))
    print("a7 model", out, predicted.tolist())


def main():
    assert os.path.isdir(REF), "reference tree not present: fixtures can only be regenerated in the build container"
    os.makedirs(OUT, exist_ok=True)
    _stub_missing()
    sys.path.insert(0, REF)
    gen_a1(); gen_a3(); gen_a4_a5(); gen_a7(); gen_a7_model()


if __name__ == "__main__":
    main()
