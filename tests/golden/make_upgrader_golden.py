"""Generates tests/golden/upgrader_state.pt: the state_dict layout (key -> shape, dtype) of a small LLaMA-style and a small OPT-style
TransformerBlock after the UNMODIFIED reference's four-pass sparse upgrade (naive_gpt/utils/adapter.py via
ModuleUpgrader + SparseLoRAHandler, stages lora -> ffn -> mha_v1 -> mha_v2, as in script/0-profile.py:182-189).
Run in the authoring container only (needs /root/reference).

    python tests/golden/make_upgrader_golden.py
"""
import contextlib
import io
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.modules["naive_gpt.ext"] = types.ModuleType("naive_gpt.ext")          # never called on this path
sys.modules["naive_gpt.loaders"] = types.ModuleType("naive_gpt.loaders")  # needs lightning / torchtext
from naive_gpt import layers, utils  # noqa: E402


def build(kind: str):
    torch.manual_seed(7)
    if kind == "llama":
        return layers.TransformerBlock(
            d_model=128, n_heads=2, layernorm_fn=layers.LlamaRMSNorm(128),
            attention_fn=layers.RotaryAttention(d_head=64, p_dropout=0.0),
            feedforward_fn=layers.LLaMaFeedforward(d_model=128, d_feedforward=512, activation=nn.SiLU()),
            attention_bias=False, pre_norm=True)
    return layers.TransformerBlock(
        d_model=128, n_heads=2, layernorm_fn=nn.LayerNorm(128),
        attention_fn=layers.VanillaAttention(d_head=64, p_dropout=0.0),
        feedforward_fn=layers.Feedforward(d_model=128, d_feedforward=512, p_dropout=0.0, activation=nn.ReLU()),
        attention_bias=True, pre_norm=True)


out = {}
for kind in ("llama", "opt"):
    model = build(kind)
    with contextlib.redirect_stdout(io.StringIO()):
        for stage in ("lora", "ffn", "mha_v1", "mha_v2"):
            model = utils.ModuleUpgrader(handler=utils.SparseLoRAHandler(d_lora=4, stage=stage)).visit(model)
    out[kind] = {"state_dict": {k: (tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()},
                 "trainable": sorted(n for n, p in model.named_parameters() if p.requires_grad),
                 "classes": {n: type(m).__name__ for n, m in model.named_modules()}}
    print(kind, len(out[kind]["state_dict"]), "tensors,", len(out[kind]["trainable"]), "trainable")
torch.save(out, os.path.join(HERE, "upgrader_state.pt"))
