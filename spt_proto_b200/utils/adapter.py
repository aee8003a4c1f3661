"""Swap dense modules of a model for their LoRA / sparse counterparts (reference naive_gpt/utils/adapter.py).

`ModuleUpgrader(handler).visit(model)` walks `model.named_modules()`; for a module of class `Cls` it calls
`handler.onCls(name=..., child=...)` if the handler has such a method, else `handler.default(...)`; a returned
module replaces the visited one in its parent (adapter.py:187-223).  The reference's fine-tuning recipe runs
four passes (script/4-sparse-tuning-0.py:33-39, script/0-profile.py:182-189):

    lora    nn.Linear / nn.Embedding              -> LoRALinear / LoRAEmbedding (frozen base + rank-d_lora factors)
    ffn     Feedforward / LLaMaFeedforward        -> LoRARoutedFFN / LoRARoutedLLaMaFFN, block = d_feedforward // 4
    mha_v1  VanillaAttention / RotaryAttention    -> Sparse*AttentionV1 (dense attention + PQ loss; PQ 8 x 16)
    mha_v2  Sparse*AttentionV1                    -> Sparse*AttentionV2 (the sparse hot path), codebook carried over

The hyper-parameters the reference hard-codes in its handler (d_codeword 8, 16 codewords, d_head // 8
subspaces, block = d_ff // 4; adapter.py:94-97,163) are constructor arguments here with those defaults."""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

from torch import nn

from .. import layers


class LoRAHandler:
    """Upgrades every nn.Linear / nn.Embedding to its LoRA form (adapter.py:6-42)."""

    def __init__(self, d_lora: int, verbose: bool = True):
        self.d_lora, self.verbose = d_lora, verbose

    def _log(self, tag: str, name: str, child: nn.Module, new: Optional[nn.Module] = None) -> None:
        if self.verbose:
            tail = () if new is None else ("->", type(new).__name__)
            print(f"[{tag}]", name, type(child).__name__, *tail)

    def default(self, name: str, child: nn.Module):
        self._log("SKIP", name, child)

    def _lora(self, name: str, child: nn.Module, target) -> nn.Module:
        new = target.from_pretrained(d_lora=self.d_lora, source=child)
        self._log("UPGRADE", name, child, new)
        return new

    def onLinear(self, name: str, child: nn.Linear):
        return self._lora(name, child, layers.LoRALinear)

    def onEmbedding(self, name: str, child: nn.Embedding):
        return self._lora(name, child, layers.LoRAEmbedding)


class SparseLoRAHandler(LoRAHandler):
    """One of the four passes of the sparse fine-tuning recipe, selected by `stage` (adapter.py:45-184)."""

    STAGES = ("lora", "ffn", "mha_v1", "mha_v2")

    def __init__(self, d_lora: int, stage: str, d_codeword: int = 8, n_codewords: int = 16, ffn_blocks: int = 4,
                 verbose: bool = True):
        super().__init__(d_lora=d_lora, verbose=verbose)
        if stage not in self.STAGES:
            raise AssertionError(f"stage must be one of {self.STAGES}")
        self.stage, self.d_codeword, self.n_codewords, self.ffn_blocks = stage, d_codeword, n_codewords, ffn_blocks
        # class name -> (stage in which it is rewritten, builder)
        self._rules: Dict[str, Tuple[str, Callable[[nn.Module], nn.Module]]] = {
            "Linear": ("lora", lambda m: layers.LoRALinear.from_pretrained(d_lora=d_lora, source=m)),
            "Embedding": ("lora", lambda m: layers.LoRAEmbedding.from_pretrained(d_lora=d_lora, source=m)),
            "Feedforward": ("ffn", lambda m: layers.LoRARoutedFFN.from_pretrained(
                d_lora=d_lora, block_size=m.d_feedforward // ffn_blocks, source=m)),
            "LLaMaFeedforward": ("ffn", lambda m: layers.LoRARoutedLLaMaFFN.from_pretrained(
                d_lora=d_lora, block_size=m.d_feedforward // ffn_blocks, source=m)),
            "VanillaAttention": ("mha_v1", lambda m: self._v1(layers.SparseVanillaAttentionV1, m)),
            "RotaryAttention": ("mha_v1", lambda m: self._v1(layers.SparseRotaryAttentionV1, m)),
            "SparseVanillaAttentionV1": ("mha_v2", lambda m: layers.SparseVanillaAttentionV2.from_pretrained(source=m)),
            "SparseRotaryAttentionV1": ("mha_v2", lambda m: layers.SparseRotaryAttentionV2.from_pretrained(source=m)),
        }

    def _v1(self, target, child):
        return target(d_head=child.d_head, p_dropout=child.p_dropout, d_codeword=self.d_codeword,
                      n_codewords=self.n_codewords, n_subspaces=child.d_head // self.d_codeword)

    def _apply(self, name: str, child: nn.Module):
        stage, build = self._rules[type(child).__name__]
        if stage != self.stage:
            self._log("SKIP", name, child)
            return None
        new = build(child)
        self._log("UPGRADE", name, child, new)
        return new

    # the dispatch convention of ModuleUpgrader: one on<ClassName> method per rewritable class
    def onLinear(self, name, child): return self._apply(name, child)
    def onEmbedding(self, name, child): return self._apply(name, child)
    def onFeedforward(self, name, child): return self._apply(name, child)
    def onLLaMaFeedforward(self, name, child): return self._apply(name, child)
    def onVanillaAttention(self, name, child): return self._apply(name, child)
    def onRotaryAttention(self, name, child): return self._apply(name, child)
    def onSparseVanillaAttentionV1(self, name, child): return self._apply(name, child)
    def onSparseRotaryAttentionV1(self, name, child): return self._apply(name, child)


class ModuleUpgrader:
    """Visitor over a module tree (adapter.py:187-223).  Replacements are collected first and applied after the
    walk, so a freshly inserted module is never visited in the same pass."""

    def __init__(self, handler: object):
        if not hasattr(handler, "default"):
            raise RuntimeError("requires default handler")
        self.handler = handler

    def visit(self, root: nn.Module) -> nn.Module:
        pending = {}
        for path, module in root.named_modules():
            fn = getattr(self.handler, "on" + type(module).__name__, self.handler.default)
            new = fn(name=path, child=module)
            if new is not None and new is not module:
                assert isinstance(new, nn.Module)
                pending[path] = new
        for path, new in pending.items():
            parent_path, _, leaf = path.rpartition(".")
            parent = root.get_submodule(parent_path) if parent_path else root
            parent.add_module(leaf, new)
        return root
