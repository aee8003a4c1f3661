"""Layer constructors with the reference's names (naive_gpt/layers/__init__.py:1-36)."""
from .basic import (Feedforward, LLaMaFeedforward, LlamaRMSNorm, MultiheadAttention, PQV1, PQV2, RotaryAttention,
                    RotaryEmbedding, TransformerBlock, VanillaAttention)
from .lora import LoRAEmbedding, LoRALinear, LoRARoutedFFN, LoRARoutedLLaMaFFN
from .routed_ffn import RoutedFFN, RoutedLLaMaFFN
from .sparse_attention import (SparseRotaryAttentionV1, SparseRotaryAttentionV2,
                               SparseVanillaAttentionV1, SparseVanillaAttentionV2)

__all__ = [
    "Feedforward", "LLaMaFeedforward", "LlamaRMSNorm", "MultiheadAttention", "TransformerBlock", "PQV1", "PQV2", "RotaryAttention", "RotaryEmbedding",
    "VanillaAttention", "LoRAEmbedding", "LoRALinear", "LoRARoutedFFN", "LoRARoutedLLaMaFFN", "RoutedFFN", "RoutedLLaMaFFN", "SparseRotaryAttentionV1", "SparseRotaryAttentionV2",
    "SparseVanillaAttentionV1", "SparseVanillaAttentionV2",
]
