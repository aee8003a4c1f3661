"""Module-tree rewriting (reference naive_gpt/utils/__init__.py:1-4)."""
from .adapter import LoRAHandler, ModuleUpgrader, SparseLoRAHandler

__all__ = ["LoRAHandler", "ModuleUpgrader", "SparseLoRAHandler"]
