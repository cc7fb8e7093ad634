"""Architecture table of the benchmark models (pure Python: importing it does NOT load libqserve_b200.so, so the CPU
baseline arm of bench.py can use it without touching the product library).  Dimensions: SURVEY.md section 8."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ModelConfig:
    name: str
    hidden: int
    intermediate: int
    heads: int
    kv_heads: int
    layers: int
    vocab: int
    rope_theta: float
    eps: float = 1e-5
    head_dim: int = 128
    max_pos: int = 8192


MODELS = {
    "llama-3-8b": ModelConfig("Llama-3-8B", 4096, 14336, 32, 8, 32, 128256, 500000.0),
    "mistral-7b": ModelConfig("Mistral-7B", 4096, 14336, 32, 8, 32, 32000, 10000.0, max_pos=32768),
    "llama-2-7b": ModelConfig("Llama-2-7B", 4096, 11008, 32, 32, 32, 32000, 10000.0, max_pos=4096),
    "qwen1.5-72b": ModelConfig("Qwen1.5-72B", 8192, 24576, 64, 64, 80, 152064, 1000000.0, eps=1e-6, max_pos=32768),
    "tiny": ModelConfig("tiny-test", 512, 1024, 4, 2, 2, 1024, 10000.0),
}

PRECISIONS = ("w4a8kv4", "w4a8kv4-g128", "w8a8kv8", "w4a8kv8", "w8a8kv4")
