"""Model zoo used by the benchmark configurations of BASELINE.json (ResNet-50 data parallel, BERT-large with the
reduce-scatter + all-gather parameter path).  The reference ships no models (its tests use a synthetic two-layer
network); these are plain PyTorch definitions with random-init weights, plus `gpt`: Megatron-style
transformer blocks on the tensor + sequence parallel layers (token rows split between blocks, heads / hidden units inside)."""
from .bert import BertConfig, BertEncoderModel, bert_large
from .gpt import ParallelTransformer, ParallelTransformerBlock
from .mlp import MLP
from .resnet import resnet50

__all__ = ["BertConfig", "BertEncoderModel", "bert_large", "MLP", "ParallelTransformer", "ParallelTransformerBlock", "resnet50"]
