"""BERT-large style encoder (24 x 1024 x 16 heads), the model of BASELINE config 5 (masked-LM head, random init)."""
from dataclasses import dataclass

import torch
from torch import nn
import torch.nn.functional as F


@dataclass
class BertConfig:
    vocab: int = 30522
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    max_seq: int = 512
    dropout: float = 0.0


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.h = c.heads
        self.qkv = nn.Linear(c.hidden, 3 * c.hidden)
        self.proj = nn.Linear(c.hidden, c.hidden)
        self.ln1 = nn.LayerNorm(c.hidden)
        self.fc1 = nn.Linear(c.hidden, c.ffn)
        self.fc2 = nn.Linear(c.ffn, c.hidden)
        self.ln2 = nn.LayerNorm(c.hidden)

    def forward(self, x):
        B, S, H = x.shape
        q, k, v = self.qkv(x).view(B, S, 3, self.h, H // self.h).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H)
        x = self.ln1(x + self.proj(a))
        return self.ln2(x + self.fc2(F.gelu(self.fc1(x))))


class BertEncoderModel(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.tok = nn.Embedding(c.vocab, c.hidden)
        self.pos = nn.Embedding(c.max_seq, c.hidden)
        self.ln = nn.LayerNorm(c.hidden)
        self.blocks = nn.ModuleList([Block(c) for _ in range(c.layers)])
        self.head = nn.Linear(c.hidden, c.vocab)

    def forward(self, ids):
        x = self.ln(self.tok(ids) + self.pos(torch.arange(ids.shape[1], device=ids.device))[None])
        for b in self.blocks:
            x = b(x)
        return self.head(x)


def bert_large():
    return BertEncoderModel(BertConfig())
