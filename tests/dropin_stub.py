"""A tiny deterministic stand-in for the reference's `cle_vit_backbone.CLEViTDualStream` (timm is absent),
shared by tests/golden/make_golden.py (all-reference model, CPU) and the GPU drop-in acceptance test.
Contract: ego_moment_clevit.py:76-82,148-153 / cle_vit_backbone.py:232-236,252-316."""
import torch
import torch.nn as nn


class StubDualStream(nn.Module):
    PATCH, DIM = 4, 32

    def __init__(self, model_name: str = "stub", pretrained: bool = False, drop_rate: float = 0.0):
        super().__init__()
        self.num_features = self.DIM
        self.proj = nn.Linear(3 * self.PATCH * self.PATCH, self.DIM)
        self.cls_token = nn.Parameter(torch.randn(1, 1, self.DIM) * 0.5)
        self.mix = nn.Linear(self.DIM, self.DIM)

    def _one(self, x):
        B, C, H, W = x.shape
        P = self.PATCH
        x = x.unfold(2, P, P).unfold(3, P, P)                       # [B, C, H/P, W/P, P, P]
        x = x.permute(0, 2, 3, 1, 4, 5).reshape(B, (H // P) * (W // P), C * P * P)
        t = torch.cat([self.cls_token.expand(B, -1, -1), self.proj(x)], dim=1)
        t = t + torch.tanh(self.mix(t.mean(dim=1, keepdim=True)))   # tokens see each other a little
        return {"patch_tokens": t[:, 1:].contiguous(), "global_features": t[:, 0]}

    def forward(self, anchor, positive):
        return self._one(anchor), self._one(positive)
