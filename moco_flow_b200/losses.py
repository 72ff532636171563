"""Image loss of the reference (models/losses.py:5-14): mse(rgb_coarse) + mse(rgb_fine)."""
import torch
from torch import nn


class MSELoss(nn.Module):
    def forward(self, inputs, targets):
        total = nn.functional.mse_loss(inputs['rgb_coarse'], targets)
        if 'rgb_fine' in inputs:
            total = total + nn.functional.mse_loss(inputs['rgb_fine'], targets)
        return total
