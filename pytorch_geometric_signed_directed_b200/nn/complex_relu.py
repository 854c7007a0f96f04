"""complex_relu_layer (reference nn/directed/complex_relu.py:7-34): mask = 1.0 * (real >= 0)
applied to both parts.  As a standalone module it is two elementwise torch ops; MagNetConv can
instead apply the same mask inside its dense-transform epilogue (`fused_complex_relu = True`)."""
import torch


class complex_relu_layer(torch.nn.Module):
    def complex_relu(self, real: torch.Tensor, img: torch.Tensor):
        mask = 1.0 * (real >= 0)
        return mask * real, mask * img

    def forward(self, real: torch.Tensor, img: torch.Tensor):
        return self.complex_relu(real, img)
