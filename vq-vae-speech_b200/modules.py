"""Drop-in building blocks: Conv1DBuilder, ConvTranspose1DBuilder, Residual, ResidualStack, Jitter.

Same constructors / forward signatures / state_dict keys as the reference's src/modules/ (conv1d_builder.py:28-44,
conv_transpose1d_builder.py:28-44, residual.py:31-70, residual_stack.py:34-46, jitter.py:31-70); forward and backward
run the hand-written implicit-GEMM kernels (functional.py) instead of ATen/cuDNN.
"""
import numpy as np
import torch
import torch.nn as nn

from . import functional as F


class Conv1d(nn.Conv1d):
    """nn.Conv1d (same parameters, same default init) whose forward/backward are libvqs_b200 kernels."""

    def forward(self, x, relu=False):
        if self.dilation[0] != 1 or self.groups != 1 or self.padding_mode != 'zeros':
            raise NotImplementedError('only the dense, undilated convolutions of the reference are implemented')
        return F.conv1d(x, self.weight, self.bias, self.stride[0], self.padding[0], relu=relu)


class ConvTranspose1d(nn.ConvTranspose1d):
    """nn.ConvTranspose1d, stride 1 (the only stride the reference decoder uses, deconvolutional_decoder.py:76-98)."""

    def forward(self, x, relu=False, out_len=None):
        if self.stride[0] != 1 or self.dilation[0] != 1 or self.groups != 1 or self.output_padding[0] != 0:
            raise NotImplementedError('only stride-1 transposed convolutions are implemented')
        return F.conv_transpose1d(x, self.weight, self.bias, self.padding[0], relu=relu, out_len=out_len)


class _WeightNormHook(object):
    """nn.utils.weight_norm (dim = 0) with the reparametrisation on this library's kernels: same parameters in the same
    order (`weight` is replaced by `weight_g` = ||weight|| per slice of dim 0, then `weight_v` = weight, both registered after
    the bias, torch/nn/utils/weight_norm.py), hence the same state_dict keys and optimizer order as the reference; before
    every forward `weight` is recomputed as g v / ||v|| by vqs_weight_norm_fwd (autograd: vqs_weight_norm_bwd)."""

    def __call__(self, module, inputs):
        module.weight = F.weight_norm(module.weight_v, module.weight_g)


def weight_norm(module):
    weight = module.weight
    del module._parameters['weight']
    module.register_parameter('weight_g', nn.Parameter(torch.norm_except_dim(weight, 2, 0).data))    # (init-time only)
    module.register_parameter('weight_v', nn.Parameter(weight.data))
    # a plain tensor until the first forward: the reference's builders run kaiming_normal_ on this recomputed attribute (it
    # leaves g and v alone but consumes the host RNG, which the parameters created after it depend on)
    module.weight = weight.data.clone()
    module.register_forward_pre_hook(_WeightNormHook())
    return module


class Conv1DBuilder(object):

    @staticmethod
    def build(in_channels, out_channels, kernel_size, stride=1, padding=0, use_kaiming_normal=False):
        conv = Conv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                      padding=padding)
        if use_kaiming_normal:   # weight-norm reparametrisation on this library's kernels, see weight_norm below
            conv = weight_norm(conv)
            nn.init.kaiming_normal_(conv.weight)
        return conv


class ConvTranspose1DBuilder(object):

    @staticmethod
    def build(in_channels, out_channels, kernel_size, stride=1, padding=0, use_kaiming_normal=False):
        conv = ConvTranspose1d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                               stride=stride, padding=padding)
        if use_kaiming_normal:
            conv = weight_norm(conv)
            nn.init.kaiming_normal_(conv.weight)
        return conv


class Residual(nn.Module):
    """x -> relu(x) + conv_2(relu(conv_1(relu(x)))).

    The reference block opens with nn.ReLU(inplace=True) (residual.py:36) so its skip term is relu(x), not x; that
    arithmetic is reproduced.  The caller's tensor is NOT mutated here (in the reference models the mutated tensor is
    either already non-negative -- encoder -- or never read again -- decoder -- so results are identical)."""

    def __init__(self, in_channels, num_hiddens, num_residual_hiddens, use_kaiming_normal):
        super(Residual, self).__init__()
        conv_1 = Conv1d(in_channels=in_channels, out_channels=num_residual_hiddens, kernel_size=3, stride=1,
                        padding=1, bias=False)
        if use_kaiming_normal:
            conv_1 = weight_norm(conv_1)
            nn.init.kaiming_normal_(conv_1.weight)
        conv_2 = Conv1d(in_channels=num_residual_hiddens, out_channels=num_hiddens, kernel_size=1, stride=1,
                        bias=False)
        if use_kaiming_normal:
            conv_2 = weight_norm(conv_2)
            nn.init.kaiming_normal_(conv_2.weight)
        # same container and indices as the reference so that state_dict keys are `_block.1.weight`, `_block.3.weight`
        self._block = nn.Sequential(nn.ReLU(True), conv_1, nn.ReLU(True), conv_2)

    def forward(self, x):
        a = F.relu(x)
        h = self._block[1](a, relu=True)
        return F.add(a, self._block[3](h))


class ResidualStack(nn.Module):
    """The SAME Residual instance applied num_residual_layers times, then ReLU (residual_stack.py:40-46)."""

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens, use_kaiming_normal):
        super(ResidualStack, self).__init__()
        self._num_residual_layers = num_residual_layers
        self._layers = nn.ModuleList(
            [Residual(in_channels, num_hiddens, num_residual_hiddens, use_kaiming_normal)] * self._num_residual_layers)

    def forward(self, x):
        for i in range(self._num_residual_layers):
            x = self._layers[i](x)
        return F.relu(x)


def jitter_plan(length, probability, rng=np.random):
    """src[t] = the column of the original tensor that ends up at position t.  Consumes `np.random` in exactly the
    reference's order (jitter.py:55-67): one choice([1, 0], p) per t and, only for a replaced interior t, one
    choice([-1, 1]) -- so the same seed gives the same plan as the reference."""
    src = np.arange(length, dtype=np.int32)
    for i in range(length):
        replace = [True, False][rng.choice([1, 0], p=[probability, 1 - probability])]
        if replace:
            if i == 0:
                neighbor_index = i + 1
            elif i == length - 1:
                neighbor_index = i - 1
            else:
                neighbor_index = i + rng.choice([-1, 1], p=[0.5, 0.5])
            src[i] = neighbor_index
    return src


class Jitter(nn.Module):
    """[Chorowski et al., 2019] time jitter: each latent column is replaced by a neighbour with `probability`.
    The plan is drawn on the host (same RNG stream as the reference), the copy is one gather kernel; replaced columns
    pass no gradient (the reference overwrites them in place from a detached clone, jitter.py:49,68)."""

    def __init__(self, probability=0.12):
        super(Jitter, self).__init__()
        self._probability = probability

    def forward(self, quantized):
        src = jitter_plan(quantized.size(2), self._probability)
        self.last_plan = src
        return F.jitter(quantized, torch.from_numpy(src).to(quantized.device, non_blocking=True))
