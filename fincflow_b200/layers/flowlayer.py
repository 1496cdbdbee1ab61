"""Operator API of the flow layers -- the drop-in boundary.

Mirrors the reference's fastflow/layers/flowlayer.py:7-56 (same class names, same abstract
methods, same defaults) so that containers written against the reference
(`module(input, context)` / `module.reverse(input, context)`,
fastflow/layers/flowsequential.py:29,99) work unchanged.
"""
from abc import ABCMeta, abstractmethod

import torch.nn as nn


class FlowLayer(nn.Module, metaclass=ABCMeta):
    """forward(input, context=None) -> (output, logdet); reverse(input, context=None) -> output."""

    @abstractmethod
    def forward(self, input, context=None):
        pass

    @abstractmethod
    def reverse(self, input, context=None):
        pass

    @abstractmethod
    def logdet(self, input, context=None):
        pass

    def reconstruct_forward(self, input, context=None):
        return self.forward(input)

    def reconstruct_reverse(self, input, context=None):
        return self.reverse(input)


class ModifiedGradFlowLayer(FlowLayer):
    @abstractmethod
    def forward(self, input, context=None, compute_expensive=False):
        pass

    @abstractmethod
    def reverse(self, input, context=None, compute_expensive=False):
        pass

    @abstractmethod
    def logdet(self, input, context=None, compute_expensive=False):
        pass


class PreprocessingFlowLayer(FlowLayer):
    pass


def mark_expensive(func):
    func._expensive_computation = True
    return func
