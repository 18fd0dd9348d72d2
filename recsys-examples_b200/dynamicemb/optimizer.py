"""Sparse optimizer descriptors.  State layout per row follows the reference
(/root/reference/corelib/dynamicemb/dynamicemb/optimizer.py:36-57): value row = [emb(D) | state],
Adagrad state D, Adam 2D (m then v), row-wise Adagrad 16 bytes (4 fp32, accumulator first), SGD none."""
from dataclasses import dataclass

import torch

from .dynamicemb_extensions import OptimizerType
from .types import DTYPE_NUM_BYTES, EmbOptimType


def get_optimizer_state_dim(optimizer_type: EmbOptimType, dim: int, dtype: torch.dtype = torch.float32) -> int:
    if optimizer_type == EmbOptimType.EXACT_ROWWISE_ADAGRAD:
        return 16 // DTYPE_NUM_BYTES[dtype]
    if optimizer_type == EmbOptimType.ADAM:
        return dim * 2
    if optimizer_type == EmbOptimType.EXACT_ADAGRAD:
        return dim
    return 0


_KERNEL_TYPE = {
    EmbOptimType.SGD: OptimizerType.SGD, EmbOptimType.EXACT_SGD: OptimizerType.SGD, EmbOptimType.ADAM: OptimizerType.ADAM,
    EmbOptimType.EXACT_ADAGRAD: OptimizerType.ADAGRAD, EmbOptimType.EXACT_ROWWISE_ADAGRAD: OptimizerType.ROWWISE_ADAGRAD,
    EmbOptimType.NONE: OptimizerType.NONE,
}


@dataclass
class OptimizerArgs:
    learning_rate: float = 0.01
    eps: float = 1.0e-8
    initial_accumulator_value: float = 0.0
    beta1: float = 0.9
    beta2: float = 0.999
    weight_decay: float = 0.0
    gradient_clipping: bool = False
    max_gradient: float = 1.0


class SparseOptimizer:
    """Holds hyper-parameters + step count and turns them into kernel arguments."""

    def __init__(self, optimizer_type: EmbOptimType, args: OptimizerArgs):
        self.optimizer_type = optimizer_type
        self.kernel_type = int(_KERNEL_TYPE[optimizer_type])
        self.args = args
        self.iter = 0

    def get_state_dim(self, dim: int, dtype: torch.dtype = torch.float32) -> int:
        return get_optimizer_state_dim(self.optimizer_type, dim, dtype)

    @property
    def initial_state_value(self) -> float:
        if self.optimizer_type in (EmbOptimType.EXACT_ADAGRAD, EmbOptimType.EXACT_ROWWISE_ADAGRAD):
            return self.args.initial_accumulator_value
        return 0.0

    def step(self) -> None:
        self.iter += 1

    def kernel_kwargs(self) -> dict:
        a = self.args
        it = max(self.iter, 1)
        return dict(opt_type=self.kernel_type, lr=a.learning_rate, eps=a.eps, beta1=a.beta1, beta2=a.beta2, weight_decay=a.weight_decay,
                    bc1=1.0 - a.beta1 ** it, bc2=1.0 - a.beta2 ** it)

    def set_learning_rate(self, lr: float) -> None:
        self.args.learning_rate = lr
