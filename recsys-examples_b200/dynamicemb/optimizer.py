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

    # ------------------------------------------------------------------ checkpoint metadata (reference optimizer.py:155-170, :248-260,
    # :330-350, :410-430, :487-512): the `<table>_opt_args.json` a dump writes and a load validates / restores
    _CKPT_NAME = {EmbOptimType.SGD: "sgd", EmbOptimType.EXACT_SGD: "sgd", EmbOptimType.ADAM: "adam",
                  EmbOptimType.EXACT_ADAGRAD: "exact_adagrad", EmbOptimType.EXACT_ROWWISE_ADAGRAD: "exact_row_wise_adagrad"}

    def get_opt_args(self) -> dict:
        a, t = self.args, self.optimizer_type
        if t == EmbOptimType.NONE:
            return {}
        out = {"opt_type": self._CKPT_NAME[t], "lr": a.learning_rate}
        if t == EmbOptimType.ADAM:
            out.update(iters=self.iter, beta1=a.beta1, beta2=a.beta2, eps=a.eps, weight_decay=a.weight_decay)
        elif t in (EmbOptimType.EXACT_ADAGRAD, EmbOptimType.EXACT_ROWWISE_ADAGRAD):
            out.update(eps=a.eps, initial_accumulator_value=a.initial_accumulator_value)
        return out

    def set_opt_args(self, args: dict) -> None:
        def need(key):
            if key not in args:
                raise ValueError(f"Input args does not contain required optimizer argument: {key}")
            return args[key]
        a, t = self.args, self.optimizer_type
        if t == EmbOptimType.NONE:
            return
        a.learning_rate = need("lr")
        if t == EmbOptimType.ADAM:
            self.iter = int(need("iters"))
            a.beta1, a.beta2, a.eps, a.weight_decay = need("beta1"), need("beta2"), need("eps"), need("weight_decay")
        elif t in (EmbOptimType.EXACT_ADAGRAD, EmbOptimType.EXACT_ROWWISE_ADAGRAD):
            a.eps, a.initial_accumulator_value = need("eps"), need("initial_accumulator_value")

    def get_ckpt_state_dim(self, dim: int) -> int:
        """State elements per row in a checkpoint file: row-wise Adagrad keeps 16 bytes at run time but only its accumulator is stored
        (reference optimizer.py:60-73)."""
        if self.optimizer_type == EmbOptimType.EXACT_ROWWISE_ADAGRAD:
            return 1
        return self.get_state_dim(dim)
