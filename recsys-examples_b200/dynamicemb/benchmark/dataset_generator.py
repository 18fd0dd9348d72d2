"""Synthetic id streams of the reference's benchmarks (corelib/dynamicemb/benchmark/dataset_generator.py:4-103): the power-law
("Zipf 1.05") key stream of the embedding benchmark — the stream bench.py feeds — and the rank-shuffled Zipf sampler the HSTU example
uses for item ids and sequence lengths (examples/commons/datasets/hstu_batch.py:156-170).  Pure torch; runs on whatever device is asked."""
import torch


def translateToPowerLaw(min, max, alpha, x):          # noqa: A002  (the reference's parameter names)
    """Inverse CDF of a power law with exponent `alpha` on [min, max): uniform x in [0, 1) -> value (float, same dtype as x)."""
    g = 1.0 - float(alpha)
    lo, hi = float(min) ** g, float(max) ** g
    y = torch.pow(x * (hi - lo) + lo, 1.0 / g)
    return torch.where(y >= max, torch.full_like(y, float(max) - 1), y)


def PowerLaw(min, max, alpha, N, device=None, permute=None):          # noqa: A002
    """N int64 ids in [min, max) following the power law; `permute` (a lookup tensor) optionally relabels them."""
    device = torch.device("cuda") if device is None else device
    y = translateToPowerLaw(min, max, alpha, torch.rand(N, device=device, dtype=torch.float64)).to(torch.int64)
    return permute[y] if permute is not None else y


def gen_key(batch, hotness, alpha, N, device, permute=None):
    """`hotness * batch` ids of a table with N rows."""
    return PowerLaw(1, N, alpha, hotness * batch, device, permute)


def zipf(min_val, max_val, exponent, size, device):
    """`size` samples of [min_val, max_val) whose popularity follows Zipf(exponent) over a random ranking of the values (the most
    popular value is a random one, not `min_val`), drawn on `device` with torch.multinomial."""
    n = int(max_val) - int(min_val)
    ranks = torch.arange(1, n + 1, dtype=torch.float64, device=device)
    probs = ranks.pow(-float(exponent))
    probs = (probs / probs.sum()).to(torch.float32)
    labels = torch.randperm(n, device=device) + int(min_val)
    return labels[torch.multinomial(probs, int(size), replacement=True)]
