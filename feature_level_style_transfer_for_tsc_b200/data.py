"""Synthetic UCR/UEA/HAR-shaped batches: what the benchmark and the examples feed the step when no
dataset is available.  Shapes and dtypes are those ``DataSource.TrainData`` yields after the trainer's casts
(``x.float()``: fp32 ``[B, C, L]``; labels int64, DataSource.py:30, train_and_test.py:543-546)."""
import torch


def synthetic_batch(B: int, C: int, L: int, n_class: int, domain_id: int = 0):
    """Seeded ``randn(B, C, L)``, z-normalised per series over L (the UCR/UEA convention; the reference's own commented
    check, multi_source_voting.py:104-115), and ``randint`` labels.  Host tensors; one generator per domain."""
    g = torch.Generator().manual_seed(1234 + domain_id)
    x = torch.randn(B, C, L, generator=g)
    x = (x - x.mean(-1, keepdim=True)) / x.std(-1, keepdim=True)
    y = torch.randint(0, n_class, (B,), generator=g)
    return x.float(), y.long()
