import collections.abc
import torch.nn as nn
from torch.nn.init import trunc_normal_  # noqa: F401


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return (x, x)


class DropPath(nn.Identity):
    def __init__(self, drop_prob=0.0, *a, **k):
        super().__init__()
