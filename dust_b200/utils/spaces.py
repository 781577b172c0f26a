"""Box space (API of dust/utils/spaces.py:4-65: `dim`, `shape`, `low`, `high`, `dtype`)."""
import torch


class Box:
    def __init__(self, dim, low=None, high=None, dtype=torch.float):
        assert dtype is not None, "Data type must be explicitly provided."
        assert isinstance(dtype, torch.dtype), "Data type must be of class `torch.dtype`."
        assert dim > 0, "Dimension must be a strictly positive integer."
        self.dtype = dtype
        self._dim = int(dim)
        self._shape = torch.Size([self._dim])
        self.low = self._bound(low, -float("inf"), "Lower")
        self.high = self._bound(high, float("inf"), "Higher")

    def _bound(self, value, default, which):
        if value is None:
            return torch.full(self._shape, default, dtype=torch.float)
        value = torch.as_tensor(value)
        if value.ndim == 0:
            return torch.full(self._shape, float(value), dtype=torch.float)
        assert value.shape == self._shape, f"{which} boundary must have same dimensions as space Box."
        return value

    @property
    def dim(self):
        return self._dim

    @property
    def shape(self):
        return self._shape
