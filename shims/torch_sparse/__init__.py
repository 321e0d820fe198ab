"""Functional stand-in for the torch_sparse surface of Label_propagation_model/outcome_correlation.py:39-55,139
(TEST INFRASTRUCTURE, see shims/README.md): ``SparseTensor(row, col, sparse_sizes)``, ``.sum(dim=1)``, scaling by a
[N, 1] / [1, N] dense vector, ``adj @ dense`` and ``.to(device)`` -- enough for the reference's own
``process_adj`` / ``gen_normalized_adjs`` / ``label_propagation`` to run unmodified as the oracle of the fused
label-propagation kernel.  COO triplets and ``index_add_``; nothing of torch-sparse's CSR machinery."""
import torch


class SparseTensor:
    def __init__(self, row=None, col=None, value=None, sparse_sizes=None):
        self.row, self.col, self.value = row, col, value
        self.sizes = tuple(sparse_sizes) if sparse_sizes is not None else (int(row.max()) + 1, int(col.max()) + 1)

    def _val(self, dtype=torch.float32):
        return self.value if self.value is not None else torch.ones(self.row.numel(), dtype=dtype, device=self.row.device)

    def sum(self, dim):
        idx, n = (self.row, self.sizes[0]) if dim == 1 else (self.col, self.sizes[1])
        return torch.zeros(n, dtype=self._val().dtype, device=idx.device).index_add_(0, idx, self._val())

    def __mul__(self, other):
        other = torch.as_tensor(other)
        if other.dim() == 2 and other.shape == (self.sizes[0], 1):
            v = self._val(other.dtype) * other[self.row, 0]
        elif other.dim() == 2 and other.shape == (1, self.sizes[1]):
            v = self._val(other.dtype) * other[0, self.col]
        else:
            raise NotImplementedError('shim SparseTensor: only [N, 1] and [1, N] scalings')
        return SparseTensor(self.row, self.col, v, self.sizes)

    __rmul__ = __mul__

    def to(self, device):
        return SparseTensor(self.row.to(device), self.col.to(device),
                            None if self.value is None else self.value.to(device), self.sizes)

    def __matmul__(self, dense):
        out = torch.zeros((self.sizes[0],) + tuple(dense.shape[1:]), dtype=dense.dtype, device=dense.device)
        return out.index_add_(0, self.row, self._val(dense.dtype).reshape(-1, *([1] * (dense.dim() - 1))) * dense[self.col])


def matmul(a, b):
    return a @ b
