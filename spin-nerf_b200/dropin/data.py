"""Drop-in for DS_NeRF/data.py (the `RayDataset` the reference trainer wraps in four DataLoaders, run_nerf.py:1340-1348;
SURVEY.md section 8 row f1).  Same class, constructor and per-item behaviour (data.py:4-15); in addition it implements the
batched fetch protocol of torch's DataLoader (`__getitems__`): a batch of N_rand rays is ONE gather from the ray array
instead of N_rand Python `__getitem__` calls each building a tensor — the host hot loop of the unmodified script
(3-4 loaders x N_rand items per step).  Batches are bit-identical to the reference's for the same sampler / generator.

Resolved ahead of the reference's own data.py by the module-resolution recipe of INTEGRATION.md section 1 (this
directory precedes $REF/DS_NeRF on PYTHONPATH).  The fully device-resident sampler is spin-nerf_b200/raypool.py."""
import numpy as np
import torch
import torch.utils.data as data
from torch.utils.data._utils.collate import default_collate_fn_map


class _Row:
    """Element type of a pre-collated batch (see _Rows)."""


class _Rows:
    """What `__getitems__` hands to the DataLoader's default_collate: looks like a sequence of rows, but the rows are
    already one gathered [N,...] tensor; the collate function registered for its element type returns that tensor
    instead of stacking N row views (torch's documented extension point for custom batch element types)."""

    def __init__(self, tensor):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, i):
        return _Row()

    def __iter__(self):
        return iter(self.tensor.unbind(0))


default_collate_fn_map[_Row] = lambda batch, *, collate_fn_map=None: batch.tensor


class RayDataset(data.Dataset):
    def __init__(self, ray_data):
        super(RayDataset, self).__init__()
        self.rayData = ray_data
        self.length = ray_data.shape[0]
        self._rows = None

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        return torch.Tensor(self.rayData[index])          # data.py:14-15

    def __getitems__(self, indices):
        if self._rows is None:
            self._rows = np.asarray(self.rayData)
        # one numpy gather; float32 like torch.Tensor(row) (data.py:15 converts each row the same way)
        return _Rows(torch.as_tensor(self._rows[np.asarray(indices, dtype=np.int64)], dtype=torch.float32))
