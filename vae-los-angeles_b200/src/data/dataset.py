"""`MultiModalDataset` with the reference's constructor, `from_numpy` and public arrays
(reference src/data/dataset.py:10-51), plus batched access and a device-resident view.

The reference builds three fresh tensors per *sample* (`torch.tensor(row)` x3) and lets the default collate
stack them; here `__getitems__` (used by torch's DataLoader fetcher when present) slices a whole batch at once,
and `to_device()` hands the arrays to the fused trainer without any per-step host work."""
import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset


class MultiModalDataset(Dataset):
    """Columns: tpm_unstranded (RNA, list per row), beta_value (DNA methylation, list per row),
    primary_site_encoded (int)."""

    def __init__(self, dataframe):
        self.dataframe = dataframe
        self.tpm_data = np.ascontiguousarray(np.asarray(dataframe['tpm_unstranded'].tolist(), dtype=np.float32))
        self.beta_data = np.ascontiguousarray(np.asarray(dataframe['beta_value'].tolist(), dtype=np.float32))
        self.primary_site = np.asarray(dataframe['primary_site_encoded']).astype(np.int64)
        self._tpm_t = torch.from_numpy(self.tpm_data)
        self._beta_t = torch.from_numpy(self.beta_data)
        self._site_t = torch.from_numpy(self.primary_site)

    def __len__(self):
        return len(self.dataframe)

    def __getitem__(self, idx):
        return self._tpm_t[idx].clone(), self._beta_t[idx].clone(), self._site_t[idx].clone()

    def __getitems__(self, indices):
        idx = torch.as_tensor(indices, dtype=torch.long)
        tpm, beta, site = self._tpm_t[idx], self._beta_t[idx], self._site_t[idx]
        return [(tpm[i], beta[i], site[i]) for i in range(len(indices))]

    @classmethod
    def from_numpy(cls, tpm_data, beta_data, primary_site):
        return cls(pd.DataFrame({'tpm_unstranded': list(tpm_data), 'beta_value': list(beta_data),
                                 'primary_site_encoded': primary_site}))

    def to_device(self, device="cuda"):
        """Device-resident copy for vla_b200.Trainer."""
        from vla_b200.engine import DeviceDataset
        return DeviceDataset(self.tpm_data, self.beta_data, self.primary_site, device)
