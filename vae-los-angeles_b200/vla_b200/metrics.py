"""Reconstruction metrics on the device (reference compare_directional_imputation.py:167-210, `compute_metrics`):
one streaming kernel instead of five scikit-learn / scipy passes over host copies and an N x N cosine matrix."""
import ctypes as C

import torch

from . import _lib
from .core import _ptr, _stream


def recon_metrics(y_true, y_pred, modality_name=None, model_name=None, per_sample=True):
    """Same keys as the reference's result dict: MAE, MSE, RMSE, R2, CosineSimilarity, PearsonMean, PearsonStd, and
    `_pearson_all` (the per-sample correlations with undefined ones dropped, as a CPU list for plotting) when
    `per_sample`.  Inputs: CUDA tensors [N, D] (any float dtype; converted to fp32).  One device->host read."""
    if not (y_true.is_cuda and y_pred.is_cuda):
        raise RuntimeError("vla_b200: recon_metrics needs CUDA tensors; there is no CPU fallback")
    if y_true.shape != y_pred.shape or y_true.dim() < 2:
        raise ValueError("y_true / y_pred must have the same [N, ...] shape")
    dev = y_true.device
    t = y_true.reshape(y_true.shape[0], -1).to(torch.float32).contiguous()
    p = y_pred.reshape(y_pred.shape[0], -1).to(device=dev, dtype=torch.float32).contiguous()
    n, d = t.shape
    L = _lib.lib()
    out = torch.empty(8, dtype=torch.float64, device=dev)
    ws = torch.zeros(L.vla_metrics_workspace_bytes(n), dtype=torch.uint8, device=dev)
    cos = torch.empty(n, dtype=torch.float32, device=dev) if per_sample else None
    pear = torch.empty(n, dtype=torch.float32, device=dev) if per_sample else None
    args = _lib.MetricsArgs(y_true=_ptr(t), y_pred=_ptr(p), rows=n, dim=d, cosine=_ptr(cos), pearson=_ptr(pear), out=_ptr(out),
                            workspace=_ptr(ws))
    with torch.cuda.device(dev):
        _lib.check(L.vla_recon_metrics(C.byref(args), _stream()), "vla_recon_metrics")
    mae, mse, rmse, r2, cs, pm, ps, cnt = out.tolist()
    res = {"Modality": modality_name, "Model": model_name, "MAE": mae, "MSE": mse, "RMSE": rmse, "R2": r2,
           "CosineSimilarity": cs, "PearsonMean": pm, "PearsonStd": ps}
    if per_sample:
        r = pear.cpu()
        res["_pearson_all"] = r[~torch.isnan(r)].tolist()
        res["_cosine_all"] = cos
    return res
