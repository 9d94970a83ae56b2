"""timm.optim.create_optimizer_v2 (benchmark.py:24; only the train bench calls it)."""
import torch


def create_optimizer_v2(model_or_params, opt="sgd", lr=None, weight_decay=0.0, momentum=0.9, **kwargs):
    params = model_or_params.parameters() if hasattr(model_or_params, "parameters") else model_or_params
    opt = (opt or "sgd").lower()
    if opt == "adamw":
        return torch.optim.AdamW(params, lr=lr or 1e-3, weight_decay=weight_decay, eps=kwargs.get("eps") or 1e-8)
    if opt == "adam":
        return torch.optim.Adam(params, lr=lr or 1e-3, weight_decay=weight_decay)
    return torch.optim.SGD(params, lr=lr or 1e-2, momentum=momentum, weight_decay=weight_decay)
