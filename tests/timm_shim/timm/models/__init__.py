"""timm.models registry + factory with timm 0.9's calling convention (timm/models/_factory.py create_model):
None-valued kwargs are dropped, `scriptable/exportable/no_jit` are layer-config switches (not forwarded), the entrypoint is
called as fn(pretrained=, pretrained_cfg=, pretrained_cfg_overlay=, **kwargs), `checkpoint_path` loads a state_dict."""
import fnmatch

import torch

_ENTRYPOINTS = {}


def register_model(fn):
    _ENTRYPOINTS[fn.__name__] = fn
    return fn


def is_model(name):
    return name in _ENTRYPOINTS


def list_models(filter="", module="", pretrained=False, exclude_filters="", name_matches_cfg=False, include_tags=None):
    names = sorted(_ENTRYPOINTS)
    if filter:
        names = [n for n in names if fnmatch.fnmatch(n, filter)]
    for ex in ([exclude_filters] if isinstance(exclude_filters, str) else exclude_filters or []):
        if ex:
            names = [n for n in names if not fnmatch.fnmatch(n, ex)]
    return names


def model_entrypoint(name):
    return _ENTRYPOINTS[name]


def create_model(model_name, pretrained=False, pretrained_cfg=None, pretrained_cfg_overlay=None, checkpoint_path="",
                 scriptable=None, exportable=None, no_jit=None, **kwargs):
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    if not is_model(model_name):
        raise RuntimeError("Unknown model (%s)" % model_name)
    model = _ENTRYPOINTS[model_name](pretrained=pretrained, pretrained_cfg=pretrained_cfg,
                                     pretrained_cfg_overlay=pretrained_cfg_overlay, **kwargs)
    if checkpoint_path:
        ckpt = torch.load(checkpoint_path, map_location="cpu")
        for key in ("state_dict_ema", "model_ema", "state_dict", "model"):
            if isinstance(ckpt, dict) and key in ckpt:
                ckpt = ckpt[key]
                break
        model.load_state_dict(ckpt)
    return model
