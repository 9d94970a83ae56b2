"""timm.data.resolve_data_config, as called at benchmark.py:431 (kwargs, model=, use_test_size=)."""


def resolve_data_config(args=None, default_cfg=None, model=None, use_test_size=False, verbose=False, pretrained_cfg=None):
    args = args or {}
    cfg = dict(pretrained_cfg or default_cfg or getattr(model, "pretrained_cfg", None) or getattr(model, "default_cfg", None) or {})
    in_chans = args.get("in_chans") or args.get("chans") or 3
    input_size = (in_chans, 224, 224)
    if args.get("input_size") is not None:
        input_size = tuple(args["input_size"])
    elif args.get("img_size") is not None:
        input_size = (in_chans, args["img_size"], args["img_size"])
    elif use_test_size and cfg.get("test_input_size") is not None:
        input_size = tuple(cfg["test_input_size"])
    elif cfg.get("input_size") is not None:
        input_size = tuple(cfg["input_size"])
    out = {"input_size": input_size}
    for key, default in (("interpolation", "bicubic"), ("mean", (0.485, 0.456, 0.406)), ("std", (0.229, 0.224, 0.225)), ("crop_pct", 0.875)):
        out[key] = args.get(key) or cfg.get(key, default)
    return out
