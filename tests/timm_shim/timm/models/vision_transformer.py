"""timm.models.vision_transformer._cfg (models/lemevit.py:22,866)."""


def _cfg(url="", **kwargs):
    d = {"url": url, "num_classes": 1000, "input_size": (3, 224, 224), "pool_size": None, "crop_pct": 0.9,
         "interpolation": "bicubic", "fixed_input_size": True, "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5),
         "first_conv": "patch_embed.proj", "classifier": "head"}
    d.update(kwargs)
    return d
