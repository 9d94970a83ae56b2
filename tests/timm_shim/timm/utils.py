"""timm.utils helpers imported at benchmark.py:25-31."""
import argparse
import ast
import logging
import math


def setup_default_logging(default_level=logging.INFO, log_path=""):
    logging.basicConfig(level=default_level, format="%(message)s")


def set_jit_fuser(fuser):
    pass


def decay_batch_step(batch_size, num_intra_steps=2, no_odd=False):
    """Next smaller batch size to retry with: power-of-two bases with `num_intra_steps` steps in between (256 -> 192 -> 128 ...)."""
    if batch_size <= 1:
        return 0
    base = int(2 ** (math.log(batch_size - 1) // math.log(2)))
    step = max(base // num_intra_steps, 1)
    batch_size = base + ((batch_size - base - 1) // step) * step
    if no_odd and batch_size % 2:
        batch_size -= 1
    return batch_size


def check_batch_size_retry(error_str):
    """timm/utils/decay_batch.py: only out-of-memory style failures are worth a smaller batch."""
    error_str = error_str.lower()
    if "required rank" in error_str:
        return False
    return any(s in error_str for s in ("cudnn", "cuda", "out of memory", "illegal memory access"))


class ParseKwargs(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):
        kw = {}
        for value in values:
            key, value = value.split("=")
            try:
                kw[key] = ast.literal_eval(value)
            except (ValueError, SyntaxError):
                kw[key] = str(value)
        setattr(namespace, self.dest, kw)
