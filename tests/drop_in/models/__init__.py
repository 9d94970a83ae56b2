"""What a maintainer switching to lemevit_b200 puts in place of the reference's `models/__init__.py:1`
(`from .lemevit import lemevit_tiny, lemevit_small, lemevit_base, ...`): the drivers' `from models import *`
(benchmark.py:70, main.py:39, validate.py:32) then registers the native entrypoints with timm.  TEST INFRASTRUCTURE."""
from lemevit_b200 import lemevit_base, lemevit_small, lemevit_tiny  # noqa: F401  (registration happens on import)
