"""Stand-in for the timm symbols the reference's drivers use (see ../README.md).  TEST INFRASTRUCTURE."""
__version__ = "0.9.16-shim"
_lemevit_stub = True
