"""B200-native drop-in for the MP-HSIR block-stack hot path (net/MP_HSIR.py of the reference)."""
from .config import NetConfig  # noqa: F401

__all__ = ["NetConfig", "MP_HSIR_Net"]


def __getattr__(name):
    if name == "MP_HSIR_Net":
        from .model import MP_HSIR_Net
        return MP_HSIR_Net
    raise AttributeError(name)
