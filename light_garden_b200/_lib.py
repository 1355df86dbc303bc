"""Loads the CUDA library.  There is no fallback: if the shared object is
missing or does not export the whole C ABI, importing the product fails."""
import ctypes
import os

from . import abi

# LG_LIB_PATH: an alternate build of the same library (kernel tuning experiments); still no fallback
LIB_PATH = os.environ.get("LG_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib",
                                                         "liblight_garden_b200.so")
_lib = None


class LightGardenError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{abi.ERROR_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(light_garden_b200 has no CPU or PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        abi.bind(lib)
        if lib.lg_abi_version() != abi.LG_ABI_VERSION:
            raise ImportError("liblight_garden_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(ctx, rc):
    if rc != 0:
        msg = load().lg_last_error(ctx) if ctx else b""
        raise LightGardenError(rc, (msg or b"").decode("utf-8", "replace"))
