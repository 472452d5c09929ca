"""data_compressor_b200 -- B200-native (sm_100a) implementation of Data Compressor's DEGA hot path.

normalize -> diff -> seg -> bac (adaptive) and its inverse, plus A-XDR (normalize only) and LZMH, as hand-written
CUDA kernels behind a plain C ABI (include/dc_cuda.h).  See DESIGN.md and INTEGRATION.md.
"""
from ._lib import (DcError, Layout, Params, build, load, OK, ERR_INVALID_VALUE, ERR_INVALID_FORMAT, ERR_INVALID_MODE,
                   ERR_MEMORY, ERR_LIBRARY_INIT, ERR_LIBRARY_CALL)

__all__ = ["DcError", "Layout", "Params", "build", "load", "Context", "Encoded", "OK", "ERR_INVALID_VALUE",
           "ERR_INVALID_FORMAT", "ERR_INVALID_MODE", "ERR_MEMORY", "ERR_LIBRARY_INIT", "ERR_LIBRARY_CALL"]


def __getattr__(name):
    # codec imports torch; keep `import data_compressor_b200` cheap for the symbol-table tests
    if name in ("Context", "Encoded"):
        from . import codec
        return getattr(codec, name)
    if name in ("plugins", "synth", "container"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
