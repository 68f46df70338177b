"""ctypes binding of liblrce_b200.so (the C ABI declared in include/lrce_b200.h).

There is deliberately no fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C vqa-lrce-kbs-2023_b200/csrc``) loading raises, and every entry point raises RuntimeError on a
non-zero return code with the library's own message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblrce_b200.so")

_c = ctypes
_vp, _i, _f = _c.c_void_p, _c.c_int, _c.c_float

# name -> argtypes (every function returns int unless listed in _RESTYPES)
_SIGNATURES = {
    "lrce_abi_version": [],
    "lrce_last_error": [],
    "lrce_gemm_bf16": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _vp],
}
_RESTYPES = {"lrce_last_error": _c.c_char_p}

_lib = None


class LrceError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LrceError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()'). There is no CPU or PyTorch fallback for the LRCE hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, _c.c_int)
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = lib().lrce_last_error()
        raise LrceError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")
