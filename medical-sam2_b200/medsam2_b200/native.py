"""ctypes binding of the C-ABI declared in include/medsam2_b200.h.

The prototypes are parsed from the header itself, so the binding cannot drift from the declared
boundary.  There is NO fallback: if the shared library is missing or a call fails, an exception is
raised (the product path must fail loudly without its CUDA extension).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmedsam2_b200.so")
_HEADER_CANDIDATES = [
    os.path.join(HERE, "..", "..", "include", "medsam2_b200.h"),
    os.path.join(HERE, "include", "medsam2_b200.h"),
]

_CTYPES = {
    "int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float,
    "int32_t": ctypes.c_int32, "ms2_stream_t": ctypes.c_void_p,
}


def header_path():
    for p in _HEADER_CANDIDATES:
        if os.path.exists(p):
            return os.path.abspath(p)
    raise FileNotFoundError("include/medsam2_b200.h not found")


def parse_header(path=None):
    """-> {name: (restype, [argtypes], [argnames])} for every `ms2_*` prototype."""
    src = open(path or header_path()).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//.*", "", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int|long|void)\s+(ms2_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = (ctypes.c_char_p if "char" in ret else None if ret == "void" else ctypes.c_long if ret == "long"
                   else ctypes.c_int)
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(a.split("*")[-1].strip())
                else:
                    ty, nm = a.rsplit(" ", 1)
                    argtypes.append(_CTYPES[ty.replace("const ", "").strip()])
                    argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


class NativeError(RuntimeError):
    pass


_lib = None
launch_count = 0


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes, _) in parse_header().items():
            fn = getattr(l, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = l
    return _lib


_fns = {}


def call(name, *args):
    """Call an `int ms2_*` entry point; non-zero -> NativeError(ms2_last_error())."""
    global launch_count
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(lib(), name)
    rc = fn(*args)
    launch_count += 1
    if rc != 0:
        raise NativeError(f"{name} failed ({rc}): {lib().ms2_last_error().decode()}")
