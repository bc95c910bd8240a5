"""ctypes binding of the C-ABI kernel library ``libdafk.so`` (``include/dafk.h``).

The argument types of every entry point are parsed from the header itself, so the Python
binding can never drift from the C declaration.  There is NO fallback: if the shared library
is missing (or a call fails) this module raises -- the product path never silently runs on
anything but the hand-written sm_100a kernels.
"""
import ctypes
import os
import re

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
LIB_PATH = os.path.join(_PKG_DIR, "libdafk.so")
HEADER_PATH = os.path.join(_REPO_DIR, "include", "dafk.h")

DAFK_F32 = 0
DAFK_BF16 = 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3


class DafkError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """mirror of ``dafk_conv_desc``"""
    _fields_ = [(n, ctypes.c_int32) for n in
                ("N", "H", "W", "Cin", "Cout", "KH", "KW", "stride", "pad", "Ho", "Wo")]


_CTYPE = {
    "int": ctypes.c_int,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "size_t": ctypes.c_size_t,
}


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [(argname, ctype)])} for every function declared in dafk.h."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"typedef struct.*?\}\s*\w+;", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"(const char\*|int64_t|int)\s+(dafk_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        restype = {"const char*": ctypes.c_char_p, "int64_t": ctypes.c_int64, "int": ctypes.c_int}[ret]
        arglist = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    arglist.append((a.split("*")[-1].strip(), ctypes.c_void_p))
                else:
                    parts = a.replace("const ", "").split()
                    arglist.append((parts[-1], _CTYPE[parts[0]]))
        decls[name] = (restype, arglist)
    return decls


DECLS = parse_header()


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise DafkError(
                "libdafk.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C multimodal_segmentation_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.fn = {}
        for name, (restype, arglist) in DECLS.items():
            f = getattr(self.cdll, name)   # AttributeError if the symbol is not exported
            f.restype = restype
            f.argtypes = [t for _, t in arglist]
            self.fn[name] = f

    def last_error(self):
        return self.fn["dafk_last_error_string"]().decode()


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, ctypes.Structure):
        return ctypes.addressof(x)
    return x


PROFILE_ALL = False     # bench.py --profile-all: CUDA-event time of EVERY entry point (by name)


def call(name, *args):
    """Call ``dafk_<name>``; tensors become device pointers; raises DafkError on a non-zero status."""
    L = lib()
    full = name if name.startswith("dafk_") else "dafk_" + name
    f = L.fn[full]
    from . import instrument as _ins
    if _ins.trace is not None:        # diagnostic: ordered list of calls with their tensor operands, event-timed
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = f(*[_ptr(a) for a in args])
        e1.record()
        _ins.trace.append((name, [(tuple(a.shape), str(a.dtype)[6:]) for a in args if hasattr(a, "numel")], e0, e1))
        if f.restype is ctypes.c_int and rc != 0:
            raise DafkError("%s failed (%d): %s" % (full, rc, L.last_error()))
        return rc
    if PROFILE_ALL:
        from . import instrument
        if instrument.enabled:
            ptrs = [_ptr(a) for a in args]
            label = "all:" + name
            if instrument.by_shape:      # diagnostic: split every entry point by its largest operand
                ts = [a for a in args if hasattr(a, "numel")]
                if ts:
                    big = max(ts, key=lambda t: t.numel())
                    label += " %s %.2fM" % (str(big.dtype)[6:], big.numel() / 1e6)
            rc = instrument.timed(label, 0, 0, lambda: f(*ptrs))
            if f.restype is ctypes.c_int and rc != 0:
                raise DafkError("%s failed (%d): %s" % (full, rc, L.last_error()))
            return rc
    rc = f(*[_ptr(a) for a in args])
    if f.restype is ctypes.c_int and rc != 0:
        raise DafkError("%s failed (%d): %s" % (full, rc, L.last_error()))
    return rc


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(lib().fn["dafk_launch_count"]())
