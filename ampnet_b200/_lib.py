"""ctypes binding of libampconv.so (the C ABI declared in include/ampconv.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# AMPNET_B200_LIB: an alternative build of the same library (kernel A/B experiments, tools/); the product is libampconv.so
LIB_PATH = os.environ.get("AMPNET_B200_LIB") or os.path.join(_HERE, "libampconv.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ampconv.h")

_lib = None


class AmpConvError(RuntimeError):
    def __init__(self, fn, status, detail=""):
        self.status = status
        super().__init__(f"{fn} failed with status {status}: {detail}")


def declared_symbols():
    """Names of every entry point declared in include/ampconv.h."""
    text = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"AMPCONV_API\s+[\w\s\*]+?\b(ampconv_\w+)\s*\(", text)))


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C ampnet_b200/csrc`. ampnet_b200 has no CPU or eager fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ampconv_strerror.restype = ctypes.c_char_p
        _lib.ampconv_strerror.argtypes = [ctypes.c_int]
    return _lib


def _as_arg(a):
    # tensors -> device pointer, None -> NULL, python ints/floats are passed explicitly typed by the caller
    if a is None:
        return ctypes.c_void_p(0)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    return a


def call(name, *args):
    lib = load()
    fn = getattr(lib, name)
    fn.restype = ctypes.c_int
    status = fn(*[_as_arg(a) for a in args])
    if status != 0:
        detail = lib.ampconv_strerror(status).decode()
        if status == -5:
            detail += f" (cudaError {lib.ampconv_last_cuda_error()})"
        raise AmpConvError(name, status, detail)
    return status


def launch_count():
    lib = load()
    lib.ampconv_launch_count.restype = ctypes.c_uint64
    return int(lib.ampconv_launch_count())


def i64(v):
    return ctypes.c_int64(int(v))


def i32(v):
    return ctypes.c_int(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def size_t(v):
    return ctypes.c_size_t(int(v))


def stream_ptr(stream):
    return ctypes.c_void_p(int(stream.cuda_stream))
