"""ctypes binding of libt2b200.so (C-ABI declared in include/t2b200.h).

There is deliberately NO CPU / eager fallback: if the shared library is missing or a call fails, the
product path raises. PyTorch is used only for device memory, streams and torch.distributed.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("T2_LIB") or os.path.join(_HERE, "libt2b200.so")     # T2_LIB: experiment builds (tools/)
_lib = None


class T2Error(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise T2Error("libt2b200.so is not built: run `python __graft_entry__.py build` "
                          "(there is no CPU fallback for the CUDA hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.t2_last_error.restype = ctypes.c_char_p
        _lib.t2_launch_count.restype = ctypes.c_longlong
    return _lib


def check(rc):
    if rc != 0:
        raise T2Error("libt2b200 error %d: %s" % (rc, load().t2_last_error().decode()))


def ptr(t):
    """Device (or host) pointer of a torch tensor as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
