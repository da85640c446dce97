"""ctypes binding of libregda_b200.so (include/regda_b200.h) -- the only way the Python host
layer reaches the CUDA kernels.  There is no fallback: if the library is missing or the
call fails, an exception is raised.

PyTorch is used for device memory and streams only: tensors are passed as raw device
pointers, the stream is torch's current stream handle.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libregda_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "regda_b200.h")

FLAG_LABEL_RANGE = 1
FLAG_REGION_RANGE = 2
FLAG_PROB_RANGE = 4

_lib = None
launch_count = 0  # kernels-launching C-ABI calls made through this module (bench.py reads it)

_CTYPES = {
    "int": ctypes.c_int, "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "size_t": ctypes.c_size_t,
    "double": ctypes.c_double, "float": ctypes.c_float, "unsigned": ctypes.c_uint, "uint64_t": ctypes.c_uint64,
}


class RegdaError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"regda_b200 error {code}: {text}")
        self.code = code


def declared_symbols(header_path: str = HEADER_PATH):
    """[(name, restype, [argtype,...])] parsed from the public header (it is plain C)."""
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"#if 0.*?#endif", " ", src, flags=re.S)
    out = []
    for m in re.finditer(r"\b(int|size_t|const char \*|void)\s*\*?\s*(regda_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argtypes = []
        for a in [x.strip() for x in args.split(",")]:
            if a in ("void", ""):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
            else:
                base = a.replace("const ", "").split()[0]
                argtypes.append(_CTYPES[base])
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "void": None}.get(ret, ctypes.c_char_p)
        out.append((name, restype, argtypes))
    return out


def lib():
    """Load the library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m regda_b200.build` "
                "(regda_b200 has no CPU / eager fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, restype, argtypes in declared_symbols():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        if L.regda_abi_version() != 3:
            raise RuntimeError("regda_b200 ABI version mismatch")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().regda_last_error().decode()


def check(rc: int):
    if rc != 0:
        raise RegdaError(rc, last_error())


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "regda_b200 takes contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def ptr_any(t):
    """device pointer of a tensor whose memory (in whatever stride order) is dense"""
    assert t.is_cuda
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Call an int-returning entry point that launches work on the current stream."""
    global launch_count
    launch_count += 1
    check(getattr(lib(), name)(*args))


class Workspace:
    """Grow-only device scratch buffer owned by the caller side (the library never allocates).  A buffer that has been handed
    out may already be baked into a captured CUDA graph (its raw pointer is a kernel argument of the captured launches), so a
    buffer that is outgrown is RETIRED, never freed: replays of older graphs keep reading and writing memory that is still
    theirs (a freed workspace would be silently reused by the caching allocator under such a graph)."""

    def __init__(self):
        self._buf = {}
        self._retired = []

    def get(self, nbytes: int, device):
        nbytes = max(int(nbytes), 256)
        key = torch.device(device).index or 0
        b = self._buf.get(key)
        if b is None or b.numel() < nbytes:
            if b is not None:
                self._retired.append(b)
            b = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._buf[key] = b
        return b


workspace = Workspace()


class ZeroPool:
    """Pre-zeroed float32 scratch for the many small accumulators of a step (BatchNorm sums, backward reductions):
    the trainer zeroes ONE buffer per step (`reset()`, a single memset node in the captured graph) and the ops take
    consecutive slices instead of issuing one tiny memset each.  Outside an armed step `take` returns (tensor, False)
    and the callee zeroes it itself."""

    def __init__(self, nfloats=8 << 20):
        self.nfloats = nfloats
        self._buf = {}
        self._off = 0
        self._armed = False

    def reset(self, device):
        key = torch.device(device).index or 0
        b = self._buf.get(key)
        if b is None:
            b = self._buf[key] = torch.empty(self.nfloats, dtype=torch.float32, device=device)
        b.zero_()
        self._off = 0
        self._armed = True

    def disarm(self):
        self._armed = False

    def take(self, shape, device):
        n = 1
        for d in shape:
            n *= int(d)
        key = torch.device(device).index or 0
        b = self._buf.get(key)
        n4 = (n + 3) // 4 * 4
        if self._armed and b is not None and self._off + n4 <= self.nfloats:
            t = b[self._off:self._off + n].view(shape)
            self._off += n4
            return t, True
        return torch.empty(shape, dtype=torch.float32, device=device), False


zero_pool = ZeroPool()
