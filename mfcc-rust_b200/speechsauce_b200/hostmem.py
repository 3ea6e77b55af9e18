"""Pinned host memory for callers that want the DMA-in-place path (ss_host_alloc / ss_host_register).

Every entry point also accepts ordinary (pageable) numpy arrays; those are staged through the library's
pinned rings by its copy threads."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from ._lib import check, lib


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """np.empty in page-locked memory (cudaHostAlloc, portable); freed when the array is collected."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    ptr = C.c_void_p()
    check(lib.ss_host_alloc(C.byref(ptr), max(nbytes, 16)))
    buf = (C.c_char * max(nbytes, 16)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)
    weakref.finalize(buf, lib.ss_host_free, C.c_void_p(ptr.value))
    return arr


def pin(arr: np.ndarray) -> np.ndarray:
    """Page-lock an existing C-contiguous array in place (cudaHostRegister); undo with `unpin`."""
    if not arr.flags.c_contiguous:
        raise ValueError("pin() needs a C-contiguous array")
    check(lib.ss_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes))
    return arr


def unpin(arr: np.ndarray) -> None:
    check(lib.ss_host_unregister(C.c_void_p(arr.ctypes.data)))


def is_pinned(arr: np.ndarray) -> bool:
    return bool(lib.ss_host_is_pinned(C.c_void_p(arr.ctypes.data)))
