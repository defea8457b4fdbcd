"""ctypes driver for the CPU warp emulator (tests/simt_emu) -- TEST INFRASTRUCTURE ONLY.

It compiles the *kernel headers themselves* (lz4net_b200/csrc/*.cuh) with g++ in emulation mode so that the
warp-level algorithms can be checked against the oracle in a container without a GPU.  Nothing here is a product
code path: liblz4b200.so contains no CPU codec.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_DIR = os.path.join(_HERE, "simt_emu")
_SO = os.path.join(_DIR, "libsimt_emu.so")
_lib = None


def _sources():
    csrc = os.path.join(_ROOT, "lz4net_b200", "csrc")
    return [os.path.join(_DIR, f) for f in ("simt_emu.cpp", "emu_harness.cpp", "simt_emu.h")] + \
           [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in _sources()):
            subprocess.check_call(["g++", "-O2", "-g", "-shared", "-fPIC", "-std=c++17", "-DLZ4B200_SIMT_EMU",
                                   "-I" + _DIR, "-I" + os.path.join(_ROOT, "lz4net_b200", "csrc"), "-o", _SO,
                                   os.path.join(_DIR, "simt_emu.cpp"), os.path.join(_DIR, "emu_harness.cpp")])
        _lib = C.CDLL(_SO)
    return _lib


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def decode(blocks, caps, lanes=32, known=True, sched_seed=1, src_skew=0, dst_skew=0):
    """blocks: list of bytes (compressed); caps: list of int. Returns (results, outputs).  src_skew shifts the
    compressed bytes inside their buffer to exercise unaligned stream starts."""
    n = len(blocks)
    srcs = []
    for b in blocks:
        a = np.zeros(src_skew + len(b) + 64, np.uint8)
        a[src_skew:src_skew + len(b)] = np.frombuffer(b, np.uint8)
        srcs.append(a)
    dsts = [np.full(max(c, 0) + 96 + dst_skew, 0xCD, np.uint8) for c in caps]
    isz = np.array([len(b) for b in blocks], np.int32)
    cps = np.array(caps, np.int32)
    res = np.zeros(n, np.int32)
    sp = (C.c_void_p * n)(*[a.ctypes.data + src_skew for a in srcs])
    dp = (C.c_void_p * n)(*[a.ctypes.data + 32 + dst_skew for a in dsts])       # 32-byte red zone in front
    if lanes in (1, 2):      # lane-per-block decoder (1: the kernel's geometry, 2: larger rings / 64-byte runs / one sequence per iteration)
        lib().emu_decode_lpb(lanes - 1, int(known), n, sp, isz.ctypes.data_as(C.c_void_p), dp, cps.ctypes.data_as(C.c_void_p),
                             res.ctypes.data_as(C.c_void_p), C.c_uint64(sched_seed))
    else:
        lib().emu_decode(lanes, int(known), n, sp, isz.ctypes.data_as(C.c_void_p), dp, cps.ctypes.data_as(C.c_void_p),
                         res.ctypes.data_as(C.c_void_p), C.c_uint64(sched_seed))
    outs = []
    lo = 32 + dst_skew
    for d, c in zip(dsts, caps):
        assert (d[:lo] == 0xCD).all() and (d[lo + max(c, 0):] == 0xCD).all(), "decoder wrote outside [dst, dst+cap)"
        outs.append(d[lo:lo + max(c, 0)].tobytes())
    return res.tolist(), outs


def encode(blocks, caps=None, sched_seed=1, src_skew=0, dst_skew=0, variant=2, tune=(12, 8, 24)):
    """variant: how the encoder finds same-hash lanes inside a round (1 always exact votes, 2 optimistic = the default)."""
    n = len(blocks)
    lib().emu_set_encode_variant(variant)
    if caps is None:
        caps = [len(b) + len(b) // 255 + 16 for b in blocks]
    srcs = []
    for b in blocks:
        a = np.zeros(src_skew + len(b) + 64, np.uint8)
        a[src_skew:src_skew + len(b)] = np.frombuffer(b, np.uint8)
        srcs.append(a)
    dsts = [np.full(max(c, 0) + 96 + dst_skew, 0xCD, np.uint8) for c in caps]
    ns = np.array([len(b) for b in blocks], np.int32)
    cps = np.array(caps, np.int32)
    res = np.zeros(n, np.int32)
    sp = (C.c_void_p * n)(*[a.ctypes.data + src_skew for a in srcs])
    dp = (C.c_void_p * n)(*[a.ctypes.data + 32 + dst_skew for a in dsts])
    lib().emu_encode(n, sp, ns.ctypes.data_as(C.c_void_p), dp, cps.ctypes.data_as(C.c_void_p),
                     res.ctypes.data_as(C.c_void_p), C.c_uint64(sched_seed))
    outs = []
    for d, c, r in zip(dsts, caps, res.tolist()):
        lo = 32 + dst_skew
        assert (d[:lo] == 0xCD).all() and (d[lo + max(c, 0):] == 0xCD).all(), "encoder wrote outside [dst, dst+cap)"
        outs.append(d[lo:lo + max(r, 0)].tobytes())
    return res.tolist(), outs


def encode_hc(block, cap=None, src_skew=0):
    n = len(block)
    if cap is None:
        cap = n + n // 255 + 16
    a = np.zeros(src_skew + n + 64, np.uint8)
    a[src_skew:src_skew + n] = np.frombuffer(block, np.uint8)
    d = np.full(max(cap, 0) + 96, 0xCD, np.uint8)
    r = lib().emu_encode_hc(C.c_void_p(a.ctypes.data + src_skew), n, C.c_void_p(d.ctypes.data + 32), cap)
    assert (d[:32] == 0xCD).all() and (d[32 + max(cap, 0):] == 0xCD).all(), "HC encoder wrote outside [dst, dst+cap)"
    return int(r), d[32:32 + max(r, 0)].tobytes()


HCW_FALLBACK = -2 ** 31


def encode_hcw(block, cap=None, src_skew=0, dst_skew=0, sched_seed=1, smem=True):
    """The warp-per-block HC encoder (lz4hc_warp.cuh).  Returns (result, bytes); result == HCW_FALLBACK when the kernel hands
    the block to the thread-per-block one."""
    n = len(block)
    if cap is None:
        cap = n + n // 255 + 16
    a = np.zeros(src_skew + n + 64, np.uint8)            # smem: block staged in shared memory; else read through the read-only path
    a[src_skew:src_skew + n] = np.frombuffer(block, np.uint8)
    lo = 32 + dst_skew
    d = np.full(max(cap, 0) + 96 + dst_skew, 0xCD, np.uint8)
    f = lib().emu_encode_hcw
    f.restype = C.c_int
    r = f(C.c_void_p(a.ctypes.data + src_skew), n, C.c_void_p(d.ctypes.data + lo), cap, C.c_uint64(sched_seed), int(smem))
    assert (d[:lo] == 0xCD).all() and (d[lo + max(cap, 0):] == 0xCD).all(), "HC warp encoder wrote outside [dst, dst+cap)"
    return int(r), d[lo:lo + max(r, 0)].tobytes()


def encode_guarded(block, sched_seed=1, variant=2, end_pad=0):
    """Encode one block whose last byte is the last byte of a readable page: the page after it is PROT_NONE, so any
    read of a word that holds no input byte (an over-read past the end of the input) kills the process with SIGSEGV.
    end_pad (0..3) bytes stay readable behind the block: the rest of its last aligned 32-bit word (the kernels read whole
    aligned words).  Returns (result, bytes)."""
    import mmap
    page = mmap.PAGESIZE
    n = len(block)
    assert 0 <= end_pad < 4
    npages = (n + page - 1) // page + 1
    m = mmap.mmap(-1, (npages + 1) * page)
    base = C.addressof(C.c_char.from_buffer(m))
    libc = C.CDLL(None, use_errno=True)
    libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    assert libc.mprotect(base + npages * page, page, 0) == 0
    start = npages * page - n - end_pad
    assert (start + n + 3) // 4 * 4 <= npages * page
    m[start:start + n] = block
    cap = n + n // 255 + 16
    dst = np.full(cap + 96, 0xCD, np.uint8)
    ns = np.array([n], np.int32); cps = np.array([cap], np.int32); res = np.zeros(1, np.int32)
    sp = (C.c_void_p * 1)(base + start); dp = (C.c_void_p * 1)(dst.ctypes.data + 32)
    lib().emu_set_encode_variant(variant)
    lib().emu_encode(1, sp, ns.ctypes.data_as(C.c_void_p), dp, cps.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p),
                     C.c_uint64(sched_seed))
    r = int(res[0])
    out = dst[32:32 + max(r, 0)].tobytes()
    assert libc.mprotect(base + npages * page, page, 3) == 0
    del sp
    return r, out
