"""Loads tests/emu/_build/libsimka_b200_emu.so (the kernel sources stepped on the CPU, see cuda_runtime.h) in place of the CUDA
library for the duration of a test.  Test infrastructure only: nothing outside tests/ imports this module."""
from __future__ import annotations

import contextlib
import ctypes as C
import os

from simka_b200 import _lib

from . import build_emu

_emu = None


def load_emulator() -> C.CDLL:
    global _emu
    if _emu is None:
        L = C.CDLL(str(build_emu.build()))
        for name, (res, args) in _lib.SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _emu = L
    return _emu


def use_emulator():
    """process-wide switch (developer runs with SIMKA_B200_EMU=1): the ctypes front end and, through LD_LIBRARY_PATH, the
    command-line binaries started by the tests"""
    _lib._lib = load_emulator()
    os.environ["LD_LIBRARY_PATH"] = cli_env()["LD_LIBRARY_PATH"]


@contextlib.contextmanager
def emulated():
    """`with emulated():` -- the ctypes front end of simka_b200 talks to the emulated library inside the block"""
    saved = _lib._lib
    _lib._lib = load_emulator()
    try:
        yield _lib._lib
    finally:
        _lib._lib = saved


def cli_env() -> dict:
    """environment for the command-line binaries (linked against libsimka_b200.so through a RUNPATH): a directory holding
    the emulated library under that name comes first on LD_LIBRARY_PATH"""
    d = build_emu.OUT / "lib"
    d.mkdir(exist_ok=True)
    link = d / "libsimka_b200.so"
    target = build_emu.build()
    if not link.exists() or os.path.realpath(link) != str(target):
        if link.is_symlink() or link.exists():
            link.unlink()
        link.symlink_to(target)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = str(d) + (":" + env["LD_LIBRARY_PATH"] if env.get("LD_LIBRARY_PATH") else "")
    return env
