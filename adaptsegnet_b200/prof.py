"""Per-kernel CUDA-event timing of libasn_b200's own launches (asn_prof_enable / asn_prof_report)."""
from __future__ import annotations

import ctypes as C
import json

from . import _lib


def enable(on: bool = True) -> None:
    _lib.check(_lib.load().asn_prof_enable(1 if on else 0), "asn_prof_enable")


def launch_count() -> int:
    """kernels launched by libasn_b200 so far (counted inside the library at every launch site)"""
    return int(_lib.load().asn_launch_count())


def report() -> dict:
    """{"kernel": {"launches", "ms", "flops", "bytes"}}; synchronises the recorded events."""
    lib = _lib.load()
    need = lib.asn_prof_report(None, 0)
    buf = C.create_string_buffer(int(need) + 64)
    lib.asn_prof_report(buf, len(buf))
    return json.loads(buf.value.decode())


def sequence() -> list:
    """names of the recorded kernel scopes in launch order"""
    lib = _lib.load()
    need = lib.asn_prof_sequence(None, 0)
    buf = C.create_string_buffer(int(need) + 64)
    lib.asn_prof_sequence(buf, len(buf))
    return [ln for ln in buf.value.decode().split("\n") if ln]
