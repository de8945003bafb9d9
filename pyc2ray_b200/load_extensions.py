"""Loader for the two extension modules (reference: pyc2ray/load_extensions.py:9-47).

The reference falls back to CPU-only operation when libasora is missing (:41-44); this build has no
CPU path, so a missing CUDA library is an error."""

_c2ray_lib = None
_asora_lib = None


def load_c2ray():
    global _c2ray_lib
    if _c2ray_lib is None:
        try:
            from .lib import libc2ray
        except ImportError as e:
            raise RuntimeError(f"Could not load c2ray library ({e})")
        _c2ray_lib = libc2ray
    return _c2ray_lib


def load_asora():
    global _asora_lib
    if _asora_lib is None:
        try:
            from .lib import libasora
        except ImportError as e:
            raise RuntimeError(f"Could not load ASORA library ({e}); there is no CPU fallback in this build")
        _asora_lib = libasora
    return _asora_lib
