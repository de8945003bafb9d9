"""Access to the two extension modules of the boundary, imported once and cached.

Same entry points as the reference (pyc2ray/load_extensions.py:9-47: ``load_c2ray``, ``load_asora``), different
policy: the reference degrades to CPU-only operation when libasora cannot be imported (:41-44); this build has
no CPU path, so an unusable CUDA library is a RuntimeError that names the cause."""
import functools
import importlib


@functools.lru_cache(maxsize=None)
def _load(name, what):
    try:
        return importlib.import_module(f"{__package__}.lib.{name}")
    except ImportError as exc:
        raise RuntimeError(f"Could not load {what} library ({exc}); there is no CPU fallback in this build") from exc


def load_c2ray():
    """Chemistry half of the boundary (stands in for the f2py module libc2ray)."""
    return _load("libc2ray", "c2ray")


def load_asora():
    """Ray-tracing half of the boundary (stands in for the CPython module libasora)."""
    return _load("libasora", "ASORA")
