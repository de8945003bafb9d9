"""Initialisation state of the ASORA library (reference: pyc2ray/asora_core.py)."""
from .load_extensions import load_asora

libasora = load_asora()

__all__ = ["cuda_is_init", "device_init", "device_close", "photo_table_to_device"]

cuda_init = False


def cuda_is_init():
    return cuda_init


def device_init(N, source_batch_size):
    """Initialise the GPU and allocate the grids for mesh size N (asora_core.py:20-37).

    ``source_batch_size`` is kept for compatibility; column densities stay on-chip here, so it does
    not bound memory any more."""
    global cuda_init
    if libasora is None:
        raise RuntimeError("Could not initialize GPU: ASORA library not loaded")
    libasora.device_init(N, source_batch_size)
    cuda_init = True


def device_close():
    """asora_core.py:39-47"""
    global cuda_init
    if not cuda_init:
        raise RuntimeError("GPU not initialized. Please initialize it by calling device_init(N)")
    libasora.device_close()
    cuda_init = False


def photo_table_to_device(thin_table, thick_table):
    """asora_core.py:49-58"""
    if not cuda_init:
        raise RuntimeError("GPU not initialized. Please initialize it by calling device_init(N)")
    libasora.photo_table_to_device(thin_table, thick_table, thin_table.shape[0])
