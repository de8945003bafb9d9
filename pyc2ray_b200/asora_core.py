"""Device life-cycle of the ASORA library as pyc2ray exposes it (reference: pyc2ray/asora_core.py:14-58):
``device_init`` / ``device_close`` / ``photo_table_to_device`` plus the ``cuda_is_init`` guard the evolve and
ray-tracing entry points check before touching the GPU."""
from .load_extensions import load_asora

__all__ = ["cuda_is_init", "device_init", "device_close", "photo_table_to_device"]

libasora = load_asora()
_NOT_INIT = "GPU not initialized. Please initialize it by calling device_init(N)"  # asora_core.py:47,58


class _DeviceState:
    ready = False


def cuda_is_init():
    """True between device_init() and device_close()."""
    return _DeviceState.ready


def _require_device():
    if not _DeviceState.ready:
        raise RuntimeError(_NOT_INIT)


def device_init(N, source_batch_size):
    """Bind to the current CUDA device and allocate the grids for an N^3 mesh (asora_core.py:20-37).
    ``source_batch_size`` is accepted for compatibility: column densities stay on-chip here, so it no longer
    bounds memory use."""
    libasora.device_init(N, source_batch_size)
    _DeviceState.ready = True


def device_close():
    """Free all device memory (asora_core.py:39-47)."""
    _require_device()
    libasora.device_close()
    _DeviceState.ready = False


def photo_table_to_device(thin_table, thick_table):
    """Upload the optically thin / thick photo-ionisation tables (asora_core.py:49-58); the table length is
    taken from ``thin_table``, as in the reference."""
    _require_device()
    libasora.photo_table_to_device(thin_table, thick_table, thin_table.shape[0])
