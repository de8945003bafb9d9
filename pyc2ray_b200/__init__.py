"""pyc2ray_b200 -- B200-native (sm_100a) ASORA ray tracing + ionisation chemistry behind pyc2ray's
Python boundary.  Importing the package loads libasora_b200.so; there is no CPU fallback."""
from .asora_core import cuda_is_init, device_init, device_close, photo_table_to_device
from .evolve import evolve3D, evolve3D_MPI, evolve3D_dist
from .raytracing import do_raytracing
from .chemistry import hydrogenODE
from .radiation import make_tau_table, BlackBodySource, blackbody_tables
from .utils.sourceutils import format_sources, generate_test_sources, read_test_sources
from .c2ray_base import C2Ray, C2Ray_Test
from .c2ray_244paper import C2Ray_244Test
from .utils.logutils import printlog
from . import evolve, raytracing, chemistry, asora_core, radiation, utils
