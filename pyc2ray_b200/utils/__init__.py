from .sourceutils import format_sources, generate_test_sources, generate_test_sourcefile, read_test_sources
from .logutils import printlog
from . import c2ray_files
from .c2ray_files import save_cbin, read_cbin, get_source_redshifts, get_redshifts_from_output, find_bins
