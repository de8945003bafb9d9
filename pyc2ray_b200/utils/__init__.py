from .sourceutils import format_sources, generate_test_sources, generate_test_sourcefile, read_test_sources
from .logutils import printlog
