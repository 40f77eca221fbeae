"""B200-native hot path of X-LXMERT (encoder, cluster head, generator). See DESIGN.md."""
from .config import LxmertDims, DEFAULT_DIMS, TINY_DIMS  # noqa: F401

__version__ = "0.1.0"
