"""B200-native force/energy path of np-shape-lab."""
from . import mesh  # noqa: F401
