from .graphsage_t import GraphSAGE_T  # noqa: F401
from . import lstm  # noqa: F401
