from nerrf_b200.ai.models import GraphSAGE_T, lstm  # noqa: F401
