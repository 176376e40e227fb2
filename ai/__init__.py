"""Drop-in alias: the reference's README.md:72-76 names a top-level `ai/` package
(ai/models/GraphSAGE-T.py, ai/models/lstm.py, ai/planner/mcts.py, ai/planner/rewards.py).
Everything lives in nerrf_b200.ai; this package only re-exports it under the reference's names."""
from nerrf_b200.ai import models, planner  # noqa: F401
