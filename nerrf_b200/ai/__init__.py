"""Python surface named by the reference: ai.models.{GraphSAGE_T, lstm}, ai.planner.{mcts, rewards}
(README.md:72-76)."""
from . import models, planner  # noqa: F401
