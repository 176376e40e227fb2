from nerrf_b200.ai.planner.mcts import *  # noqa: F401,F403
from nerrf_b200.ai.planner.mcts import search, plan, SearchResult  # noqa: F401
