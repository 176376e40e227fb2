from nerrf_b200.ai.planner.rewards import *  # noqa: F401,F403
from nerrf_b200.ai.planner.rewards import score, reward_bounds, Actions  # noqa: F401
