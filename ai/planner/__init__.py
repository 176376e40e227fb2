from nerrf_b200.ai.planner import mcts, rewards  # noqa: F401
