from nerrf_b200.ai.planner import emit, mcts, rewards  # noqa: F401
