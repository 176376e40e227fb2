from . import emit, mcts, rewards  # noqa: F401
