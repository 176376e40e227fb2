from . import mcts, rewards  # noqa: F401
