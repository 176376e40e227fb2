from nerrf_b200.ai.models.lstm import *  # noqa: F401,F403
from nerrf_b200.ai.models.lstm import LSTMScorer, Model, forward  # noqa: F401
