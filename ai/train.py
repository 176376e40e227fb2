"""Reference-named module ai/train.py (README.md:75) -> nerrf_b200.ai.train."""
from nerrf_b200.ai.train import *  # noqa: F401,F403
from nerrf_b200.ai.train import evaluate, main, roc_auc, train  # noqa: F401

if __name__ == "__main__":
    main()
