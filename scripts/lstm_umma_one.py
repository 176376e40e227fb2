"""One tensor-core LSTM forward (for ncu): python scripts/lstm_umma_one.py [B] [T]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nerrf_b200.ai.models.lstm import LSTMScorer  # noqa: E402

os.environ["NERRF_LSTM_ALGO"] = "umma"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2304
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
m = LSTMScorer().cuda()
seq = torch.randn(B, T, 16, device="cuda"); ln = torch.randint(T // 2, T + 1, (B,), device="cuda")
print(m(seq, ln).sum().item())
