import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerrf_b200.ai.models.lstm import LSTMScorer
os.environ["NERRF_LSTM_PROF"] = "1"
m = LSTMScorer().cuda()
B, T = 4096, 100
seq = torch.randn(B, T, 16, device="cuda"); ln = torch.randint(T // 2, T + 1, (B,), device="cuda")
m(seq, ln); torch.cuda.synchronize()
