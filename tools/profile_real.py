import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, numpy as np, jampack_b200 as jp
from real_text import corpus
T = corpus(64 << 20)
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(T.size + 480, dtype=torch.uint8, device="cuda")
jp.forward_device(d_T, d_B); jp.forward_device(d_T, d_B)
print(jp.last_stats().asdict())
