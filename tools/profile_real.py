"""ncu target: one warm + one measured forward of a 64 MiB block of the box's own source text (tools/real_text.py corpus)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, jampack_b200 as jp
from real_text import corpus
T = corpus((int(sys.argv[1]) if len(sys.argv) > 1 else 64) << 20)
d_T = torch.from_numpy(T).cuda(); d_B = torch.zeros(T.size + 480, dtype=torch.uint8, device="cuda")
jp.forward_device(d_T, d_B); torch.cuda.synchronize()
torch.cuda.profiler.start()
jp.forward_device(d_T, d_B); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(jp.last_stats().asdict())
