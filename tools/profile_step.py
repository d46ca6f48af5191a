"""One warm + one measured step of the hot path (inverse, then forward) with device-resident blocks: the
command ncu wraps (see profiles/README.md). Prints the library's own per-phase event times for comparison."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import jampack_b200 as jp  # noqa: E402
import synth  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
kind = sys.argv[2] if len(sys.argv) > 2 else "markov2"
what = sys.argv[3] if len(sys.argv) > 3 else "both"
n = mib << 20
T = synth.gen(kind, n, {"uniform": 2, "repetitive": 3, "alla": 0}.get(kind, 1))
d_T = torch.from_numpy(T).cuda()
d_B = torch.zeros(n + 480, dtype=torch.uint8, device="cuda")
d_back = torch.zeros(n, dtype=torch.uint8, device="cuda")
jp.forward_device(d_T, d_B)           # warm (also produces the inverse's input)
jp.inverse_device(d_B, d_back)        # warm
torch.cuda.synchronize()
torch.cuda.profiler.start()            # ncu --profile-from-start off captures the measured step only
if what in ("both", "inv"):
    jp.inverse_device(d_B, d_back)
    print("inverse", jp.last_stats().asdict())
if what in ("both", "fwd"):
    jp.forward_device(d_T, d_B)
    print("forward", jp.last_stats().asdict())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
assert torch.equal(d_back, d_T)
