"""A/B of two builds of the engine on one workload: kernel time of the C2 bench sweep (1024 trajectories, min of 3)
per library.   python tools/ab_lib.py sequencing_b200/_seq_engine.so sequencing_b200/_seq_engine_x.so [c1|c2|c3|c4]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and not sys.argv[1].startswith("--one"):
    which = sys.argv[3] if len(sys.argv) > 3 else "c2"
    for lib in sys.argv[1:3]:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--one", lib, which])
    sys.exit(0)
sys.path.insert(0, ROOT)
from sequencing_b200.solver import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[2])
from sequencing_b200 import workloads as w
from sequencing_b200.solver import Engine

bp = {"c2": lambda: w.c2_selective(32, 32), "c4": lambda: w.c4_two_transmons_bus(512), "c1": lambda: w.c1_rabi(65536),
      "c3": lambda: w.c3_drag(4096)}[sys.argv[3]]()
eng = Engine.get(0)
tens = eng.to_device(bp)
for _ in range(2):
    eng.solve_device(tens, bp)
    eng.torch.cuda.synchronize()
ms = []
for _ in range(4):
    eng.solve_device(tens, bp)
    eng.torch.cuda.synchronize()
    ms.append(eng.kernel_ms())
out = eng.solve_device(tens, bp)
st = out["out"]["stats"].cpu().numpy()
print(f"   steps acc {st[:, 0].mean():.2f} rej {st[:, 1].mean():.2f} rhs {st[:, 2].mean():.2f}")
print(f"{sys.argv[2]}: {sys.argv[3]} kernel_ms min {min(ms):.3f} all {[round(m, 3) for m in ms]}  -> {bp.B / min(ms) * 1e3:.0f} trajectories/s")
