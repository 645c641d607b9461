"""Wall time of a reference-API sweep (calibration.tune_rabi, transmon(3) with T1/T2 collapse operators = config C1)
through the per-point host compilation vs device_synthesis=True.   python tools/sweep_bench.py [points]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sequencing_b200 import System, Transmon
from sequencing_b200.calibration import tune_rabi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for dev in (False, True, False, True):
    q = Transmon("qubit", levels=3, kerr=-0.2, t1=20e3, t2=15e3)
    system = System("system", modes=[q])
    t = time.perf_counter()
    _, old, new = tune_rabi(system, q.fock(0), amp_range=(-2, 2, n), progbar=False, plot=False, update=False, verify=False, device_synthesis=dev)
    dt = time.perf_counter() - t
    print(f"tune_rabi {n} points device_synthesis={dev}: {dt:.3f} s wall = {n / dt:.0f} sequences/s, new_amp {new:.12f}", flush=True)
