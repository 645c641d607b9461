# cuBLAS DGEMM / ZGEMM throughput through torch (library reference for the FP64 roofline).
import json, torch, time
dev = torch.device("cuda:0")
out = {}
for name, dt, fl in (("dgemm", torch.float64, 2.0), ("zgemm", torch.complex128, 8.0)):
    for n in (2048, 4096):
        a = torch.randn(n, n, dtype=dt, device=dev); b = torch.randn(n, n, dtype=dt, device=dev)
        for _ in range(3): (a @ b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[f"{name}_{n}_tflops"] = round(fl * n**3 / best * 1e-9, 2)
# batched small zgemm (the shape class of the hot path)
for n, B in ((48, 4096), (96, 1024)):
    a = torch.randn(B, n, n, dtype=torch.complex128, device=dev); b = torch.randn(B, n, n, dtype=torch.complex128, device=dev)
    for _ in range(3): torch.bmm(a, b)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = torch.bmm(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[f"zbmm_{n}x{B}_tflops"] = round(8.0 * n**3 * B / best * 1e-9, 2)
print(json.dumps(out))
