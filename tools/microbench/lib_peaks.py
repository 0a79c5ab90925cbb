"""Library FP64 peaks on the box (cuBLAS DGEMM, cuSOLVER batched potrf) via torch. Prints JSON lines."""
import json, time, torch
def ev(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = ev(lambda: torch.matmul(a, b)); out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms / 1e9
    del a, b
a = torch.randn(16, 3000, 3000, dtype=torch.float64, device="cuda"); b = torch.randn(16, 3000, 3000, dtype=torch.float64, device="cuda")
ms = ev(lambda: torch.bmm(a, b)); out["dgemm_batched16_3000_tflops"] = 16 * 2 * 3000**3 / ms / 1e9
ms = ev(lambda: torch.bmm(a, b.transpose(1, 2))); out["dgemm_batched16_3000_nt_tflops"] = 16 * 2 * 3000**3 / ms / 1e9
spd = torch.bmm(a, a.transpose(1, 2)) + 3000 * torch.eye(3000, dtype=torch.float64, device="cuda")
ms = ev(lambda: torch.linalg.cholesky(spd), reps=3); out["potrf_batched16_3000_tflops"] = 16 * 3000**3 / 3 / ms / 1e9
ms = ev(lambda: torch.linalg.cholesky(spd[0]), reps=3); out["potrf_single_3000_tflops"] = 3000**3 / 3 / ms / 1e9
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
ms = ev(lambda: y.copy_(x)); out["copy_gbs"] = 2 * x.numel() * 8 / ms / 1e6
print(json.dumps(out))
