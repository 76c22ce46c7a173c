"""raw tcgen05 GEMM throughput (asn_gemm_bf16_tn) on large square problems: how far is the mainloop from the
tensor roofline when nothing else (epilogue share, tile quantisation, tiny K) is in the way?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptsegnet_b200 import ops, prof

res = {}
shapes = [(4096, 4096, 4096), (8192, 8192, 4096), (14400, 688, 2048), (14400, 688, 1024), (8192, 688, 2048), (8192, 688, 1024),
          (2048, 14400, 688), (1024, 8192, 688), (8192, 256, 8192), (8192, 128, 8192)]
for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    ops.gemm_bf16_tn(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_bf16_tn(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = None
    e0.record()
    for _ in range(10):
        ref = a @ b.t()
    e1.record()
    torch.cuda.synchronize()
    ms_cublas = e0.elapsed_time(e1) / 10
    res[f"{M}x{N}x{K}"] = {"asn_ms": round(ms, 4), "asn_tflops": round(2 * M * N * K / ms / 1e9, 1),
                           "cublas_bf16_ms": round(ms_cublas, 4), "cublas_tflops": round(2 * M * N * K / ms_cublas / 1e9, 1)}
print(json.dumps(res, indent=1))
