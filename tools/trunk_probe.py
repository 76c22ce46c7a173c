"""How the (unchanged) ResNet-101 trunk behaves on B200 under different execution modes:
NCHW vs channels_last, TF32 vs bf16 autocast, eager vs CUDA graph.  Times trunk fwd+bwd only."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti

torch.backends.cudnn.benchmark = True
dev = "cuda"
res = {}
for cl in (False, True):
    for amp in (False, True):
        torch.manual_seed(0)
        m = DeeplabMulti(19).to(dev).train()
        x = torch.randn(1, 3, 720, 1280, device=dev)
        if cl:
            m = m.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                f3, f4 = m.trunk(x)
            (f3.float().mean() + f4.float().mean()).backward()

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        t_cpu = (time.perf_counter() - t0) / 5 * 1e3
        torch.cuda.synchronize()
        r = {"eager_gpu_ms": e0.elapsed_time(e1) / 5, "cpu_issue_ms": t_cpu}
        for p in m.parameters():
            p.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        r["graph_ms"] = e0.elapsed_time(e1) / 5
        res[f"cl={cl},bf16={amp}"] = r
        del m, x, g
        torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
