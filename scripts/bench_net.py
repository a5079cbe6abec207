"""Quick look at the bf16 forward throughput of the 40x256 net at lockstep batch sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechesszero_b200.net import Net, BatchedEvaluator, FLOP_PER_POSITION

torch.manual_seed(0)
net = Net().cuda().eval()
for bench_mode in (False, True):
    torch.backends.cudnn.benchmark = bench_mode
    ev = BatchedEvaluator(net, fused_epilogue=True)
    for g in (2048, 4096, 8192):
        x = (torch.rand(g, 17, 7, 10, 9, device="cuda") > 0.9).to(torch.bfloat16)
        for _ in range(3):
            ev.forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ev.forward(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(json.dumps({"cudnn_benchmark": bench_mode, "batch": g, "ms": ms, "tflops": g * FLOP_PER_POSITION / ms / 1e9,
                          "pos_per_s": g / ms * 1e3}))
