"""Forward-pass experiments on the 40x256 tower (one GPU): steady-state time of the two fused
conv flavours, whole-forward time when the 4096-leaf batch is evaluated in L2-sized chunks
(with / without CUDA-graph replay), cudnn.benchmark on/off."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechesszero_b200.net import Net, BatchedEvaluator, FLOP_PER_POSITION

torch.manual_seed(0)
net = Net().cuda().eval()
CONV_FLOP = 2 * 90 * 256 * 2304


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for bench_mode in (False, True):
    torch.backends.cudnn.benchmark = bench_mode
    ev = BatchedEvaluator(net)
    for g in (512, 1024, 4096):
        x = torch.randn(g, 256, 10, 9, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        s = torch.randn_like(x)
        c1, c2 = ev.blocks[0]
        n = 400 * 4096 // g
        t1 = timed(lambda: ev._conv_relu(x, c1), n)
        t2 = timed(lambda: ev._conv_add_relu(x, c2, s), n)
        print(json.dumps({"cudnn_benchmark": bench_mode, "g": g, "conv_relu_us": t1 * 1e3, "conv_relu_tflops": g * CONV_FLOP / t1 / 1e9,
                          "conv_add_relu_us": t2 * 1e3, "conv_add_relu_tflops": g * CONV_FLOP / t2 / 1e9}), flush=True)

torch.backends.cudnn.benchmark = False
ev = BatchedEvaluator(net)
G = 4096
planes = (torch.rand(G, 17, 7, 10, 9, device="cuda") > 0.9).to(torch.bfloat16)
for chunk in (4096, 2048, 1024, 512, 256):
    def fwd():
        outs = [ev.forward(planes[i:i + chunk]) for i in range(0, G, chunk)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
    ms_eager = timed(fwd, 12)
    # graph replay
    sstream = torch.cuda.Stream()
    with torch.cuda.stream(sstream):
        for _ in range(2):
            fwd()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = fwd()
    ms_graph = timed(gr.replay, 12)
    print(json.dumps({"chunk": chunk, "ms_eager": ms_eager, "ms_graph": ms_graph,
                      "tflops_graph": G * FLOP_PER_POSITION / ms_graph / 1e9, "tflops_eager": G * FLOP_PER_POSITION / ms_eager / 1e9}), flush=True)
