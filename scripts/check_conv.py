"""K9 (tcgen05 3x3 conv) against a plain fp32 torch reference, plus timing against cuDNN's fused conv."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from chinesechesszero_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[128, 256, 100, 37])
ap.add_argument("--cg", type=int, nargs="+", default=[1, 2])
ap.add_argument("--time", type=int, default=0, help="boards for the timing section (0 = skip)")
args = ap.parse_args()
torch.manual_seed(0)
cl = torch.channels_last
w = (torch.randn(256, 256, 3, 3, device="cuda") * 0.03).to(torch.bfloat16).contiguous(memory_format=cl)
bias = torch.randn(256, device="cuda") * 0.1
ok = True
for n in args.n:
    x = torch.randn(n, 256, 10, 9, device="cuda").to(torch.bfloat16).contiguous(memory_format=cl)
    skip = torch.randn(n, 256, 10, 9, device="cuda").to(torch.bfloat16).contiguous(memory_format=cl)
    ref0 = F.conv2d(x.float(), w.float(), bias, padding=1)
    for cg in args.cg:
        for sk in (None, skip):
            ref = torch.relu(ref0 + (sk.float() if sk is not None else 0))
            y = _lib.conv3x3_c256(x, w, bias, sk, cta_group=cg)
            torch.cuda.synchronize()
            err = (y.float() - ref).abs().max().item()
            tol = 2e-2 * max(1.0, ref.abs().max().item())
            good = err <= tol
            ok &= good
            print(json.dumps({"n": n, "cg": cg, "skip": sk is not None, "max_abs_err": err, "ref_max": ref.abs().max().item(), "ok": good}), flush=True)
if args.time:
    n = args.time
    x = torch.randn(n, 256, 10, 9, device="cuda").to(torch.bfloat16).contiguous(memory_format=cl)
    skip = torch.randn_like(x)
    out = torch.empty_like(x)
    bb = bias.to(torch.bfloat16)
    flop = 2 * n * 90 * 256 * 2304

    def timed(fn, iters=200):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    res = {}
    res["cudnn_conv_relu"] = timed(lambda: torch.cudnn_convolution_relu(x, w, bb, (1, 1), (1, 1), (1, 1), 1))
    res["cudnn_conv_add_relu"] = timed(lambda: torch.cudnn_convolution_add_relu(x, w, skip, 1.0, bb, (1, 1), (1, 1), (1, 1), 1))
    for cg in args.cg:
        res[f"k9_cg{cg}"] = timed(lambda: _lib.conv3x3_c256(x, w, bias, None, out=out, cta_group=cg))
        res[f"k9_cg{cg}_skip"] = timed(lambda: _lib.conv3x3_c256(x, w, bias, skip, out=out, cta_group=cg))
    print(json.dumps({"n": n, **{k: {"us": v * 1e3, "tflops": flop / v / 1e9} for k, v in res.items()}}), flush=True)
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
