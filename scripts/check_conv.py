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
ap.add_argument("--tcg", type=int, nargs="+", default=None, help="cta_group codes for the timing section")
ap.add_argument("--no-cudnn", action="store_true")
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
            y = _lib.conv3x3_c256(x, w, bias, sk, variant=cg)
            torch.cuda.synchronize()
            err = (y.float() - ref).abs().max().item()
            tol = 2e-2 * max(1.0, ref.abs().max().item())
            good = err <= tol
            ok &= good
            print(json.dumps({"n": n, "cg": cg, "skip": sk is not None, "max_abs_err": err, "ref_max": ref.abs().max().item(), "ok": good}), flush=True)
if args.time:
    # The part is power-capped: clocks sag over hundreds of milliseconds of dense MMA work, so every variant
    # gets its own long warm-up, a long timed run, and the list is run forwards and backwards.
    n = args.time
    x = (torch.randn(n, 256, 10, 9, device="cuda").relu_()).to(torch.bfloat16).contiguous(memory_format=cl)  # post-ReLU like the tower
    skip = torch.randn_like(x).relu_()
    out = torch.empty_like(x)
    bb = bias.to(torch.bfloat16)
    flop = 2 * n * 90 * 256 * 2304

    def timed(fn, iters=1500, warm=400):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    variants = {
        "cudnn_conv_relu": lambda: torch.cudnn_convolution_relu(x, w, bb, (1, 1), (1, 1), (1, 1), 1),
        "cudnn_conv_add_relu": lambda: torch.cudnn_convolution_add_relu(x, w, skip, 1.0, bb, (1, 1), (1, 1), (1, 1), 1),
    }
    if args.no_cudnn:
        variants = {}
    for cg in (args.tcg or args.cg):
        variants[f"k9_cg{cg}"] = (lambda cg=cg: _lib.conv3x3_c256(x, w, bias, None, out=out, variant=cg))
        variants[f"k9_cg{cg}_skip"] = (lambda cg=cg: _lib.conv3x3_c256(x, w, bias, skip, out=out, variant=cg))
    names = list(variants)
    res = {k: [] for k in names}
    for order in (names, names[::-1]):
        for k in order:
            res[k].append(timed(variants[k]))
    print(json.dumps({"n": n, **{k: {"us": [round(t * 1e3, 1) for t in v], "tflops": round(flop / min(v) / 1e9, 1)} for k, v in res.items()}}), flush=True)
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
