"""Short launch sequence for ncu: cuDNN fused convs and K9 (both CTA-group variants, with/without skip) at n boards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chinesechesszero_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cgs = [int(a) for a in sys.argv[3:]] or [2]
torch.manual_seed(0)
cl = torch.channels_last
w = (torch.randn(256, 256, 3, 3, device="cuda") * 0.03).to(torch.bfloat16).contiguous(memory_format=cl)
bias = torch.randn(256, device="cuda") * 0.1
bb = bias.to(torch.bfloat16)
x = torch.randn(n, 256, 10, 9, device="cuda").to(torch.bfloat16).contiguous(memory_format=cl)
skip = torch.randn_like(x)
out = torch.empty_like(x)
for _ in range(reps):
    torch.cudnn_convolution_relu(x, w, bb, (1, 1), (1, 1), (1, 1), 1)
    torch.cudnn_convolution_add_relu(x, w, skip, 1.0, bb, (1, 1), (1, 1), (1, 1), 1)
    for cg in cgs:
        _lib.conv3x3_c256(x, w, bias, None, out=out, variant=cg)
        _lib.conv3x3_c256(x, w, bias, skip, out=out, variant=cg)
torch.cuda.synchronize()
print("ok")
