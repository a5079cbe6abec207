"""net_oracle -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; never imported by the product path).

Functional fp32 restatement of the reference network (/root/reference/net.py:15-110) over a plain
``state_dict`` with the reference's key names: stem conv3x3(119->C)+BN+ReLU, R residual blocks
(conv-BN-ReLU-conv-BN, skip, ReLU; net.py:33-41), policy head conv1x1->17, BN, ReLU, FC->2086,
log_softmax (net.py:94-99), value head conv1x1->7, BN, ReLU, FC->256, ReLU, FC->1, tanh
(net.py:101-108).  Pinned to the UNMODIFIED reference ``net.Net`` by tests/golden/net_reference.npz
(scripts/golden_net.py).  Also the fp32 forward used by the CPU baseline leg of bench.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=1e-5)


@torch.no_grad()
def forward(sd: dict, x: torch.Tensor):
    """x (N,17,7,10,9) float32 -> (log_probs (N,2086), value (N,1)); eval-mode BatchNorm."""
    n = x.shape[0]
    x = x.reshape(n, -1, 10, 9).float()
    x = F.relu(_bn(F.conv2d(x, sd["conv_block.weight"], sd["conv_block.bias"], padding=1), sd, "conv_block_bn"))
    i = 0
    while f"res_blocks.{i}.conv1.weight" in sd:
        p = f"res_blocks.{i}."
        y = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1), sd, p + "conv1_bn"))
        y = _bn(F.conv2d(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1), sd, p + "conv2_bn")
        x = F.relu(x + y)
        i += 1
    pol = F.relu(_bn(F.conv2d(x, sd["policy_conv.weight"], sd["policy_conv.bias"]), sd, "policy_bn"))
    pol = F.linear(pol.reshape(n, -1), sd["policy_fc.weight"], sd["policy_fc.bias"])
    val = F.relu(_bn(F.conv2d(x, sd["value_conv.weight"], sd["value_conv.bias"]), sd, "value_bn"))
    val = F.relu(F.linear(val.reshape(n, -1), sd["value_fc1.weight"], sd["value_fc1.bias"]))
    val = torch.tanh(F.linear(val, sd["value_fc2.weight"], sd["value_fc2.bias"]))
    return F.log_softmax(pol, dim=1), val


def perturb_(sd: dict, seed: int = 1) -> dict:
    """Deterministically give every BatchNorm non-trivial statistics / affine terms so that BN
    folding is actually exercised (a fresh net has mean 0, var 1, gamma 1, beta 0)."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd):
        t = sd[k]
        if k.endswith("running_mean") or (k.endswith(".bias") and k.replace(".bias", ".running_var") in sd):
            t.copy_(torch.randn(t.shape, generator=g) * 0.05)
        elif k.endswith("running_var"):
            t.copy_(1.0 + 0.2 * torch.rand(t.shape, generator=g))
        elif k.endswith(".weight") and k.replace(".weight", ".running_var") in sd:
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
    return sd


def search_planes(records: np.ndarray) -> torch.Tensor:
    """(n,17,7,10,9) float32 search-time input of board records via the C oracle (net.py:160-177)."""
    from . import cchess_shim as cs

    _, _, _, planes = cs.batch_movegen_encode(records)
    f = (planes.astype(np.uint32) << 16).view(np.float32)
    return torch.from_numpy(f.reshape(-1, 17, 7, 10, 9).copy())
