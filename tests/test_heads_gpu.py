"""K11 ``heads_pack_kernel`` (csrc/ccz_heads.cuh): ReLU + NHWC -> channel-major packing of the 1x1 head outputs into the
K-padded FC operands, against the torch expression it replaces (net.py:96-97,103-104) -- exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 3, 257, 4096])
def test_heads_pack_equals_relu_transpose(n):
    from chinesechesszero_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(n)
    h = torch.randn(n * 90, 32, device="cuda", generator=g).to(torch.bfloat16)
    kp, kv = 1536, 640
    buf = torch.full((n, kp + kv), 7.0, dtype=torch.bfloat16, device="cuda")  # pad columns must stay untouched
    _lib.heads_pack(h, buf, kp)
    ref = torch.full_like(buf, 7.0)
    hr = torch.relu(h).view(n, 90, 32)
    ref[:, :17 * 90] = hr[:, :, :17].transpose(1, 2).reshape(n, 17 * 90)          # x.view(-1, 17*90) of an NCHW tensor
    ref[:, kp:kp + 7 * 90] = hr[:, :, 17:24].transpose(1, 2).reshape(n, 7 * 90)
    assert torch.equal(buf, ref)


def test_heads_pack_argument_checks():
    from chinesechesszero_b200 import _lib

    h = torch.zeros(90, 32, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.CczError):
        _lib.heads_pack(h, torch.zeros(1, 2000, dtype=torch.bfloat16, device="cuda"), 1536)   # row too short for the value operand
    with pytest.raises(_lib.CczError):
        _lib.heads_pack(h.float(), torch.zeros(1, 2176, dtype=torch.bfloat16, device="cuda"), 1536)
    with pytest.raises(_lib.CczError):
        _lib.heads_pack(h, torch.zeros(1, 2176, dtype=torch.bfloat16, device="cuda"), 1000)  # value operand overlaps the policy operand
