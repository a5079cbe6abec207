"""Network parity.  CPU: the fp32 oracle and the product ``Net`` reproduce the UNMODIFIED reference
net.Net outputs stored in tests/golden/net_reference.npz; BN folding is exact in fp32.  GPU: the bf16
BatchedEvaluator stays within the north-star tolerance (1e-2 abs on probabilities and value)."""
import os

import numpy as np
import pytest
import torch

from oracle import net_oracle


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "net_reference.npz"))


def _product_net(perturbed: bool):
    from chinesechesszero_b200.net import Net

    torch.manual_seed(0)
    net = Net()
    if perturbed:
        net_oracle.perturb_(net.state_dict(), seed=1)
    return net.eval()


@pytest.mark.parametrize("name", ["seed0", "perturbed"])
def test_oracle_and_product_net_match_reference_golden(gold, name):
    net = _product_net(name == "perturbed")
    x = net_oracle.search_planes(gold["records"])
    logp_o, v_o = net_oracle.forward(net.state_dict(), x)
    # same machine -> bit-equal; other CPUs may reorder fp32 sums slightly
    assert np.allclose(logp_o.numpy(), gold[f"{name}_logp"], atol=2e-5)
    assert np.allclose(v_o.numpy(), gold[f"{name}_value"], atol=2e-5)
    with torch.no_grad():
        logp_p, v_p = net(x)
    assert np.allclose(logp_p.numpy(), gold[f"{name}_logp"], atol=2e-5)
    assert np.allclose(v_p.numpy(), gold[f"{name}_value"], atol=2e-5)


def test_state_dict_keys_are_the_reference_ones():
    net = _product_net(False)
    keys = list(net.state_dict().keys())
    assert keys[0] == "conv_block.weight" and "res_blocks.39.conv2_bn.running_var" in keys
    assert "policy_fc.weight" in keys and "value_fc2.bias" in keys
    assert sum(p.numel() for p in net.parameters()) == 50_883_979


def test_bn_folding_is_exact_in_fp32_on_cpu():
    from chinesechesszero_b200.net import BatchedEvaluator, Net

    torch.manual_seed(3)
    net = Net(num_channels=32, resblocks_num=3)
    net_oracle.perturb_(net.state_dict(), seed=9)
    net.eval()
    ev = BatchedEvaluator(net, device="cpu", dtype=torch.float32, fused_epilogue=False)
    x = torch.rand(5, 17, 7, 10, 9)
    logits, v = ev.forward(x)
    logp_o, v_o = net_oracle.forward(net.state_dict(), x)
    assert torch.allclose(torch.log_softmax(logits, 1), logp_o, atol=1e-4)
    assert torch.allclose(v, v_o.view(-1), atol=1e-5)


def test_refresh_is_in_place_on_cpu():
    """refresh() must not rebind the folded tensors (captured CUDA graphs and other owners of
    the evaluator hold their addresses); the outputs must follow the new weights."""
    from chinesechesszero_b200.net import BatchedEvaluator, Net

    torch.manual_seed(4)
    net = Net(num_channels=32, resblocks_num=2).eval()
    ev = BatchedEvaluator(net, device="cpu", dtype=torch.float32, fused_epilogue=False)
    ptrs = {k: t.data_ptr() for k, t in ev._params.items()}
    blocks_before = [[t.data_ptr() for t in c] for pair in ev.blocks for c in pair]
    x = torch.rand(3, 17, 7, 10, 9)
    before = ev.forward(x)
    net_oracle.perturb_(net.state_dict(), seed=11)
    ev.refresh(net)
    assert ev.version == 2
    assert {k: t.data_ptr() for k, t in ev._params.items()} == ptrs
    assert [[t.data_ptr() for t in c] for pair in ev.blocks for c in pair] == blocks_before
    logits, v = ev.forward(x)
    logp_o, v_o = net_oracle.forward(net.state_dict(), x)
    assert torch.allclose(torch.log_softmax(logits, 1), logp_o, atol=1e-4) and torch.allclose(v, v_o.view(-1), atol=1e-5)
    assert not torch.allclose(logits, before[0])
    # the evaluator does not alias the module's parameters: training the net does not leak in before refresh()
    with torch.no_grad():
        net.value_fc2.weight.add_(1.0)
    assert torch.equal(ev.forward(x)[1], v)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["seed0", "perturbed"])
@pytest.mark.parametrize("conv_impl", ["k9", "k9_skip", "cudnn", "torch"])
def test_bf16_evaluator_within_tolerance(gold, name, conv_impl):
    from chinesechesszero_b200 import _lib
    from chinesechesszero_b200.net import BatchedEvaluator

    net = _product_net(name == "perturbed").cuda()
    ev = BatchedEvaluator(net, conv_impl=conv_impl)
    assert ev.conv_impl == conv_impl  # the 256-channel product net takes the hand-written tcgen05 path
    boards = torch.from_numpy(gold["records"]).cuda()
    _, _, _, planes = _lib.movegen_encode(boards)
    logits, v = ev.forward(planes)
    probs = torch.softmax(logits, 1).cpu().numpy()
    ref_p = np.exp(gold[f"{name}_logp"])
    # north-star tolerance: 1e-2 abs on policy and value in bf16
    assert np.abs(probs - ref_p).max() < 1e-2
    assert np.abs(v.cpu().numpy() - gold[f"{name}_value"].reshape(-1)).max() < 1e-2
    # tighter sanity bound so the check is not vacuous for near-uniform policies
    logp = torch.log_softmax(logits, 1).cpu().numpy()
    assert np.abs(logp - gold[f"{name}_logp"]).max() < 5e-2


def test_stem_tables_reproduce_the_stem_convolution_on_cpu():
    """K10's lookup tables (net.stem_tables): bias_turn[turn][border class] + the table rows of the occupied
    3x3 neighbourhood equal conv_block + BN on the search-time planes (net.py:84, 160-177) -- host arithmetic only."""
    import torch.nn.functional as F

    from chinesechesszero_b200.net import Net, _fold, stem_tables
    from oracle import cchess_shim as cs
    from tests.positions import random_playout_positions

    torch.manual_seed(5)
    net = Net(num_channels=256, resblocks_num=0)
    net_oracle.perturb_(net.state_dict(), seed=2)
    net.eval()
    w, b = _fold(net.conv_block, net.conv_block_bn)
    w = w.to(torch.bfloat16)  # the evaluator's operand precision
    table, bias_turn = stem_tables(w, b)
    table, bias_turn = table.float().numpy(), bias_turn.numpy()
    recs = random_playout_positions(2, 40, seed=1, every=3)[:24]
    planes = cs.batch_movegen_encode(recs)[3]  # bf16 bit patterns (n, 10710)
    x = torch.from_numpy((planes.astype(np.uint32) << 16).view(np.float32).reshape(-1, 119, 10, 9))
    ref = F.conv2d(x, w.float(), b, padding=1).numpy()
    assert set(recs[:, 90].tolist()) == {0, 1}
    for i, rec in enumerate(recs):
        turn = int(rec[90] != 0)
        for h in range(10):
            for wc in range(9):
                cls = (0 if h == 0 else 2 if h == 9 else 1) * 3 + (0 if wc == 0 else 2 if wc == 8 else 1)
                acc = bias_turn[turn, cls].copy()
                for r in range(3):
                    for s in range(3):
                        nh, nw = h + r - 1, wc + s - 1
                        if 0 <= nh <= 9 and 0 <= nw <= 8:
                            acc += table[r * 3 + s, int(rec[nh * 9 + nw]) & 15]
                assert np.abs(acc - ref[i, :, h, wc]).max() < 2e-5
