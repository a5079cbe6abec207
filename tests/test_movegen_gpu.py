"""K1 parity: ccz_movegen_encode (CUDA, through the C ABI) vs the CPU oracle, bit-exact.

Covers ordered move-id lists, counts, flag bytes and the bf16 (17,7,10,9) planes on perft
positions, random-playout positions (checks, captures, endgames, clocks, repetition counts), hand
made edge cases, ragged batch sizes (the kernel works in quads of 4) and the planes=NULL variant.
"""
import numpy as np
import pytest
import torch

from oracle import cchess_shim as cs
from tests import positions

pytestmark = pytest.mark.gpu


def _run(boards_np, planes=True):
    from chinesechesszero_b200 import _lib

    boards = torch.from_numpy(np.ascontiguousarray(boards_np)).cuda()
    n = boards.shape[0]
    # poison the outputs: the kernel must write every byte it owns
    ids = torch.full((n, 128), 0x7F7F, dtype=torch.int16, device="cuda")
    counts = torch.full((n,), 0x7F7F, dtype=torch.int16, device="cuda")
    flags = torch.full((n,), 0xFF, dtype=torch.uint8, device="cuda")
    pl = torch.full((n, 17, 7, 10, 9), float("nan"), dtype=torch.bfloat16, device="cuda") if planes else None
    _lib.movegen_encode(boards, planes=planes, out=(ids, counts, flags, pl))
    torch.cuda.synchronize()
    pl_bits = pl.view(torch.int16).cpu().numpy().view(np.uint16).reshape(n, -1) if planes else None
    return ids.cpu().numpy(), counts.cpu().numpy(), flags.cpu().numpy(), pl_bits


def _check(boards_np, planes=True):
    ids, counts, flags, pl = _run(boards_np, planes)
    o_ids, o_counts, o_flags, o_pl = cs.batch_movegen_encode(boards_np, want_planes=planes)
    assert np.array_equal(counts, o_counts)
    assert np.array_equal(flags, o_flags)
    bad = np.nonzero((ids != o_ids).any(axis=1))[0]
    assert bad.size == 0, f"move lists differ at {bad[:5]}: {ids[bad[0]][:48]} vs {o_ids[bad[0]][:48]}"
    if planes:
        assert np.array_equal(pl, o_pl)


def test_edge_cases():
    _check(positions.edge_case_records())


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 31, 33])
def test_ragged_batches(n):
    recs = positions.perft_leaves(2)[:n]
    _check(recs)


def test_empty_batch():
    from chinesechesszero_b200 import _lib

    empty = torch.empty((0, 96), dtype=torch.uint8, device="cuda")
    ids, counts, flags, pl = _lib.movegen_encode(empty)
    assert ids.shape == (0, 128) and pl.shape[0] == 0


def test_perft3_leaves_all():
    recs = positions.perft_leaves(3)
    assert recs.shape[0] == 79666
    _check(recs)


def test_planes_null_variant():
    recs = positions.perft_leaves(2)
    _check(recs, planes=False)


def test_random_playout_positions():
    recs = positions.random_playout_positions(n_games=200, max_plies=300, seed=1234)
    assert recs.shape[0] > 20000
    ids, counts, flags, _ = cs.batch_movegen_encode(recs, want_planes=False)
    # the set must exercise the interesting flags
    assert (flags & cs.FLAG_CHECK).any() and (counts > 60).any()
    _check(recs)


def test_perft4_sample_and_count():
    """Depth-4 leaves: a seeded sample checked bit-exact, and sum(counts) over ALL perft-3 leaves
    equals the published perft(4) = 3,290,240 (a size-independent checksum of the move generator)."""
    recs3 = positions.perft_leaves(3)
    _, counts, _, _ = _run(recs3, planes=False)
    assert int(counts.astype(np.int64).sum()) == 3290240
    leaves4 = positions.perft_leaves(4)
    rng = np.random.default_rng(0)
    pick = rng.choice(leaves4.shape[0], size=200000, replace=False)
    _check(leaves4[np.sort(pick)])


def test_start_boards_kernel():
    from chinesechesszero_b200 import _lib

    b = _lib.boards_start(5).cpu().numpy()
    assert np.array_equal(b, np.tile(cs.start_record(), (5, 1)))


def test_device_perft_counts_and_order():
    """GPU breadth-first expansion (K1 + K2): level sizes are the published perft numbers and the
    perft-3 level equals the oracle's leaves in generation order."""
    from chinesechesszero_b200 import positions as dev_positions

    levels = dev_positions.perft_levels(3)
    assert [lv.shape[0] for lv in levels] == [44, 1920, 79666]
    assert np.array_equal(levels[2].cpu().numpy(), positions.perft_leaves(3))
    assert dev_positions.perft_count(4) == 3290240
    assert dev_positions.perft_count(5) == 133312995


def test_device_perft_published_suite():
    """K1 + K2 on the ten published perft positions (tests/positions.PERFT_SUITE), depths 1-5: the device move
    generator and push reproduce all 50 known answers (53 M leaves for the largest)."""
    from chinesechesszero_b200 import positions as dev_positions

    for fen, expect in positions.PERFT_SUITE.items():
        rec = positions.record_from_fen(fen)
        for depth, want in enumerate(expect, start=1):
            assert dev_positions.perft_count(depth, root=rec) == want, (fen, depth)
