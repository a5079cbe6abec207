"""Sharded self-play collection under torchrun: every rank plays its own games on its own GPU and
writes `<data_dir>/rank{r}/data.h5` (+ npy triple) with disjoint game numbers; rank 0 then merges the
shards into `<data_dir>/data.h5` in the reference layout.

  python -m torch.distributed.run --nproc-per-node N scripts/collect_multi_gpu.py --games 64 --playouts 16 ...
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from chinesechesszero_b200 import distributed as D
from chinesechesszero_b200 import h5lite
from chinesechesszero_b200.collect import CollectPipeline, merge_h5_shards

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=64)
ap.add_argument("--playouts", type=int, default=16)
ap.add_argument("--max-games", type=int, default=64)
ap.add_argument("--max-game-moves", type=int, default=6)
ap.add_argument("--data-dir", default="gpurun_out/collect_demo")
ap.add_argument("--channels", type=int, default=32)
ap.add_argument("--blocks", type=int, default=2)
args = ap.parse_args()

rank, local_rank, world = D.shard_info()
torch.cuda.set_device(local_rank)
D.init("gloo")  # only a barrier is needed: no collective on the data path
torch.manual_seed(0)
pipe = CollectPipeline(n_games=args.games, n_playout=args.playouts, data_dir=args.data_dir, rank=rank, world=world,
                       max_game_moves=args.max_game_moves, node_cap=8192, states_mode="per_move",
                       net_kwargs=dict(num_channels=args.channels, resblocks_num=args.blocks))
n = pipe.run(max_games=args.max_games)
D.barrier()
if rank == 0:
    shards = [os.path.join(args.data_dir, f"rank{r}", "data.h5") if world > 1 else os.path.join(args.data_dir, "data.h5")
              for r in range(world)]
    if world > 1:
        total = merge_h5_shards(shards, os.path.join(args.data_dir, "data.h5"))
    else:
        total = n
    with h5lite.H5Reader(os.path.join(args.data_dir, "data.h5")) as r:
        g0 = r.read_group("game_0")
        print(json.dumps({"world": world, "games_merged": int(total), "iters_attr": int(r.root_attrs()["iters"]),
                          "game_0_states": list(g0["states"].shape), "game_0_probs": list(g0["mcts_probs"].shape)}))
D.barrier()
D.shutdown()
