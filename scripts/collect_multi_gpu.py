"""BASELINE configs[3] end to end: sharded self-play collection under torchrun.  Every rank plays its own
games on its own GPU through ``CollectPipeline.collect_data()`` and writes ``<data_dir>/rank{r}/data.h5``
with disjoint game numbers; rank 0 then merges the shards into ``<data_dir>/data.h5`` and checks the
merged file against the reference's layout (collect.py:146-167).  No collective on the data path (gloo
for the barrier and the max-over-ranks time only).

  # configs[3]: 8 GPUs x 8192 games x 800 playouts
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \\
      scripts/collect_multi_gpu.py --games 8192 --playouts 800 --moves 4 --warmup 1 --max-game-moves 4
  # toy run (32-channel net): --channels 32 --blocks 2 --games 64 --playouts 16
"""
import argparse
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from chinesechesszero_b200 import distributed as D
from chinesechesszero_b200 import h5lite
from chinesechesszero_b200.collect import CollectPipeline, merge_h5_shards
from chinesechesszero_b200.net import BatchedEvaluator, Net

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=8192, help="concurrent games per GPU")
ap.add_argument("--playouts", type=int, default=800)
ap.add_argument("--moves", type=int, default=4, help="timed lockstep moves")
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--warmup-playouts", type=int, default=None, help="playouts of the warm-up moves (default: --playouts)")
ap.add_argument("--max-game-moves", type=int, default=4, help="games are cut after this many moves; slots start staggered")
ap.add_argument("--data-dir", default="/tmp/ccz_config4")
ap.add_argument("--channels", type=int, default=256)
ap.add_argument("--blocks", type=int, default=40)
ap.add_argument("--keep", action="store_true", help="keep the replay files (default: delete them after the check)")
ap.add_argument("--out", default=None, help="also write the JSON line to this file")
args = ap.parse_args()

rank, local_rank, world = D.shard_info()
torch.cuda.set_device(local_rank)
D.init("gloo")
torch.manual_seed(0)
net = Net(num_channels=args.channels, resblocks_num=args.blocks).cuda().eval()
ev = BatchedEvaluator(net)
if rank == 0 and os.path.isdir(args.data_dir):
    shutil.rmtree(args.data_dir)
D.barrier()
pipe = CollectPipeline(n_games=args.games, n_playout=args.playouts, data_dir=args.data_dir, rank=rank, world=world,
                       max_game_moves=args.max_game_moves, states_mode="per_move", write_npy=False, evaluator=ev,
                       seed=1234)
pipe.load_model()
eng = pipe.engine
eng.move_count[:] = np.arange(args.games) % args.max_game_moves
if args.warmup_playouts:
    eng.n_playout = args.warmup_playouts   # first-call costs (module load, cuBLAS handles, TMA descriptors, writer) only
for _ in range(args.warmup):
    pipe.collect_data()
eng.n_playout = args.playouts
pipe.flush()
torch.cuda.synchronize()
D.barrier()
games0 = pipe.iters
t0 = time.perf_counter()
for _ in range(args.moves):
    pipe.collect_data()
pipe.flush()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3
worst = D.max_over_ranks(ms)
games = D.sum_over_ranks(pipe.iters - games0)
pool = eng.pool_events()
dropped = D.sum_over_ranks(pool["trees_dropped"] + pool["expand_failed"])
w = pipe.writer
shard_stats = {"samples": w.samples, "raw_bytes": w.raw_bytes, "busy_s": w.busy_seconds}
pipe.close()
D.barrier()
if rank == 0:
    shards = [os.path.join(args.data_dir, f"rank{r}", "data.h5") if world > 1 else os.path.join(args.data_dir, "data.h5")
              for r in range(world)]
    shard_mb = sum(os.path.getsize(p) for p in shards) / 1e6
    t1 = time.perf_counter()
    merged = os.path.join(args.data_dir, "data.h5")
    total = merge_h5_shards(shards, merged) if world > 1 else pipe.iters
    merge_s = time.perf_counter() - t1
    # the reference's reader walks game_0 .. game_{iters-1} (convert.py:38-81)
    with h5lite.H5Reader(merged) as r:
        iters = int(r.root_attrs()["iters"])
        links = r.root_links()
        assert sorted(links) == sorted(f"game_{k}" for k in range(iters)), "game numbering has holes"
        rows = 0
        rng = np.random.default_rng(0)
        for k in [0, iters - 1] + rng.integers(0, iters, size=min(64, iters)).tolist():
            g = r.read_group(f"game_{k}")
            st, pi, z = g["states"], g["mcts_probs"], g["winners"]
            t = st.shape[0] // 2
            assert st.dtype == np.float16 and st.shape == (2 * t, 17, 7, 10, 9)
            assert pi.dtype == np.float64 and pi.shape == (2 * t, 2086) and z.dtype == np.float64 and z.shape == (2 * t,)
            assert np.allclose(pi.sum(1), 1.0, atol=1e-9) and np.array_equal(z[:t], z[t:])
            # second half = file mirror of the first (collect.py:115-131)
            assert np.array_equal(st[t:], st[:t, :, :, :, ::-1])
            assert set(np.unique(st)) <= {0.0, 1.0}
            rows += 2 * t
    line = {
        "workload": f"configs[3]: {world} GPU(s) x {args.games} games x {args.playouts} playouts, per-rank data.h5 shards merged "
                    "into the reference layout", "n_gpus": world, "games_per_gpu": args.games, "n_playout": args.playouts,
        "net": f"{args.blocks}x{args.channels}", "moves_timed": args.moves, "warmup": args.warmup,
        "warmup_playouts": args.warmup_playouts or args.playouts,
        "max_game_moves": args.max_game_moves, "ms_per_move": worst / args.moves,
        "value": world * args.games * args.moves / worst * 1e3, "unit": "moves/s",
        "timing": "host clock around collect_data() x moves + flush (replay compressed, written, indexed), max over ranks",
        "games_written": int(games), "games_merged": int(total), "iters_attr": iters, "shards_mb": shard_mb,
        "merged_mb": os.path.getsize(merged) / 1e6, "merge_seconds": merge_s, "rows_checked": rows,
        "rank0_writer": shard_stats, "pool_events_all_ranks": int(dropped), "pool_rank0": pool,
    }
    print(json.dumps(line))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(json.dumps(line) + "\n")
    if not args.keep:
        shutil.rmtree(args.data_dir, ignore_errors=True)
D.barrier()
D.shutdown()
