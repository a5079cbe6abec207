"""configs[4] train-step timing alone (bench.py's `train` section), optionally with a channels_last training net."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

for cl in ([False, True] if len(sys.argv) < 2 else [bool(int(sys.argv[1]))]):
    os.environ["CCZ_TRAIN_CHANNELS_LAST"] = "1" if cl else "0"
    r = bench.bench_train(torch)
    print(json.dumps({"channels_last": cl, **{k: r[k] for k in ("value", "ms_per_step", "loss", "kl")}}), flush=True)
    torch.cuda.empty_cache()
