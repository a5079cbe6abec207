"""N>1 host logic on CPU: two gloo ranks check seed / game-index sharding and the whole-job
throughput aggregation (sum of units over the max of the ranks' times) used by bench.py; the
reference arm of bench.py lets only rank 0 work."""
import json
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORKER = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, %r)
    from chinesechesszero_b200 import distributed as D
    rank, local_rank, world = D.init("gloo")
    seed = D.rank_seed(1234, rank)
    games = [D.global_game_index(i, rank, world, start=10) for i in range(4)]
    D.barrier()
    # rank r "plays" 100*(r+1) moves in 50*(r+1) ms
    agg = D.aggregate_throughput(100.0 * (rank + 1), 50.0 * (rank + 1))
    worst = D.max_over_ranks(50.0 * (rank + 1))
    print(json.dumps({"rank": rank, "world": world, "seed": seed, "games": games, "agg": agg, "worst": worst}))
    D.shutdown()
""") % ROOT


def test_two_gloo_ranks_shard_and_aggregate(tmp_path):
    port = _free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        outs.append(json.loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda o: o["rank"])
    assert [o["world"] for o in outs] == [2, 2]
    assert outs[0]["seed"] != outs[1]["seed"]
    g0, g1 = outs[0]["games"], outs[1]["games"]
    assert g0 == [10, 12, 14, 16] and g1 == [11, 13, 15, 17] and not set(g0) & set(g1)
    # 300 moves over max(50, 100) ms = 3000 moves/s on both ranks
    assert all(abs(o["agg"] - 3000.0) < 1e-9 and o["worst"] == 100.0 for o in outs)


def test_reference_arm_only_rank0_works():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


import pytest


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_prints_contract_line(kind, tmp_path):
    """--impl reference: the UNMODIFIED reference modules when a copy is reachable (/root/reference here, the
    oracle/_ref snapshot on the GPU box), else the oracle port; --warmup is honoured either way."""
    from oracle import load_reference

    if kind == "reference" and not load_reference.available():
        pytest.skip("no reference copy reachable")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    if kind == "port":
        env["CCZ_REFERENCE_DIR"] = str(tmp_path / "nowhere")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--playouts", "6"], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "moves/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == kind and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 1


def test_reference_snapshot_recipe(tmp_path):
    """oracle/make_ref.py copies exactly the reference's self-play modules, unmodified, into the git-ignored
    oracle/_ref; the snapshot alone is enough to drive the reference arm."""
    from oracle import make_ref

    if not os.path.isfile(os.path.join(make_ref.SRC, "mcts.py")):
        pytest.skip("reference sources not present (GPU box)")
    out = make_ref.ensure_ref()
    import filecmp

    assert sorted(os.listdir(out)) == sorted(set(make_ref.MODULES) | ({"__pycache__"} & set(os.listdir(out))))
    for name in make_ref.MODULES:
        assert filecmp.cmp(os.path.join(make_ref.SRC, name), os.path.join(out, name), shallow=False)
    ignored = subprocess.run(["git", "check-ignore", "oracle/_ref/mcts.py"], cwd=ROOT, capture_output=True, text=True)
    assert ignored.returncode == 0
    code = ("import sys; sys.path.insert(0, %r); from oracle import reference_arm as r; "
            "sp = r.ReferenceSelfPlay(n_playout=4, threads=2); print(sp.play_move()[0] >= 0)" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CCZ_REFERENCE_DIR=out, CUDA_VISIBLE_DEVICES=""),
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("True"), p.stderr[-2000:]
