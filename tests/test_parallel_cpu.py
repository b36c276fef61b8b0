"""CPU: the N>1 host logic (clip sharding, max-over-ranks timing, result gathering) with world_size 2 on gloo."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
from vps_b200 import parallel as P
rank, local, world = P.init(backend="gloo")
clips = list(range(7))
mine = P.shard_clips(clips, rank, world)
t = P.max_over_ranks([1.0 + rank, 5.0 - rank])
allc = P.gather_objects(mine)
P.barrier()
print(json.dumps({"rank": rank, "world": world, "mine": mine, "t": t, "all": allc}))
'''


def test_two_rank_gloo_sharding_and_reduction(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        outs.append(__import__("json").loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["mine"] == [0, 2, 4, 6] and outs[1]["mine"] == [1, 3, 5]
    assert sorted(outs[0]["mine"] + outs[1]["mine"]) == list(range(7))          # disjoint cover
    assert outs[0]["t"] == outs[1]["t"] == [2.0, 5.0]                            # max over ranks
    assert outs[0]["all"] == outs[1]["all"] == [[0, 2, 4, 6], [1, 3, 5]]


def test_single_process_is_a_noop():
    from vps_b200 import parallel as P
    assert P.shard_clips([3, 4, 5], 0, 1) == [3, 4, 5]
    assert P.max_over_ranks([1.5]) == [1.5]
    assert P.gather_objects("x") == ["x"]
