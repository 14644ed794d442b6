"""N>1 host logic on CPU: world_size-2 gloo process group (sharding, max-over-ranks timing, result gather) and the
`bench.py --impl reference` contract under torchrun-style environments (rank 0 prints, other ranks exit 0)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import horopose_b200  # noqa: F401
    from horopose_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 13
    b, e = shard.shard_range(total, rank, world)
    local = torch.arange(b, e, dtype=torch.float32).view(-1, 1) * 2.0   # "results" of the images this rank owns
    gathered = shard.gather_results(local)
    slowest = shard.max_over_ranks(10.0 + rank)
    q.put((rank, b, e, gathered.flatten().tolist(), slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, e0, g0, s0), (r1, b1, e1, g1, s1) = res
    assert (b0, e0, b1, e1) == (0, 7, 7, 13)              # contiguous, sizes differ by <= 1, full coverage
    assert g0 == g1 == [2.0 * i for i in range(13)]        # every rank sees all results in image order
    assert s0 == s1 == 11.0                                # slowest rank's time


def test_shard_range_properties():
    import horopose_b200  # noqa: F401
    from horopose_b200.shard import shard_range
    for total in (0, 1, 7, 512, 2048):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_reference_arm_contract_under_two_ranks():
    """`bench.py --impl reference`: rank 0 prints ONE JSON line with the contract keys, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999", OMP_NUM_THREADS="8")
    out = {}
    for rank in (0, 1):
        env.update(RANK=str(rank), LOCAL_RANK=str(rank))
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "1"], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[rank] = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert out[1] == []
    assert len(out[0]) == 1
    line = json.loads(out[0][0])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("kuka_full")
