"""world_size-2 gloo tests (CPU) of the replica / batch-split plumbing used by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    from gpt4roi_b200 import dist_utils as D
    w, r, _ = D.init('gloo')
    assert (w, r) == (world, rank)
    B, L = 5, 7
    ids = torch.arange(B * L).view(B, L)
    images = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 3, 2, 2).contiguous()
    boxes = [torch.full((i + 1, 4), float(i)) for i in range(B)]
    lids, limg, lbb, (lo, hi) = D.shard_batch(ids, images, boxes, world, rank)
    assert lids.shape[0] == hi - lo == len(lbb)
    for j, b in enumerate(lbb):                       # boxes travel with their image
        assert b.shape[0] == lo + j + 1 and float(b[0, 0]) == float(limg[j, 0, 0, 0]) == lo + j
    D.barrier()
    t = D.max_over_ranks(10.0 + rank)
    sizes = [3, 2]
    rows = D.gather_rows(lids[:, :2].float(), world, rank, sizes)
    q.put((rank, t, (lo, hi), rows.tolist()))
    torch.distributed.destroy_process_group()


def test_two_rank_replica_plumbing():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [11.0, 11.0]                  # max over ranks
    assert res[0][2] == (0, 3) and res[1][2] == (3, 5)          # contiguous, covers the batch once
    want = torch.arange(35).view(5, 7)[:, :2].float().tolist()
    assert res[0][3] == want and res[1][3] == want              # gathered back in global order


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for rank in (0, 1):
        env = dict(os.environ, WORLD_SIZE='2', RANK=str(rank), LOCAL_RANK=str(rank), G4R_BENCH_TINY='1')
        r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                            '--steps', '1'], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[1] == ''
    line = json.loads(outs[0].splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'samples/s' and line['value'] > 0
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['cpu_baseline']['cores'] >= 1


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from gpt4roi_b200.train import LAYER_KEYS, LayerBucketAllReduce
    red = LayerBucketAllReduce()
    layers = []
    for i in (1, 0):                                  # backward order: last layer first
        g = {k: torch.full((3, 2), float(10 * i + j + rank)) for j, k in enumerate(LAYER_KEYS)}
        layers.append((i, g))
        red.hook(i, g)                                # async all-reduce of the layer's bucket
    top = [torch.full((4,), float(100 + rank))]
    red.reduce_now(top)
    # the trainer's layout: the four matrix gradients of a layer are views of ONE flat buffer -> one collective
    flat = torch.arange(8, dtype=torch.float32) + rank
    red.hook(0, {'flat': flat, 'wqkv': flat[:4].view(2, 2)})
    red.top_hook({'lm_head': top[0]})                 # lm_head goes out again here: 2 * (100 + 101)
    red.wait()
    assert red.calls == 2 * 6 + 1 + 1 + 1, red.calls
    assert torch.equal(flat, 2 * torch.arange(8, dtype=torch.float32) + 1)
    out = {i: {k: float(v[0, 0]) for k, v in g.items()} for i, g in layers}
    q.put((rank, out, float(top[0][0])))
    dist.destroy_process_group()


def test_layer_bucket_allreduce_sums_over_ranks():
    """DDP gradient all-reduce of the training step: one bucket per decoder layer, summed in place
    over the ranks (the 1/world average is folded into AdamW's grad_scale)."""
    from gpt4roi_b200.train import LAYER_KEYS
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_reducer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, top in res:
        for i in (0, 1):
            for j, k in enumerate(LAYER_KEYS):
                assert out[i][k] == 2 * (10 * i + j) + 1            # (v + 0) + (v + 1)
        assert top == 402.0
