"""Host-side data-parallel plumbing on CPU: world_size-2 gloo run of FlatParams + bucketed GradAllReduce
(the N>1 path of bench.py / Trainer, minus the CUDA kernels)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from hawkeye_b200 import engine
    r, l, w = engine.init_distributed(backend='gloo')
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                       # identical replicas
    backbone = torch.nn.Linear(6, 5)
    head = torch.nn.Linear(5, 3)
    flat = engine.FlatParams(None, groups=[list(backbone.parameters()), list(head.parameters())])
    ar = engine.GradAllReduce(flat, early_group=1, world=world)
    keys_before = [tuple(p.shape) for p in flat.params]
    g = torch.Generator().manual_seed(100 + rank)  # different shard per rank
    x = torch.randn(4, 6, generator=g)
    flat.zero_grad()
    loss = head(torch.relu(backbone(x))).pow(2).mean()
    loss.backward()
    ar.finish()
    avg = flat.grad / world
    # reference: mean of the per-rank gradients == gradient of the mean loss over the concatenated batch
    xs = [torch.randn(4, 6, generator=torch.Generator().manual_seed(100 + k)) for k in range(world)]
    b2, h2 = torch.nn.Linear(6, 5), torch.nn.Linear(5, 3)
    b2.load_state_dict(backbone.state_dict())
    h2.load_state_dict(head.state_dict())
    sum(h2(torch.relu(b2(xx))).pow(2).mean() for xx in xs).div(world).backward()
    ref = torch.cat([p.grad.flatten() for p in list(b2.parameters()) + list(h2.parameters())])
    got = torch.cat([avg[a:a + p.numel()] for p, a in zip(flat.params, _offsets(flat))])
    ok = torch.allclose(got, ref, atol=1e-6) and keys_before == [tuple(p.shape) for p in flat.params]
    ok = ok and all(p.data_ptr() >= flat.flat.data_ptr() for p in flat.params)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _offsets(flat):
    offs, off = [], 0
    for p in flat.params:
        offs.append(off)
        off += (p.numel() + 3) // 4 * 4
    return offs


def test_gloo_world2_grad_allreduce():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) for r in range(world)), dict(out)


def test_flat_params_keep_state_dict_surface():
    sys.path.insert(0, ROOT)
    from hawkeye_b200 import engine
    m = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Linear(5, 3))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    flat = engine.FlatParams(m)
    assert list(m.state_dict()) == list(sd) and all(torch.equal(m.state_dict()[k], sd[k]) for k in sd)
    assert flat.numel % 4 == 0 and all(p.data_ptr() % 16 == 0 for p in flat.params)


def test_average_meter_async_readback():
    """Trainer meters fold in device read-backs lazily (no host sync in the step); `.avg` drains what is pending."""
    import torch
    from hawkeye_b200.train import AverageMeter

    class Ev:
        def __init__(self, done):
            self.done, self.synced = done, False

        def query(self):
            return self.done

        def synchronize(self):
            self.done = self.synced = True

    m = AverageMeter()
    buf = torch.tensor([2.0, 16.0])
    e1, e2 = Ev(True), Ev(False)
    m.update_async(buf, 0, 1.0, 4, e1)                  # already landed: folded immediately
    assert m.count == 4 and abs(m.sum - 8.0) < 1e-9
    m.update_async(buf, 1, 100.0 / 32, 32, e2)          # still in flight: kept pending
    assert m.count == 4 and len(m._pending) == 1
    assert abs(m.avg - (8.0 + 50.0 * 32) / 36) < 1e-9   # .avg waits for it
    assert e2.synced and not m._pending
    m.update(1.0, 4)
    assert m.count == 40
