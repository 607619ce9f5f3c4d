"""CPU: the N>1 plumbing (flat gradient bucket + one all-reduce per step) with world_size-2 gloo processes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from equiformer_b200.parallel import FlatGradAllReduce, broadcast_parameters, init_distributed
    r, _l, w = init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                     # different init per rank -> broadcast must equalise
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.SiLU(), torch.nn.Linear(5, 1)).double()
    broadcast_parameters(model)
    bucket = FlatGradAllReduce(model.parameters())
    g = torch.Generator().manual_seed(7)
    data = torch.randn(8, 6, generator=g, dtype=torch.float64)   # same full batch on every rank
    shard = data[rank::world]                                     # each rank owns its molecules
    bucket.zero_grad()
    (model(shard).sum() / data.shape[0] * world).backward()       # so that the rank-average equals the full-batch grad
    bucket.reduce()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])   # the bucket itself is padded per tensor
    # single-process reference on the full batch
    ref_model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.SiLU(), torch.nn.Linear(5, 1)).double()
    ref_model.load_state_dict(model.state_dict())
    (ref_model(data).sum() / data.shape[0]).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = all(torch.allclose(t, ref, atol=1e-12) for t in gathered)
    views_ok = all(p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr() for p in model.parameters())
    if rank == 0:
        ret["ok"] = bool(ok and views_ok)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret.get("ok") is True


def test_bench_reference_arm_ranks_other_than_zero_exit_quietly(monkeypatch, capsys):
    import bench
    monkeypatch.setenv("RANK", "1")

    class A:
        gpus, steps, warmup, ref_graphs = 2, 1, 0, 1
    bench.run_reference(A())
    assert capsys.readouterr().out == ""


def test_flat_adamw_matches_torch_adamw():
    """FlatAdamW (one flat buffer, six element-wise kernels) == torch.optim.AdamW with the decay groups of the
    reference's ``add_weight_decay`` (optim_factory.py:27-42): exemption by NAME - a flat 1-D ``tp.weight`` is decayed."""
    from equiformer_b200.nets.tensor_product_rescale import LinearRS
    from equiformer_b200.parallel import FlatAdamW, FlatGradAllReduce, is_no_decay

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = LinearRS("6x0e", "5x0e", bias=True)          # lin.tp.weight is 1-D (30), lin.bias.0 exempt
            self.fc = torch.nn.Linear(5, 3)
            self.norm = torch.nn.Module()
            self.norm.affine_weight = torch.nn.Parameter(torch.ones(3))   # '*.affine_weight' is exempt
            self.scale = torch.nn.Parameter(torch.ones(3))           # 1-D, NOT exempt under the reference rule

        def forward(self, x):
            return self.fc(torch.nn.functional.silu(self.lin(x))) * self.scale * self.norm.affine_weight

    torch.manual_seed(0)
    a, b = Net().double(), Net().double()
    b.load_state_dict(a.state_dict())
    assert a.lin.tp.weight.dim() == 1
    bucket = FlatGradAllReduce(a.parameters())
    skip = {"scale_not_present"}
    opt_a = FlatAdamW(a.named_parameters(), bucket, lr=1e-2, weight_decay=0.1, no_decay=skip)
    decay = [p for n, p in b.named_parameters() if not is_no_decay(n, skip)]
    no_decay = [p for n, p in b.named_parameters() if is_no_decay(n, skip)]
    names_decay = {n for n, _ in b.named_parameters() if not is_no_decay(n, skip)}
    assert "lin.tp.weight" in names_decay and "scale" in names_decay and "fc.weight" in names_decay
    assert "lin.bias.0" not in names_decay and "fc.bias" not in names_decay and "norm.affine_weight" not in names_decay
    opt_b = torch.optim.AdamW([{"params": decay, "weight_decay": 0.1}, {"params": no_decay, "weight_decay": 0.0}], lr=1e-2)
    g = torch.Generator().manual_seed(1)
    for _ in range(5):
        x = torch.randn(7, 6, generator=g, dtype=torch.float64)
        bucket.zero_grad()
        a(x).pow(2).sum().backward()
        opt_a.step()
        opt_b.zero_grad()
        b(x).pow(2).sum().backward()
        opt_b.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-12), (pa - pb).abs().max()
    # on the headline model the rule decays every tensor-product weight (3.0 M of 3.5 M parameters)
    from equiformer_b200.nets import model_entrypoint
    m = model_entrypoint("graph_attention_transformer_nonlinear_l2")(irreps_in="5x0e", radius=5.0, num_basis=128)
    skip = m.no_weight_decay()
    n_exempt = sum(p.numel() for n, p in m.named_parameters() if is_no_decay(n, skip))
    assert n_exempt < 40000, n_exempt


def test_flat_bucket_store_matches_backward_accumulation():
    """FlatGradAllReduce.store(autograd.grad(...)) fills the bucket exactly like loss.backward() into the views."""
    import torch
    from equiformer_b200.parallel import FlatGradAllReduce
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.SiLU(), torch.nn.Linear(7, 3))
    unused = torch.nn.Parameter(torch.randn(4))
    params = list(net.parameters()) + [unused]
    bucket = FlatGradAllReduce(params)
    x = torch.randn(11, 5)
    bucket.zero_grad()
    net(x).square().sum().backward()
    ref = bucket.flat.clone()
    bucket.flat.fill_(123.0)
    bucket.store(torch.autograd.grad(net(x).square().sum(), bucket.params, allow_unused=True))
    live = torch.cat([p.grad.reshape(-1) for p in params])
    assert torch.equal(live, torch.cat([ref[o:o + p.numel()] for p, o in zip(params, bucket.offsets)]))
    assert torch.count_nonzero(unused.grad) == 0
