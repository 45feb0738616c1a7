"""Multi-process form of the model-level driver (SURVEY 8e x 8f N1) on CPU with world_size-2 gloo: every rank runs the
forwards of ITS calibration samples, Hessians are all-reduced per layer, linears are dealt to the ranks, results
broadcast.  The numpy oracle stands in for the CUDA quantiser (tests only); the collectives, the dealing and the
bookkeeping are the product's.  Checked: both ranks end with bit-identical parameters and models, each ran half of the
forwards, the linears were split between the ranks, and the result is within the model-level tolerances of the
unmodified reference's single-process ``quantize()`` (tests/golden/model_toy_seq.npz)."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    torch.set_num_threads(2)
    import parity
    import toy_model
    import tq100
    from test_model_driver_cpu import OracleGPTQ
    from tq100.sharded import ShardContext
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ran = []

        def runner(gs, ssr):
            ran.extend(id(g) for g in gs)
            for g in gs:
                g.run(ssr)

        model = toy_model.build()
        pq = tq100.PT2LLMQuantizer(model, None, model_type="llama", use_ssr=False, device="cpu",
                                   gptq_factory=OracleGPTQ, chain_runner=runner, shard=ShardContext(rank, world))
        params = pq.quantize(toy_model.samples())
        assert model.forward_calls == 16 // world and pq.layer_forwards == (16 // world) * 3
        assert 0 < len(ran) < 14                               # this rank quantised its share of the 14 linears only
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([len(ran)]))
        assert sum(int(c) for c in counts) == 14
        # identical everywhere: parameters and the overwritten model
        blob = torch.cat([params[k][f].float().reshape(-1) for k in sorted(params) for f in ("alpha", "mu", "T", "perm")]
                         + [p.detach().reshape(-1) for p in model.parameters()])
        ref = blob.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(blob, ref)
        if rank == 0:
            gold = np.load(os.path.join(HERE, "golden", "model_toy_seq.npz"))
            for name, p in params.items():
                got = {k: v.numpy() for k, v in p.items()}
                parity.assert_model_level_parity(name, got, {k: gold[f"{name}/{k}"] for k in ("alpha", "mu", "T", "perm")})
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_model_driver_world2_gloo():
    world = 2
    port = 30100 + (os.getpid() % 500)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
