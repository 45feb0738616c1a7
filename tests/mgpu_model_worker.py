"""torchrun worker: the sharded whole-model driver (PT2LLMQuantizer(shard=...)) on N GPUs -- NCCL all-reduce of the
hook-accumulated Hessians, linears dealt to ranks, results broadcast -- against the reference's own single-process
``quantize()`` on the toy model (tests/golden/model_toy_seq.npz)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import parity  # noqa: E402
import toy_model  # noqa: E402
import tq100  # noqa: E402
from tq100 import sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = sharded.ShardContext(rank, world, dev)
    model = toy_model.build().to(dev)
    pq = tq100.PT2LLMQuantizer(model, None, model_type="llama", use_ssr=False, device=str(dev), shard=ctx)
    params = pq.quantize(toy_model.samples())
    assert model.forward_calls == len(ctx.my_samples(16))
    blob = torch.cat([params[k][f].float().reshape(-1) for k in sorted(params) for f in ("alpha", "mu", "T", "perm")]).to(dev)
    ref = blob.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(blob, ref), "ranks ended with different parameters"
    if rank == 0:
        gold = np.load(os.path.join(HERE, "golden", "model_toy_seq.npz"))
        for name, p in params.items():
            got = {k: v.numpy() for k, v in p.items()}
            agree = parity.assert_model_level_parity(name, got, {k: gold[f"{name}/{k}"] for k in ("alpha", "mu", "T", "perm")})
            print(f"[mgpu-model] {name}: code agreement {agree:.6f}", flush=True)
        print("MGPU_MODEL_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
