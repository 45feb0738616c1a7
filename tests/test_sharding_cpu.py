"""Host-side logic of the multi-GPU path, on CPU with world_size-2 gloo.

The CUDA kernels need a GPU; what is checked here is the DECOMPOSITION the sharded path relies on,
with the oracle standing in for the kernels:
  * sample-sharded Hessians summed by all-reduce == the unsharded Hessian;
  * rows are independent given H: row-slab sweeps concatenated == the unsharded sweep, bit for bit;
  * SSR: the block chosen from all-reduced per-slab column statistics == the unsharded choice.
"""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shard_context_partitions():
    from tq100.sharded import ShardContext
    for world in (1, 2, 3, 8):
        samples, rows = [], []
        for r in range(world):
            c = ShardContext(r, world)
            samples += c.my_samples(128)
            lo, hi = c.row_range(4099)
            assert lo % 4 == 0 and (hi % 4 == 0 or hi == 4099)
            rows.append((lo, hi))
        assert sorted(samples) == list(range(128))
        assert rows[0][0] == 0 and rows[-1][1] == 4099
        assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in rows]
        assert max(sizes) - min(sizes) <= 4


def test_deal_linears_balances_by_chain_cost():
    from tq100.pipeline import chain_cost
    from tq100.sharded import deal_linears
    shapes = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]          # one LLaMA-2-7B layer
    assert deal_linears(shapes, 1) == [0] * 7
    for world in (2, 4, 8):
        owner = deal_linears(shapes, world)
        load = [sum(chain_cost(*s) for s, o in zip(shapes, owner) if o == r) for r in range(world)]
        widest = chain_cost(4096, 11008)
        assert owner[6] == 0                                   # the longest chain is dealt first
        assert max(load) <= max(widest, sum(load) / world * 1.35)
    assert len(set(deal_linears(shapes, 8))) == 7               # 7 linears on 7 of 8 ranks


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import oracle
    import synth
    from tq100.sharded import ShardContext
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = ShardContext(rank, world)
        n, m, samples, seq = 96, 384, 6, 64
        W = synth.make_weight(n, m, seed=5)
        X = synth.make_activations(samples, seq, m, seed=6, lam=0.5).astype(np.float32)

        # 1. sample-sharded Hessian + all-reduce
        H = np.zeros((m, m), dtype=np.float32)
        ns = 0
        for i in ctx.my_samples(samples):
            H, ns = oracle.hessian_add_batch(H, ns, X[i])
        Ht = torch.from_numpy(H)
        nst = torch.tensor([ns])
        dist.all_reduce(Ht)
        dist.all_reduce(nst)
        Hfull = np.zeros((m, m), dtype=np.float32)
        nfull = 0
        for i in range(samples):
            Hfull, nfull = oracle.hessian_add_batch(Hfull, nfull, X[i])
        assert int(nst.item()) == nfull
        np.testing.assert_allclose(Ht.numpy(), Hfull, rtol=1e-5, atol=1e-4)
        Hsum = Ht.numpy()

        # 2. row-slab sweep (sequential order: no exchange) == unsharded, bit for bit
        lo, hi = ctx.row_range(n)
        a, u, T, p = oracle.quantize_layer(W[lo:hi], Hsum, nfull, 128, 0.01, "sequential")
        fa, fu, fT, fp = oracle.quantize_layer(W, Hsum, nfull, 128, 0.01, "sequential")
        assert np.array_equal(a, fa[lo:hi]) and np.array_equal(u, fu[lo:hi]) and np.array_equal(T, fT[lo:hi])
        assert np.array_equal(p, fp)

        # 3. SSR block selection from all-reduced per-slab statistics (what tq_sweep_layer does with
        #    TQ_SWEEP_ROW_SHARD: fold local partials, all-reduce 2*rem+1 floats, select)
        rem = np.setdiff1d(np.arange(m), np.arange(3, m, 7))
        Wl = W[lo:hi][:, rem]
        wbar = Wl.mean(axis=1, keepdims=True, dtype=np.float32)        # row-local
        stats = np.concatenate([(Wl * wbar).sum(0), (Wl * Wl).sum(0), [(wbar * wbar).sum()]]).astype(np.float32)
        st = torch.from_numpy(stats)
        dist.all_reduce(st)
        st = st.numpy()
        r_ = rem.shape[0]
        sim = st[:r_] / np.maximum(np.sqrt(st[r_:2 * r_]), 1e-8) / np.maximum(np.sqrt(st[2 * r_]), 1e-8)
        ref_sim = oracle.column_similarity_to_mean(W, rem)
        np.testing.assert_allclose(sim, ref_sim, rtol=2e-4, atol=2e-6)
        order = np.argsort(-sim, kind="stable")[:128]
        blk_ref, _ = oracle.select_next_block_ssr(W, rem, 128)
        kth = np.sort(ref_sim)[-128]
        diff = set(rem[order].tolist()) ^ set(blk_ref.tolist())
        pos = {c: i for i, c in enumerate(rem.tolist())}
        assert all(abs(ref_sim[pos[c]] - kth) < 1e-5 for c in diff)
        # every rank sees the same reduced statistics, hence selects the same block
        gathered = [torch.zeros(128, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(rem[order]))
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        # 4. linear-parallel mode: linears dealt to ranks, each H reduced onto its owner, the owner sweeps the whole
        #    linear alone -> identical to the unsharded result, no collective in the sweep
        from tq100.sharded import deal_linears
        shapes = [(96, 384), (64, 256), (48, 256)]
        owner = deal_linears(shapes, world)
        assert owner == deal_linears(shapes, world) and set(owner) == set(range(world))
        for li, (n2, m2) in enumerate(shapes):
            W2 = synth.make_weight(n2, m2, seed=50 + li)
            X2 = synth.make_activations(samples, seq, m2, seed=60 + li, lam=0.5).astype(np.float32)
            H2, ns2 = np.zeros((m2, m2), dtype=np.float32), 0
            for i in ctx.my_samples(samples):
                H2, ns2 = oracle.hessian_add_batch(H2, ns2, X2[i])
            H2t, ns2t = torch.from_numpy(H2), torch.tensor([ns2])
            dist.reduce(H2t, dst=owner[li])
            dist.all_reduce(ns2t)
            if owner[li] == rank:
                Hf, nf = np.zeros((m2, m2), dtype=np.float32), 0
                for i in range(samples):
                    Hf, nf = oracle.hessian_add_batch(Hf, nf, X2[i])
                assert int(ns2t.item()) == nf
                np.testing.assert_allclose(H2t.numpy(), Hf, rtol=1e-5, atol=1e-4)
                got = oracle.quantize_layer(W2, H2t.numpy(), nf, 128, 0.01, "sequential")
                ref = oracle.quantize_layer(W2, Hf, nf, 128, 0.01, "sequential")
                assert (got[2] == ref[2]).mean() >= 0.999
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_decomposition_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 500)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
