"""Layer-level parity of the drop-in API (GPTQ.add_batch / quantize / get_quantized_weight, the
main.py quantize_layer body) against the reference's own outputs (golden fixtures) and the oracle,
at the north_star tolerances: codes >= 99.9 %, scales 1e-4 rel, reconstruction error 1e-3 rel."""

import os

import numpy as np
import pytest
import torch
import torch.nn as nn

import oracle
import parity
import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _layer(W):
    n, m = W.shape
    layer = nn.Linear(m, n, bias=False).to(DEV)
    layer.weight.data = torch.from_numpy(W).to(DEV)
    return layer


def _run(W, X, use_ssr, fp16_acts=True, **kw):
    import tq100
    g = tq100.GPTQ(_layer(W), block_size=kw.pop("block_size", 128), percdamp=0.01)
    for i in range(X.shape[0]):
        xi = torch.from_numpy(X[i]).to(DEV)
        if not fp16_acts:
            xi = xi.float()
        g.add_batch(xi[None] if i % 2 == 0 else xi)
    alpha, mu, T, perm = g.quantize(use_ssr=use_ssr, **kw)
    return g, dict(alpha=alpha.cpu().numpy(), mu=mu.cpu().numpy(), T=T.cpu().numpy(), perm=perm.cpu().numpy())


def _regen(g):
    W = synth.make_weight(int(g["n"]), int(g["m"]), seed=int(g["seed"]))
    X = synth.make_activations(int(g["samples"]), int(g["seq"]), int(g["m"]), seed=int(g["seed"]) + 1000,
                               lam=float(g["lam"]))
    assert abs(synth.checksum(W, X) - float(g["checksum"])) < 1e-6 * max(1.0, abs(float(g["checksum"])))
    return W, X


@pytest.mark.parametrize("tag", ["small", "mid"])
@pytest.mark.parametrize("order", ["seq", "ssr"])
@pytest.mark.parametrize("fp16_acts", [True, False])
def test_gptq_vs_reference_fixture(golden_dir, tag, order, fp16_acts):
    gold = np.load(os.path.join(golden_dir, f"gptq_{tag}_{order}.npz"))
    W, X = _regen(gold)
    g, got = _run(W, X, use_ssr=(order == "ssr"), fp16_acts=fp16_acts)
    assert g.nsamples == int(gold["nsamples"]) and g.info == 0
    assert got["T"].dtype == np.float32 and got["perm"].dtype == np.int64
    assert got["alpha"].shape == gold["alpha"].shape and got["T"].shape == gold["T"].shape
    ref = dict(alpha=gold["alpha"], mu=gold["mu"], T=gold["T"], perm=gold["perm"])
    parity.assert_layer_parity(got, ref, what=f"{tag}/{order}")
    Xf = X.astype(np.float64).reshape(-1, W.shape[1])
    H = Xf.T @ Xf
    Wq = g.get_quantized_weight().cpu().numpy()
    e_got = oracle.reconstruction_error(W, Wq, H)
    e_ref = oracle.reconstruction_error(W, gold["Wq"], H)
    assert abs(e_got - e_ref) <= parity.RECON_RTOL * e_ref, (e_got, e_ref)
    # dequant kernel == gptq.py:201-230 applied to our own outputs
    np.testing.assert_allclose(Wq, oracle.get_quantized_weight(got["alpha"], got["mu"], got["T"], got["perm"]),
                               rtol=0, atol=1e-7)
    assert sorted(got["perm"].tolist()) == list(range(W.shape[1]))
    assert set(np.unique(got["T"]).tolist()) <= {-1.0, 0.0, 1.0}


@pytest.mark.parametrize("order", ["seq", "ssr"])
def test_main_quantize_layer_dropin(golden_dir, order):
    """main.py:102-230 body: same inputs, same CPU dict (T int8), AGA on the raw-activation Gram."""
    import tq100
    gold = np.load(os.path.join(golden_dir, f"main_small_{order}.npz"))
    W = synth.make_weight(48, 320, seed=int(gold["seed"]))
    X = synth.make_activations(4, 160, 320, seed=int(gold["seed"]) + 1000, lam=0.5)
    res = tq100.quantize_layer(_layer(W), "l", torch.from_numpy(X).float().to(DEV), use_ssr=(order == "ssr"))
    assert res["T"].dtype == torch.int8 and all(not v.is_cuda for v in res.values())
    got = {k: v.numpy() for k, v in res.items()}
    parity.assert_layer_parity(got, dict(alpha=gold["alpha"], mu=gold["mu"], T=gold["T"], perm=gold["perm"]),
                               what=f"main/{order}")


@pytest.mark.parametrize("n,m,order,aga,lam", [
    (1024, 1024, "sequential", "hessian", 0.5),
    (1024, 1024, "sequential", "hessian", 1.0),
    (512, 2048, "sequential", "none", 0.5),
    (768, 1536, "sequential", "activations", 0.5),
    (300, 1100, "sequential", "hessian", 0.5),        # ragged rows and a ragged last block
])
def test_gptq_vs_oracle_larger(n, m, order, aga, lam):
    W = synth.make_weight(n, m, seed=n + m)
    X = synth.make_activations(8, 512, m, seed=n * 3 + m, lam=lam)
    g, got = _run(W, X, use_ssr=False, aga=aga)
    Xf = X.astype(np.float32).reshape(-1, m)
    H = Xf.T @ Xf
    ra, ru, rT, rp = oracle.quantize_layer(W, H, Xf.shape[0], 128, 0.01, order, aga=aga)
    agree = parity.assert_layer_parity(got, dict(alpha=ra, mu=ru, T=rT, perm=rp), what=f"{n}x{m}/{aga}")
    e_got = oracle.reconstruction_error(W, g.get_quantized_weight().cpu().numpy(), H)
    e_ref = oracle.reconstruction_error(W, oracle.get_quantized_weight(ra, ru, rT, rp), H)
    assert abs(e_got - e_ref) <= parity.RECON_RTOL * e_ref
    print(f"{n}x{m} {aga} lam={lam}: agreement {agree:.6f} recon {e_got:.5f}/{e_ref:.5f}")


def test_gptq_ssr_vs_oracle_per_block():
    """With SSR one swapped top-k boundary de-correlates everything after it (the reference is not
    reproducible against itself there), so SSR is validated block by block: same W in ->
    same block out, then end to end on reconstruction error."""
    n, m = 512, 1536
    W = synth.make_weight(n, m, seed=71)
    X = synth.make_activations(8, 512, m, seed=72, lam=0.5)
    g, got = _run(W, X, use_ssr=True)
    Xf = X.astype(np.float32).reshape(-1, m)
    H = Xf.T @ Xf
    ra, ru, rT, rp = oracle.quantize_layer(W, H, Xf.shape[0], 128, 0.01, "ssr")
    assert sorted(got["perm"].tolist()) == list(range(m))
    same_blocks = sum(set(got["perm"][k:k + 128].tolist()) == set(rp[k:k + 128].tolist()) for k in range(0, m, 128))
    e_got = oracle.reconstruction_error(W, g.get_quantized_weight().cpu().numpy(), H)
    e_ref = oracle.reconstruction_error(W, oracle.get_quantized_weight(ra, ru, rT, rp), H)
    print(f"SSR blocks with identical membership: {same_blocks}/{m // 128}; recon {e_got:.5f} vs {e_ref:.5f}")
    assert set(got["perm"][:128].tolist()) == set(rp[:128].tolist()), "first SSR block must match (no cascade yet)"
    assert abs(e_got - e_ref) <= parity.RECON_RTOL * e_ref
    if np.array_equal(got["perm"], rp):
        parity.assert_layer_parity(got, dict(alpha=ra, mu=ru, T=rT, perm=rp), what="ssr e2e")


def test_actorder_extension_matches_prepermuted_sequential():
    n, m = 256, 512
    W = synth.make_weight(n, m, seed=81)
    X = synth.make_activations(4, 256, m, seed=82, lam=0.5).astype(np.float32)
    X *= np.linspace(0.5, 2.0, m, dtype=np.float32)[None, None, :]
    X = X.astype(np.float16)
    g, got = _run(W, X, use_ssr=False, order="actorder", aga="none")
    Xf = X.astype(np.float32).reshape(-1, m)
    H = Xf.T @ Xf
    ra, ru, rT, rp = oracle.quantize_layer(W, H, Xf.shape[0], 128, 0.01, "actorder", aga="none")
    parity.assert_layer_parity(got, dict(alpha=ra, mu=ru, T=rT, perm=rp), what="actorder")


def test_single_block_layer_and_other_block_sizes():
    """m <= block_size crashes the reference; here it quantises the one block."""
    W = synth.make_weight(64, 96, seed=91)
    X = synth.make_activations(2, 128, 96, seed=92)
    for use_ssr in (False, True):
        g, got = _run(W, X, use_ssr=use_ssr)
        assert got["alpha"].shape == (64, 1) and np.array_equal(got["perm"], np.arange(96))
    W2 = synth.make_weight(128, 512, seed=93)
    X2 = synth.make_activations(2, 256, 512, seed=94)
    Xf = X2.astype(np.float32).reshape(-1, 512)
    for bs in (64, 256):
        g, got = _run(W2, X2, use_ssr=False, block_size=bs)
        ra, ru, rT, rp = oracle.quantize_layer(W2, Xf.T @ Xf, Xf.shape[0], bs, 0.01, "sequential")
        parity.assert_layer_parity(got, dict(alpha=ra, mu=ru, T=rT, perm=rp), block_size=bs, what=f"block {bs}")


def test_shared_hessian_state_gives_identical_results():
    """q/k/v read the same activations: one HessianState, one inverse."""
    import tq100
    m = 512
    X = synth.make_activations(4, 256, m, seed=95)
    Wq_, Wk_ = synth.make_weight(128, m, seed=96), synth.make_weight(192, m, seed=97)
    shared = tq100.HessianState(m, DEV)
    for i in range(4):
        shared.add_batch(torch.from_numpy(X[i]).to(DEV))
    outs = []
    for W in (Wq_, Wk_):
        gs = tq100.GPTQ(_layer(W), hessian=shared)
        a, u, T, p = gs.quantize(use_ssr=False)
        g, got = _run(W, X, use_ssr=False)
        assert np.array_equal(T.cpu().numpy(), got["T"]) and np.array_equal(a.cpu().numpy(), got["alpha"])
        outs.append(T)
    assert len(shared._cache) == 1


def test_get_quantized_weight_before_quantize_raises():
    import tq100
    g = tq100.GPTQ(_layer(synth.make_weight(8, 128, seed=1)))
    with pytest.raises(RuntimeError, match="Must call quantize"):
        g.get_quantized_weight()
    with pytest.raises(RuntimeError, match="before add_batch"):
        g.quantize()


def test_layer_driver_streams_match_sequential():
    """LayerDriver spreads the per-linear prologue+sweep chains over CUDA streams; results must match the plain
    one-linear-at-a-time API (to parity tolerance: the Hessian's reduce-add order is not fixed run to run), also
    when earlier results are still alive (regression: finish() used to read the Cholesky status on the caller's
    stream before the side stream had produced it)."""
    import tq100
    from tq100.pipeline import LayerDriver
    shapes = [("a", 384, 640), ("b", 640, 384), ("c", 300, 520)]
    Ws = {nm: torch.from_numpy(synth.make_weight(n, m, seed=200 + i)).to(DEV) for i, (nm, n, m) in enumerate(shapes)}
    Xs = {m: torch.from_numpy(synth.make_activations(4, 256, m, seed=300 + m)).to(DEV) for m in (640, 384, 520)}
    ref = {}
    for nm, n, m in shapes:
        g = tq100.GPTQ(_layer(Ws[nm].cpu().numpy()))
        g.add_batch(Xs[m])
        a, u, T, p = g.quantize(use_ssr=True)
        ref[nm] = dict(alpha=a.cpu().numpy(), mu=u.cpu().numpy(), T=T.cpu().numpy(), perm=p.cpu().numpy())
    keep = []
    for streams in (1, 2, 3):
        drv = LayerDriver(DEV, num_streams=streams)
        for rep in range(3):
            gs = drv.quantize([(nm, Ws[nm], Xs[m]) for nm, n, m in shapes], use_ssr=True)
            keep.append(gs)
            for g, (nm, n, m) in zip(gs, shapes):
                assert g.info == 0
                got = dict(alpha=g.alpha.cpu().numpy(), mu=g.mu.cpu().numpy(), T=g.T.cpu().numpy(), perm=g.perm.cpu().numpy())
                parity.assert_layer_parity(got, ref[nm], what=f"{nm}/streams={streams}/rep={rep}")


@pytest.mark.parametrize("share,depth", [(False, 3), (True, 2), (False, 1)])
def test_host_pipeline_matches_device_api(share, depth):
    """HostPipeline (pinned host in, pinned host out, copies double-buffered against the kernels, the group's chains
    on side streams) returns what the plain device API returns; with share_inputs the linears of a group use one
    Hessian and one inverse.  Three passes over the groups, as three transformer layers would."""
    import tq100
    from tq100.pipeline import HostPipeline
    groups_spec = [(640, [("q", 384), ("k", 256), ("v", 384)]), (384, [("o", 640)]), (520, [("gate", 300), ("up", 300)])]
    groups, ref = [], {}
    for gi, (m, lins) in enumerate(groups_spec):
        X = synth.make_activations(4, 256, m, seed=400 + gi)
        xh = torch.from_numpy(X).pin_memory()
        entry = []
        for li, (nm, n) in enumerate(lins):
            W = synth.make_weight(n, m, seed=500 + 10 * gi + li)
            entry.append((nm, torch.from_numpy(W).pin_memory()))
            g = tq100.GPTQ(_layer(W))
            g.add_batch(torch.from_numpy(X).to(DEV))
            a, u, T, p = g.quantize(use_ssr=True)
            ref[nm] = dict(alpha=a.cpu().numpy(), mu=u.cpu().numpy(), T=T.cpu().numpy(), perm=p.cpu().numpy())
        groups.append((xh, entry))
    pipe = HostPipeline(DEV, use_ssr=True, aga="hessian", share_inputs=share, num_streams=2, depth=depth)
    seen = 0
    for out_group in pipe.run_iter(groups * 3):
        pipe.synchronize()
        for out in out_group:
            assert not out["T"].is_cuda and out["T"].dtype == torch.int8 and out["perm"].dtype == torch.int64
            got = dict(alpha=out["alpha"].numpy(), mu=out["mu"].numpy(), T=out["T"].numpy().astype(np.float32),
                       perm=out["perm"].numpy())
            parity.assert_layer_parity(got, ref[out["name"]], what=f"host pipeline {out['name']} share={share}")
            seen += 1
    assert seen == 3 * 6
    res = pipe.run(groups)
    assert [r["name"] for r in res] == ["q", "k", "v", "o", "gate", "up"]
    x_bytes = sum(x.numel() * 2 for x, _ in groups)
    w_bytes = sum(w.numel() * 4 for _, e in groups for _, w in e)
    assert pipe.h2d_bytes == 4 * (x_bytes + w_bytes)


@pytest.mark.parametrize("tag", ["seq", "ssr"])
def test_non_pd_hessian_takes_the_pinv_route(tag, golden_dir):
    """gptq.py:101-106: Cholesky fails on an indefinite H, the reference continues on torch.linalg.pinv.  Here
    tq_chol_inverse reports the failing pivot through `info`, GPTQ.finish() re-runs the whole sweep on the library
    pseudo-inverse.  Compared with the unmodified reference's outputs (tests/golden/gptq_pinv.npz)."""
    import tq100
    g = np.load(os.path.join(golden_dir, "gptq_pinv.npz"))
    W, H = g["W"], g["H"]
    q = tq100.GPTQ(_layer(W), block_size=128, percdamp=0.01)
    q.H = torch.from_numpy(H.copy())                          # the reference's attributes (gptq.py:50-51)
    q.nsamples = 1
    alpha, mu, T, perm = q.quantize(use_ssr=(tag == "ssr"))
    assert q.info > 0                                         # 1-based index of the first non-positive pivot
    Hd, Hinv, info = q.state.damped_inverse(q.percdamp)       # the cache now holds the pseudo-inverse
    assert int(info.item()) == 0
    eye = (Hd.double() @ Hinv.double()).cpu().numpy()
    assert np.abs(eye - np.eye(H.shape[0])).max() < 1e-3      # pinv of a nonsingular matrix is its inverse
    got = dict(alpha=alpha.cpu().numpy(), mu=mu.cpu().numpy(), T=T.cpu().numpy(), perm=perm.cpu().numpy())
    ref = dict(alpha=g[f"{tag}_alpha"], mu=g[f"{tag}_mu"], T=g[f"{tag}_T"], perm=g[f"{tag}_perm"])
    parity.assert_layer_parity(got, ref, what=f"pinv route/{tag}")
    # a second linear sharing the state reuses the cached pseudo-inverse without failing again
    q2 = tq100.GPTQ(_layer(W), block_size=128, percdamp=0.01, hessian=q.state)
    a2, u2, T2, p2 = q2.quantize(use_ssr=(tag == "ssr"))
    assert q2.info == 0 and torch.equal(T2, T) and torch.equal(p2, perm)


# ---- canonical-GPTQ exactness options (default off), against oracle/gptq_canonical.py ---------------
@pytest.mark.parametrize("order", ["sequential", "actorder"])
def test_block_exact_feedback_vs_oracle(order):
    """feedback='block_exact': the OBS-exact block update through the Cholesky factor of H^-1, against the oracle's
    first-principles form (solve with the current inverse Hessian + Schur complement per block).  The two formulations
    differ in arithmetic (fp64 Schur recursion vs one Cholesky), so parity is judged like every layer test; the
    reconstruction error must also improve on the reference's own feedback (what the option is for)."""
    import tq100
    from oracle import gptq_canonical as gc
    n, m = 256, 768
    W = synth.make_weight(n, m, seed=811)
    X = synth.make_activations(6, 256, m, seed=812, lam=1.0)
    Xf = X.astype(np.float32).reshape(-1, m)
    H = Xf.T @ Xf
    q = tq100.GPTQ(_layer(W))
    q.add_batch(torch.from_numpy(X).to(DEV))
    alpha, mu, T, perm = q.quantize(use_ssr=False, order=order, feedback="block_exact")
    ra, ru, rT, rp = gc.quantize_layer_exact(W, H, Xf.shape[0], 128, 0.01, order, feedback="block_exact")
    got = dict(alpha=alpha.cpu().numpy(), mu=mu.cpu().numpy(), T=T.cpu().numpy(), perm=perm.cpu().numpy())
    parity.assert_layer_parity(got, dict(alpha=ra, mu=ru, T=rT, perm=rp), what=f"block_exact/{order}")
    e_exact = oracle.reconstruction_error(W, q.get_quantized_weight().cpu().numpy(), H)
    q.quantize(use_ssr=False, order=order)                                         # the reference's feedback, same inputs
    e_ref = oracle.reconstruction_error(W, q.get_quantized_weight().cpu().numpy(), H)
    assert e_exact < e_ref, (e_exact, e_ref)
    with pytest.raises(ValueError):
        q.quantize(use_ssr=True, feedback="block_exact")                           # needs a static order


def test_dead_columns_vs_oracle():
    """dead_columns=True: input features that never fire (diag(H) == 0) get H_jj = 1 and weight 0 (canonical GPTQ)."""
    import tq100
    from oracle import gptq_canonical as gc
    n, m = 128, 512
    W = synth.make_weight(n, m, seed=821)
    X = synth.make_activations(4, 256, m, seed=822, lam=0.5)
    deadc = np.array([3, 130, 131, 400])
    X[:, :, deadc] = 0
    Xf = X.astype(np.float32).reshape(-1, m)
    H = Xf.T @ Xf
    q = tq100.GPTQ(_layer(W))
    q.add_batch(torch.from_numpy(X).to(DEV))
    alpha, mu, T, perm = q.quantize(use_ssr=False, dead_columns=True)
    ra, ru, rT, rp = gc.quantize_layer_exact(W, H, Xf.shape[0], 128, 0.01, "sequential", feedback="reference", dead_columns=True)
    got = dict(alpha=alpha.cpu().numpy(), mu=mu.cpu().numpy(), T=T.cpu().numpy(), perm=perm.cpu().numpy())
    parity.assert_layer_parity(got, dict(alpha=ra, mu=ru, T=rT, perm=rp), what="dead_columns")
    # the shared accumulator is untouched, and without the option the zero diagonal stays what the reference makes of it
    assert float(q.H.diagonal()[3]) == 0.0
