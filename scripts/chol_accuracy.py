import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as G, synth, oracle
for m, lam in ((1000, 1.0), (2048, 0.5), (4096, 0.5)):
    X = synth.make_activations(4, max(256, m // 2), m, seed=41, lam=lam).astype(np.float32).reshape(-1, m)
    Hraw = (X.T @ X).astype(np.float32)
    Hd_o, Hinv_o = oracle.damped_inverse(Hraw.copy(), X.shape[0], 0.01)
    Hd, Hinv, info = G.finalize_and_invert(G.dev(Hraw), X.shape[0], 0.01)
    Hd, Hinv = Hd.cpu().numpy().astype(np.float64), Hinv.cpu().numpy().astype(np.float64)
    I = np.eye(m)
    res = np.abs(Hd @ Hinv - I).max(); res_o = np.abs(Hd_o.astype(np.float64) @ Hinv_o.astype(np.float64) - I).max()
    exact = np.linalg.inv(Hd_o.astype(np.float64))
    e = np.abs(Hinv - exact).max() / np.abs(exact).max(); e_o = np.abs(Hinv_o - exact).max() / np.abs(exact).max()
    print(f"TQ_CHOL_FFMA={os.environ.get('TQ_CHOL_FFMA','0')} m={m}: residual {res:.2e} (oracle {res_o:.2e}); rel err {e:.2e} (oracle {e_o:.2e})", flush=True)
