import sys, os, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from gf2_loader import load
gf2 = load(); synth = importlib.import_module("gf2_b200.synth"); import gf2_oracle as oracle
for prior, nl in (("anchor", 200), ("dense", 1000)):
    n = 3
    w = synth.make_windows(n, n_landmarks=nl, prior=prior)
    oracle.imu_preintegrate(w)
    opts = gf2.abi.default_opts()
    s = gf2.Solver(n, w["n_frames"], w["max_landmarks"], w["max_obs"])
    s.upload(w, preintegrate="records"); s.solve(opts, n)
    st = s.get_states(n); lam = s.get_landmarks(n)
    status, m = s.marginalize(opts, mode=0)
    print(prior, "status", status, "m", m, "ms", s.last_marginalize_ms())
    got = s.get_prior(n)
    w["para_pose"][...] = st["para_pose"]; w["para_speedbias"][...] = st["para_speedbias"]; w["inv_depth"][...] = lam
    for i in range(n):
        ref = oracle.marginalize_window(w, i, opts, mode=0)
        nn = int(got["prior_rows"][i]); nb = int(got["prior_nblocks"][i])
        g = {"n": nn, "J0": got["prior_J0"][i, :nn, :nn], "r0": got["prior_r0"][i, :nn], "blocks": got["prior_blocks"][i, :nb]}
        Hg, gg, _ = oracle.prior_information(g, 11); Hr, gr, _ = oracle.prior_information(ref, 11)
        d = np.sqrt(np.maximum(np.diag(Hr), 1e-300)); nz = np.diag(Hr) > 0
        eH = (np.abs(Hg - Hr) / np.outer(d, d))[np.ix_(nz, nz)].max()
        eg = (np.abs(gg - gr) / d)[nz].max()
        print(i, "n", nn, ref["n"], "relH(max)", np.abs(Hg - Hr).max() / np.abs(Hr).max(), "normH", eH, "norm g", eg, "rel g", np.abs(gg - gr).max() / np.abs(gr).max(),
              "min/max eig", np.linalg.eigvalsh(Hr)[[0, -1]])
    s.close()
