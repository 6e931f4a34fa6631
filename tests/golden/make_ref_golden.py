"""Generates tests/golden/ref_golden.npz: outputs of the REFERENCE's own factor sources (oracle/_ref/libgf2_ref.so — compiled unmodified
from /root/reference by `make -C oracle`, Eigen / Ceres / ROS / Sophus headers replaced by the stand-ins of oracle/shim) on the seeded
inputs of tests/ref_cases.py and on synthetic windows. Run in the build container (needs /root/reference):
    python tests/golden/make_ref_golden.py
The .npz travels to the GPU box, where tests/test_oracle_vs_ref.py checks the restated oracle (and through it the device) against it."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from gf2_loader import load   # noqa: E402
import gf2_ref as ref          # noqa: E402
import gf2_oracle as orc       # noqa: E402
import ref_cases as rc         # noqa: E402

gf2 = load(); abi = gf2.abi
synth = importlib.import_module("gf2_b200.synth")
out = {}
for s in range(4):
    c, p = rc.projection_case(s)
    r, J = ref.factor_eval(0, c, p)
    out[f"proj{s}_res"] = r
    for b, j in enumerate(J):
        out[f"proj{s}_J{b}"] = j
for s in range(3):
    smp, first, lb = rc.imu_samples(abi, s)
    rec = ref.imu_preintegrate_one(abi, smp, len(smp), first, lb, rc.IMU_NOISE)
    for f in ("sum_dt", "delta_p", "delta_q", "delta_v", "lin_ba", "lin_bg", "jacobian", "covariance"):
        out[f"imu{s}_{f}"] = rec[f][0]
    r, J = ref.factor_eval(1, rec, rc.imu_params(s, lb), extra=[9.7944])
    out[f"imu{s}_res"] = r
    for b, j in enumerate(J):
        out[f"imu{s}_J{b}"] = j
for s in range(3):
    smp, first, lin = rc.wheel_samples(abi, s)
    rec = ref.wheel_preintegrate_one(abi, smp, len(smp), first, lin, rc.WHEEL_NOISE)
    for f in ("sum_dt", "delta_p", "delta_q", "jacobian", "covariance", "vel_1", "gyr_1", "lin_vel", "lin_gyr"):
        out[f"wheel{s}_{f}"] = rec[f][0]
    for tag, dtd in (("a", 0.0), ("b", 0.004)):
        r, J = ref.factor_eval(2, rec, rc.wheel_params(s, lin, dtd))
        out[f"wheel{s}{tag}_res"] = r
        for b, j in enumerate(J):
            out[f"wheel{s}{tag}_J{b}"] = j
for s in range(3):
    for ct in (0, 1):
        c, p = rc.plane_case(s, ct)
        r, J = ref.factor_eval(3 + ct, c, p)
        out[f"plane{s}_{ct}_res"] = r
        for b, j in enumerate(J):
            out[f"plane{s}_{ct}_J{b}"] = j
x, jac = ref.pose_plus(np.array([0.1, -0.2, 0.3, 0.1, 0.2, -0.3, 0.9273618495495703]), np.array([0.01, 0.02, -0.03, 0.05, -0.04, 0.02]))
out["plus_x"] = x; out["plus_jac"] = jac

# marginalization of synthetic windows: MARGIN_OLD from the anchor prior, then MARGIN_SECOND_NEW and MARGIN_OLD on the resulting prior
opts = abi.default_opts()
for tag, kw in (("vio", dict(n_landmarks=120)), ("wheel", dict(n_landmarks=100, wheel=True, config_id=4))):
    w = synth.make_windows(2, **kw)
    orc.imu_preintegrate(w)
    if kw.get("wheel"):
        orc.wheel_preintegrate(w)
    for i in range(2):
        m0 = ref.marginalize_window(w, i, opts, mode=0)
        assert m0["status"] == 0, m0
        H, g, x0 = orc.prior_information(m0, 11)
        out[f"marg_{tag}{i}_old_H"] = H; out[f"marg_{tag}{i}_old_g"] = g; out[f"marg_{tag}{i}_old_n"] = np.array([m0["n"], m0["m"]])
        for k, v in x0.items():
            out[f"marg_{tag}{i}_old_x0_{k[0]}_{k[1]}"] = v
        # second generation: the new prior (already renamed) on the same window, both modes
        w2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
        P = 96
        w2["prior_rows"] = w2["prior_rows"].copy(); w2["prior_rows"][i] = m0["n"]
        J0 = np.zeros((2, P, P)); r0 = np.zeros((2, P)); J0[i, :m0["n"], :m0["n"]] = m0["J0"]; r0[i, :m0["n"]] = m0["r0"]
        blocks = np.zeros((2, 2 * 11 + 8), abi.PRIOR_BLOCK); blocks[i, :len(m0["blocks"])] = m0["blocks"]
        nb = w2["prior_nblocks"].copy(); nb[i] = len(m0["blocks"])
        w2["prior_J0"] = J0; w2["prior_r0"] = r0; w2["prior_blocks"] = blocks; w2["prior_nblocks"] = nb
        for mode, name in ((1, "second"), (0, "old2")):
            m1 = ref.marginalize_window(w2, i, opts, mode=mode)
            assert m1["status"] == 0, (name, m1)
            H, g, _ = orc.prior_information(m1, 11)
            out[f"marg_{tag}{i}_{name}_H"] = H; out[f"marg_{tag}{i}_{name}_g"] = g; out[f"marg_{tag}{i}_{name}_n"] = np.array([m1["n"], m1["m"]])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"), **out)
print("wrote", len(out), "arrays")
