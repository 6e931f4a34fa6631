"""Factor-sharded multi-GPU solve (BASELINE.json config 4): launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/run_sharded.py [--windows B]
Every rank holds all frame states + IMU/wheel/prior and its l mod N landmarks / planes; one NCCL all-reduce of the visual block
system per linearisation. Rank 0 also solves the unsharded windows on its own GPU and checks parity; prints one JSON line."""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import torch.distributed as dist
from gf2_loader import load

ap = argparse.ArgumentParser(); ap.add_argument("--windows", type=int, default=64); ap.add_argument("--landmarks", type=int, default=1000)
ap.add_argument("--planes", type=int, default=5000); ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--marginalize", action="store_true", help="after the solve: MARGIN_OLD on every rank (partial systems all-reduced), then a second solve with the resident prior")
ap.add_argument("--free-wheel", action="store_true", help="body_T_wheel free (estimate_wheel_extrinsic: 1): one more block row of the reduced system")
args = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gf2 = load(); synth = importlib.import_module("gf2_b200.synth"); shard = importlib.import_module("gf2_b200.shard")
B = args.windows
distinct = min(B, 8)
base = synth.make_windows(distinct, config_id=4, n_landmarks=args.landmarks, wheel=True, n_planes=args.planes)
w = {k: (np.concatenate([v] * ((B + distinct - 1) // distinct))[:B] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == distinct and k not in ("imu_noise", "wheel_noise") else v) for k, v in base.items()}

def mk(d):   # config 4: IMU + wheel (both preintegrated on the device from raw samples) + projection + LiDAR plane factors
    return gf2.Solver(B, d["n_frames"], d["max_landmarks"], d["max_obs"], max_planes=d["max_planes"], max_imu_samples=d["n_imu_samples"],
                      use_wheel=True, max_wheel_samples=d["n_wheel_samples"], device=local)
mine = shard.shard_windows(w, rank, world)
s = mk(mine)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.from_numpy(gf2.Solver.comm_unique_id()).cuda())
dist.broadcast(uid, 0)
s.comm_init(rank, world, uid.cpu().numpy())
opts = gf2.abi.default_opts()
if args.free_wheel:
    opts.const_mask = gf2.abi.CONST_EX_POSE | gf2.abi.CONST_TD | gf2.abi.CONST_WHEEL_INTRINSIC | gf2.abi.CONST_TD_WHEEL
s.upload(mine, preintegrate="device"); s.snapshot(B)
summ = s.solve(opts, B)   # warm-up
times = []
for _ in range(args.steps):
    s.restore(B); torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    summ = s.solve(opts, B)
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
tt = torch.tensor([min(times)], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
st = s.get_states(B); lam = s.get_landmarks(B)
marg = None
if args.marginalize:   # every rank marginalizes its landmark shard; the new prior is the same on all ranks
    mst, mm = s.marginalize(opts, 0, B)
    pr = s.get_prior(B)
    summ2 = s.solve(opts, B)          # the next solve uses the resident prior (same window data: the prior's frames are shifted by one, a consistency run only)
    st2 = s.get_states(B)
    marg = (mst, mm, pr, summ2, st2)
lam_t = torch.from_numpy(lam).cuda(); gathered = [torch.zeros_like(lam_t) for _ in range(world)]
dist.all_gather(gathered, lam_t)
if rank == 0:
    ref = mk(w); ref.upload(w, preintegrate="device"); ref.snapshot(B)
    rs = ref.solve(opts, B)
    t1 = []
    for _ in range(args.steps):
        ref.restore(B); torch.cuda.synchronize(); t0 = time.perf_counter(); rs = ref.solve(opts, B); torch.cuda.synchronize(); t1.append(time.perf_counter() - t0)
    rst = ref.get_states(B); rlam = ref.get_landmarks(B)
    lam_full = shard.gather_landmarks(w["n_landmarks"], [g.cpu().numpy() for g in gathered], world)
    mout = {}
    if marg is not None:
        rmst, rmm = ref.marginalize(opts, 0, B)
        rpr = ref.get_prior(B)
        rs2 = ref.solve(opts, B); rst2 = ref.get_states(B)
        mst, mm, pr, summ2, st2 = marg
        hd = gd = 0.0
        for i in range(B):
            n = int(rpr["prior_rows"][i]); assert int(pr["prior_rows"][i]) == n
            Jr, Js = rpr["prior_J0"][i, :n, :n], pr["prior_J0"][i, :n, :n]
            Hr, Hs = Jr.T @ Jr, Js.T @ Js
            gr, gs = Jr.T @ rpr["prior_r0"][i, :n], Js.T @ pr["prior_r0"][i, :n]
            hd = max(hd, float(np.abs(Hs - Hr).max() / np.abs(Hr).max())); gd = max(gd, float(np.abs(gs - gr).max() / max(np.abs(gr).max(), 1e-300)))
        mout = {"marg_status_equal": bool((mst == rmst).all() and (mst == 0).all()), "marg_m_equal": bool((mm == rmm).all()), "marg_H_rel_diff": hd, "marg_g_rel_diff": gd,
                "marg_blocks_equal": bool(all(np.array_equal(pr["prior_blocks"][f], rpr["prior_blocks"][f]) for f in ("kind", "index", "offset"))),
                "second_solve_pose_diff": float(np.abs(st2["para_pose"] - rst2["para_pose"]).max()), "second_solve_iterations_equal": bool((summ2["iterations"] == rs2["iterations"]).all())}
    out = {"n_gpus": world, "windows": B, "landmarks": args.landmarks, "planes": args.planes,
           "sharded_ms": 1e3 * float(tt.item()), "single_gpu_ms": 1e3 * min(t1), "nccl_ms_rank0": s.last_timing()["nccl_ms"],
           "sharded_solves_per_s": B / float(tt.item()), "single_solves_per_s": B / min(t1),
           "pose_diff": float(np.abs(st["para_pose"] - rst["para_pose"]).max()), "speedbias_diff": float(np.abs(st["para_speedbias"] - rst["para_speedbias"]).max()),
           "inv_depth_diff": float(np.abs(lam_full - rlam).max()), "free_wheel": bool(args.free_wheel),
           "ex_wheel_diff": float(np.abs(st["ex_pose_wheel"] - rst["ex_pose_wheel"]).max()), "ex_wheel_moved": float(np.abs(rst["ex_pose_wheel"] - w["ex_pose_wheel"]).max()),
           "iterations_equal": bool((summ["iterations"] == rs["iterations"]).all()), "termination_equal": bool((summ["termination"] == rs["termination"]).all()),
           "final_cost_rel_diff": float((np.abs(summ["final_cost"] - rs["final_cost"]) / rs["final_cost"]).max())}
    out.update(mout)
    print(json.dumps(out))
    if mout:
        assert mout["marg_status_equal"] and mout["marg_m_equal"] and mout["marg_blocks_equal"] and mout["marg_H_rel_diff"] <= 1e-8 and mout["marg_g_rel_diff"] <= 1e-5 and mout["second_solve_pose_diff"] < 1e-6, out
    assert out["iterations_equal"] and out["termination_equal"] and out["pose_diff"] < 1e-6 and out["inv_depth_diff"] < 1e-6, out
s.close()
dist.destroy_process_group()
