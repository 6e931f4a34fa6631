"""Factor sharding for the multi-GPU mode of SURVEY.md 8(e): every rank keeps every window's frame states, IMU / wheel
factors and prior, but only its share of the landmarks (with all their observations) and LiDAR plane factors. Landmark l goes
to rank l mod nranks — track lengths cycle with l, so the load is even — and plane k to rank k mod nranks. Pure index
bookkeeping (bit-exact by construction); the reduced systems of the shards add up to the full one."""
import numpy as np


def shard_windows(w, rank, nranks):
    """Return a window dict holding rank's landmarks / planes of every window of `w` (same capacities as `w`)."""
    out = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in w.items()}
    n = w["para_pose"].shape[0]
    for k in ("inv_depth", "start_frame", "track_len", "fixed", "obs"):
        out[k][...] = 0
    out["n_landmarks"] = np.zeros(n, np.int32)
    for i in range(n):
        L = int(w["n_landmarks"][i])
        tl = w["track_len"][i, :L]
        beg = np.concatenate([[0], np.cumsum(tl)[:-1]]).astype(np.int64)
        mine = np.arange(rank, L, nranks)
        m = len(mine)
        out["n_landmarks"][i] = m
        for k in ("inv_depth", "start_frame", "track_len", "fixed"):
            out[k][i, :m] = w[k][i, mine]
        o = 0
        for l in mine:
            out["obs"][i, o:o + tl[l]] = w["obs"][i, beg[l]:beg[l] + tl[l]]
            o += tl[l]
    if w.get("max_planes", 0) > 0:
        out["planes"][...] = 0
        out["n_planes"] = np.zeros(n, np.int32)
        for i in range(n):
            P = int(w["n_planes"][i])
            mine = np.arange(rank, P, nranks)
            out["n_planes"][i] = len(mine)
            out["planes"][i, :len(mine)] = w["planes"][i, mine]
    return out


def gather_landmarks(full_n_landmarks, shard_inv_depths, nranks):
    """Inverse of the landmark partition: per-rank inverse depth arrays [nranks][n][Lm] -> full [n][Lm]."""
    n = len(full_n_landmarks)
    out = np.zeros_like(shard_inv_depths[0])
    for i in range(n):
        L = int(full_n_landmarks[i])
        for r in range(nranks):
            mine = np.arange(r, L, nranks)
            out[i, mine] = shard_inv_depths[r][i, :len(mine)]
    return out
