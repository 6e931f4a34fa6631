"""CPU tests of the LIO factor-construction oracle (oracle/gf2o_lio.cpp: searchNeighbors / computeNeighborhoodDistribution /
addSurfCostFactor restated) against an independent numpy restatement: brute-force neighbour search over the same voxels and
numpy.linalg.eigh. The reference holds no known-answer test for this function (parity unpinned at the reference level)."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def synth(gf2):
    return importlib.import_module("gf2_b200.synth")


def numpy_factors(scene, o):
    """Independent restatement; returns per keypoint (sorted neighbours, factor dict or None)."""
    size = o.size_voxel_map
    table = {tuple(k): scene["points"][i, :scene["n_points"][i]] for i, k in enumerate(scene["keys"].tolist())}
    lam_w, lam_n = abs(o.weight_alpha), abs(o.weight_neighborhood)
    lam_w, lam_n = lam_w / (lam_w + lam_n), lam_n / (lam_w + lam_n)
    tb = np.array(list(o.translation_begin)); R_IL = np.array(list(o.R_IL)).reshape(3, 3); t_IL = np.array(list(o.t_IL))
    out = []
    for kp in scene["keypoints"]:
        p = kp["point"]
        k0 = [int(c / size) for c in p]
        nb = o.nb_voxels_visited
        cand = []
        for x in range(k0[0] - nb, k0[0] + nb + 1):
            for y in range(k0[1] - nb, k0[1] + nb + 1):
                for z in range(k0[2] - nb, k0[2] + nb + 1):
                    blk = table.get((x, y, z))
                    if blk is not None and len(blk) >= o.threshold_voxel_capacity:
                        cand.append(blk)
        if not cand:
            out.append((np.zeros((0, 3)), None)); continue
        cand = np.concatenate(cand)
        d = np.linalg.norm(cand - p, axis=1)
        nbs = cand[np.argsort(d, kind="stable")[: o.max_number_neighbors]]
        fac = None
        if len(nbs) >= o.min_number_neighbors:
            c = nbs - nbs.mean(0)
            ev, V = np.linalg.eigh(c.T @ c)
            n = V[:, 0] / np.linalg.norm(V[:, 0])
            s1, s2, s3 = np.sqrt(abs(ev[2])), np.sqrt(abs(ev[1])), np.sqrt(abs(ev[0]))
            a2d = (s2 - s3) / s1
            loc = R_IL @ kp["raw_point"] + t_IL
            if n @ (tb - loc) < 0:
                n = -n
            w = lam_w * a2d ** o.power_planarity + lam_n * np.exp(-np.linalg.norm(nbs[0] - p) / (o.max_dist_to_plane_icp * o.min_number_neighbors))
            if abs((p - nbs[0]) @ n) < o.max_dist_to_plane_icp:
                fac = {"normal": n, "offset": -n @ nbs[0], "weight": w}
        out.append((nbs, fac))
    return out


@pytest.mark.parametrize("nb_visited,thr", [(1, 1), (2, 1), (1, 3)])
def test_lio_oracle_matches_numpy(gf2, oracle, synth, nb_visited, thr):
    scene = synth.lio_scene(1, n_map_points=15000, n_keypoints=400)
    o = gf2.abi.default_lio_opts(nb_voxels_visited=nb_visited, threshold_voxel_capacity=thr, translation_begin=scene["translation_begin"], rotation=scene["rotation"],
                                 translation=scene["translation"], max_num_residuals=100000)
    fac, alpha, nbs, nn = oracle.lio_build_factors(scene, o, want_neighbors=True)
    ref = numpy_factors(scene, o)
    k_fac = 0
    n_with, n_gate = 0, 0
    for k, (rn, rf) in enumerate(ref):
        assert nn[k] == len(rn)
        assert np.abs(nbs[k, :nn[k]] - rn).max(initial=0.0) == 0.0          # same neighbours in the same (distance) order
        if len(rn) >= o.min_number_neighbors:
            n_with += 1
        if rf is None:
            continue
        n_gate += 1
        f = fac[k_fac]; assert f["frame"] == k
        assert np.abs(f["normal"] - rf["normal"]).max() < 1e-7 and abs(f["offset"] - rf["offset"]) < 1e-7 and abs(f["weight"] - rf["weight"]) < 1e-9
        assert np.array_equal(f["p_body"], scene["keypoints"][k]["raw_point"]) and alpha[k_fac] == scene["keypoints"][k]["alpha_time"]
        k_fac += 1
    assert k_fac == len(fac) and 0 < n_gate <= n_with < len(ref)
    if nb_visited == 2:
        assert n_gate < n_with                                         # the point-to-plane gate rejected something (off-surface keypoints)


def test_lio_oracle_residual_cap_and_point_to_plane_model(gf2, oracle, synth):
    scene = synth.lio_scene(2, n_map_points=15000, n_keypoints=300)
    kw = dict(translation_begin=scene["translation_begin"], rotation=scene["rotation"], translation=scene["translation"])
    full, _, _, _ = oracle.lio_build_factors(scene, gf2.abi.default_lio_opts(max_num_residuals=100000, **kw))
    cap, _, _, _ = oracle.lio_build_factors(scene, gf2.abi.default_lio_opts(max_num_residuals=50, **kw))
    assert len(full) > 50 and len(cap) == 50 and np.array_equal(cap, full[:50])      # keypoint order, hard cap (:1058-1061)
    ptp, _, _, _ = oracle.lio_build_factors(scene, gf2.abi.default_lio_opts(max_num_residuals=100000, icp_model=gf2.abi.ICP_POINT_TO_PLANE, **kw))
    assert len(ptp) == len(full)
    # point_end = R^-1 (point - t) equals raw_point by construction of the scene
    assert np.abs(ptp["p_body"] - scene["keypoints"]["raw_point"][ptp["frame"]]).max() < 1e-12
