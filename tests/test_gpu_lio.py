"""GPU tests of the LIO factor construction (gf2_lio_build_factors: searchNeighbors + computeNeighborhoodDistribution + the
residual gate of lidarodom::addSurfCostFactor) through the C ABI against the CPU oracle (oracle/gf2o_lio.cpp, itself pinned against
numpy in tests/test_oracle_lio.py). Neighbour lists (index work) must be identical; normals / offsets / weights within 1e-10."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth(gf2):
    if gf2.device_count() < 1:
        pytest.fail("no CUDA device: the hot path has no CPU fallback")
    return importlib.import_module("gf2_b200.synth")


def _opts(gf2, scene, **kw):
    return gf2.abi.default_lio_opts(translation_begin=scene["translation_begin"], rotation=scene["rotation"], translation=scene["translation"], **kw)


def _lio(gf2, scene, n_kp=None):
    h = gf2.Lio(max_voxels=len(scene["keys"]) + 8, max_keypoints=n_kp or len(scene["keypoints"]), max_points_per_voxel=scene["max_points_per_voxel"])
    h.set_map(scene["keys"], scene["n_points"], scene["points"])
    return h


def _compare(fac, alpha, rf, ra):
    assert len(fac) == len(rf)
    assert np.array_equal(fac["frame"], rf["frame"])                     # which keypoints produced a residual: index work, exact
    assert np.array_equal(fac["p_body"], rf["p_body"]) or np.abs(fac["p_body"] - rf["p_body"]).max() < 1e-12
    assert np.array_equal(alpha, ra)
    assert np.abs(fac["normal"] - rf["normal"]).max() < 1e-10
    assert np.abs(fac["offset"] - rf["offset"]).max() < 1e-10
    assert np.abs(fac["weight"] - rf["weight"]).max() < 1e-12


@pytest.mark.parametrize("nb_visited,thr,model", [(1, 1, 0), (2, 1, 0), (1, 3, 1), (3, 1, 0)])
def test_lio_factors_match_oracle(gf2, oracle, synth, nb_visited, thr, model):
    scene = synth.lio_scene(3, n_map_points=40000, n_keypoints=2000)
    o = _opts(gf2, scene, nb_voxels_visited=nb_visited, threshold_voxel_capacity=thr, icp_model=model, max_num_residuals=100000)
    h = _lio(gf2, scene)
    fac, alpha, nbs, nn = h.build_factors(scene["keypoints"], o, want_neighbors=True)
    rf, ra, rnbs, rnn = oracle.lio_build_factors(scene, o, want_neighbors=True)
    assert np.array_equal(nn, rnn)
    for k in range(len(nn)):
        assert np.array_equal(nbs[k, :nn[k]], rnbs[k, :nn[k]]), k         # same neighbours in the same order (nb 3: first 200 voxels only)
    _compare(fac, alpha, rf, ra)
    assert 100 < len(fac) < len(scene["keypoints"])
    assert h.last_timing()["keypoints"] == len(scene["keypoints"])
    h.close()


def test_lio_residual_cap_small_neighbourhoods_and_edge_cases(gf2, oracle, synth):
    scene = synth.lio_scene(4, n_map_points=20000, n_keypoints=500)
    h = _lio(gf2, scene)
    for kw in (dict(max_num_residuals=37), dict(max_number_neighbors=8, min_number_neighbors=5, max_num_residuals=100000),
               dict(num_closest_neighbors=3, max_num_residuals=100000), dict(max_num_residuals=0)):
        o = _opts(gf2, scene, **kw)
        fac, alpha, _, _ = h.build_factors(scene["keypoints"], o)
        rf, ra, _, _ = oracle.lio_build_factors(scene, o)
        _compare(fac, alpha, rf, ra)
    # empty map: no neighbours, no residuals
    h.set_map(np.zeros((0, 3), np.int16), np.zeros(0, np.int32), np.zeros((0, scene["max_points_per_voxel"], 3)))
    fac, _, _, nn = h.build_factors(scene["keypoints"], _opts(gf2, scene), want_neighbors=True)
    assert len(fac) == 0 and not nn.any()
    with pytest.raises(gf2.Gf2Error, match="appears twice"):
        h.set_map(np.array([[1, 2, 3], [1, 2, 3]], np.int16), np.array([1, 1], np.int32), np.zeros((2, scene["max_points_per_voxel"], 3)))
    with pytest.raises(gf2.Gf2Error, match="max_number_neighbors"):
        h.build_factors(scene["keypoints"], _opts(gf2, scene, max_number_neighbors=64))
    h.close()
