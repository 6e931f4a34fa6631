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


def test_device_resident_map_equals_sequential_add_point_to_map(gf2, oracle, synth):
    """gf2_lio_add_points == lidarodom::addPointToMap over the scan in order (LIO/liw/lio/lidarodom.cpp:1167-1237): three scans, the last
    with min_num_points = 3 (no new voxels, only voxels that already hold 3 points grow); the device map must hold exactly the same
    points in the same per-voxel order, and factors built on it must equal the oracle's on the downloaded snapshot."""
    scans = [synth.lio_scan(0, 30000), synth.lio_scan(1, 20000), synth.lio_scan(2, 25000, noise=0.03)]
    mins = [0, 0, 3]
    vox = {}
    h = gf2.Lio(max_voxels=40000, max_keypoints=2000)
    sizes = []
    for pts, mn in zip(scans, mins):
        synth.voxel_map_insert(vox, pts, 0.2, 20, 0.05, mn)
        h.add_points(pts, 0.2, 0.05, mn)
        sizes.append((len(vox), h.last_timing()["voxels"], h.last_timing()["new_voxels"]))
    assert sizes[0][0] == sizes[0][1] == sizes[0][2] and sizes[1][0] == sizes[1][1] and sizes[2][2] == 0 and sizes[2][1] == sizes[1][1]
    keys, npts, pts = h.get_map()
    assert len(keys) == len(vox)
    assert (np.diff(keys[:, 0].astype(np.int64) * 2 ** 32 + keys[:, 1].astype(np.int64) * 2 ** 16 + keys[:, 2].astype(np.int64)) > 0).all()   # ascending keys
    grew = 0
    for k, n, p in zip(keys.tolist(), npts, pts):
        ref = np.array(vox[tuple(k)])
        assert n == len(ref) and np.array_equal(p[:n], ref), k          # same points, same insertion order, bit-equal
        grew += n > 1
    assert grew > 1000 and npts.max() == 20                                 # voxels filled up to the capacity
    # factors on the resident map == oracle on its snapshot
    scene = synth.lio_scene(5, n_map_points=1000, n_keypoints=2000)
    scene.update(keys=keys, n_points=npts, points=pts)
    o = _opts(gf2, scene, max_num_residuals=100000)
    fac, alpha, nbs, nn = h.build_factors(scene["keypoints"], o, want_neighbors=True)
    rf, ra, rnbs, rnn = oracle.lio_build_factors(scene, o, want_neighbors=True)
    assert np.array_equal(nn, rnn) and all(np.array_equal(nbs[k, :nn[k]], rnbs[k, :nn[k]]) for k in range(len(nn)))
    _compare(fac, alpha, rf, ra)
    assert len(fac) > 500
    # capacity overflow is reported, not silently dropped
    small = gf2.Lio(max_voxels=100, max_keypoints=8)
    with pytest.raises(gf2.Gf2Error, match="capacity"):
        small.add_points(scans[0])
    small.close(); h.close()
