"""Classical predictors.  Parity vs upstream socialforce / rvo2 / pykalman is UNPINNED (packages
not vendored, not installable, no reference tests): the checker is the CPU restatement in
oracle/ (numpy float64 / plain C float), anchored on the reference's call sites."""
import numpy as np
import pytest

from oracle import classical_oracle as C


def _scenes(num_scenes, max_peds, seed, min_peds=1):
    rng = np.random.RandomState(seed)
    sizes = rng.randint(min_peds, max_peds + 1, size=num_scenes)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    A = int(offs[-1])
    pos = rng.randn(A, 2) * 2.0
    ang = rng.rand(A) * 2 * np.pi
    spd = 0.4 + rng.rand(A) * 1.2
    vel = np.stack([spd * np.cos(ang), spd * np.sin(ang)], axis=1)
    goal = pos + vel * 4.8 + rng.randn(A, 2) * 0.3
    return offs, pos, vel, goal, spd


# ---- Kalman: host C++ behind the C ABI vs the numpy restatement (CPU test) --------------------
def test_kalman_matches_numpy_restatement():
    from trajnetplusplusbaselines_b200.classical import kalman
    rng = np.random.RandomState(3)
    tracks = []
    for T in (9, 9, 8, 5, 3, 2, 9):
        v = rng.randn(2) * 0.3
        tracks.append(np.arange(T)[:, None] * v[None] + rng.randn(2) + rng.randn(T, 2) * 0.03)
    pred = kalman.predict_tracks(tracks, n_predict=12, n_samples=0)
    for p, t in zip(pred, tracks):
        ref = C.kalman_predict_track(t, n_predict=12)
        assert np.abs(p - ref).max() < 1e-8          # float64 both sides; pinv vs Gauss-Jordan


def test_kalman_host_threads_do_not_change_results(monkeypatch):
    """Tracks are independent: the host-thread fan-out of tb2_kalman_predict (>= 64 tracks per thread) returns the
    single-thread result bit for bit, and a bad track is reported whichever thread meets it."""
    from trajnetplusplusbaselines_b200.classical import kalman
    rng = np.random.RandomState(7)
    tracks = [np.cumsum(rng.randn(rng.randint(2, 10), 2) * 0.3, axis=0) + rng.randn(2) * 5 for _ in range(700)]
    monkeypatch.setenv("TB2_KALMAN_THREADS", "1")
    one = kalman.predict_tracks(tracks, n_predict=12, n_samples=0)
    monkeypatch.setenv("TB2_KALMAN_THREADS", "5")
    many = kalman.predict_tracks(tracks, n_predict=12, n_samples=0)
    assert all(np.array_equal(a, b) for a, b in zip(one, many))
    tracks[650] = tracks[650][:1]                       # a single observation (kalman.py:28-29)
    with pytest.raises(RuntimeError):
        kalman.predict_tracks(tracks, n_predict=12, n_samples=0)


def test_kalman_config0_64_scenes_of_5_peds():
    """BASELINE configs[0]: 64 synthetic 5-ped scenes, obs=9 pred=12, through `predict`."""
    from trajnetplusplusbaselines_b200.classical import kalman
    from trajnetplusplusbaselines_b200.data import TrackRow
    rng = np.random.RandomState(0)
    for scene in range(64):
        paths = []
        for p in range(5):
            v = rng.randn(2) * 0.3
            x0 = rng.randn(2) * 2
            first = 0 if p == 0 else rng.randint(0, 4)
            paths.append([TrackRow(f, p, *(x0 + v * f + rng.randn(2) * 0.02)) for f in range(first, 9)])
        out = kalman.predict(paths, n_predict=12, obs_length=9, n_samples=0)
        prim, neigh = out[0]
        assert prim.shape == (12, 2) and neigh.shape == (12, 4, 2)
        ref = C.kalman_predict_track(np.array([(r.x, r.y) for r in paths[0]]))
        assert np.abs(prim - ref).max() < 1e-8
    # the reference's mean-of-5 sampled variant is a random variable around that expectation
    np.random.seed(1)
    out_s = kalman.predict(paths, n_predict=12, obs_length=9, n_samples=5)
    assert np.abs(out_s[0][0] - prim).max() < 0.5


def test_constant_velocity_exact():
    from trajnetplusplusbaselines_b200.classical import constant_velocity
    from trajnetplusplusbaselines_b200.data import TrackRow
    paths = [[TrackRow(f, 0, 0.1 * f, 1.0 - 0.2 * f) for f in range(9)],
             [TrackRow(f, 1, 3.0, 0.5 * f) for f in range(9)]]
    out = constant_velocity.predict(paths)
    prim, neigh = out[0]
    xy = np.array([[[0.1 * f, 1.0 - 0.2 * f], [3.0, 0.5 * f]] for f in range(9)])
    ref = C.constant_velocity(xy)
    assert np.array_equal(prim, ref[:, 0]) and np.array_equal(neigh, ref[:, 1:])


def test_adapter_initial_state_matches_restatement():
    from trajnetplusplusbaselines_b200.classical.common import initial_states
    from trajnetplusplusbaselines_b200.data import TrackRow
    rng = np.random.RandomState(5)
    xy = np.cumsum(rng.randn(9, 4, 2) * 0.3, axis=0)
    xy[:6, 2] = np.nan          # late entry: 3 observations -> stride 2
    xy[:8, 3] = np.nan          # single observation -> stride 0, destination = position
    paths = [[TrackRow(f, p, xy[f, p, 0], xy[f, p, 1]) for f in range(9) if not np.isnan(xy[f, p, 0])]
             for p in range(4)]
    st, sp = initial_states(paths, 8, 12)
    st_o, sp_o, keep = C.adapter_initial_states(xy, 12)
    assert np.allclose(st, st_o, atol=1e-12) and np.allclose(sp, sp_o, atol=1e-12)


# ---- GPU: social force / ORCA persistent kernels vs the CPU restatements ----------------------
@pytest.mark.gpu
def test_social_force_matches_restatement():
    from trajnetplusplusbaselines_b200.classical import socialforce
    offs, pos, vel, goal, spd = _scenes(40, 12, seed=1)
    state = np.concatenate([pos, vel, goal], axis=1)
    out = socialforce.simulate_batch(state, offs.tolist()).cpu().numpy()
    assert out.shape == (12, len(pos), 2)
    worst = 0.0
    for b in range(len(offs) - 1):
        s, e = offs[b], offs[b + 1]
        ref = C.sf_simulate(state[s:e])
        worst = max(worst, np.abs(out[:, s:e] - ref).max())
    # float64 on both sides; exp/cos implementations differ in the last ulp
    assert worst < 1e-9, worst


@pytest.mark.gpu
@pytest.mark.parametrize("params", [(0.5, 2.1, 0.3), (0.5, 5.0, 0.3)])
def test_social_force_predict_boundary(params):
    from trajnetplusplusbaselines_b200.classical import socialforce
    from trajnetplusplusbaselines_b200.data import TrackRow
    paths = [[TrackRow(f, 1, 0.1, 6.2 - 0.4 * (f - 1)) for f in range(1, 10)],
             [TrackRow(f, 2, 0.0, -6.2 + 0.4 * (f - 1)) for f in range(1, 10)],
             [TrackRow(f, 3, 2.0 + 0.1 * f, 1.0) for f in range(5, 10)]]
    out = socialforce.predict(paths, sf_params=list(params))
    prim, neigh = out[0]
    assert prim.shape == (12, 2) and neigh.shape == (12, 2, 2)
    xy = np.full((9, 3, 2), np.nan)
    for p, path in enumerate(paths):
        for r in path:
            xy[r.frame - 1, p] = (r.x, r.y)
    st, _, _ = C.adapter_initial_states(xy, 12)
    ref = C.sf_simulate(st, tau=params[0], v0=params[1], sigma=params[2])
    assert np.abs(prim - ref[:, 0]).max() < 1e-9
    # first sample is after ONE 0.05 s step (reference quirk, socialforce.py:93-95)
    assert abs(prim[0, 1] - 3.0) < 0.06


@pytest.mark.gpu
def test_orca_matches_c_restatement_bit_exact():
    from oracle.build_c import orca_simulate
    from trajnetplusplusbaselines_b200.classical import orca
    offs, pos, vel, goal, spd = _scenes(60, 14, seed=2)
    out = orca.simulate_batch(pos, vel, goal, spd, offs.tolist()).cpu().numpy()
    assert out.shape == (12, len(pos), 2)
    mism = 0
    worst = 0.0
    for b in range(len(offs) - 1):
        s, e = offs[b], offs[b + 1]
        ref = orca_simulate(pos[s:e], vel[s:e], goal[s:e], spd[s:e])
        mism += int((out[:, s:e] != ref).sum())
        worst = max(worst, float(np.abs(out[:, s:e] - ref).max()))
    # float on both sides, FMA contraction off on both: identical operation sequence
    assert worst < 1e-5, worst
    assert mism == 0, "%d of %d coordinates differ (max %.3g)" % (mism, out.size, worst)


@pytest.mark.gpu
def test_orca_predict_boundary_head_on():
    from trajnetplusplusbaselines_b200.classical import orca
    from trajnetplusplusbaselines_b200.data import TrackRow
    paths = [[TrackRow(f, 1, 0.1, 6.2 - 0.4 * (f - 1)) for f in range(1, 10)],
             [TrackRow(f, 2, 0.0, -6.2 + 0.4 * (f - 1)) for f in range(1, 10)]]
    prim, neigh = orca.predict(paths)[0]
    assert prim.shape == (12, 2) and neigh.shape == (12, 1, 2)
    # the two agents must not collide (radius 0.4 each)
    d = np.linalg.norm(prim - neigh[:, 0], axis=1)
    assert d.min() > 0.75


@pytest.mark.gpu
def test_classical_large_batch_properties():
    """Many scenes in lockstep: result of a scene is independent of its batch neighbours."""
    from trajnetplusplusbaselines_b200.classical import orca, socialforce
    offs, pos, vel, goal, spd = _scenes(2000, 20, seed=3, min_peds=20)
    state = np.concatenate([pos, vel, goal], axis=1)
    full = socialforce.simulate_batch(state, offs.tolist(), n_steps=40).cpu().numpy()
    cut = int(offs[1000])
    part = socialforce.simulate_batch(state[cut:], (offs[1000:] - cut).tolist(), n_steps=40).cpu().numpy()
    assert np.array_equal(full[:, cut:], part)
    o_full = orca.simulate_batch(pos, vel, goal, spd, offs.tolist(), n_steps=41).cpu().numpy()
    o_part = orca.simulate_batch(pos[:cut], vel[:cut], goal[:cut], spd[:cut], offs[:1001].tolist(), n_steps=41).cpu().numpy()
    assert np.array_equal(o_full[:, :cut], o_part)
    assert np.isfinite(o_full).all()
