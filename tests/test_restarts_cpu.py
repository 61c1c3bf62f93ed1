"""Host logic of the restart orchestration (scripts/run_mmctm.jl:86-147), no GPU needed."""
import numpy as np

from mmsig import restarts


def test_dense_rank_and_picks():
    assert restarts.dense_rank([3.0, 1.0, 3.0, 2.0]).tolist() == [3, 1, 3, 2]
    ll = np.array([[-4.0, -3.5], [-3.9, -3.6], [-4.1, -3.4]])
    assert restarts.pick_optimal_modality_models(ll).tolist() == [1, 2]
    # ranks of |ll|: col0 [2,1,3], col1 [2,3,1] -> means [2,2,2] -> first minimum
    assert restarts.pick_optimal_model(ll) == 0
    ll2 = np.array([[-4.0, -3.5], [-3.9, -3.4], [-4.1, -3.6]])
    assert restarts.pick_optimal_model(ll2) == 1


def test_slices_cover_all_restarts():
    for R in (1, 7, 64):
        for world in (1, 2, 8):
            got = sorted(r for k in range(world) for r in restarts.my_slice(R, k, world))
            assert got == list(range(R))


def test_fit_model_with_a_fake_model():
    class Fake:
        K, V, M, G = [2, 1], [3, 2], 2, 8

        def set_state(self, g):
            self.g = np.asarray(g, float).copy()

        def fit(self, maxiter, tol, verbose):
            self.elbo = -float(self.g.sum())
            return np.array([[-self.g[:6].sum(), -self.g[6:].sum()]])

        gamma = property(lambda s: s.g + 0.5)

    g0 = np.array([[1.0] * 6 + [9.0] * 2, [5.0] * 6 + [1.0] * 2, [3.0] * 8])
    out = restarts.fit_model(Fake(), g0)
    assert out["winners"].tolist() == [0, 1]
    # stage 2 starts from modality 0 of restart 0 and modality 1 of restart 1 (their final gammas)
    assert np.allclose(out["stage2_ll"], [-(1.5 * 6), -(1.5 * 2)])
    # two "ranks" emulated through the gather hook give the same result as one rank
    ll_all, gam_all, nit_all = restarts.fit_seed_models(Fake(), g0)
    for rank in (0, 1):
        other = {r: (ll_all[r], gam_all[r], int(nit_all[r])) for r in restarts.my_slice(3, 1 - rank, 2)}
        ll, gam, nit = restarts.fit_seed_models(Fake(), g0, rank=rank, world=2, gather=lambda obj: [obj, other])
        assert np.array_equal(ll, ll_all) and np.array_equal(gam, gam_all) and np.array_equal(nit, nit_all)
