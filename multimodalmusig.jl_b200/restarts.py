"""Random-restart orchestration on resident counts (BASELINE config 5).

Mirrors scripts/run_mmctm.jl:75-182 of the reference:
  stage 1  fit R restarts (maxiter=1000, tol=1e-4), pick per modality the restart with the best
           final log-likelihood (pick_optimal_modality_models, :86-97);
  stage 2  build a model whose γ / Elnϕ / ϕ of modality m come from that modality's winner and
           refit (maxiter=1000, tol=1e-5) (seed_and_fit_restart, :113-134);
           pick_optimal_model (:136-147) ranks the stage-2 models by the mean dense rank of |ll|.
In the reference every stage-2 restart starts from the same γ, λ = 0, ν = 1 and LD_MMA is
deterministic, so the R stage-2 fits are identical; one is run here.

Restarts are independent: with several GPUs each rank (one process per GPU, each holding the full
counts) fits its slice of the restarts with NO communication during the fits, and the per-restart
(ll, γ) are gathered once at the end (`torch.distributed.all_gather_object`).
The README's simpler recipe ("fit many models and pick the best one", README.md:42) is
`MMCTM.fit_restarts` / `mmsig_mmctm_restarts` (arg-max ELBO).
"""
import numpy as np


def dense_rank(x):
    """StatsBase.denserank: 1-based rank, ties share a rank, no gaps."""
    _, inv = np.unique(np.asarray(x), return_inverse=True)
    return inv + 1


def pick_optimal_modality_models(ll):
    """ll: (R, M) final log-likelihoods -> per modality the index of the best restart (:86-97)."""
    return np.argmax(np.asarray(ll), axis=0)


def pick_optimal_model(ll):
    """(:136-147) arg-min over restarts of the mean dense rank of |ll| across modalities."""
    ll = np.asarray(ll)
    ranks = np.stack([dense_rank(np.abs(ll[:, i])) for i in range(ll.shape[1])], axis=1).astype(float)
    return int(np.argmin(ranks.mean(axis=1)))


def my_slice(R, rank, world):
    per = -(-R // world)
    return range(min(R, rank * per), min(R, (rank + 1) * per))


def fit_seed_models(model, gamma0s, maxiter=1000, tol=1e-4, rank=0, world=1, gather=None):
    """Stage 1 on `model` (an mmsig.MMCTM holding the full counts).  gamma0s: (R, G).
    Returns (ll (R, M), gammas (R, G), n_iter (R,)) for ALL restarts (gathered when world > 1;
    `gather(obj) -> list of objs` defaults to torch.distributed.all_gather_object)."""
    gamma0s = np.asarray(gamma0s, dtype=np.float64).reshape(-1, model.G)
    R = gamma0s.shape[0]
    mine = {}
    for r in my_slice(R, rank, world):
        model.set_state(gamma0s[r])
        hist = model.fit(maxiter=maxiter, tol=tol, verbose=False)
        mine[r] = (hist[-1].copy(), model.gamma, len(hist))
    if world > 1:
        if gather is None:
            import torch.distributed as dist

            def gather(obj):
                out = [None] * world
                dist.all_gather_object(out, obj)
                return out
        parts = gather(mine)
        mine = {}
        for p in parts:
            mine.update(p)
    ll = np.stack([mine[r][0] for r in range(R)])
    gammas = np.stack([mine[r][1] for r in range(R)])
    nit = np.asarray([mine[r][2] for r in range(R)])
    return ll, gammas, nit


def seed_gamma(model, gammas, winners):
    """γ whose modality-m block comes from restart winners[m] (:123-129)."""
    go = np.cumsum([0] + [k * v for k, v in zip(model.K, model.V)])
    return np.concatenate([gammas[winners[m], go[m]:go[m + 1]] for m in range(model.M)])


def fit_model(model, gamma0s, rank=0, world=1, gather=None, stage1_tol=1e-4, stage2_tol=1e-5, maxiter=1000):
    """fit_model (:163-182).  Leaves the stage-2 fit in `model`; returns a dict of the stage results."""
    ll1, gammas, nit1 = fit_seed_models(model, gamma0s, maxiter=maxiter, tol=stage1_tol, rank=rank, world=world,
                                        gather=gather)
    winners = pick_optimal_modality_models(ll1)
    g2 = seed_gamma(model, gammas, winners)
    model.set_state(g2)
    hist2 = model.fit(maxiter=maxiter, tol=stage2_tol, verbose=False)
    return {"stage1_ll": ll1, "stage1_iterations": nit1, "winners": winners, "stage2_ll": hist2[-1].copy(),
            "stage2_iterations": len(hist2), "elbo": model.elbo}
