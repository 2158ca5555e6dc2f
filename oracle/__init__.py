"""CPU oracle for the rllab hot path (TEST INFRASTRUCTURE -- not product code).

Every function in this package is a float64 NumPy restatement of the reference
(rll/rllab @ ba78e4c) for the path named in BASELINE.json, citing the reference
file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``rllab_b200`` never does (it fails loudly without the CUDA
library instead).

Pinning status (see DESIGN.md "Oracle"):
  * pinned against the reference's own code run in the build container
    (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``):
    discount_cumsum/GAE/process_samples, center_advantages, explained_variance,
    LinearFeatureBaseline, krylov.cg, ConjugateGradientOptimizer.optimize,
    DiagonalGaussian kl/log_likelihood/entropy, NormalizedEnv, PointEnv,
    rollout(), truncate_paths.
  * PARITY UNPINNED (third-party arithmetic absent from /root/reference and not
    installable here): Theano autodiff of the MLP (checked instead against
    finite differences and torch.autograd in tests), Lasagne Adam, Box2D
    CartPole, gym Pendulum-v0, MuJoCo-1.31 Swimmer/Hopper (their MODEL constants
    are pinned to vendor/mujoco_models/*.xml by tests/golden/reference_mujoco_models.json;
    the dynamics algorithm is a restatement of the published pipeline).
"""
