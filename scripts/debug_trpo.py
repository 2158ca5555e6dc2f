import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import optim as OPT, policy as P, sampler as S
from rllab_b200 import ops, _lib as L
from rllab_b200.misc import logger
logger.set_quiet(True)
sys.path.insert(0, 'tests')
import test_gpu_algos as TA
algo = TA._algo("cartpole", "trpo", 1024, 50, 32)
algo.start_worker(); algo.init_opt()
paths = algo.sampler.obtain_samples(0)
sd = algo.sampler.process_samples(0, paths)
b = sd.lane_batch
pol = algo.policy
theta0 = pol.theta32.double().cpu().numpy()
batch = S.batch_from_traj(b.to_numpy(), b.adv.cpu().numpy())
dims = P.Dims(b.O, (32, 32), b.A)
dev = b.device
z = lambda n=dims.P: torch.zeros(n, dtype=torch.float64, device=dev)
g, x, r, p, zz, Hx, step, st, info, out = z(), z(), z(), z(), z(), z(), z(), z(4), z(2), z(3)
sc = 1.0 / b.B
ops.grad(L.LOSS_TRPO, pol.theta32, pol.dims, 1e-6, b, g)
g_ref = P.grad_surr(theta0, batch, dims, "trpo")
print("g rel", np.abs(g.cpu().numpy() - g_ref).max() / np.abs(g_ref).max())
ops.cg_init(g, x, r, p, st)
xr = np.zeros(dims.P); rr = g_ref.copy(); pr = g_ref.copy(); rd = rr.dot(rr)
for i in range(10):
    ops.fvp(pol.theta32, pol.dims, 1e-6, b, p, 1e-5, 1.0, zz)
    zr = P.fvp(theta0, batch, pr, dims, 1e-5)
    print(i, "Hp rel (dev vs oracle at dev p)", np.abs(zz.cpu().numpy() - P.fvp(theta0, batch, p.cpu().numpy(), dims, 1e-5)).max() / np.abs(zr).max(),
          "p rel", np.abs(p.cpu().numpy() - pr).max() / np.abs(pr).max())
    ops.cg_step(zz, x, r, p, st)
    v = rd / pr.dot(zr); xr += v * pr; rr -= v * zr; nrd = rr.dot(rr); pr = rr + nrd / rd * pr; rd = nrd
print("x rel", np.abs(x.cpu().numpy() - xr).max() / np.abs(xr).max(), "rdotr dev/ref", st.cpu().numpy()[0], rd)
ops.fvp(pol.theta32, pol.dims, 1e-6, b, x, 1e-5, 1.0, Hx)
ops.trpo_step_size(x, Hx, 0.01, step, info)
Hxr = P.fvp(theta0, batch, xr, dims, 1e-5)
print("beta dev/ref", info.cpu().numpy(), np.sqrt(0.02 / (xr.dot(Hxr) + 1e-8)), xr.dot(Hxr))
prev = pol.theta64.clone()
for k in range(4):
    ratio = 0.8 ** k
    ops.axpy_params(prev, step, ratio, pol.theta64, pol.theta32)
    ops.loss_kl(L.LOSS_TRPO, pol.theta32, pol.dims, 1e-6, b, out)
    th = theta0 - ratio * np.sqrt(0.02 / (xr.dot(Hxr) + 1e-8)) * xr
    th_dev = pol.theta32.double().cpu().numpy()
    print(k, "dev loss/kl", out.cpu().numpy()[:2], "oracle@ref-theta", P.surr_loss_trpo(th, batch, dims), P.kl_stats(th, batch, dims)[0],
          "oracle@dev-theta32", P.surr_loss_trpo(th_dev, batch, dims), P.kl_stats(th_dev, batch, dims)[0])
print("loss_before dev", end=" ")
ops.axpy_params(prev, step, 0.0, pol.theta64, pol.theta32)
ops.loss_kl(L.LOSS_TRPO, pol.theta32, pol.dims, 1e-6, b, out); print(out.cpu().numpy(), P.surr_loss_trpo(theta0, batch, dims))
