import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import relearn_b200 as R
from relearn_b200 import _lib as L
from oracle import tensor_oracle as TO
from tests.test_gpu_update import _collect, _rel
ctx = R.Context(0)
for seed, E, T, reg in [(3, 64, 64, 1e-5), (4, 200, 100, 1e-5), (5, 33, 257, 1e-5), (3, 64, 64, 1e-1), (4, 200, 100, 1e-2)]:
    env, traj, net, params, host, valid = _collect(ctx, E, T, seed=seed, scale=1.5)
    rng = np.random.default_rng(seed + 100)
    adv = rng.normal(size=(T, E)).astype(np.float32)
    adv_d = ctx.to_device(adv)
    cfg = R.TrpoConfig(optimizer_config=R.ConjugateGradientOptimizerConfig(hpv_reg_coeff=reg))
    policy = R.Trpo(net, cfg)
    log = {}
    status = policy.update(traj, adv_d, log)
    new = net.get_weights()
    obs, act, a = host["obs"][valid], host["action"][valid], adv[valid]
    ocfg = TO.CgConfig(hpv_reg_coeff=reg)
    new64, log64 = TO.trpo_update(params, 5, 128, 2, obs, act, a, cfg=ocfg, dtype=torch.float64)
    new32, log32 = TO.trpo_update(params, 5, 128, 2, obs, act, a, cfg=ocfg, dtype=torch.float32)
    d, d64, d32 = new - params, new64 - params.astype(np.float64), new32 - params
    print(f"--- seed {seed} E {E} T {T} reg {reg} N {valid.sum()} status {status}")
    for k in ("entropy", "step_size", "cg_iterations", "num_backtracks", "loss_initial", "loss_final", "constraint_val_final"):
        print(f"  {k:22s} kernel {log[k]!r:24} f64 {log64[k]!r:24} f32 {log32[k]!r}")
    print(f"  delta rel err vs f64: kernel {_rel(d, d64):.3e}  torch-f32 {_rel(d32, d64):.3e}   kernel vs f32 {_rel(d, d32):.3e}")
    # direction comparison independent of backtracking scale
    def unit(v): return v / np.linalg.norm(v)
    print(f"  direction cos: kernel/f64 {float(unit(d.astype(np.float64)) @ unit(d64)):.8f} f32/f64 {float(unit(d32.astype(np.float64)) @ unit(d64)):.8f}")
