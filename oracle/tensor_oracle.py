"""Torch-CPU restatement of the tensor side of relearn's update path.  TEST INFRASTRUCTURE ONLY.

The reference computes these with `tch` -> libtorch 1.12 (not vendored, cannot be built here:
"parity unpinned" for MLP/Adam numerics beyond the reference's own closed-form KATs, which
tests/test_oracle_golden.py replays).  Every function follows the cited reference code op for op,
on torch 2.11 CPU tensors, with autograd doing the first- and second-order backward passes exactly
as `Tensor::run_backward` does in the reference.

dtype: the reference runs f32 tensors with f64 host scalars; pass dtype=torch.float64 to get a
high-precision run of the same algorithm (used to bound the f32 rounding noise of both sides).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

F32_MIN = float(torch.finfo(torch.float32).min)


# ------------------------------------------------------------------------------------------------
# modules (src/torch/modules/ff/mlp.rs:139-151, linear.rs:118-123)
# ------------------------------------------------------------------------------------------------
def unflatten_mlp(flat: torch.Tensor, n_in: int, hidden, n_out: int):
    """Module::variables() order: [W, b] per Linear, layers in order (linear.rs:108-110, mlp.rs:126-128); `hidden` is
    MlpConfig::hidden_sizes (an int for the default single hidden layer)."""
    sizes = [hidden] if isinstance(hidden, (int, np.integer)) else list(hidden)
    o = 0
    shapes = []
    prev = n_in
    for h in sizes + [n_out]:
        shapes += [(h, prev), (h,)]
        prev = h
    out = []
    for s in shapes:
        n = int(np.prod(s))
        out.append(flat[o:o + n].reshape(s))
        o += n
    assert o == flat.numel(), (o, flat.numel())
    return out


# hidden activation of mlp_forward (Activation enum, src/torch/modules/ff/activation.rs:11; MlpConfig default Relu,
# mlp.rs:25-34); tests of non-default modules switch it with `with mlp_activation("tanh"):`
_MLP_ACTIVATION = "relu"
_ACTIVATIONS = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "identity": lambda t: t}


class mlp_activation:
    def __init__(self, name: str):
        assert name in _ACTIVATIONS, name
        self.name = name

    def __enter__(self):
        global _MLP_ACTIVATION
        self.prev, _MLP_ACTIVATION = _MLP_ACTIVATION, self.name
        return self

    def __exit__(self, *exc):
        global _MLP_ACTIVATION
        _MLP_ACTIVATION = self.prev
        return False


def mlp_forward(params, x):
    """Mlp::forward (mlp.rs:139-151): the activation between Linear layers, none on the output (output_activation Identity)."""
    h = x
    n_layers = len(params) // 2
    for l in range(n_layers):
        h = torch.nn.functional.linear(h, params[2 * l], params[2 * l + 1])
        if l + 1 < n_layers:
            h = _ACTIVATIONS[_MLP_ACTIVATION](h)
    return h


# ------------------------------------------------------------------------------------------------
# Categorical (src/torch/distributions/categorical.rs:29-76, distributions/mod.rs:25-31)
# ------------------------------------------------------------------------------------------------
class Categorical:
    def __init__(self, unnormalized_log_probs: torch.Tensor):
        self.log_probs = torch.log_softmax(unnormalized_log_probs, dim=-1)  # :29-33

    def _clamp_min(self, x):
        lo = F32_MIN if x.dtype == torch.float32 else float(torch.finfo(torch.float64).min)
        return x.clamp_min(lo)

    def log_prob(self, elements):  # :56-60
        return self.log_probs.gather(-1, elements.unsqueeze(-1)).squeeze(-1)

    def entropy(self):  # :62-68
        return -(self._clamp_min(self.log_probs) * self.log_probs.exp()).sum(-1)

    def kl_divergence_from(self, other: "Categorical"):  # :70-76  KL(self || other)
        return (self._clamp_min(self.log_probs - other.log_probs) * self.log_probs.exp()).sum(-1)


# ------------------------------------------------------------------------------------------------
# utils (src/torch/utils.rs:31-97)
# ------------------------------------------------------------------------------------------------
def flatten_tensors(ts):
    return torch.cat([t.reshape(-1) for t in ts])


def unflatten_tensors(flat, shapes):
    out, o = [], 0
    for s in shapes:
        n = int(np.prod(s))
        out.append(flat[o:o + n].reshape(s))
        o += n
    return out


def flat_dot(a, b):
    return torch.dot(a.reshape(-1), b.reshape(-1))


# ------------------------------------------------------------------------------------------------
# ConjugateGradientOptimizer (src/torch/optimizers/conjugate_gradient.rs:41-403)
# ------------------------------------------------------------------------------------------------
@dataclass
class CgConfig:  # :55-64
    iterations: int = 10
    max_backtracks: int = 15
    backtrack_ratio: float = 0.8
    hpv_reg_coeff: float = 1e-5
    accept_violation: bool = False


class HessianVectorProduct:  # :262-339
    def __init__(self, output, params, reg_coeff):
        self.params = params
        self.reg_coeff = reg_coeff
        self.shapes = [tuple(p.shape) for p in params]
        grads = torch.autograd.grad(output, params, retain_graph=True, create_graph=True, allow_unused=True)
        self.grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, params)]

    def mat_vec_mul(self, vector):
        vs = unflatten_tensors(vector, self.shapes)
        gvp = torch.stack([flat_dot(g, v) for g, v in zip(self.grads, vs)]).sum()
        hvp = torch.autograd.grad(gvp, self.params, retain_graph=True)
        return flatten_tensors(hvp) + self.reg_coeff * vector


class MatrixProduct:
    def __init__(self, m):
        self.m = m

    def mat_vec_mul(self, v):
        return self.m.mv(v)


def solve_conjugate_gradient(f_ax, b, iterations, residual_tol):  # :371-403
    x = torch.zeros_like(b)
    residual = b.clone()
    step = b.clone()
    rns = residual.dot(residual)
    iters = 0
    for _ in range(iterations):
        iters += 1
        z = f_ax.mat_vec_mul(step)
        alpha = rns / step.dot(z)
        x.addcmul_(alpha, step)
        residual.addcmul_(-alpha, z)
        new_rns = residual.dot(residual)
        if float(new_rns) < residual_tol:
            break
        mu = new_rns / rns
        step.mul_(mu)
        step.add_(residual)
        rns = new_rns
    return x, iters


class OptimizerStepError(Exception):
    def __init__(self, kind, **kw):
        super().__init__(kind)
        self.kind = kind
        self.info = kw


def trust_region_backward_step(params, loss_distance_fn, max_distance, cfg: CgConfig, log: dict):  # :115-179
    loss, distance = loss_distance_fn()
    loss_grads = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
    used = [(p, g) for p, g in zip(params, loss_grads) if g is not None]
    params = [p for p, _ in used]
    flat_loss_grads = flatten_tensors([g for _, g in used])
    hvp_fn = HessianVectorProduct(distance, params, cfg.hpv_reg_coeff)
    step_dir, cg_iters = solve_conjugate_gradient(hvp_fn, flat_loss_grads, cfg.iterations, 1e-10)
    step_dir = torch.nan_to_num(step_dir, nan=0.0)
    val = 1.0 / (float(step_dir.dot(hvp_fn.mat_vec_mul(step_dir))) + 1e-8) * max_distance * 2.0
    step_size = math.sqrt(val) if val >= 0 else float("nan")
    if math.isnan(step_size):
        step_size = 1.0
    log["step_size"] = step_size
    log["cg_iterations"] = cg_iters
    log["flat_grad"] = flat_loss_grads.detach().clone()
    log["step_dir"] = step_dir.detach().clone()
    descent_step = step_size * step_dir
    initial_loss = float(loss.detach())
    backtracking_line_search(params, descent_step, loss_distance_fn, max_distance, initial_loss, cfg, log)
    return initial_loss


def backtracking_line_search(params, descent_step, loss_constraint_fn, max_constraint_value, initial_loss,
                             cfg: CgConfig, log: dict):  # :183-254
    prev_params = [p.detach().clone() for p in params]
    shapes = [tuple(p.shape) for p in params]
    steps = unflatten_tensors(descent_step.detach(), shapes)
    loss = initial_loss
    constraint_val = float("inf")
    log["loss_initial"] = loss
    log["num_backtracks"] = -1
    for i in range(cfg.max_backtracks):
        ratio = cfg.backtrack_ratio ** i
        with torch.no_grad():
            for step, prev, p in zip(steps, prev_params, params):
                p.copy_(prev - ratio * step)
        with torch.no_grad():
            lt, ct = loss_constraint_fn()
        loss, constraint_val = float(lt), float(ct)
        if loss < initial_loss and constraint_val <= max_constraint_value:
            log["num_backtracks"] = i
            log["step_scale"] = ratio
            break
    log["loss_final"] = loss
    log["constraint_val_final"] = constraint_val
    err = None
    if math.isnan(loss):
        err = OptimizerStepError("NaNLoss")
    elif math.isnan(constraint_val):
        err = OptimizerStepError("NaNConstraint")
    elif loss >= initial_loss:
        err = OptimizerStepError("LossNotImproving", loss=loss, loss_before=initial_loss)
    elif constraint_val >= max_constraint_value and not cfg.accept_violation:
        err = OptimizerStepError("ConstraintViolated", constraint_val=constraint_val)
    if err is not None:
        with torch.no_grad():
            for p, prev in zip(params, prev_params):
                p.copy_(prev)
        raise err


# ------------------------------------------------------------------------------------------------
# Trpo::update (src/torch/agents/policies/trpo.rs:97-164)
# ------------------------------------------------------------------------------------------------
def trpo_update(flat_params: np.ndarray, n_in, hidden, n_out, obs, actions, advantages, max_kl=0.01,
                cfg: CgConfig | None = None, dtype=torch.float32):
    """Returns (new_flat_params, log dict with the reference's log keys + 'error')."""
    cfg = cfg or CgConfig()
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, n_out)]
    obs_t = torch.tensor(np.asarray(obs), dtype=dtype)
    act_t = torch.tensor(np.asarray(actions), dtype=torch.int64)
    adv_t = torch.tensor(np.asarray(advantages), dtype=dtype)
    log = {}
    with torch.no_grad():
        dist0 = Categorical(mlp_forward(params, obs_t))
        logp0 = dist0.log_prob(act_t)
        log["entropy"] = float(dist0.entropy().mean())

    def loss_distance_fn():
        dist = Categorical(mlp_forward(params, obs_t))
        ratio = (dist.log_prob(act_t) - logp0).exp()
        loss = -(ratio * adv_t).mean()
        distance = dist0.kl_divergence_from(dist).mean()
        return loss, distance

    log["error"] = None
    try:
        trust_region_backward_step(params, loss_distance_fn, max_kl, cfg, log)
    except OptimizerStepError as e:
        log["error"] = e.kind
    new_flat = flatten_tensors([p.detach() for p in params]).numpy().copy()
    return new_flat, log


def policy_loss_kl_grad_fvp(flat_params, n_in, hidden, n_out, obs, actions, advantages, vector, reg=0.0,
                            dtype=torch.float64):
    """Pieces of the TRPO step for kernel-level parity: loss, KL, flat loss gradient, (H + reg I) v."""
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, n_out)]
    obs_t = torch.tensor(np.asarray(obs), dtype=dtype)
    act_t = torch.tensor(np.asarray(actions), dtype=torch.int64)
    adv_t = torch.tensor(np.asarray(advantages), dtype=dtype)
    with torch.no_grad():
        dist0 = Categorical(mlp_forward(params, obs_t))
        logp0 = dist0.log_prob(act_t)
        entropy = float(dist0.entropy().mean())
    dist = Categorical(mlp_forward(params, obs_t))
    loss = -((dist.log_prob(act_t) - logp0).exp() * adv_t).mean()
    kl = dist0.kl_divergence_from(dist).mean()
    g = flatten_tensors(torch.autograd.grad(loss, params, retain_graph=True))
    hvp = HessianVectorProduct(kl, params, reg)
    hv = hvp.mat_vec_mul(torch.tensor(np.asarray(vector), dtype=dtype))
    return float(loss.detach()), float(kl.detach()), entropy, g.detach().numpy(), hv.detach().numpy()


# ------------------------------------------------------------------------------------------------
# ValuesOpt::update + n_backward_steps + Adam (critics/opt.rs:100-127, torch/agents/mod.rs:35-72,
# optimizers/coptimizer.rs:13-27,136-168; libtorch Adam defaults eps=1e-8, amsgrad=false)
# ------------------------------------------------------------------------------------------------
class Adam112:
    """libtorch 1.12 `torch::optim::Adam::step` (torch/csrc/api/src/optim/adam.cpp), the optimizer the
    reference's COptimizer::adam constructs: mul_/add_ moment updates (not the lerp_ of newer Python
    torch.optim.Adam), eps 1e-8, no amsgrad."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0, eps=1e-8):
        self.params = params
        self.lr, self.betas, self.wd, self.eps = lr, betas, weight_decay, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.step_count = 0

    @torch.no_grad()
    def step(self, grads):
        self.step_count += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.step_count
        bc2 = 1 - b2 ** self.step_count
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if self.wd != 0:
                g = g.add(p, alpha=self.wd)
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-(self.lr / bc1))


def value_update(flat_params, n_in, hidden, obs, targets, n_steps=80, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.0,
                 eps=1e-8, dtype=torch.float32, opt=None):
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, 1)]
    if opt is None:
        opt = Adam112(params, lr, betas, weight_decay, eps)
    else:
        opt.params = params
    obs_t = torch.tensor(np.asarray(obs), dtype=dtype)
    tgt_t = torch.tensor(np.asarray(targets), dtype=dtype)
    losses = []
    for _ in range(n_steps):
        loss = torch.nn.functional.mse_loss(mlp_forward(params, obs_t).squeeze(-1), tgt_t, reduction="mean")
        grads = torch.autograd.grad(loss, params)
        opt.step(grads)
        losses.append(float(loss.detach()))
    new_flat = flatten_tensors([p.detach() for p in params]).numpy().copy()
    return new_flat, losses, opt


def q_update(flat_params, n_in, hidden, n_out, obs, actions, targets, n_steps=1, lr=1e-3, dtype=torch.float32):
    """DQN loss (dqn.rs:316-326): mse(Q(obs).gather(actions), targets) + Adam steps."""
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, n_out)]
    opt = Adam112(params, lr)
    obs_t = torch.tensor(np.asarray(obs), dtype=dtype)
    act_t = torch.tensor(np.asarray(actions), dtype=torch.int64).unsqueeze(-1)
    tgt_t = torch.tensor(np.asarray(targets), dtype=dtype)
    losses = []
    for _ in range(n_steps):
        q = mlp_forward(params, obs_t).gather(-1, act_t).squeeze(-1)
        loss = torch.nn.functional.mse_loss(q, tgt_t, reduction="mean")
        opt.step(torch.autograd.grad(loss, params))
        losses.append(float(loss.detach()))
    return flatten_tensors([p.detach() for p in params]).numpy().copy(), losses


def dqn_update(flat_params, n_in, hidden, n_out, minibatches, discount, one_step_td=False, lr=1e-3,
               dtype=torch.float32):
    """DqnAgent::batch_update_slice_refs (dqn.rs:263-337) given the sampled episodes of every optimizer step.

    `minibatches` is a list (one per optimizer step) of lists of episodes; an episode is a dict with
    obs [L, F], action [L], reward [L], last_succ (1 = Terminate, 2 = Interrupt) and next_obs [F].
    Targets follow StepValueTarget::targets (critics/mod.rs:203-229): RewardToGo = discounted return without
    bootstrap; OneStepTd = r + gamma * max_a Q(next) under no_grad with the CURRENT parameters, 0 after a
    terminal step (masked_fill_, critics/mod.rs:127-129).  Returns (new flat params, per-step losses)."""
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, n_out)]
    opt = Adam112(params, lr)
    g = torch.tensor(float(np.float32(discount)), dtype=dtype)
    losses = []
    for episodes in minibatches:
        obs = torch.tensor(np.concatenate([ep["obs"] for ep in episodes]), dtype=dtype)
        act = torch.tensor(np.concatenate([ep["action"] for ep in episodes]), dtype=torch.int64).unsqueeze(-1)
        tgts = []
        with torch.no_grad():
            for ep in episodes:
                r = torch.tensor(np.asarray(ep["reward"]), dtype=dtype)
                n = r.shape[0]
                if one_step_td:
                    o = torch.tensor(np.asarray(ep["obs"]), dtype=dtype)
                    v = mlp_forward(params, o).amax(dim=-1)
                    nxt = torch.zeros(n, dtype=dtype)
                    nxt[:-1] = v[1:]
                    if ep["last_succ"] == 2:
                        no = torch.tensor(np.asarray(ep["next_obs"]), dtype=dtype).unsqueeze(0)
                        nxt[-1] = mlp_forward(params, no).amax(dim=-1)[0]
                    tgts.append(r + g * nxt)
                else:
                    y = torch.zeros(n, dtype=dtype)
                    carry = torch.zeros((), dtype=dtype)
                    for t in range(n - 1, -1, -1):
                        carry = r[t] + carry * g  # packed.rs:336
                        y[t] = carry
                    tgts.append(y)
        tgt = torch.cat(tgts)
        q = mlp_forward(params, obs).gather(-1, act).squeeze(-1)
        loss = torch.nn.functional.mse_loss(q, tgt, reduction="mean")
        opt.step(torch.autograd.grad(loss, params))
        losses.append(float(loss.detach()))
    return flatten_tensors([p.detach() for p in params]).numpy().copy(), losses


# ------------------------------------------------------------------------------------------------
# Chain<Gru, Linear> (src/torch/modules/chain.rs:145-186, seq/rnn/gru.rs:30-39,72-102, ff/linear.rs:118-123)
# ------------------------------------------------------------------------------------------------
def unflatten_gru_linear(flat, n_in, hidden, n_out):
    """Module::variables() order: w_ih[3H,in], w_hh[3H,H], b_ih[3H], b_hh[3H], kernel[out,H], bias[out]."""
    flat = torch.as_tensor(np.asarray(flat))
    sizes = [(3 * hidden, n_in), (3 * hidden, hidden), (3 * hidden,), (3 * hidden,), (n_out, hidden), (n_out,)]
    out, off = [], 0
    for sh in sizes:
        n = int(np.prod(sh))
        out.append(flat[off:off + n].reshape(sh))
        off += n
    assert off == flat.numel()
    return out


def gru_linear_episode(flat, n_in, hidden, n_out, obs, activation="relu", dtype=torch.float32):
    """SeqIterative::step over one episode (rnn/mod.rs:166-184 -> gru.rs:30-39 gru_cell, chain.rs:176-186):
    h0 = 0, per step h = gru_cell(x, h), logits = Linear(act(h)).  obs [L, F] -> logits [L, out]."""
    w_ih, w_hh, b_ih, b_hh, lw, lb = [t.to(dtype) for t in unflatten_gru_linear(flat, n_in, hidden, n_out)]
    x = torch.tensor(np.asarray(obs), dtype=dtype)
    h = torch.zeros(1, hidden, dtype=dtype)
    outs = []
    act = {"relu": torch.relu, "identity": lambda t: t, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[activation]
    for t in range(x.shape[0]):
        h = torch.gru_cell(x[t:t + 1], h, w_ih, w_hh, b_ih, b_hh)
        outs.append(torch.nn.functional.linear(act(h), lw, lb)[0])
    return torch.stack(outs).numpy() if outs else np.zeros((0, n_out), np.float32)


def gru_packed_episodes(flat, n_in, hidden, n_out, episodes, activation="relu", dtype=torch.float32):
    """SeqPacked::seq_packed (gru.rs:72-102: Tensor::gru_data on the packed sequence) for a list of episodes
    [L_i, F]; returns the per-episode logits (unpacked again) so that packed == iterated can be checked
    (modules/testing.rs:124-157)."""
    w_ih, w_hh, b_ih, b_hh, lw, lb = [t.to(dtype) for t in unflatten_gru_linear(flat, n_in, hidden, n_out)]
    order = sorted(range(len(episodes)), key=lambda i: -len(episodes[i]))
    seqs = [torch.tensor(np.asarray(episodes[i]), dtype=dtype) for i in order]
    packed = torch.nn.utils.rnn.pack_sequence(seqs, enforce_sorted=True)
    h0 = torch.zeros(1, len(seqs), hidden, dtype=dtype)
    out, _ = torch._VF.gru(packed.data, packed.batch_sizes, h0, [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, True, False)
    act = {"relu": torch.relu, "identity": lambda t: t, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[activation]
    logits = torch.nn.functional.linear(act(out), lw, lb)
    unpacked, lens = torch.nn.utils.rnn.pad_packed_sequence(
        torch.nn.utils.rnn.PackedSequence(logits, packed.batch_sizes), batch_first=True)
    res = [None] * len(episodes)
    for k, i in enumerate(order):
        res[i] = unpacked[k, :int(lens[k])].numpy()
    return res


# ------------------------------------------------------------------------------------------------
# PPO / REINFORCE (src/torch/agents/policies/ppo.rs:97-147, reinforce.rs:64-89)
# ------------------------------------------------------------------------------------------------
def ppo_update(flat_params, n_in, hidden, n_out, obs, actions, advantages, opt_steps=10, clip_distance=0.2, lr=1e-3,
               dtype=torch.float32):
    """no_grad initial log-probs + entropy, then opt_steps x Adam on
    -mean(min(ratio * adv, clip(ratio, 1 - eps, 1 + eps) * adv)).  Returns (new flat params, losses, entropy)."""
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, n_out)]
    opt = Adam112(params, lr)
    obs_t = torch.tensor(np.asarray(obs), dtype=dtype)
    act_t = torch.tensor(np.asarray(actions), dtype=torch.int64)
    adv_t = torch.tensor(np.asarray(advantages), dtype=dtype)
    with torch.no_grad():
        d0 = Categorical(mlp_forward(params, obs_t))
        lp0 = d0.log_prob(act_t)
        entropy = float(d0.entropy().mean())
    losses = []
    for _ in range(opt_steps):
        lp = Categorical(mlp_forward(params, obs_t)).log_prob(act_t)
        ratio = (lp - lp0).exp()
        clipped = ratio.clip(1.0 - clip_distance, 1.0 + clip_distance)
        loss = -torch.minimum(ratio * adv_t, clipped * adv_t).mean()
        opt.step(torch.autograd.grad(loss, params))
        losses.append(float(loss.detach()))
    return flatten_tensors([p.detach() for p in params]).numpy().copy(), losses, entropy


def reinforce_update(flat_params, n_in, hidden, n_out, obs, actions, advantages, lr=1e-3, dtype=torch.float32):
    """One backward_step on -(log_probs * advantages).mean(); logs the mean entropy of that evaluation."""
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    params = [p.clone().requires_grad_(True) for p in unflatten_mlp(flat, n_in, hidden, n_out)]
    opt = Adam112(params, lr)
    obs_t = torch.tensor(np.asarray(obs), dtype=dtype)
    act_t = torch.tensor(np.asarray(actions), dtype=torch.int64)
    adv_t = torch.tensor(np.asarray(advantages), dtype=dtype)
    dist = Categorical(mlp_forward(params, obs_t))
    loss = -(dist.log_prob(act_t) * adv_t).mean()
    entropy = float(dist.entropy().mean().detach())
    opt.step(torch.autograd.grad(loss, params))
    return flatten_tensors([p.detach() for p in params]).numpy().copy(), float(loss.detach()), entropy


# ------------------------------------------------------------------------------------------------
# The same updates with a recurrent module (Chain<Gru, Linear>) in place of the MLP: Trpo::update
# (trpo.rs:97-164; cuDNN disabled :104-108 so autograd differentiates the composed gru_cell twice),
# ValuesOpt::update (opt.rs:100-127), eval_extended_state_values (critics/mod.rs:116-131).
# The reference feeds PackedTensor batches through SeqPacked::seq_packed; the loss is a mean over all
# steps, so episodes are padded to a common length here and the valid steps selected afterwards.
# ------------------------------------------------------------------------------------------------
class EpisodeBatch:
    """episodes: list of [L_i, F] observation arrays.  `select(x[B, Lmax, ...])` returns the valid steps
    episode-major (episode 0 steps 0..L0-1, episode 1 ...)."""

    def __init__(self, episodes, dtype):
        self.lens = [len(e) for e in episodes]
        B, Lmax = len(episodes), max(self.lens) if episodes else 0
        F = np.asarray(episodes[0]).shape[1]
        obs = np.zeros((B, Lmax, F), np.float64)
        for i, e in enumerate(episodes):
            obs[i, :len(e)] = e
        self.obs = torch.tensor(obs, dtype=dtype)
        mask = np.zeros((B, Lmax), bool)
        for i, n in enumerate(self.lens):
            mask[i, :n] = True
        self.mask = torch.tensor(mask)

    def select(self, x):
        return x[self.mask]


def gru_linear_forward(params, batch: EpisodeBatch, activation="relu"):
    """h0 = 0 per episode (gru.rs:23-28); h = gru_cell(x, h); out = Linear(act(h)) (chain.rs:157-168)."""
    w_ih, w_hh, b_ih, b_hh, lw, lb = params
    act = {"relu": torch.relu, "identity": lambda t: t, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[activation]
    B, Lmax, _ = batch.obs.shape
    h = torch.zeros(B, w_hh.shape[1], dtype=batch.obs.dtype)
    outs = []
    for t in range(Lmax):
        # composed gru_cell (what libtorch's CPU gru_cell computes, gate order [r, z, n])
        gi = torch.nn.functional.linear(batch.obs[:, t], w_ih, b_ih)
        gh = torch.nn.functional.linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(h_r + i_r)
        z = torch.sigmoid(h_z + i_z)
        n = torch.tanh(i_n + r * h_n)
        h = (h - n) * z + n
        outs.append(torch.nn.functional.linear(act(h), lw, lb))
    return batch.select(torch.stack(outs, dim=1))


def _seq_params(flat_params, n_in, hidden, n_out, dtype):
    flat = torch.tensor(np.asarray(flat_params), dtype=dtype)
    return [p.clone().requires_grad_(True) for p in unflatten_gru_linear(flat, n_in, hidden, n_out)]


def _cat(xs, dtype):
    return torch.tensor(np.concatenate([np.asarray(x) for x in xs]), dtype=dtype)


def seq_policy_loss_kl_grad_fvp(flat_params, n_in, hidden, n_out, episodes, actions, advantages, vector, reg=0.0,
                                dtype=torch.float64, activation="relu"):
    """episodes / actions / advantages: per-episode arrays.  Returns loss, kl, entropy, grad, (H + reg I) v."""
    params = _seq_params(flat_params, n_in, hidden, n_out, dtype)
    batch = EpisodeBatch(episodes, dtype)
    act_t, adv_t = _cat(actions, torch.int64), _cat(advantages, dtype)
    with torch.no_grad():
        dist0 = Categorical(gru_linear_forward(params, batch, activation))
        logp0 = dist0.log_prob(act_t)
        entropy = float(dist0.entropy().mean())
    dist = Categorical(gru_linear_forward(params, batch, activation))
    loss = -((dist.log_prob(act_t) - logp0).exp() * adv_t).mean()
    kl = dist0.kl_divergence_from(dist).mean()
    g = flatten_tensors(torch.autograd.grad(loss, params, retain_graph=True))
    hv = HessianVectorProduct(kl, params, reg).mat_vec_mul(torch.tensor(np.asarray(vector), dtype=dtype))
    return float(loss.detach()), float(kl.detach()), entropy, g.detach().numpy(), hv.detach().numpy()


def seq_trpo_update(flat_params, n_in, hidden, n_out, episodes, actions, advantages, max_kl=0.01,
                    cfg: CgConfig | None = None, dtype=torch.float32, activation="relu"):
    cfg = cfg or CgConfig()
    params = _seq_params(flat_params, n_in, hidden, n_out, dtype)
    batch = EpisodeBatch(episodes, dtype)
    act_t, adv_t = _cat(actions, torch.int64), _cat(advantages, dtype)
    log = {}
    with torch.no_grad():
        dist0 = Categorical(gru_linear_forward(params, batch, activation))
        logp0 = dist0.log_prob(act_t)
        log["entropy"] = float(dist0.entropy().mean())

    def loss_distance_fn():
        dist = Categorical(gru_linear_forward(params, batch, activation))
        loss = -((dist.log_prob(act_t) - logp0).exp() * adv_t).mean()
        return loss, dist0.kl_divergence_from(dist).mean()

    log["error"] = None
    try:
        trust_region_backward_step(params, loss_distance_fn, max_kl, cfg, log)
    except OptimizerStepError as e:
        log["error"] = e.kind
    return flatten_tensors([p.detach() for p in params]).numpy().copy(), log


def seq_value_update(flat_params, n_in, hidden, episodes, targets, n_steps=80, lr=1e-3, dtype=torch.float32,
                     activation="relu"):
    params = _seq_params(flat_params, n_in, hidden, 1, dtype)
    opt = Adam112(params, lr)
    batch = EpisodeBatch(episodes, dtype)
    tgt_t = _cat(targets, dtype)
    losses = []
    for _ in range(n_steps):
        v = gru_linear_forward(params, batch, activation).squeeze(-1)
        loss = torch.nn.functional.mse_loss(v, tgt_t, reduction="mean")
        opt.step(torch.autograd.grad(loss, params))
        losses.append(float(loss.detach()))
    return flatten_tensors([p.detach() for p in params]).numpy().copy(), losses
