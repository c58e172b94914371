"""TEST-ONLY engine: plays the role of libmcpc_b200.so with the numpy oracle so the host logic of
PCTrainer (segments, zero_grad windows, normalisation, results, data-parallel reduction) can be
tested on CPU, including world_size-2 gloo runs.  Never shipped, never imported by the package."""
import numpy as np
import torch

from golden_util import orc
from montecarlopredictivecoding_b200 import _native as N


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


class OracleEngine:
    name = "oracle-test-double"

    def __init__(self, dtype=np.float32):
        self.dtype = dtype
        self.calls = []

    def _net(self, c_plan, top, W, b):
        return orc.OracleNet(W=[_np(w).astype(self.dtype) for w in W],
                             b=[None if v is None else _np(v).astype(self.dtype) for v in b],
                             n_layers=c_plan.L, act=list(c_plan.act), energy_scale=list(c_plan.energy_scale),
                             top=top.kind, top_var=1.0 / top.inv_var, mask_start_col=top.mask_start, dtype=self.dtype)

    def fill_noise(self, seed, t_begin, n_steps, chain_offset, B, n_units, noise_scale, device):
        out = np.stack([orc.langevin_normals(seed, t_begin + s, chain_offset, B, n_units) for s in range(n_steps)])
        return torch.from_numpy((out * noise_scale).astype(np.float32))

    def infer(self, c):
        self.calls.append(("infer", c.t_begin, c.n_steps, c.save_begin, c.save_end))
        plan = c.plan
        net = self._net(plan, c.top, c.W, c.b)
        B = c.B
        inputs = np.zeros((B, plan.d_in), self.dtype) if c.inputs is None else _np(c.inputs)
        target = _np(c.target)
        offs = np.cumsum([0] + plan.dims)
        noise = None
        if c.noise_mode == N.NOISE_SUPPLIED:
            nz = _np(c.noise)
            noise = [[nz[t][:, offs[l]:offs[l + 1]] for l in range(plan.L)] for t in range(c.n_steps)]
        elif c.noise_mode == N.NOISE_PHILOX:
            nz = self.fill_noise(c.seed, c.t_begin, c.n_steps, c.chain_offset, B, plan.SD, c.noise_scale, None).numpy()
            noise = [[nz[t][:, offs[l]:offs[l + 1]] for l in range(plan.L)] for t in range(c.n_steps)]
        adam_state = None
        if c.optimizer == N.OPT_ADAM and c.adam_m is not None:
            adam_state = orc.AdamState([_np(m).copy() for m in c.adam_m], [_np(v).copy() for v in c.adam_v], c.adam_step0)
        xs = [_np(x).copy() for x in c.x]
        for s in range(c.n_steps):
            ro = orc.readouts(net, xs, inputs, target)
            if c.energy is not None:
                c.energy[s] = ro.energy
            if c.loss is not None:
                c.loss[s] = ro.loss or 0.0
            if c.traj_every > 0 and s % c.traj_every == 0:
                r = s // c.traj_every
                for l in range(plan.L):
                    if c.traj_x and c.traj_x[l] is not None:
                        c.traj_x[l][r] = torch.from_numpy(xs[l].copy())
                if c.traj_out is not None:
                    c.traj_out[r] = torch.from_numpy(ro.out.copy())
            if c.save_g is not None and c.save_begin <= s < c.save_end:
                k = c.energy_coefficient
                G = [(-self.dtype(k * plan.energy_scale[l])) * ro.eps[l] for l in range(plan.L)]
                if plan.d_out > 0:
                    G.append(ro.e_out)
                c.save_g[s - c.save_begin] = torch.from_numpy(np.concatenate(G, axis=1).astype(np.float32))
                c.save_f[s - c.save_begin] = torch.from_numpy(np.concatenate(ro.fx, axis=1).astype(np.float32))
            if c.x_grad is not None and s == c.n_steps - 1:
                gs = orc.latent_grads(net, xs, ro, c.energy_coefficient)
                for l in range(plan.L):
                    c.x_grad[l].copy_(torch.from_numpy(gs[l]))
            r1 = orc.infer(net, xs, inputs, target, 1, optimizer="adam" if c.optimizer == N.OPT_ADAM else "sgd",
                           lr=c.lr, betas=c.betas, adam_eps=c.adam_eps, adam_state=adam_state,
                           noise=None if noise is None else [noise[s]], update_x=c.update_x,
                           energy_coefficient=c.energy_coefficient)
            xs = r1.xs
        for l in range(plan.L):
            c.x[l].copy_(torch.from_numpy(xs[l]))
            if adam_state is not None:
                c.adam_m[l].copy_(torch.from_numpy(adam_state.m[l]))
                c.adam_v[l].copy_(torch.from_numpy(adam_state.v[l]))

    def traj_stats(self, traj, n_rec, count_before, mean, m2):
        """Same contract as NativeEngine.traj_stats (mcpc_traj_stats_update): merge a block of records (Chan et al.)."""
        blk = traj[:n_rec].to(torch.float64).reshape(n_rec, -1)
        nb, na = float(n_rec), float(count_before)
        mb = blk.mean(0)
        m2b = ((blk - mb) ** 2).sum(0)
        mean64, m264 = mean.to(torch.float64).reshape(-1), m2.to(torch.float64).reshape(-1)
        delta = mb - mean64
        tot = na + nb
        mean.copy_((mean64 + delta * nb / tot).to(torch.float32).view_as(mean))
        m2.copy_((m264 + m2b + delta * delta * na * nb / tot).to(torch.float32).view_as(m2))

    def weight_grad(self, plan, top, energy_coefficient, B, n_save, save_g, save_f, inputs, gW, gb, precision):
        self.calls.append(("weight_grad", n_save))
        G = save_g.numpy().reshape(n_save * B, -1).astype(np.float64)
        F = save_f.numpy().reshape(n_save * B, -1).astype(np.float64)
        offs = np.cumsum([0] + plan.dims)
        n_lin = plan.L + (1 if plan.d_out > 0 else 0)
        has_grad = top.kind in (N.TOP_GAUSS, N.TOP_BERNOULLI)
        for l in range(n_lin):
            if l == plan.L and not has_grad:
                continue
            Gl = G[:, offs[l]:offs[l + 1]] if l < plan.L else G[:, offs[plan.L]:]
            if l == 0:
                below = None if inputs is None else np.tile(inputs.numpy().astype(np.float64), (n_save, 1))
            else:
                below = F[:, offs[l - 1]:offs[l]]
            if gW[l] is not None and below is not None:
                gW[l] += torch.from_numpy((Gl.T @ below).astype(np.float32))
            if gb[l] is not None:
                gb[l] += torch.from_numpy(Gl.sum(0).astype(np.float32))
