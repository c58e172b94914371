"""Pin the oracle (oracle/mcpc_oracle.py) against golden vectors recorded from the real
reference implementation (tests/golden/make_golden.py).  CPU only.

Tolerances: the Langevin/SGD path is contractive: 1e-5 relative on latents per
step over the whole horizon.  Adam-on-x trajectories are chaotic w.r.t. fp32 op order, so the
horizons recorded in the fixtures are short (<= 60 steps) and use the same 1e-5 bound, except
where noted.
"""
import numpy as np
import pytest
import torch

from golden_util import ALL_CASES, GoldenCase, orc, rel_err

TOL_X = 1e-5
TOL_SCALAR = 1e-5


def _p_stepper(net, call):
    tr = call["trainer"]
    params = []
    for W, b in zip(net.W, net.b):
        params.append(torch.nn.Parameter(torch.from_numpy(W)))      # shares memory with net.W
        if b is not None:
            params.append(torch.nn.Parameter(torch.from_numpy(b)))
    opt_cls = {"sgd": torch.optim.SGD, "adam": torch.optim.Adam}[tr.get("opt_p", "sgd")]
    opt = opt_cls(params, **tr.get("opt_p_kwargs", {"lr": 0.0}))

    def p_step(gW, gb):
        it = iter(params)
        for i, (W, b) in enumerate(zip(net.W, net.b)):
            next(it).grad = torch.from_numpy(np.ascontiguousarray(gW[i]))
            if b is not None:
                next(it).grad = torch.from_numpy(np.ascontiguousarray(gb[i]))
        opt.step()
    return p_step


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_matches_reference(name):
    gc = GoldenCase(name)
    grads_prev = None
    adam_prev = {}
    for ci, call in enumerate(gc.calls):
        tr = call["trainer"]
        net = gc.oracle_net(ci)
        upd_x, upd_p, acc = gc.step_lists(ci)
        rows = gc.rows(ci)
        carry = None if call.get("reset_opt_x", True) or call.get("sample_x", True) else adam_prev.get(call.get("trainer_of", ci))
        res = orc.train_on_batch(
            net, gc.x0(ci), gc.inputs[:rows], gc.target[:rows], gc.steps_run(ci), update_p_at=upd_p, accumulate_p_at=acc,
            p_step=_p_stepper(net, call), grads_in=grads_prev, optimizer=tr["opt_x"], lr=tr["lr_x"],
            noise=gc.noise(ci), update_x_at=upd_x, energy_coefficient=tr.get("energy_coefficient", 1.0),
            adam_state_in=carry)
        adam_prev[call.get("trainer_of", ci)] = res.adam
        for l in range(gc.L):
            got = np.stack([res.traj_xs[t][l] for t in range(gc.steps_run(ci))])
            assert rel_err(got, gc.traj(ci, l)) < TOL_X, (name, ci, l)
            assert rel_err(res.xs[l], gc.x_final(ci)[l]) < TOL_X, (name, ci, l, "final")
        assert rel_err(np.stack(res.traj_out), gc.z[f"c{ci}_outputs"]) < TOL_X
        assert rel_err(res.energy, gc.z[f"c{ci}_energy"]) < TOL_SCALAR
        assert rel_err(res.overall, gc.z[f"c{ci}_overall"]) < TOL_SCALAR
        ref_loss = gc.z[f"c{ci}_loss"]
        assert len(res.loss) == len(ref_loss)
        if len(ref_loss) and np.max(np.abs(ref_loss)) > 0:
            assert rel_err(res.loss, ref_loss) < TOL_SCALAR
        gW_ref, gb_ref = gc.grads(ci)
        for i in range(gc.n_lin):
            if gW_ref[i] is not None:
                scale = max(float(np.max(np.abs(gW_ref[i]))), 1e-6)
                assert np.max(np.abs(res.gW[i] - gW_ref[i])) < 2e-5 * max(scale, 1.0), (name, ci, i)
            if gb_ref[i] is not None:
                scale = max(float(np.max(np.abs(gb_ref[i]))), 1.0)
                assert np.max(np.abs(res.gb[i] - gb_ref[i])) < 2e-5 * scale, (name, ci, i)
        W_after, b_after = gc.weights(ci, "after")
        for i in range(gc.n_lin):
            assert np.max(np.abs(net.W[i] - W_after[i])) < 2e-6, (name, ci, i, "W after p-step")
            if b_after[i] is not None:
                assert np.max(np.abs(net.b[i] - b_after[i])) < 2e-6
        grads_prev = (res.gW, res.gb)


def test_posterior_linear_gaussian_analytic():
    """figure_2.py:40-48,79 -- Langevin samples of the 1-D linear-Gaussian model must follow
    the analytic posterior N(0.44, 0.2) up to Monte-Carlo error and the O(lr) bias."""
    rng = np.random.default_rng(0)
    net = orc.OracleNet(W=[np.zeros((1, 1)), np.full((1, 1), 2.0)], b=[np.full((1,), 0.2), None],
                        n_layers=1, act=[orc.ACT_IDENTITY], energy_scale=[1.0], top=orc.TOP_GAUSS,
                        top_var=1.0, dtype=np.float64)
    B, T, lr = 64, 3000, 0.02
    noise = [[rng.standard_normal((B, 1)) * np.sqrt(2.0 / lr)] for _ in range(T)]
    res = orc.infer(net, [np.full((B, 1), 3.0)], np.zeros((B, 1)), np.ones((B, 1)), T, optimizer="sgd", lr=lr,
                    noise=noise, record_traj=True)
    s = np.stack([res.traj_xs[t][0] for t in range(500, T)])
    assert abs(s.mean() - 0.44) < 0.02
    assert abs(s.var() - 0.2) < 0.02


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    r = orc.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = orc.philox4x32_10([0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], [0xFFFFFFFF], 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v[0]) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = orc.philox4x32_10([0x243F6A88], [0x85A308D3], [0x13198A2E], [0x03707344], 0xA4093822, 0x299F31D0)
    assert [int(v[0]) for v in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_langevin_normals_moments():
    xi = orc.langevin_normals(seed=1234, t=7, chain0=0, n_chains=4096, n_units=64)
    assert xi.shape == (4096, 64)
    assert abs(xi.mean()) < 0.01
    assert abs(xi.var() - 1.0) < 0.02
    # chain-offset invariance: rows are keyed by the GLOBAL chain id
    xi2 = orc.langevin_normals(seed=1234, t=7, chain0=1024, n_chains=16, n_units=64)
    assert np.array_equal(xi[1024:1040], xi2)
