"""Environment bookkeeping around the solver.

CPU: the oracle restatement (oracle/env_oracle.py) against the fixtures recorded from the UNMODIFIED reference
env + solver classes (tests/golden/env_{csmri,spi}.npz, oracle/make_golden.py), the host-side containers, and
the loud failure on CPU tensors.  GPU (-m gpu): tfpnp_b200.env through the C ABI against the same fixtures, the
oracle env on PR / CT episodes, and the stand-alone gather / scatter / channel-pack kernels bit-exactly.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, weights
from oracle import env_oracle as E
from oracle import pnp_oracle as O
from oracle import synth


def _fixture_episode(task):
    g = load_golden(f"env_{task}")
    data = {k[5:]: v for k, v in g.items() if k.startswith("data_")}
    n_steps = int(g["n_steps"])
    actions = []
    for s in range(n_steps):
        pre = f"step{s}_action_"
        actions.append({k[len(pre):]: v for k, v in g.items() if k.startswith(pre)})
    return g, data, actions, int(g["max_episode_step"])


@pytest.mark.parametrize("task", ["csmri", "spi"])
def test_oracle_env_matches_reference_fixture(task):
    g, data, actions, steps = _fixture_episode(task)
    env = E.EnvOracle(task, weights("he"), steps)
    ob = env.reset({k: v.clone() for k, v in data.items()})
    assert rel_err(env.policy_ob(ob), g["reset_policy_ob"])[1] <= 2e-6
    for s, a in enumerate(actions):
        ob, ob_m, reward, all_done, info = env.step({k: v.clone() for k, v in a.items()})
        assert rel_err(env.policy_ob(ob), g[f"step{s}_ob_policy_ob"])[1] <= 2e-6
        pm = env.policy_ob(ob_m)
        assert tuple(pm.shape) == tuple(g[f"step{s}_masked_policy_ob"].shape)
        if pm.shape[0]:
            assert rel_err(pm, g[f"step{s}_masked_policy_ob"])[1] <= 2e-6
        assert torch.allclose(reward, g[f"step{s}_reward"], rtol=1e-4, atol=1e-4)
        assert bool(all_done) == bool(g[f"step{s}_all_done"])
        assert torch.equal(info["done"], g[f"step{s}_done"])


def test_batch_container():
    from tfpnp_b200.env import Batch
    b = Batch(gt=torch.arange(6.).view(3, 2), T=torch.ones(3, 1))
    assert b.gt is b["gt"] and b.shape == [3]
    sub = b[torch.tensor([0, 2])]
    assert torch.equal(sub.gt, torch.tensor([[0., 1.], [4., 5.]]))
    with pytest.raises(AttributeError):
        b.missing


def test_env_refuses_cpu_state():
    import tfpnp_b200 as T
    from tfpnp_b200.env import CSMRIEnv

    class _Stub:                         # no native handle is touched before the device check
        def reset(self, data):
            raise AssertionError("unreachable")
    env = CSMRIEnv(None, _Stub(), 3)
    with pytest.raises(TypeError):
        env.to("cuda")
    _, data, _, _ = _fixture_episode("csmri")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.reset(data=data)
    assert T.CSMRIEnv is CSMRIEnv and CSMRIEnv.ob_base_dim == 6 and T.PREnv.ob_base_dim == 14
    assert T.CTEnv.ob_base_dim == 4 and T.SPIEnv.ob_base_dim == 3


# ------------------------------------------------------------------ GPU ---------------------------------------
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _cu(d, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _native_env(task, dev, steps, prec="fp16x3", opnorm=None):
    import tfpnp_b200 as T
    den = T.UNetDenoiser2D(state_dict=weights("he"), precision=prec)
    solver = {"csmri": T.ADMMSolver_CSMRI, "pr": T.IADMMSolver_PR, "ct": T.IADMMSolver_CT, "spi": T.ADMMSolver_SPI}[task](den)
    if opnorm is not None:
        solver.opnorm_override = opnorm
    env = {"csmri": T.CSMRIEnv, "pr": T.PREnv, "ct": T.CTEnv, "spi": T.SPIEnv}[task](None, solver, steps)
    return env.to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["csmri", "spi"])
def test_native_env_matches_reference_fixture(task, dev):
    g, data, actions, steps = _fixture_episode(task)
    env = _native_env(task, dev, steps)
    ob = env.reset(data={k: v.clone() for k, v in data.items()})
    assert torch.equal(env.get_policy_ob(ob).cpu(), g["reset_policy_ob"])          # pure data movement: bit-exact
    # SPI: the 10-step bisection prox has a 1e-3 resolution, a last-ulp difference flips isolated pixels by one cell
    # (tests/test_gpu_parity.py::test_spi_golden) -> judged by the relative L2 error; CS-MRI by the max error
    k, t = (0, 5e-4) if task == "spi" else (1, 1e-4)
    for s, a in enumerate(actions):
        ob, ob_m, reward, all_done, info = env.step(_cu(a, dev))
        assert rel_err(ob.variables, g[f"step{s}_ob_variables"])[k] <= t
        assert rel_err(env.get_policy_ob(ob), g[f"step{s}_ob_policy_ob"])[k] <= t
        pm = env.get_policy_ob(ob_m)
        assert tuple(pm.shape) == tuple(g[f"step{s}_masked_policy_ob"].shape)
        if pm.shape[0]:
            assert rel_err(pm, g[f"step{s}_masked_policy_ob"])[k] <= t
        assert torch.allclose(reward.cpu(), g[f"step{s}_reward"], rtol=1e-3, atol=2e-3)   # dB
        assert bool(all_done) == bool(g[f"step{s}_all_done"])
        assert torch.equal(info["done"].cpu(), g[f"step{s}_done"])


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["pr", "ct"])
def test_native_env_matches_oracle_env(task, dev):
    B, n, steps, pack = 3, 32, 3, 2
    if task == "pr":
        d = synth.pr_batch(B, n, steps * pack)
        opnorm = None
    else:
        d = synth.ct_batch(B, n, 12, steps * pack)
        opnorm = d["opnorm"]
    data = E.env_data(task, d)
    actions = E.episode_actions(task, d, B, steps, pack)
    ora = E.EnvOracle(task, weights("he"), steps, opnorm=opnorm or 0.0)
    env = _native_env(task, dev, steps, opnorm=opnorm)
    ob_o = ora.reset({k: v.clone() for k, v in data.items()})
    ob = env.reset(data={k: v.clone() for k, v in data.items()})
    assert torch.equal(env.get_policy_ob(ob).cpu(), ora.policy_ob(ob_o))
    assert env.get_policy_ob(ob).shape[1] == type(env).ob_base_dim + 3
    for a in actions:
        o = ora.step({k: v.clone() for k, v in a.items()})
        r = env.step(_cu(a, dev))
        assert rel_err(r[0].variables, o[0]["variables"])[1] <= 1e-4
        assert rel_err(env.get_policy_ob(r[0]), ora.policy_ob(o[0]))[1] <= 1e-4
        assert tuple(r[1].variables.shape) == tuple(o[1]["variables"].shape)
        assert torch.allclose(r[2].cpu(), o[2], rtol=1e-3, atol=2e-3)
        assert bool(r[3]) == bool(o[3]) and torch.equal(r[4]["done"].cpu(), o[4]["done"])
        if o[3]:
            break


@pytest.mark.gpu
def test_env_kernels_bit_exact(dev):
    """gather / scatter / channel pack are pure data movement: bit-identical to torch indexing."""
    from tfpnp_b200.env import gather_rows, CSMRIEnv
    g = torch.Generator().manual_seed(0)
    a = torch.randn(7, 3, 16, 16, 2, generator=g).to(dev)
    m = (torch.rand(7, 1, 16, 16, generator=g) > 0.5).to(dev)
    odd = torch.randn(7, 5, generator=g).to(dev)                     # 20-byte rows: the unvectorised path
    idx = torch.tensor([6, 0, 3], device=dev)
    ga, gm, go = gather_rows([a, m, odd], idx)
    assert torch.equal(ga, a[idx]) and torch.equal(gm, m[idx]) and torch.equal(go, odd[idx])
    ia, = gather_rows([a], None)
    assert torch.equal(ia, a) and ia.data_ptr() != a.data_ptr()
    e0, = gather_rows([a], idx[:0])
    assert e0.shape[0] == 0
    # scatter: state['solver'][idx] = s; state['output'][idx] = Re(x)
    env = CSMRIEnv(None, None, 3).to(dev)
    env.state = {"solver": a.clone(), "output": torch.zeros(7, 1, 16, 16, device=dev)}
    s = torch.randn(3, 3, 16, 16, 2, generator=g).to(dev)
    env._scatter_state(s, idx)
    ref = a.clone(); ref[idx] = s
    assert torch.equal(env.state["solver"], ref)
    assert torch.equal(env.state["output"][idx], s[:, :1, ..., 0])
    assert float(env.state["output"][1].abs().max()) == 0.0
    # solvers with fewer variables (HQS / APG: 2, PG: 1; tfpnp/pnp/solver/base.py:118-214): the row stride follows num_var
    for nv in (2, 1):
        st = torch.randn(7, nv, 16, 16, 2, generator=g).to(dev)
        env.state = {"solver": st.clone(), "output": torch.zeros(7, 1, 16, 16, device=dev)}
        s = torch.randn(3, nv, 16, 16, 2, generator=g).to(dev)
        env._scatter_state(s, idx)
        ref = st.clone(); ref[idx] = s
        assert torch.equal(env.state["solver"], ref)
        assert torch.equal(env.state["output"][idx], s[:, :1, ..., 0])
    from tfpnp_b200.env import CTEnv
    renv = CTEnv(None, None, 3).to(dev)
    st = torch.randn(5, 1, 16, 16, generator=g).to(dev)
    renv.state = {"solver": st.clone(), "output": torch.zeros(5, 1, 16, 16, device=dev)}
    s = torch.randn(2, 1, 16, 16, generator=g).to(dev)
    renv._scatter_state(s, idx[:2] % 5)
    ref = st.clone(); ref[idx[:2] % 5] = s
    assert torch.equal(renv.state["solver"], ref) and torch.equal(renv.state["output"][idx[:2] % 5], s)
    with pytest.raises(ValueError):                      # a solver whose rows do not match the environment's state
        renv._scatter_state(torch.zeros(2, 3, 16, 16, device=dev), idx[:2] % 5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hqs", "pg"])
def test_env_step_with_fewer_solver_variables(dev, name):
    """PnPEnv is generic in solver.num_var (the reference's index_put is): an episode step with the 2-variable HQS and the
    1-variable PG CS-MRI solvers equals calling the solver by hand and scattering with torch indexing."""
    import tfpnp_b200 as T
    from tfpnp_b200.env import CSMRIEnv
    d = synth.csmri_batch(4, 32, 2)
    den = T.UNetDenoiser2D(state_dict=weights("he"), precision="fp16x3")
    solver = {"hqs": T.HQSSolver_CSMRI, "pg": T.PGSolver_CSMRI}[name](den)
    env = CSMRIEnv(None, solver, 3).to(dev)
    data = dict(gt=d["gt"], y0=d["y0"], mask=d["mask"], x0=d["x0"], ATy0=d["x0"], output=d["x0"][..., 0],
                sigma_n=torch.full((4, 1, 32, 32, 2), 15 / 255))
    ob = env.reset({k: v.clone() for k, v in data.items()})
    nv = solver.num_var
    assert tuple(env.state["solver"].shape) == (4, nv, 32, 32, 2)
    keys = solver._param_keys
    g = torch.Generator().manual_seed(1)
    action = {k: (torch.rand(4, 2, generator=g) * 0.5 + 0.1).to(dev) for k in keys}
    action["idx_stop"] = torch.tensor([0, 1, 0, 0], device=dev)
    before = env.state["solver"].clone()
    env.step(action)
    with torch.no_grad():
        want = solver((before, (env.state["y0"], env.state["mask"])), solver.filter_hyperparameter(action))
    assert torch.equal(env.state["solver"], want)
    assert torch.equal(env.state["output"], solver.get_output(want))
    # second step on the 3 images left: rows idx_left = [0, 2, 3] only
    action2 = {k: (torch.rand(3, 2, generator=g) * 0.5 + 0.1).to(dev) for k in keys}
    action2["idx_stop"] = torch.tensor([0, 0, 0], device=dev)
    before = env.state["solver"].clone()
    env.step(action2)
    left = torch.tensor([0, 2, 3], device=dev)
    with torch.no_grad():
        want = solver((before[left], (env.state["y0"][left], env.state["mask"][left])),
                      solver.filter_hyperparameter(action2))
    ref = before.clone(); ref[left] = want
    assert torch.equal(env.state["solver"], ref)
    assert torch.equal(env.state["solver"][1], before[1])
