"""The other CS-MRI solvers of the reference's _solver_map: HQS, PG, APG, RED-ADMM
(tasks/csmri/solver.py:60-201).  Fixture tests/golden/csmri_variants.npz was recorded from the UNMODIFIED reference
classes (oracle/make_golden.py); CPU: the oracle restatement against it; GPU (-m gpu): the native solvers against it
(fp16x3: 1e-4, fp16: 5e-3 on the variance-preserving 'he' weights) and the reset / get_output / num_var interface."""
import pytest
import torch

from conftest import load_golden, rel_err, weights
from oracle import pnp_oracle as O

SPECS = {"hqs": (O.hqs_csmri, ("sigma_d", "mu"), 2), "pg": (O.pg_csmri, ("sigma_d", "tau"), 1),
         "apg": (O.apg_csmri, ("sigma_d", "tau", "beta"), 2), "redadmm": (O.redadmm_csmri, ("sigma_d", "mu", "lamda"), 3)}


@pytest.mark.parametrize("name", list(SPECS))
def test_oracle_variants_match_reference_fixture(name):
    g = load_golden("csmri_variants")
    fn, keys, nvar = SPECS[name]
    out = fn(weights("he"), g[name + "_state0"], g["y0"], g["mask"], *[g[k] for k in keys])
    assert rel_err(out, g[name + "_out"])[1] <= 2e-6
    assert g[name + "_state0"].shape[1] == nvar


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp16x3", 1e-4), ("fp16", 5e-3)])
@pytest.mark.parametrize("name", list(SPECS))
def test_native_variants_match_reference_fixture(dev, name, prec, tol):
    import tfpnp_b200 as T
    g = load_golden("csmri_variants")
    _, keys, nvar = SPECS[name]
    cls = {"hqs": T.HQSSolver_CSMRI, "pg": T.PGSolver_CSMRI, "apg": T.APGSolver_CSMRI, "redadmm": T.REDADMMSolver_CSMRI}[name]
    s = cls(T.UNetDenoiser2D(state_dict=weights("he"), precision=prec))
    assert s.num_var == nvar
    state0 = s.reset({"x0": g["x0"].to(dev)})
    assert torch.equal(state0.cpu(), g[name + "_state0"])
    action = {k: g[k].to(dev) for k in keys}
    with torch.no_grad():
        out = s((state0, s.filter_aux_inputs({"y0": g["y0"].to(dev), "mask": g["mask"].to(dev)})), s.filter_hyperparameter(action))
    l2, mx = rel_err(out, g[name + "_out"])
    assert l2 <= tol and mx <= tol, (name, prec, l2, mx)
    ref_out = O.complex2real(torch.split(g[name + "_out"], g[name + "_out"].shape[1] // nvar, dim=1)[0])
    assert rel_err(s.get_output(out), ref_out)[1] <= tol
    # zero iterations: the reference returns the state unchanged
    with torch.no_grad():
        same = s((state0, (g["y0"].to(dev), g["mask"].to(dev))), tuple(a[:, :0] for a in s.filter_hyperparameter(action)), iter_num=0)
    assert torch.equal(same, state0)


@pytest.mark.gpu
def test_pg_ct_vs_oracle(dev):
    """PGSolver_CT (tasks/ct/solver.py:56-87) on this build's Radon pair against the CPU restatement."""
    import tfpnp_b200 as T
    from oracle import synth
    d = synth.ct_batch(2, 64, 24, 3, seed=5)
    ref = O.pg_ct(weights("he"), d["x0"], d["y0"], d["views"], d["opnorm"], d["sigma_d"], d["tau"])
    s = T.PGSolver_CT(T.UNetDenoiser2D(state_dict=weights("he"), precision="fp16x3"))
    s.opnorm_override = d["opnorm"]
    state0 = s.reset({"x0": d["x0"].to(dev)})
    with torch.no_grad():
        out = s((state0, (d["y0"].to(dev), d["view"].to(dev))), (d["sigma_d"].to(dev), d["tau"].to(dev)))
    assert rel_err(out, ref)[1] <= 1e-4
    assert s.num_var == 1 and torch.equal(s.get_output(out), out)
