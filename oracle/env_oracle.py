"""CPU restatement of the environment bookkeeping around the solver (TEST INFRASTRUCTURE ONLY).

Follows tfpnp/env/base.py:121-191 (``PnPEnv.reset`` / ``step``), :225-242 (metric, reward,
``torch_psnr``) and the task environments tasks/{csmri,pr,ct,spi}/env.py (``_observation``,
``get_policy_ob``) in plain PyTorch on the CPU, with the inner loop taken from
``oracle/pnp_oracle.py``.  Pinned against the unmodified reference classes by
``oracle/make_golden.py`` (fixture tests/golden/env_csmri.npz, env_spi.npz).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import pnp_oracle as O

Tensor = torch.Tensor


def complex2channel(x: Tensor) -> Tensor:        # transforms.py:20-26
    N, C, H, W, _ = x.shape
    return x.permute(0, 1, 4, 2, 3).contiguous().view(N, C * 2, H, W)


class EnvOracle:
    """One environment for all four tasks; `task` in {'csmri','pr','ct','spi'}."""

    OB_KEYS = {
        "csmri": ("gt", "y0", "ATy0", "mask", "sigma_n", "T"),      # tasks/csmri/env.py:49-57
        "pr": ("gt", "y0", "x0", "mask", "sigma_n", "T"),           # tasks/pr/env.py:47-56
        "ct": ("gt", "y0", "ATy0", "view", "sigma_n", "T"),         # tasks/ct/env.py:45-54
        "spi": ("gt", "x0", "K", "T"),                              # tasks/spi/env.py:42-50
    }

    def __init__(self, task: str, sd, max_episode_step: int, opnorm: float = 0.0):
        self.task, self.sd, self.max_episode_step, self.opnorm = task, sd, max_episode_step, opnorm
        self.complex = task in ("csmri", "pr")

    # -- solver side ---------------------------------------------------------------------------
    def _solver_reset(self, data):
        return O.pr_reset(data["x0"]) if self.task == "pr" else O.admm_reset(data["x0"])

    def _solve(self, variables, aux, action):
        sd = self.sd
        if self.task == "csmri":
            return O.admm_csmri(sd, variables, aux[0], aux[1], action["sigma_d"], action["mu"])
        if self.task == "pr":
            return O.iadmm_pr(sd, variables, aux[0], aux[1], action["sigma_d"], action["mu"], action["tau"])
        if self.task == "ct":
            views = int(aux[1][0, 0, 0, 0].item() * 120)           # tasks/ct/solver.py:26
            return O.iadmm_ct(sd, variables, aux[0], views, self.opnorm, action["sigma_d"], action["mu"],
                              action["tau"])
        return O.admm_spi(sd, variables, aux[0], aux[1], action["sigma_d"], action["mu"])

    def _aux(self, state):
        return {"csmri": ("y0", "mask"), "pr": ("y0", "mask"), "ct": ("y0", "view"), "spi": ("x0", "K")}[self.task]

    # -- PnPEnv --------------------------------------------------------------------------------
    def reset(self, data: Dict[str, Tensor]):
        self.cur_step = 0
        data = dict(data)
        data["solver"] = self._solver_reset(data)                   # base.py:141-142
        B, _, W, H = data["gt"].shape
        data["T"] = torch.ones([B, 1, W, H]) * self.cur_step / self.max_episode_step
        self.state = data
        self.idx_left = torch.arange(0, B)
        self.last_metric = O.psnr(data["output"], data["gt"])
        return self.observation()

    def step(self, action):
        self.cur_step += 1
        idx = self.idx_left
        aux = tuple(self.state[k][idx] for k in self._aux(self.state))
        with torch.no_grad():
            s = self._solve(self.state["solver"][idx], aux, action)
        self.state["T"] = torch.ones_like(self.state["T"]) * self.cur_step / self.max_episode_step
        self.state["output"][idx] = O.get_output(s, self.complex)
        self.state["solver"][idx] = s
        metric = O.psnr(self.state["output"], self.state["gt"])
        reward = metric - self.last_metric
        self.last_metric = metric
        ob = self.observation()
        idx_stop = action["idx_stop"]
        self.idx_left = self.idx_left[idx_stop == 0]
        all_done = len(self.idx_left) == 0
        done = idx_stop.detach()
        if self.cur_step == self.max_episode_step:
            all_done = True
            done = torch.ones_like(idx_stop)
        return ob, self.observation(), reward, all_done, {"done": done}

    def observation(self):
        idx = self.idx_left
        ob = {"variables": self.state["solver"][idx]}
        for k in self.OB_KEYS[self.task]:
            v = self.state[k][idx]
            ob[k] = v.float() if k == "mask" else v
        return ob

    def policy_ob(self, ob):
        c2r = O.complex2real
        if self.task == "csmri":                                     # tasks/csmri/env.py:14-23
            parts = [c2r(ob["variables"]), complex2channel(ob["y0"]), c2r(ob["ATy0"]), ob["mask"], ob["T"],
                     c2r(ob["sigma_n"])]
        elif self.task == "pr":                                      # tasks/pr/env.py:14-21
            parts = [c2r(ob["variables"]), ob["y0"], complex2channel(ob["mask"]), ob["T"], ob["sigma_n"]]
        elif self.task == "ct":                                      # tasks/ct/env.py:12-19
            parts = [ob["variables"], ob["ATy0"], ob["view"], ob["T"], ob["sigma_n"]]
        else:                                                        # tasks/spi/env.py:12-18
            parts = [ob["variables"], ob["x0"], ob["K"], ob["T"]]
        return torch.cat(parts, 1)


def env_data(task: str, d: Dict[str, Tensor], sigma_n: float = 15 / 255) -> Dict[str, Tensor]:
    """The dict a task dataset hands to PnPEnv.reset, built from an oracle/synth.py batch
    (tasks/csmri/dataset.py:62-74, tasks/pr/dataset.py:57-66, tasks/ct/dataset.py:95-104,
    tasks/spi/dataset.py:55-66)."""
    gt = d["gt"]
    if task == "csmri":
        aty0 = d["x0"]
        return dict(y0=d["y0"], x0=d["x0"].clone(), ATy0=aty0.clone(), gt=gt, mask=d["mask"].bool(),
                    sigma_n=torch.ones_like(d["y0"]) * sigma_n, output=O.complex2real(aty0).clone(),
                    input=d["x0"].clone())
    if task == "pr":
        return dict(y0=d["y0"], x0=d["x0"].clone(), gt=gt, mask=d["mask"], output=d["x0"].clone(),
                    sigma_n=torch.ones_like(gt) * (27 / 255), input=d["x0"].clone())
    if task == "ct":
        return dict(y0=d["y0"], x0=d["x0"].clone(), ATy0=d["x0"].clone(), gt=gt, view=d["view"],
                    output=d["x0"].clone(), sigma_n=torch.ones_like(gt) * 0.05, input=d["x0"].clone())
    return dict(x0=d["x0"].clone(), gt=gt, K=d["K"], output=d["x0"].clone(), input=d["x0"].clone())


def episode_actions(task: str, d: Dict[str, Tensor], B: int, steps: int, pack: int, seed: int = 11):
    """A deterministic action sequence with images dropping out of the episode (idx_stop) at
    different steps, as the policy's termination head would produce (base.py:179-181)."""
    g = torch.Generator().manual_seed(seed)
    alive = B
    actions = []
    for s in range(steps):
        sl = slice(s * pack, (s + 1) * pack)
        idx_stop = (torch.rand(alive, generator=g) < 0.35).long() if s > 0 else torch.zeros(alive, dtype=torch.long)
        rows = torch.randperm(B, generator=g)[:alive].sort().values      # parameters of the surviving rows
        a = {"sigma_d": d["sigma_d"][rows][:, sl].contiguous(), "mu": d["mu"][rows][:, sl].contiguous(),
             "idx_stop": idx_stop}
        if "tau" in d:
            a["tau"] = d["tau"][rows][:, sl].contiguous()
        actions.append(a)
        alive = int((idx_stop == 0).sum())
        if alive == 0:
            break
    return actions
