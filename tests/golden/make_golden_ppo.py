"""Generates ppo_ref.npz by RUNNING THE REFERENCE's networks and PPO.update
(/root/reference/soa/agent/net/all_net.py, soa/agent/PPO.py) on the CPU, through ref_shim.py plus
no-op stubs for tensorboardX / seaborn (logging and plotting only; no arithmetic).  Run only where
the reference exists:   python tests/golden/make_golden_ppo.py

Fixture contents
    x_s, x_p, x_g                 inputs (B=6): state matrices, positions, goal
    actor_prob, critic_v          reference outputs for networks built under torch.manual_seed(0)
    actor_sums, critic_sums       per-parameter float64 sums right after construction
    buf_*                         a 96-sample synthetic rollout buffer (the reference's record fields)
    upd_actor_sums, upd_critic_sums, upd_losses
                                  after PPO.update(buffer) with K_epochs=2, batch_size=32,
                                  torch.manual_seed(1) (drives SubsetRandomSampler)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
tbx = types.ModuleType("tensorboardX")


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass


tbx.SummaryWriter = SummaryWriter
sys.modules["tensorboardX"] = tbx
sns = types.ModuleType("seaborn")
sys.modules["seaborn"] = sns
import matplotlib.pyplot as plt  # the ref_shim stub  # noqa: E402

from agent.net.all_net import Net_PPO_actor, Net_PPO_critic  # noqa: E402
import agent.PPO as ref_ppo  # noqa: E402

ref_ppo.heatmap = lambda *a, **k: None  # plotting side effect at the end of update (PPO.py:161)


def sums(net):
    return np.array([p.detach().double().sum().item() for p in net.parameters()])


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    actor, critic = Net_PPO_actor(), Net_PPO_critic()
    rng = np.random.RandomState(3)
    lut = np.array([0.9, -0.9, -0.5, 0.3], np.float32)
    B = 6
    x_s = lut[rng.randint(0, 4, size=(B, 4, 289))]
    x_p = rng.randint(1, 16, size=(B, 4, 2)).astype(np.float32)
    x_g = np.tile(np.array([[2.0, 14.0]], np.float32), (B, 1))
    with torch.no_grad():
        prob = actor(torch.from_numpy(x_s), torch.from_numpy(x_p), torch.from_numpy(x_g)).numpy()
        val = critic(torch.from_numpy(x_s), torch.from_numpy(x_p), torch.from_numpy(x_g)).numpy()
    out = dict(x_s=x_s, x_p=x_p, x_g=x_g, actor_prob=prob, critic_v=val, actor_sums=sums(actor), critic_sums=sums(critic))

    # one PPO.update on a synthetic buffer
    torch.manual_seed(0)
    agent = ref_ppo.PPO()
    agent.K_epochs, agent.batch_size = 2, 32
    n = 96
    tr = np.dtype([("s", np.float32, (5, 289)), ("a", np.int64, (1,)), ("p", np.float32, (5, 2)), ("g", np.float32, (2,)),
                   ("r", np.float32, (1,)), ("d", np.float32, (1,)), ("a_logp", np.float32, (1,))])
    buf = np.empty(n, dtype=tr)
    buf["s"] = lut[rng.randint(0, 4, size=(n, 5, 289))]
    buf["a"] = rng.randint(0, 5, size=(n, 1))
    buf["p"] = rng.randint(1, 16, size=(n, 5, 2)).astype(np.float32)
    buf["g"] = np.array([2.0, 14.0], np.float32)
    buf["r"] = rng.choice(np.array([-0.01, -0.1, -0.9, 0.2, 0.9], np.float32), size=(n, 1))
    buf["d"] = 0
    buf["a_logp"] = np.log(rng.uniform(0.1, 0.4, size=(n, 1))).astype(np.float32)
    torch.manual_seed(1)
    agent.update(buf, torch.device("cpu"), 0)
    out.update({f"buf_{k}": buf[k] for k in tr.names})
    out.update(upd_actor_sums=sums(agent.actor), upd_critic_sums=sums(agent.critic))
    np.savez_compressed(os.path.join(HERE, "ppo_ref.npz"), **out)
    print("wrote ppo_ref.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
