"""Generates predictor_ref.npz by RUNNING THE REFERENCE's predictor networks and ppo_predictor
(soa/agent/net/all_net.py:7-137,249-305, soa/agent/PPO_Predictor.py) on the CPU through ref_shim.py and
no-op stubs for tensorboardX / seaborn.  Run only where the reference exists.

Contents: per-parameter sums of every network built in the reference's order under
torch.manual_seed(0); pred_states output and the actor / critic outputs on fixed inputs; parameter sums
after one ppo_predictor.update (K_epochs=1, batch_size=16, torch.manual_seed(1)) on a synthetic buffer of
the reference's 9-frame records."""
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
tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda s, *a, **k: None, "add_scalar": lambda s, *a, **k: None})
sys.modules["tensorboardX"] = tbx
sys.modules["seaborn"] = types.ModuleType("seaborn")
import agent.PPO_Predictor as ref  # noqa: E402

ref.heatmap = lambda *a, **k: None


def sums(net):
    return np.array([p.detach().double().sum().item() for p in net.parameters()])


def main():
    torch.set_num_threads(1)
    torch.manual_seed(0)
    agent = ref.ppo_predictor()
    dev = torch.device("cpu")
    agent.encoder.device = dev
    agent.predictor.device = dev
    out = {f"{n}_sums": sums(getattr(agent, n)) for n in ("actor", "critic", "encoder", "decoder", "predictor")}
    rng = np.random.RandomState(5)
    lut = np.array([0.9, -0.9, -0.5, 0.3], np.float32)
    B = 3
    x = lut[rng.randint(0, 4, size=(B, 4, 289))]
    p = rng.randint(1, 16, size=(B, 4, 2)).astype(np.float32)
    g = np.tile(np.array([[2.0, 14.0]], np.float32), (B, 1))
    with torch.no_grad():
        pred, _, _ = agent.pred_states(torch.from_numpy(x))
        cat = torch.cat([torch.from_numpy(x), pred], 1)
        agent.actor.eval(); agent.critic.eval()
        prob = agent.actor(cat, torch.from_numpy(p), torch.from_numpy(g)).numpy()
        val = agent.critic(cat, torch.from_numpy(p), torch.from_numpy(g)).numpy()
    out.update(x=x, p=p, g=g, pred=pred.numpy(), actor_prob=prob, critic_v=val)
    n = 32
    tr = np.dtype([("s", np.float64, (9, 289)), ("a", np.int64, (5, 1)), ("p", np.float64, (9, 2)), ("g", np.float64, (2,)),
                   ("r", np.float64, (5, 1)), ("d", np.int64, (5, 1)), ("a_logp", np.float64, (5, 1))])
    buf = np.zeros(n, dtype=tr)
    buf["s"] = lut[rng.randint(0, 4, size=(n, 9, 289))]
    buf["a"] = rng.randint(0, 5, size=(n, 5, 1))
    buf["p"] = rng.randint(1, 16, size=(n, 9, 2))
    buf["g"] = np.array([2.0, 14.0])
    buf["r"] = rng.choice(np.array([-0.01, -0.1, 0.2, 0.9]), size=(n, 5, 1))
    buf["a_logp"] = np.log(rng.uniform(0.1, 0.4, size=(n, 5, 1)))
    agent.K_epochs, agent.batch_size = 1, 16
    torch.manual_seed(1)
    agent.update(buf, dev, 0)
    out.update({f"buf_{k}": buf[k] for k in ("s", "a", "p", "g", "r", "a_logp")})
    out.update(upd_actor_sums=sums(agent.actor), upd_critic_sums=sums(agent.critic))
    np.savez_compressed(os.path.join(HERE, "predictor_ref.npz"), **out)
    print("wrote predictor_ref.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
