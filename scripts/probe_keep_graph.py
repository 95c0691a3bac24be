"""Development probe: parameters after 3 updates with the step graph kept across update() calls vs re-captured, and the
run-to-run difference of the re-capturing mode itself."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0")
B, mb = 2048, 512
flat = lambda net: torch.cat([p.detach().reshape(-1).float() for p in net.parameters()])
def run(keep, updates=3, graph=True):
    torch.manual_seed(0)
    agent = P.PPO(device=dev)
    agent.K_epochs, agent.keep_graph, agent.use_graph = 2, keep, graph
    g = torch.Generator(device=dev).manual_seed(1)
    buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8),
           "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
           "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
           "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
    outs = []
    for it in range(updates):
        torch.manual_seed(10 + it)
        agent.update(buf, minibatch=mb)
        outs.append(torch.cat([flat(agent.actor), flat(agent.critic)]).clone())
        buf["r"].copy_(torch.rand(B, 1, generator=g, device=dev) - 0.5)
        buf["s"][:, :, ::7] = (buf["s"][:, :, ::7] + 1) % 3
    return outs
a, b, c, e = run(True), run(False), run(False), run(False, graph=False)
for name, x, y in (("keep vs recapture", a, b), ("recapture vs recapture", b, c), ("recapture vs eager", b, e), ("keep vs eager", a, e)):
    print(name, [(round(float((u == v).float().mean()), 5), float((u - v).abs().max())) for u, v in zip(x, y)])
