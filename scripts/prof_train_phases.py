"""Device time of the phases of one configs[4] train step (batch 512, bf16 autocast)."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from chinesechesszero_b200.train import TrainPipeline

torch.manual_seed(0)
pipe = TrainPipeline(batch_size=512)
g = torch.Generator(device="cuda").manual_seed(0)
states = (torch.rand((512, 17, 7, 10, 9), device="cuda", generator=g) > 0.95).float()
pi = torch.rand((512, 2086), device="cuda", generator=g) ** 8
pi = pi / pi.sum(1, keepdim=True)
z = torch.randint(-1, 2, (512,), device="cuda", generator=g).float()
for _ in range(3):
    pipe.train_step(states, pi, z)
pv = pipe.policy_value_net
net, opt = pv.policy_value_net, pv.optimizer


def timed(name, fn, n=3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} device {e0.elapsed_time(e1) / n:8.2f} ms   wall {(time.perf_counter() - t0) / n * 1e3:8.2f} ms", flush=True)
    return out


timed("eval forward fp32 (x2/step)", lambda: pipe._policy_value_tensor(states))
net.train()
timed("weights clone", lambda: {k: v.clone() for k, v in net.state_dict().items()})
timed("optimizer deepcopy", lambda: copy.deepcopy(opt.state_dict()))


def fwd_bwd():
    opt.zero_grad()
    loss, *_ = pipe.loss_terms(states, pi, z)
    loss.backward()
    return loss


timed("fwd+bwd bf16 autocast", fwd_bwd)
timed("clip_grad_norm", lambda: torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0))
timed("adam step", lambda: opt.step())
timed("whole train_step", lambda: pipe.train_step(states, pi, z))
