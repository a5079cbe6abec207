"""Reference-shaped façade of ``train.py`` (train.py:36-422): the consumer of the self-play data.

SURVEY.md §8(f) rank 2 / BASELINE config 5: GPU-generated replay -> the PolicyValueNet training step
with the reference's semantics, in bf16 autocast (the reference uses fp16 autocast + GradScaler on
CUDA and plain fp32 on CPU, train.py:164-207; bf16 needs no loss scaling).  Pure PyTorch -- no custom
kernels on this row.  Per batch, exactly as train.py:130-279:

  validate pi rows; lr = learning_rate * lr_multiplier; old policy eval (KL baseline); backup of weights
  and optimizer state; forward in train mode; loss = mse(value, z) + CE(label-smoothed pi, log p)
  (smoothing 0.05); backward; clip_grad_norm 5.0; Adam step; NaN/inf -> rollback and halve lr_multiplier;
  new policy eval, KL(old||new) batchmean; entropy guard (< 1.0 with non-one-hot targets -> rollback);
  KL > 4*kl_targ -> lr_multiplier /= 1.5; after the pass the KL-adaptive multiplier update.
"""
from __future__ import annotations

import copy
import os
import pickle

import numpy as np
import torch
import torch.nn.functional as F

from .net import PolicyValueNet
from .parameters import BATCH_SIZE, CHECK_FREQ, DATA_DIR, KL_TARG, MODEL_DIR


class NpyReplayDataset:
    """states.npy / mcts.npy / winners.npy, memory-mapped (dataset.py:6-73)."""

    def __init__(self, path: str):
        self.states = np.load(os.path.join(path, "states.npy"), mmap_mode="r")
        self.mcts = np.load(os.path.join(path, "mcts.npy"), mmap_mode="r")
        self.winners = np.load(os.path.join(path, "winners.npy"), mmap_mode="r")
        if not (len(self.states) == len(self.mcts) == len(self.winners)):
            raise ValueError("states / mcts / winners lengths differ")  # dataset.py:57-60

    def __len__(self):
        return len(self.states)

    def batches(self, batch_size: int, shuffle: bool = True, generator=None):
        n = len(self)
        order = torch.randperm(n, generator=generator).numpy() if shuffle else np.arange(n)
        for i in range(0, n, batch_size):
            idx = np.sort(order[i:i + batch_size])
            yield (torch.tensor(np.asarray(self.states[idx]), dtype=torch.float32),
                   torch.tensor(np.asarray(self.mcts[idx]), dtype=torch.float32),
                   torch.tensor(np.asarray(self.winners[idx]), dtype=torch.float32))


class _StepBackup:
    """The rollback copy of train.py:155-163 (weights + optimizer state before the step) kept in preallocated
    buffers and refreshed with multi-tensor copies: same values as ``{k: v.clone()}`` / ``copy.deepcopy(
    opt.state_dict())``, without ~1000 small allocations and launches per step (24 ms -> <1 ms at 50.9 M parameters)."""

    def __init__(self):
        self.w_live = self.w_copy = None
        self.s_live = self.s_copy = None
        self.steps = None
        self.had_state = False

    def save(self, net, opt):
        live = [t for t in net.state_dict().values()]
        if self.w_copy is None or len(self.w_copy) != len(live):
            self.w_copy = [torch.empty_like(t) for t in live]
        self.w_live = live
        torch._foreach_copy_(self.w_copy, live)
        self.had_state = len(opt.state) > 0
        self.steps, s_live = [], []
        for group in opt.param_groups:
            for p in group["params"]:
                st = opt.state.get(p)
                if not st:
                    continue
                self.steps.append((p, st["step"].clone() if torch.is_tensor(st["step"]) else st["step"]))
                s_live.extend(v for k, v in st.items() if k != "step" and torch.is_tensor(v))
        if self.s_copy is None or len(self.s_copy) != len(s_live):
            self.s_copy = [torch.empty_like(t) for t in s_live]
        self.s_live = s_live
        if s_live:
            torch._foreach_copy_(self.s_copy, s_live)

    def restore(self, net, opt):
        with torch.no_grad():
            torch._foreach_copy_(self.w_live, self.w_copy)
            if not self.had_state:
                opt.state.clear()  # the step being undone created the state
                return
            if self.s_live:
                torch._foreach_copy_(self.s_live, self.s_copy)
            for p, step in self.steps:
                if torch.is_tensor(opt.state[p]["step"]):
                    opt.state[p]["step"].copy_(step)
                else:
                    opt.state[p]["step"] = step


class TrainPipeline:
    def __init__(self, init_model: str | None = None, data_dir: str = DATA_DIR, batch_size: int = BATCH_SIZE,
                 autocast_dtype=torch.bfloat16, net_kwargs=None):
        self.learning_rate = 1e-3
        self.lr_multiplier = 1.0
        self.batch_size = batch_size
        self.kl_targ = KL_TARG
        self.check_freq = CHECK_FREQ
        self.label_smoothing = 0.05   # train.py:52
        self.min_entropy_guard = 1.0  # train.py:53
        self.train_iters = 0
        self.data_dir = data_dir
        self.autocast_dtype = autocast_dtype
        self.current_policy_path = os.path.join(MODEL_DIR, "current_policy.pkl")
        net_kwargs = net_kwargs or {}
        try:
            self.policy_value_net = PolicyValueNet(model=init_model, **net_kwargs) if init_model else PolicyValueNet(**net_kwargs)
        except Exception:
            self.policy_value_net = PolicyValueNet(**net_kwargs)  # train.py:66-78
        # NHWC weights: cuDNN's bf16 tensor-core kernels are channels-last; with NCHW parameters every convolution
        # of the forward and backward pass is wrapped in layout transposes.  state_dict keys / values are unchanged.
        if self.policy_value_net.device.type == "cuda" and os.environ.get("CCZ_TRAIN_CHANNELS_LAST", "1") != "0":
            self.policy_value_net.policy_value_net.to(memory_format=torch.channels_last)
        self.dataset = None
        self.last_kl = 0.0
        self._backup = _StepBackup()

    # ---- one batch (train.py:130-267) -------------------------------------------------------
    def _policy_value_tensor(self, state_batch):
        net = self.policy_value_net.policy_value_net
        net.eval()
        with torch.no_grad():
            log_p, v = net(state_batch)
        return torch.exp(log_p.float()), v.float()

    def loss_terms(self, state_batch, mcts_probs_batch, winner_batch):
        """Forward in the current mode -> (loss, policy_loss, value_loss, log_act_probs)."""
        net = self.policy_value_net.policy_value_net
        dev = self.policy_value_net.device
        with torch.autocast(dev.type, dtype=self.autocast_dtype, enabled=dev.type == "cuda"):
            log_act_probs, value = net(state_batch)
            value = value.flatten()
            value_loss = F.mse_loss(value, winner_batch)
            eps = self.label_smoothing
            smooth = (1 - eps) * mcts_probs_batch + eps / mcts_probs_batch.size(1) if eps > 0 else mcts_probs_batch
            policy_loss = -torch.mean(torch.sum(smooth * log_act_probs, dim=1))
            loss = value_loss + policy_loss
        return loss, policy_loss, value_loss, log_act_probs

    def train_step(self, state_batch, mcts_probs_batch, winner_batch) -> dict:
        pv = self.policy_value_net
        net, opt, dev = pv.policy_value_net, pv.optimizer, pv.device
        sums = mcts_probs_batch.sum(dim=1)
        if not ((sums > 0.99) & (sums < 1.01)).all():
            raise ValueError("mcts_probs_batch rows must sum to 1 (±0.01)")  # train.py:134-136
        if torch.isnan(mcts_probs_batch).any():
            raise ValueError("mcts_probs_batch contains NaN")
        for g in opt.param_groups:
            g["lr"] = self.learning_rate * self.lr_multiplier
        state_batch = state_batch.float().to(dev)
        mcts_probs_batch = mcts_probs_batch.float().to(dev)
        winner_batch = winner_batch.float().to(dev)
        old_probs, old_v = self._policy_value_tensor(state_batch)
        net.train()
        opt.zero_grad()
        self._backup.save(net, opt)  # train.py:155-163
        loss, policy_loss, value_loss, log_act_probs = self.loss_terms(state_batch, mcts_probs_batch, winner_batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
        opt.step()
        out = {"loss": float(loss.detach()), "policy_loss": float(policy_loss.detach()), "value_loss": float(value_loss.detach()),
               "rolled_back": False, "counted": False, "kl": None, "entropy": None}
        if (torch.isnan(loss) or torch.isinf(loss) or torch.isnan(log_act_probs).any()
                or torch.isinf(log_act_probs).any()):
            self._backup.restore(net, opt)
            self.lr_multiplier = max(0.05, self.lr_multiplier / 2)
            out["rolled_back"] = True
            return out
        new_probs, new_v = self._policy_value_tensor(state_batch)
        net.train()
        new_log = new_probs.clamp_min(1e-10).log()
        self.last_kl = out["kl"] = float(F.kl_div(new_log, old_probs, reduction="batchmean"))
        with torch.no_grad():
            lp = log_act_probs.float()
            out["entropy"] = float(-torch.mean(torch.sum(torch.exp(lp) * lp, dim=1)))
        out["counted"] = True  # train.py:247-254: the epoch totals take the batch BEFORE the entropy guard looks at it
        if out["entropy"] < self.min_entropy_guard:
            if float((mcts_probs_batch > 0).sum(dim=1).float().mean()) > 1.5:
                self._backup.restore(net, opt)
                self.lr_multiplier = max(0.1, self.lr_multiplier / 2)
                out["rolled_back"] = True
                return out
        if self.last_kl > self.kl_targ * 4:
            self.lr_multiplier = max(0.05, self.lr_multiplier / 1.5)
        out["old_v"], out["new_v"] = old_v, new_v
        return out

    # ---- one pass over the data (train.py:80-311) ---------------------------------------------
    def policy_update(self, generator=None):
        if self.dataset is None:
            self.dataset = NpyReplayDataset(self.data_dir)
        tot = {"loss": 0.0, "entropy": 0.0}
        n = 0
        for batch in self.dataset.batches(self.batch_size, shuffle=True, generator=generator):
            r = self.train_step(*batch)
            if not r["counted"]:  # NaN / inf roll-back (train.py:211-221); entropy-guard roll-backs are counted
                continue
            tot["loss"] += r["loss"]
            tot["entropy"] += r["entropy"]
            n += 1
        if self.last_kl > self.kl_targ * 2 and self.lr_multiplier > 0.05:
            self.lr_multiplier = max(0.05, self.lr_multiplier / 1.2)
        elif self.last_kl < self.kl_targ / 2 and self.lr_multiplier < 2.0:
            self.lr_multiplier = min(2.0, self.lr_multiplier * 1.2)
        return tot["loss"] / max(1, n), tot["entropy"] / max(1, n)

    def save_train_state(self, path=None):
        path = path or os.path.join(MODEL_DIR, "train_state.pkl")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:  # train.py:321-340
            pickle.dump({"train_iters": self.train_iters, "lr_multiplier": self.lr_multiplier}, f)

    def run(self, iterations: int):
        """train.py:360-422 without the endless loop: iterate, save current_policy.pkl, numbered copies
        every CHECK_FREQ iterations."""
        os.makedirs(MODEL_DIR, exist_ok=True)
        for _ in range(iterations):
            self.policy_update()
            self.train_iters += 1
            self.policy_value_net.save_model(self.current_policy_path)
            self.save_train_state()
            if self.train_iters % self.check_freq == 0:
                self.policy_value_net.save_model(os.path.join(MODEL_DIR, f"current_policy_batch{self.train_iters}.pkl"))
