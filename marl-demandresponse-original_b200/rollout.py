"""Device-resident rollout collection.

The reference's learners step the env one python dict per agent per step:
`{k: agent.select_action(normStateDict(obs_dict[k], config_dict))}` -> `env.step(action)` ->
`agent.store_transition(Transition(state, action, a_log_prob, reward, next_state, done), k)`
(train_ppo.py:62-116, agents/ppo.py:68-92).  Here the whole loop stays on the GPU:

* the step kernel writes the already-normalised observation `[E, N, F]` *directly into the rollout storage*
  (zero copy: `states[t + 1]` is the kernel's output buffer);
* the policy is evaluated once per step on the flattened `[E*N, F]` batch (any callable returning `[M, A]` action
  probabilities, e.g. the reference's `agents.network.Actor`, which ends in a softmax);
* ONE kernel (`mdr_sample_actions`) draws `Categorical(probs).sample()` for every agent (Philox), writes the uint8
  action where the step kernel reads it and the chosen probability into the storage;
* the T-step loop (policy -> sample -> step, 3 + the policy's kernels per step) is captured once into a CUDA graph and
  replayed; Philox counters live on the device (`MdrStepInputs.step_counter`), so replays do not repeat their draws;
* `ppo_batch()` returns the transitions in exactly the layout `PPO.update` builds from its per-agent buffers
  (agents/ppo.py:92-107): agent after agent, time inside.

No host synchronisation anywhere in `collect()`.
"""
import ctypes as C

import torch

from . import _lib


class ActorMLP(torch.nn.Module):
    """The reference's actor architecture (agents/network.py:14-33): Linear-ReLU stack ending in a softmax over the
    actions.  Restated (the reference module is not importable on the GPU box); `Actor(51, 2, [100, 100])` is
    BASELINE config 2's policy."""

    def __init__(self, num_state, num_action, layers):
        super().__init__()
        dims = [int(num_state)] + [int(x) for x in layers]
        self.fc = torch.nn.ModuleList([torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
        self.fc.append(torch.nn.Linear(dims[-1], int(num_action)))

    def forward(self, x):
        for layer in self.fc[:-1]:
            x = torch.relu(layer(x))
        return torch.softmax(self.fc[-1](x), dim=1)


class DeviceRolloutCollector:
    def __init__(self, env, n_steps, episode_steps=None, seed=0, use_graph=True):
        self.env, self.n_steps = env, int(n_steps)
        e, n, f, dev, dt = env.n_envs, env.n_houses, env.n_features, env.device, env.dtype
        if env.obs is None:
            raise ValueError("the collector needs an env built with with_obs=True")
        # every states[t] must be a legal output buffer of the step kernel: 16-byte aligned (bulk stores).  The per-step
        # stride E*N*F*itemsize is padded up to a multiple of 16 bytes (e.g. 1 x 50 x 51 fp32 = 10 200 B -> 10 208 B).
        item = torch.empty((), dtype=dt).element_size()
        row = e * n * f
        self._stride = (row * item + 15) // 16 * 16 // item
        self._state_buf = torch.empty((self.n_steps + 1) * self._stride, dtype=dt, device=dev)
        # states[t] is the observation the action of step t was chosen on; states[t + 1] its successor
        self.states = torch.as_strided(self._state_buf, (self.n_steps + 1, e, n, f), (self._stride, n * f, f, 1))
        self.actions = torch.empty(self.n_steps, e, n, dtype=torch.uint8, device=dev)
        self.action_probs = torch.empty(self.n_steps, e, n, dtype=torch.float32, device=dev)
        self.rewards = torch.empty(self.n_steps, e, n, dtype=dt, device=dev)
        self.dones = torch.zeros(self.n_steps, e, n, dtype=torch.bool, device=dev)
        self.power = torch.empty(self.n_steps, e, dtype=torch.float64, device=dev)
        self.signal = torch.empty(self.n_steps, e, dtype=torch.float64, device=dev)
        self.episode_steps = None if episode_steps is None else int(episode_steps)
        self.seed = int(seed)
        self.total_steps = 0
        self._have_first = False
        self._counter = torch.zeros(1, dtype=torch.int64, device=dev)  # device-side Philox counter (graph replays)
        self.use_graph = bool(use_graph)
        self._graph, self._graph_policy = None, None

    # ------------------------------------------------------------------ pieces of one step
    def _sample(self, probs, t):
        m, a = probs.shape
        if probs.dtype != torch.float32 or not probs.is_contiguous():
            probs = probs.float().contiguous()
        env = self.env
        with torch.cuda.device(env.device):
            _lib.check(env.lib.mdr_sample_actions(
                C.c_void_p(probs.data_ptr()), m, a, self.seed, t, C.c_void_p(self._counter.data_ptr()),
                C.c_void_p(self.actions[t].data_ptr()), C.c_void_p(self.action_probs[t].data_ptr()), env._stream()),
                "mdr_sample_actions")
        return probs  # keep alive until the launch is enqueued (stream order protects the memory afterwards)

    def _step(self, t, step_kwargs=None):
        env = self.env
        kw = step_kwargs(t) if step_kwargs is not None else {}
        _, _, p, s = env.step_tensor(self.actions[t], obs_out=self.states[t + 1], reward_out=self.rewards[t],
                                     step_counter=self._counter, **kw)
        self.power[t].copy_(p)
        self.signal[t].copy_(s)

    def _rollout_body(self, policy, step_kwargs=None):
        e, n, f = self.env.n_envs, self.env.n_houses, self.env.n_features
        keep = []
        for t in range(self.n_steps):
            probs = policy(self.states[t].reshape(e * n, f).float())
            keep.append(self._sample(probs, t))
            self._step(t, step_kwargs)
        return keep

    # ------------------------------------------------------------------ API
    @torch.no_grad()
    def collect(self, policy=None, reset=False, forced=None, step_kwargs=None):
        """Runs n_steps env steps; returns a dict of views of the rollout storage.
        `policy(obs_flat [M, F] fp32) -> probs [M, A]`; `forced = (actions [T, E, N], probs [T, E, N])` replays recorded
        decisions instead (trace tests); `step_kwargs(t)` supplies replayed noise for step t (parity mode, eager)."""
        env = self.env
        if reset or not self._have_first:
            self.states[0].copy_(env.reset_tensor() if reset or env.step_index == 0 else env.observe_tensor())
            self._have_first = True
        else:
            self.states[0].copy_(self.states[self.n_steps])  # continue the episode where the last rollout ended
        if self.episode_steps:
            # train_ppo.py:86: done = t % time_steps_per_episode == time_steps_per_episode - 1 (an episode boundary of the
            # training loop; the env itself never terminates, make_dones_dict :375-390)
            t_abs = torch.arange(self.total_steps, self.total_steps + self.n_steps, device=env.device)
            self.dones.copy_((t_abs % self.episode_steps == self.episode_steps - 1)[:, None, None].expand_as(self.dones))
        if forced is not None:
            fa, fp = forced
            self.actions.copy_(torch.as_tensor(fa).to(env.device).reshape(self.actions.shape))
            self.action_probs.copy_(torch.as_tensor(fp).to(env.device).reshape(self.action_probs.shape))
            for t in range(self.n_steps):
                self._step(t, step_kwargs)
        elif self.use_graph and step_kwargs is None:
            self._collect_graph(policy)
        else:
            self._rollout_body(policy, step_kwargs)
        self._counter.add_(self.n_steps)
        self.total_steps += self.n_steps
        return self.views()

    def _collect_graph(self, policy):
        if self._graph is None or self._graph_policy is not policy:
            # one eager rollout warms up lazy initialisations (cuBLAS handles, kernel attributes) on a side stream, with
            # the env state saved and restored, then the same T steps are captured
            env = self.env
            saved = env.state_dict()
            side = torch.cuda.Stream(device=env.device)
            side.wait_stream(torch.cuda.current_stream(env.device))
            with torch.cuda.stream(side):
                self._rollout_body(policy)
            torch.cuda.current_stream(env.device).wait_stream(side)
            env.load_state_dict(saved)
            step0 = env.step_index
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._keep = self._rollout_body(policy)
            env.load_state_dict(saved)        # capture does not execute; keep the host-side step counter where it was
            env.step_index = step0
            self._graph, self._graph_policy = graph, policy
            self._graph_step0 = step0
        self._graph.replay()
        self.env.step_index += self.n_steps

    def views(self):
        return dict(state=self.states[:-1], action=self.actions, a_log_prob=self.action_probs, reward=self.rewards,
                    next_state=self.states[1:], done=self.dones, cluster_hvac_power=self.power, reg_signal=self.signal)

    def ppo_batch(self):
        """The tensors `PPO.update` builds from its buffers (agents/ppo.py:92-107): `sequential_buffer` is agent 0's
        transitions in time order, then agent 1's, ... -> row index = (env * N + agent) * T + t.
        Returns dict(state [M, F] fp32, next_state [M, F] fp32, action [M, 1] int64, old_action_log_prob [M, 1] fp32,
        reward [M], done [M]) with M = E * N * T."""
        t, e, n, f = self.n_steps, self.env.n_envs, self.env.n_houses, self.env.n_features
        agent_major = lambda x: x.permute(1, 2, 0, *range(3, x.dim())).reshape(e * n * t, *x.shape[3:])
        return dict(state=agent_major(self.states[:-1]).float(), next_state=agent_major(self.states[1:]).float(),
                    action=agent_major(self.actions).long().view(-1, 1),
                    old_action_log_prob=agent_major(self.action_probs).float().view(-1, 1),
                    reward=agent_major(self.rewards), done=agent_major(self.dones))
