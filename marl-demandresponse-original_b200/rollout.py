"""Device-resident rollout collection.

The reference's learners step the env one python dict per agent per step:
`{k: agent.select_action(normStateDict(obs_dict[k], config_dict))}` -> `env.step(action)` ->
`agent.store_transition(Transition(state, action, prob, reward, next_state, done), k)`
(train_ppo.py:62-116, agents/ppo.py:68-92).  Here the whole loop stays on the GPU: the step kernel
writes the already-normalised observation `[E, N, F]` *directly into the rollout storage* (zero copy),
the policy is evaluated once per step on the flattened `[E*N, F]` batch, actions are sampled on the
device, and the transition tensors are laid out the way `PPO.update` consumes them
(state, action, action probability, reward, next_state, done).  No host synchronisation per step.

The policy is any callable `probs = policy(obs_flat)` returning `[M, n_actions]` action probabilities
(e.g. the reference's `agents.network.Actor`, which ends in a softmax).
"""
import torch


class DeviceRolloutCollector:
    def __init__(self, env, n_steps):
        self.env, self.n_steps = env, int(n_steps)
        e, n, f, dev, dt = env.n_envs, env.n_houses, env.n_features, env.device, env.dtype
        # states[t] is the observation the action of step t was chosen on; states[t + 1] its successor
        self.states = torch.empty(self.n_steps + 1, e, n, f, dtype=dt, device=dev)
        self.actions = torch.empty(self.n_steps, e, n, dtype=torch.uint8, device=dev)
        self.action_probs = torch.empty(self.n_steps, e, n, dtype=torch.float32, device=dev)
        self.rewards = torch.empty(self.n_steps, e, n, dtype=dt, device=dev)
        self.dones = torch.zeros(self.n_steps, e, n, dtype=torch.bool, device=dev)  # make_dones_dict: never done
        self.power = torch.empty(self.n_steps, e, dtype=torch.float64, device=dev)
        self.signal = torch.empty(self.n_steps, e, dtype=torch.float64, device=dev)
        self._have_first = False

    @torch.no_grad()
    def collect(self, policy, generator=None, reset=False):
        """Runs n_steps env steps under `policy`; returns a dict of views in PPO's transition layout."""
        env = self.env
        e, n, f = env.n_envs, env.n_houses, env.n_features
        if reset or not self._have_first:
            self.states[0].copy_(env.reset_tensor() if reset or env.step_index == 0 else env.observe_tensor())
            self._have_first = True
        else:
            self.states[0].copy_(self.states[self.n_steps])  # continue the episode where the last rollout ended
        for t in range(self.n_steps):
            probs = policy(self.states[t].reshape(e * n, f).float())
            act = torch.multinomial(probs, 1, generator=generator).squeeze(1)  # Categorical(probs).sample()
            self.action_probs[t].copy_(probs.gather(1, act[:, None]).squeeze(1).reshape(e, n))
            self.actions[t].copy_(act.reshape(e, n).to(torch.uint8))
            _, _, p, s = env.step_tensor(self.actions[t], obs_out=self.states[t + 1], reward_out=self.rewards[t])
            self.power[t].copy_(p)
            self.signal[t].copy_(s)
        return dict(state=self.states[:-1], action=self.actions, a_log_prob=self.action_probs, reward=self.rewards,
                    next_state=self.states[1:], done=self.dones, cluster_hvac_power=self.power, reg_signal=self.signal)
