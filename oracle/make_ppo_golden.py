"""TEST INFRASTRUCTURE -- generator of tests/golden/mc_ppo_rollout.npz (build container only: needs /root/reference).

Runs the reference's UNMODIFIED training loop `train_ppo.train_ppo` (train_ppo.py:27-152: select_action per agent ->
env.step -> store_transition per agent -> agent.update at the epoch boundary) on the unmodified env, with a stub agent
whose `select_action` returns forced (seeded) actions / probabilities and whose `update` executes the batch-building
lines of the reference's own `PPO.update` (agents/ppo.py, from `sequential_buffer =  []` to `done = [...]`, located by
text and run from source).  Records the env's random draws (like oracle/make_golden.py) and the resulting
(state, action, a_log_prob, reward, next_state, done) batch: the fixture that pins the layout the device-resident
rollout collector hands to a PPO learner."""
import os
import random
import sys
import textwrap
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, Recorder, base_config, config_to_json, snapshot  # noqa: E402

T, N, SEED = 24, 12, 21


class ForcedAgent:
    """Duck-typed stand-in for agents.ppo.PPO: forced decisions in, the reference's own batch construction out."""
    FIRST, LAST = "sequential_buffer =  []", "done = [t.done for t in sequential_buffer]"

    def __init__(self, nb_agents, rng):
        self.nb_agents, self.rng = nb_agents, rng
        self.batch_size = 1
        self.device = "cpu"
        self.buffer = {a: [] for a in range(nb_agents)}      # PPO.reset_buffer, agents/ppo.py:84-87
        self.forced_actions, self.forced_probs, self.seen_states = [], [], []
        src = open(os.path.join(ref_stubs.REFERENCE_ROOT, "agents", "ppo.py"), encoding="utf-8").read().split("\n")
        i0 = next(k for k, l in enumerate(src) if l.strip() == self.FIRST)
        i1 = next(k for k, l in enumerate(src) if l.strip() == self.LAST)
        self.code = compile(textwrap.dedent("\n".join(src[i0:i1 + 1])), "agents/ppo.py:update", "exec")
        self.batch = None

    def select_action(self, state):                           # agents/ppo.py:68-75, decisions forced
        a = int(self.rng.random() < 0.5)
        p = float(self.rng.uniform(0.05, 0.95))
        self.forced_actions.append(a)
        self.forced_probs.append(p)
        self.seen_states.append(np.asarray(state, dtype=np.float64))
        return a, p

    def store_transition(self, transition, agent):            # agents/ppo.py:89-90
        self.buffer[agent].append(transition)

    def update(self, t):                                      # agents/ppo.py:92-107, executed from the reference source
        import torch
        ns = {"self": self, "np": np, "torch": torch}
        exec(self.code, ns)
        self.batch = {k: ns[k] for k in ("state", "next_state", "action", "old_action_log_prob", "reward", "done")}


def main():
    Env, norm, cfg, ref_utils = ref_stubs.import_reference()
    import train_ppo as ref_train  # the unmodified loop
    c = base_config(cfg, N)
    c["training_prop"] = dict(cfg["training_prop"])
    # one episode, one epoch, one train log; no test / actor saving inside the recorded window
    c["training_prop"].update(nb_time_steps=T, nb_tr_episodes=1, nb_tr_epochs=1, nb_tr_logs=1, nb_test_logs=0.25,
                              nb_inter_saving_actor=0)
    opt = types.SimpleNamespace(save_actor_name=None, render_after=0)
    random.seed(SEED)
    rec = Recorder(ref_utils.Perlin, np_seed=SEED + 1000)
    try:
        env = Env(c)
        # train_ppo's first act is env.reset(): wrap it to snapshot the population it starts from
        state = {}
        orig_reset = env.reset

        def reset():
            obs = orig_reset()
            if "snap" not in state:
                rec.drain()
                state["snap"] = snapshot(env)
            return obs

        env.reset = reset
        agent = ForcedAgent(N, np.random.default_rng(SEED + 1))
        # record the env's draws step by step
        orig_step = env.step
        draws = dict(od=[], sig=[])

        def step(action):
            out = orig_step(action)
            gauss, choices, samples, rands, perlin = rec.drain()
            assert len(gauss) == 1
            draws["od"].append(gauss[0])
            draws["sig"].append(perlin[0] if perlin else 0.0)
            return out

        env.step = step
        ref_train.train_ppo(env, agent, opt, c, False, False, None)
    finally:
        rec.close()
    b = agent.batch
    assert b is not None and b["state"].shape == (N * T, 51)
    data = dict(config_json=np.array(config_to_json(c)), seed=np.int64(SEED), steps=np.int64(T),
                od_noise=np.array(draws["od"][:T]), sig_noise=np.array(draws["sig"][:T]),
                forced_actions=np.array(agent.forced_actions, np.uint8).reshape(T, N),
                forced_probs=np.array(agent.forced_probs, np.float64).reshape(T, N),
                state=b["state"].numpy(), next_state=b["next_state"].numpy(), action=b["action"].numpy(),
                old_action_log_prob=b["old_action_log_prob"].numpy(), reward=np.array(b["reward"], np.float64),
                done=np.array(b["done"], np.bool_))
    for k, v in state["snap"].items():
        data["snap_" + k] = v
    path = os.path.join(GOLDEN_DIR, "mc_ppo_rollout.npz")
    np.savez_compressed(path, **data)
    print("mc_ppo_rollout: N=%d T=%d batch %s  %.1f KB" % (N, T, b["state"].shape, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
