"""TEST INFRASTRUCTURE ONLY -- golden decisions of the UNMODIFIED reference GreedyMyopic controller
(agents/greedy_myopic_controller.py:29-49) on random observation dicts with distinct temperatures, to pin
oracle.mdr_oracle.greedy_myopic_actions (and through it the on-device action source MDR_ACT_GREEDY).

    python oracle/make_greedy_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_stubs  # noqa: E402


def main():
    ref_stubs.import_reference()
    import agents.greedy_myopic_controller as gm  # type: ignore
    rng = np.random.default_rng(7)
    cases = []
    for n, sig_frac in ((5, 0.5), (12, 0.3), (30, 0.7), (50, 0.05), (50, 1.5), (17, 0.0)):
        for rep in range(4):
            t_air = 20 + rng.normal(0, 3, n)
            target = 20 + np.abs(rng.normal(0, 1, n))
            cap = rng.choice([10000.0, 12500.0, 15000.0, 17500.0, 20000.0], n)
            cop = 2.5
            lockout = rng.random(n) < 0.3
            signal = float(sig_frac * (cap / cop).sum())
            obs = {i: {"house_temp": float(t_air[i]), "house_target_temp": float(target[i]), "hvac_cooling_capacity": float(cap[i]),
                       "hvac_COP": cop, "hvac_lockout": bool(lockout[i]), "reg_signal": signal} for i in range(n)}
            gm.global_myopic_memory[0], gm.global_myopic_memory[1] = None, None
            agents = {i: gm.GreedyMyopic({"id": i}, {}) for i in range(n)}
            act = np.array([int(agents[i].act(obs)) for i in range(n)], np.uint8)
            cases.append(dict(t_air=t_air, target=target, cap=cap, cop=cop, lockout=lockout, signal=signal, action=act))
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "mc_greedy_myopic.npz")
    flat = {}
    for k, c in enumerate(cases):
        for name, v in c.items():
            flat["%d_%s" % (k, name)] = np.asarray(v)
    np.savez_compressed(out, n_cases=len(cases), **flat)
    print("saved", out, len(cases), "cases; on fractions", [round(float(c["action"].mean()), 2) for c in cases][:8])


if __name__ == "__main__":
    main()
