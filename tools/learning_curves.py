"""Qualitative learning check: train from the reference's pretrained weights with one adversary and report
the team return / critic estimate per block for H = 0 and H = 1.  Reference results (simulation_results, README):
malicious attacker: H=0 -> team return -7.20 with the critic estimate poisoned to +4.24; H=1 -> -5.46 (estimate -5.65).

    python tools/learning_curves.py [n_blocks] [n_envs] [scenario]     (scenario: malicious | greedy | faulty | coop)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "resilient-consensus-based-marl_b200"))
import bench                                   # noqa: E402
from rcmarl.trainer import Trainer             # noqa: E402

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
scenario = sys.argv[3] if len(sys.argv) > 3 else "malicious"
common = len(sys.argv) > 4 and sys.argv[4] == "global"          # the reference's *_global runs: --common_reward True
label = {"malicious": "Malicious", "greedy": "Greedy", "faulty": "Faulty", "coop": "Cooperative"}[scenario]
w, desired, _ = bench.load_pretrained()
labels = ["Cooperative"] * 4 + [label]
out = {}
for H in (0, 1):
    tr = Trainer(labels=labels, in_nodes=bench.IN_NODES5, weights=w, desired=desired, n_envs=n_envs, H=H, seed=100,
                 common_reward=common, **bench.HYPER)
    coop = [i for i, l in enumerate(labels) if l == "Cooperative"]
    rows = []
    for b in range(n_blocks):
        est, ret = tr.rollout_block()
        tr.update_round()
        rows.append((float(ret[:, coop].mean()), float(est[:, coop].mean()), float(ret[:, 4].mean())))
        if b % 5 == 0 or b == n_blocks - 1:
            print(f"{scenario}{'_global' if common else ''} H={H} block {b:3d} (episodes {50 * b}-{50 * b + 49}): team return {rows[-1][0]:7.3f}  "
                  f"critic estimate {rows[-1][1]:7.3f}  adversary return {rows[-1][2]:7.3f}", flush=True)
    out[f"H={H}"] = dict(team_return=[r[0] for r in rows], critic_estimate=[r[1] for r in rows], adv_return=[r[2] for r in rows])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(scenario=scenario, n_envs=n_envs, n_blocks=n_blocks, curves=out),
          open(os.path.join(ROOT, "gpurun_out", f"learning_{scenario}{'_global' if common else ''}.json"), "w"))
