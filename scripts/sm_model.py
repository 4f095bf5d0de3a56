#!/usr/bin/env python3
"""Model of ONE SM holding 14 units of unequal work (BASELINE config 5's spread) under four placement policies, from the
measured per-warp rates with k = 1..4 warps on a sub-partition (profiles/r02b_warps_per_subpartition.jsonl, scaled to
the shipped kernel's lone-warp time).  DESIGN.md section 3a quotes its output next to the measured times.

    P0  fixed placement (one-warp CTAs: a unit stays on the sub-partition its warp landed on)
    P1  rotation: the active units advance at the average rate of the active warps, spread as evenly as they go
    P2  longest-remaining-first: continuously re-sorted, the unit with most work left on the least crowded sub-partition
    P3  P2, and the two units with least work left wait while 13 or 14 are live (12 active warps = 3 per sub-partition)
"""
import numpy as np

T = {0: 0.0, 1: .01145, 2: .0181, 3: .02206, 4: .0250}      # units of 1 MiB text per ms, per sub-partition with k warps
S = {k: (T[k] / k if k else 0.0) for k in T}


def loads(n):
    base, rem = divmod(n, 4)
    return [base + (i < rem) for i in range(4)]


def sim(w, policy, dt=0.05):
    rem = np.array(w, float)
    t = 0.0
    while (rem > 1e-9).any():
        act = np.where(rem > 1e-9)[0]
        run = act
        if policy == "P3" and len(act) >= 13:
            run = act[np.argsort(-rem[act])][:12]
        ld = loads(len(run))
        speeds = sorted((S[ld[sp]] for sp in range(4) for _ in range(ld[sp])), reverse=True)
        if policy in ("P2", "P3"):
            for u, v in zip(run[np.argsort(-rem[run])], speeds):
                rem[u] -= v * dt
        elif policy == "P1":
            rem[run] -= sum(speeds) / len(run) * dt
        else:
            for sp in range(4):
                mine = [u for u in run if u % 4 == sp]
                for u in mine:
                    rem[u] -= S[len(mine)] * dt
        t += dt
    return t


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    res = {p: [] for p in ("P0", "P1", "P2", "P3")}
    ideal = []
    for _ in range(20):
        ratios = rng.uniform(0.29, 0.56, 14)                 # compressed / plain of config 5's streams
        w = (206 * ratios + 33) / (206 * 0.29 + 33) * 4      # work of a 4 MiB unit, in 1 MiB text units
        for p in res:
            res[p].append(sim(w, p))
        ideal.append(w.sum() / (2 * T[4] + 2 * T[3]))
    for p in res:
        print(p, round(float(np.mean(res[p])), 1), "ms")
    print("perfectly divisible load:", round(float(np.mean(ideal)), 1), "ms")
