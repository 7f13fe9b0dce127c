"""How long is a launch of the persistent Heyvaerts kernel for a given hand-out order of the points?

The per-point cost (rule applications) of the seeded pitchy power-law batch is measured with the host build of
the kernels (tests/hostemu, all host cores), then a list scheduler with W warps replays the launch for the
cost classes of k_classify (rimphony_b200.cu) and for alternatives.  The makespan is printed in units of the
ideal (total work / W).  usage: python tools/schedule_sim.py [N=65536]   (about 2.5 min per 65 536 points on 8 cores)
Output of the run behind DESIGN.md section 4: profiles/r02_schedule_sim.log"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import emu_sweep as E  # noqa: E402
import rimphony_b200 as R  # noqa: E402

W = 148 * 32  # resident warps of k_heyvaerts_fast: 148 SMs x 8 CTAs x 4 warps


def costs(n):
    kind, s, th, params = R.synthetic_batch("pitchy_pl", n, seed=1)
    fx = {"kind": kind, "s": np.asarray(s), "theta": np.asarray(th),
          "params": np.asarray([np.broadcast_to(p, (n,)) for p in params])}
    _, _, info = E.run_fixture(E.load_emu(), fx, which=2, mode=2)
    return fx["s"], fx["theta"], info[2].astype(float)


def makespan(apps, order, w=W):
    heap = [0.0] * w
    heapq.heapify(heap)
    for i in order:
        heapq.heappush(heap, heapq.heappop(heap) + apps[i])
    return max(heap) / (apps.sum() / w)


def three_classes(s, sigma0):   # round 1
    r = np.where(s < 1, 2, np.where(sigma0 < 3, 1, 0))
    r[s < 0.18] = -1            # the reference's NaN region: free, last
    return r


def four_classes(s, sigma0):    # round 2: the points that can run into the application budget first
    r = three_classes(s, sigma0)
    r[(sigma0 < 0.1) & (s >= 0.18) & (s < 10)] = 3
    return r


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    s0, th0, apps0 = costs(n)
    print(f"{n} points: mean {apps0.mean():.0f} applications, max {apps0.max():.0f}, above 15 k: {(apps0 > 15000).sum()}, "
          f"of those with s sin(theta) < 0.1 and s < 10: {((apps0 > 15000) & (s0 * np.sin(th0) < 0.1) & (s0 < 10)).sum()}")
    for rep in (1, 2, 4):
        s, th, apps = np.tile(s0, rep), np.tile(th0, rep), np.tile(apps0, rep)
        sigma0 = s * np.sin(th)
        idx = np.arange(len(s))
        row = {"index order": makespan(apps, idx),
               "three classes": makespan(apps, np.lexsort((idx, -three_classes(s, sigma0)))),
               "four classes": makespan(apps, np.lexsort((idx, -four_classes(s, sigma0)))),
               "sorted by true cost": makespan(apps, np.argsort(-apps)),
               "longest point / ideal": apps.max() / (apps.sum() / W)}
        print(f"n = {len(s):7d}: " + ", ".join(f"{k} {v:.3f}" for k, v in row.items()))


if __name__ == "__main__":
    main()
