#!/usr/bin/env python3
"""Kernel timeline of the training schedules (GPU box): LOC_TIMELINE=1 python scripts/timeline.py [solo|ring|lockstep] [G]
Prints, for a window of steps in the middle of an epoch, every step kernel's start / end (us, relative) per model."""
import ctypes
import os
import sys

import numpy as np

os.environ["LOC_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import torch  # noqa: E402
from locator_b200 import model, _cabi  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "solo"
G = int(sys.argv[2]) if len(sys.argv) > 2 else (1 if mode == "solo" else 4)
K, ntr, nva = 100_000, 810, 90
x, y = bench.synth(ntr + nva, K, 1002)
rng = np.random.default_rng(0)
ctas = model.spare_cluster_l1_ctas() if mode == "ring" else None
ms = [model.LocatorModel(K, seed=10 + g, max_epochs=8, l1_ctas=ctas) for g in range(G)]
gtr = model.PackedGenotypes.from_counts(x[:ntr])
gva = model.PackedGenotypes.from_counts(x[ntr:])
for m in ms:
    m.bind_train(gtr, y[:ntr])
    m.bind_val(gva, y[ntr:])
    m.set_schedule(patience=10 ** 6)
stream = torch.cuda.current_stream().cuda_stream
lib = _cabi.lib
buf = np.zeros(2 * (1 << 16), dtype=np.uint64)


def run(ne):
    if mode == "solo":
        ms[0].train_epochs(np.stack([rng.permutation(ntr) for _ in range(ne)]).astype(np.int32))
    else:
        os.environ["LOC_GROUP_SCHEDULE"] = mode
        perms = [torch.as_tensor(np.stack([rng.permutation(ntr) for _ in range(ne)]).astype(np.int32)).cuda() for _ in ms]
        handles = (ctypes.c_void_p * G)(*[m._h for m in ms])
        pp = (ctypes.c_void_p * G)(*[p.data_ptr() for p in perms])
        _cabi.check(lib.loc_group_train_epochs(handles, G, pp, ne, stream), "group")
        run.keep = perms
    torch.cuda.synchronize()


run(1)
lib.loc_debug_timeline(buf.ctypes.data, 1 << 16)  # drop the warm-up
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run(1)
e1.record()
torch.cuda.synchronize()
n = int(lib.loc_debug_timeline(buf.ctypes.data, 1 << 16))
rec = buf[: 2 * n].reshape(n, 2)
t = rec[:, 0].astype(np.int64)
tag = (rec[:, 1] >> np.uint64(32)).astype(np.int64)
mid = (rec[:, 1] & np.uint64(0xffffffff)).astype(np.int64)
order = np.argsort(t, kind="stable")
t, tag, mid = t[order], tag[order], mid[order]
names = {1: "B.start", 2: "B.end", 3: "H.start", 4: "H.end", 5: "U.start", 6: "U.end", 7: "F.start", 8: "F.end", 9: "W.start",
         10: "W.end", 17: "B.start(last)", 18: "B.end(last)", 19: "H.start(last)", 20: "H.end(last)", 21: "U.start(last)",
         22: "U.end(last)", 23: "H.tiles_ready", 24: "H.published", 25: "B.end(other)"}
print(f"mode {mode} G {G}: epoch of {-(-ntr // 32)} steps x {G} models in {e0.elapsed_time(e1) * 1000:.1f} us "
      f"({e0.elapsed_time(e1) * 1000 / (-(-ntr // 32)) / G:.1f} us per replicate-step); {n} records")
base_id = mid.min()
# per kernel instance: pair starts and ends
t0 = t[0]
lo = int(0.4 * n)
shown = 0
for i in range(lo, n):
    if tag[i] == 25:
        continue
    shown += 1
    if shown > 16 * 3 * G + 8:
        break
    print(f"{(t[i] - t0) / 1000.0:10.2f} us  model {mid[i] - base_id}  {names.get(int(tag[i]), tag[i])}")
# summary: mean duration of B (first block start -> last end), H, U and the mean idle gap between consecutive B kernels
def spans(s_tags, e_tags):
    out = []
    open_ = {}
    for ti, tg, mi in zip(t, tag, mid):
        if tg in s_tags:
            open_.setdefault(mi, []).append(ti)
        elif tg in e_tags and open_.get(mi):
            st = open_[mi]
            if tg == max(e_tags) or len(e_tags) == 1:
                pass
            out.append((min(st), ti, mi))
    return out
bs = [(ti, mi) for ti, tg, mi in zip(t, tag, mid) if tg in (1, 17)]
be = [(ti, mi) for ti, tg, mi in zip(t, tag, mid) if tg in (2, 18, 25)]
if bs and be:
    # group per launch: consecutive records of the same model
    def launches(recs):
        out = []
        for ti, mi in recs:
            if out and out[-1][2] == mi and ti - out[-1][1] < 30000:
                out[-1][0] = min(out[-1][0], ti); out[-1][1] = max(out[-1][1], ti)
            else:
                out.append([ti, ti, mi])
        return out
    ls, le = launches(bs), launches(be)
    m_ = min(len(ls), len(le))
    dur = [(le[i][1] - ls[i][0]) / 1000.0 for i in range(m_)]
    gap = [(ls[i + 1][0] - le[i][1]) / 1000.0 for i in range(m_ - 1)]
    # hand-over latencies of the chained step
    pub = [ti for ti, tg in zip(t, tag) if tg == 24]
    rdy = [ti for ti, tg in zip(t, tag) if tg == 23]
    if pub and rdy and len(le) > 2:
        ends = np.array([e[1] for e in le])
        d1 = [(r_ - ends[ends <= r_].max()) / 1000.0 for r_ in rdy if (ends <= r_).any()]
        d2 = []
        for p_ in pub:
            later = ends[ends > p_]
            if len(later):
                d2.append((later.min() - p_) / 1000.0)
        hs = [(p_ - max([r_ for r_ in rdy if r_ <= p_], default=p_)) / 1000.0 for p_ in pub]
        print(f"chain: last B CTA done -> H sees tiles {np.median(d1):.1f} us; H tiles-ready -> published {np.median(hs):.1f} us; "
              f"H published -> that step's B all done {np.median(d2):.1f} us")
    # per-block end times of every backward launch relative to the publication it waited for (chained step)
    if pub:
        ends_all = np.array(sorted(ti for ti, tg in zip(t, tag) if tg in (2, 18, 25)))
        rows = []
        for a_, b_ in zip(pub[:-1], pub[1:]):
            e = ends_all[(ends_all > a_) & (ends_all <= b_)]
            if len(e) >= 100:
                d = (e - a_) / 1000.0
                rows.append([len(e), d.min(), np.percentile(d, 25), np.median(d), np.percentile(d, 75), np.percentile(d, 95), d.max()])
        if rows:
            r_ = np.median(np.array(rows), axis=0)
            print(f"B block ends after H's publication (median over launches; {int(r_[0])} blocks): min {r_[1]:.1f} p25 {r_[2]:.1f} "
                  f"median {r_[3]:.1f} p75 {r_[4]:.1f} p95 {r_[5]:.1f} max {r_[6]:.1f} us")
    print(f"B launches {m_}: mean duration {np.mean(dur):.1f} us (first block start -> last block end); "
          f"idle between consecutive B: mean {np.mean(gap):.1f} median {np.median(gap):.1f} us")
